// CPU test hooks: compiles the PRODUCT's host+device headers with g++ so that the
// introselect replay and the 4x4 SVD inverse that run on the GPU can be checked on the
// CPU (against the real std::nth_element and cv2 fixtures) without a device.
#include <stdint.h>
#include "vs_introselect.cuh"
#include "vs_linalg4.cuh"

extern "C" {

// keys: (abs_delta << th_key_shift() | index); replays std::nth_element(keys, keys+nth, keys+n)
int th_key_shift() { return vs_sel::KEY_SHIFT; }
void th_nth_element(uint32_t* keys, int n, int nth) { vs_sel::nth_element_serial(keys, n, nth); }
int th_selected_count(int n, float fraction) { return vs_sel::selected_count(n, fraction); }
// forces the heap-select fallback from the first round (depth limit 0)
void th_introselect_depth(uint32_t* keys, int n, int nth, int depth) { vs_sel::introselect_from(keys, 0, nth, n, depth); }

void th_svd4(const double* H, double* w, double* u, double* vt) { vs_svd4(H, w, u, vt); }
double th_condition_and_invert(double* H, double* Hinv) { return vs_condition_and_invert(H, Hinv); }

}
