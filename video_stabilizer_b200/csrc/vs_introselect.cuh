// vs_introselect.cuh — exact emulation of libstdc++ 13 std::nth_element on packed keys.
//
// alignment.cpp:460-486 keeps the (size_t)(N*0.8f) DeltaPixels with the smallest
// abs_delta by std::nth_element.  abs_delta only takes a few hundred distinct values, so
// WHICH of the tied elements survive is an artefact of libstdc++'s introselect
// (bits/stl_algo.h: __introselect / __unguarded_partition_pivot / __move_median_to_first /
// __heap_select / __insertion_sort, bits/stl_heap.h: __make_heap / __pop_heap /
// __adjust_heap / __push_heap) and it changes the recovered transform by more than the
// 0.01 px parity budget (SURVEY.md finding 3).  The functions here replay that algorithm
// move for move on keys = (abs_delta << KEY_SHIFT | tile_index); only the abs_delta bits take
// part in comparisons, exactly like the reference's comparator.  KEY_SHIFT = 17 leaves 17 bits
// for the tile (131 071 tiles per level: 8K has 82 944) and 15 for abs_delta: |template - sample|
// of u8 images stays below 1024 (the normalised Lanczos-2 weights sum to at most 1.4 in absolute
// value per axis), so the clamp at 32 767 never acts.
//
// Host+device so the emulation is checked against the real std::nth_element on the CPU
// (tests/test_introselect.py) before it ever runs on a GPU.
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef __CUDACC__
#define VS_SEL_HD __host__ __device__ __forceinline__
#else
#define VS_SEL_HD inline
#endif

namespace vs_sel {

constexpr int KEY_SHIFT = 17;
constexpr uint32_t KEY_TILE_MASK = (1u << KEY_SHIFT) - 1u;
constexpr uint32_t KEY_VALUE_MAX = (1u << (32 - KEY_SHIFT)) - 1u;
VS_SEL_HD uint32_t make_key(uint32_t abs_delta, uint32_t tile)
{
    return ((abs_delta < KEY_VALUE_MAX ? abs_delta : KEY_VALUE_MAX) << KEY_SHIFT) | tile;
}
VS_SEL_HD uint32_t key_value(uint32_t k) { return k >> KEY_SHIFT; }
VS_SEL_HD uint32_t key_tile(uint32_t k) { return k & KEY_TILE_MASK; }
VS_SEL_HD bool less(uint32_t a, uint32_t b) { return (a >> KEY_SHIFT) < (b >> KEY_SHIFT); }
VS_SEL_HD void swp(uint32_t* a, int i, int j) { uint32_t t = a[i]; a[i] = a[j]; a[j] = t; }

// std::__lg
VS_SEL_HD int lg(int n) { int k = 0; while (n > 1) { n >>= 1; k++; } return k; }

// std::__move_median_to_first(result, a, b, c)
VS_SEL_HD void move_median_to_first(uint32_t* v, int result, int a, int b, int c)
{
    if (less(v[a], v[b])) {
        if (less(v[b], v[c])) swp(v, result, b);
        else if (less(v[a], v[c])) swp(v, result, c);
        else swp(v, result, a);
    } else if (less(v[a], v[c])) swp(v, result, a);
    else if (less(v[b], v[c])) swp(v, result, c);
    else swp(v, result, b);
}

// std::__unguarded_partition(first, last, pivot)
VS_SEL_HD int unguarded_partition(uint32_t* v, int first, int last, int pivot)
{
    const uint32_t pv = v[pivot];
    while (true) {
        while (less(v[first], pv)) ++first;
        --last;
        while (less(pv, v[last])) --last;
        if (!(first < last)) return first;
        swp(v, first, last);
        ++first;
    }
}

// std::__push_heap(first, holeIndex, topIndex, value)
VS_SEL_HD void push_heap(uint32_t* v, int first, int hole, int top, uint32_t value)
{
    int parent = (hole - 1) / 2;
    while (hole > top && less(v[first + parent], value)) {
        v[first + hole] = v[first + parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    v[first + hole] = value;
}

// std::__adjust_heap(first, holeIndex, len, value)
VS_SEL_HD void adjust_heap(uint32_t* v, int first, int hole, int len, uint32_t value)
{
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (less(v[first + child], v[first + (child - 1)])) child--;
        v[first + hole] = v[first + child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        v[first + hole] = v[first + (child - 1)];
        hole = child - 1;
    }
    push_heap(v, first, hole, top, value);
}

// std::__heap_select(first, middle, last)
VS_SEL_HD void heap_select(uint32_t* v, int first, int middle, int last)
{
    const int len = middle - first;
    if (len >= 2) {   // std::__make_heap
        int parent = (len - 2) / 2;
        while (true) {
            uint32_t value = v[first + parent];
            adjust_heap(v, first, parent, len, value);
            if (parent == 0) break;
            parent--;
        }
    }
    for (int i = middle; i < last; ++i)
        if (less(v[i], v[first])) {   // std::__pop_heap(first, middle, i)
            uint32_t value = v[i];
            v[i] = v[first];
            adjust_heap(v, first, 0, len, value);
        }
}

// std::__insertion_sort(first, last)
VS_SEL_HD void insertion_sort(uint32_t* v, int first, int last)
{
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
        uint32_t val = v[i];
        if (less(val, v[first])) {
            for (int j = i; j > first; --j) v[j] = v[j - 1];   // move_backward
            v[first] = val;
        } else {                                               // __unguarded_linear_insert
            int lastp = i, next = i - 1;
            while (less(val, v[next])) { v[lastp] = v[next]; lastp = next; --next; }
            v[lastp] = val;
        }
    }
}

// One serial round of __introselect's loop body on [first,last).  Returns the cut.
VS_SEL_HD int partition_pivot(uint32_t* v, int first, int last)
{
    int mid = first + (last - first) / 2;
    move_median_to_first(v, first, first + 1, mid, last - 1);
    return unguarded_partition(v, first + 1, last, first);
}

// std::__introselect continued from an arbitrary state (used by the parallel kernel for
// the short tail and by the serial reference path).
VS_SEL_HD void introselect_from(uint32_t* v, int first, int nth, int last, int depth_limit)
{
    while (last - first > 3) {
        if (depth_limit == 0) {
            heap_select(v, first, nth + 1, last);
            swp(v, first, nth);
            return;
        }
        --depth_limit;
        int cut = partition_pivot(v, first, last);
        if (cut <= nth) first = cut; else last = cut;
    }
    insertion_sort(v, first, last);
}

// std::nth_element(v, v+nth, v+n)
VS_SEL_HD void nth_element_serial(uint32_t* v, int n, int nth)
{
    if (n == 0 || nth == n) return;
    introselect_from(v, 0, nth, n, lg(n) * 2);
}

// selected count of alignment.cpp:464-465: (size_t)(size * fraction), float product
// alignment.cpp:464-465: (size_t)(N * fraction) in float; callers validate 0 <= fraction <= 1, the clamp keeps an
// out-of-contract value from indexing past the keys
VS_SEL_HD int selected_count(int n, float fraction)
{
    const int k = (int)(size_t)((float)n * fraction);
    return k < 0 ? 0 : (k > n ? n : k);
}

}  // namespace vs_sel
