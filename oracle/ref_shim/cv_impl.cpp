// cv_impl.cpp — CPU definitions of the cv:: functions that include/compat/opencv2/opencv.hpp
// only declares.  TEST INFRASTRUCTURE ONLY: linked into oracle/_ref so the reference's own
// host sources (alignment.cpp, imgproc.cpp, stabilizer.cpp) run against them; never part of
// the product.  Each function forwards to the oracle restatement that is pinned against
// cv2 4.13 golden fixtures (tests/golden/).
#include <opencv2/opencv.hpp>

#include <math.h>

#include "../vs_oracle.h"

namespace cv {

// cv::cvtColor(BGR2GRAY) — alignment.cpp:212
void cvtColor(const Mat& src, Mat& dst, int code)
{
    if (code != COLOR_BGR2GRAY || src.type() != CV_8UC3) throw std::runtime_error("oracle cv shim: cvtColor supports CV_8UC3 BGR2GRAY only");
    Mat out(src.rows, src.cols, CV_8UC1);
    for (int r = 0; r < src.rows; r++) vo_bgr2gray(src.ptr(r), src.cols, 1, out.ptr(r));
    dst = out;
}

// cv::warpAffine(INTER_LINEAR, BORDER_CONSTANT) — imgproc.cpp:473-481.  Classic fixed-point
// path: f64 inverse of M (unless WARP_INVERSE_MAP), AB_BITS=10 coordinates rounded to 1/32 px,
// 15-bit weights, taps outside the image read the border value (0).
void warpAffine(const Mat& src, Mat& dst, const Mat& M, Size dsize, int flags, int borderMode, const Scalar& bv)
{
    if (src.depth() != CV_8U || (src.channels() != 3 && src.channels() != 1)) throw std::runtime_error("oracle cv shim: warpAffine supports 8-bit 1/3 channel images only");
    if ((flags & 7) != INTER_LINEAR || borderMode != BORDER_CONSTANT) throw std::runtime_error("oracle cv shim: warpAffine supports INTER_LINEAR + BORDER_CONSTANT only");
    if (M.rows != 2 || M.cols != 3 || M.type() != CV_64F) throw std::runtime_error("oracle cv shim: warpAffine needs a 2x3 CV_64F matrix");
    const int cn = src.channels();
    double m[6] = {M.at<double>(0, 0), M.at<double>(0, 1), M.at<double>(0, 2), M.at<double>(1, 0), M.at<double>(1, 1), M.at<double>(1, 2)};
    if (!(flags & WARP_INVERSE_MAP)) {
        double D = m[0] * m[4] - m[1] * m[3];
        D = D != 0 ? 1. / D : 0;
        double A11 = m[4] * D, A22 = m[0] * D;
        m[0] = A11; m[1] *= -D; m[3] *= -D; m[4] = A22;
        double b1 = -m[0] * m[2] - m[1] * m[5];
        double b2 = -m[3] * m[2] - m[4] * m[5];
        m[2] = b1; m[5] = b2;
    }
    Mat out(dsize.height, dsize.width, src.type());
    std::vector<int> adelta(dsize.width), bdelta(dsize.width);
    for (int x = 0; x < dsize.width; x++) {
        adelta[x] = (int)lrint(m[0] * x * 1024);
        bdelta[x] = (int)lrint(m[3] * x * 1024);
    }
    for (int y = 0; y < dsize.height; y++) {
        const int X0 = (int)lrint((m[1] * y + m[2]) * 1024) + 16;
        const int Y0 = (int)lrint((m[4] * y + m[5]) * 1024) + 16;
        uint8_t* d = out.ptr(y);
        for (int x = 0; x < dsize.width; x++) {
            const int X = (X0 + adelta[x]) >> 5, Y = (Y0 + bdelta[x]) >> 5;
            const int sx = X >> 5, sy = Y >> 5, fx = X & 31, fy = Y & 31;
            const int w00 = (32 - fx) * (32 - fy) * 32, w10 = fx * (32 - fy) * 32;
            const int w01 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
            for (int c = 0; c < cn; c++) {
                auto tap = [&](int xx, int yy) -> int {
                    if (xx < 0 || xx >= src.cols || yy < 0 || yy >= src.rows) return (int)bv.val[c];
                    return src.ptr(yy)[xx * cn + c];
                };
                int v = w00 * tap(sx, sy) + w10 * tap(sx + 1, sy) + w01 * tap(sx, sy + 1) + w11 * tap(sx + 1, sy + 1);
                d[x * cn + c] = (uint8_t)((v + 16384) >> 15);
            }
        }
    }
    dst = out;
}

static void load4x4(const Mat& m, double H[16])
{
    if (m.rows != 4 || m.cols != 4 || m.type() != CV_64F) throw std::runtime_error("oracle cv shim: SVD / inv support 4x4 CV_64F only");
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) H[r * 4 + c] = m.at<double>(r, c);
}

// cv::SVD(H) — alignment.cpp:558
SVD::SVD(const Mat& src, int)
{
    double H[16], sw[4], su[16], svt[16];
    load4x4(src, H);
    vo_svd4(H, sw, su, svt);
    w = Mat(4, 1, CV_64F); u = Mat(4, 4, CV_64F); vt = Mat(4, 4, CV_64F);
    for (int i = 0; i < 4; i++) w.at<double>(i, 0) = sw[i];
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) { u.at<double>(r, c) = su[r * 4 + c]; vt.at<double>(r, c) = svt[r * 4 + c]; }
}

// H.inv(cv::DECOMP_SVD) — alignment.cpp:582
Mat Mat::inv(int method) const
{
    if (method != DECOMP_SVD) throw std::runtime_error("oracle cv shim: Mat::inv supports DECOMP_SVD only");
    double H[16], Hinv[16];
    load4x4(*this, H);
    vo_inv4_svd(H, Hinv);
    Mat out(4, 4, CV_64F);
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) out.at<double>(r, c) = Hinv[r * 4 + c];
    return out;
}

// Hinv * b — alignment.cpp:624 (cv::gemm on f64: plain row-by-column accumulation)
Mat operator*(const Mat& a, const Mat& b)
{
    if (a.type() != CV_64F || b.type() != CV_64F || a.cols != b.rows) throw std::runtime_error("oracle cv shim: operator* needs conforming CV_64F matrices");
    Mat out(a.rows, b.cols, CV_64F);
    for (int r = 0; r < a.rows; r++)
        for (int c = 0; c < b.cols; c++) {
            double s = 0;
            for (int k = 0; k < a.cols; k++) s += a.at<double>(r, k) * b.at<double>(k, c);
            out.at<double>(r, c) = s;
        }
    return out;
}

// alignment.cpp:374 — the oracle's restatement of cv::phaseCorrelate (direct f64 DFT, see vs_oracle.cpp)
Point2d phaseCorrelate(const Mat& src1, const Mat& src2, _NoArray, double* response)
{
    if (src1.type() != CV_32F || src2.type() != CV_32F || src1.rows != src2.rows || src1.cols != src2.cols)
        throw std::runtime_error("oracle cv shim: phaseCorrelate needs two CV_32F images of one size");
    if (src1.step[0] != src2.step[0]) throw std::runtime_error("oracle cv shim: phaseCorrelate needs equal row steps");
    double out[3];
    vo_phase_correlate(src1.ptr<float>(0), src2.ptr<float>(0), src1.cols, src1.rows, (int64_t)(src1.step[0] / sizeof(float)), out);
    if (response) *response = out[2];
    return Point2d(out[0], out[1]);
}

}  // namespace cv
