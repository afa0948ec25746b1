// cv_io.cpp — file I/O behind the cv:: functions that include/compat/opencv2/opencv.hpp only declares (imread, imwrite,
// VideoCapture, VideoWriter).  TEST INFRASTRUCTURE ONLY: linked into oracle/_ref/align_test_ref and video_test_ref, the
// reference's own drivers (align_test.cpp, video_test.cpp) compiled UNMODIFIED against the drop-in headers.
//
// There is no image or video codec in this image, so "files" are raw dumps with a small header, whatever their extension:
//   "VSRAW1\n" <width> <height> <channels> <frames> "\n" then frames x height x width x channels bytes.
// The tests write input.png / template.png / recordings/*.mp4 in that form and read the drivers' outputs back.
#include <opencv2/opencv.hpp>

#include <execinfo.h>
#include <signal.h>
#include <stdio.h>
#include <unistd.h>

#include <fstream>

namespace {

struct RawHeader {
    int w = 0, h = 0, c = 0, n = 0;
    long data_offset = 0;
};

bool read_header(FILE* f, RawHeader& hd)
{
    char magic[8] = {0};
    if (fread(magic, 1, 7, f) != 7 || std::string(magic, 7) != "VSRAW1\n") return false;
    if (fscanf(f, "%d %d %d %d", &hd.w, &hd.h, &hd.c, &hd.n) != 4) return false;
    if (fgetc(f) != '\n') return false;
    hd.data_offset = ftell(f);
    return hd.w > 0 && hd.h > 0 && (hd.c == 1 || hd.c == 3) && hd.n > 0;
}

// the drivers run as child processes of the tests: line-buffered output and a backtrace on a crash make a failure readable
void on_crash(int sig)
{
    void* frames[64];
    const int n = backtrace(frames, 64);
    const char msg[] = "[cv_io] fatal signal, backtrace:\n";
    (void)!write(2, msg, sizeof(msg) - 1);
    backtrace_symbols_fd(frames, n, 2);
    _exit(128 + sig);
}
struct CrashReport {
    CrashReport()
    {
        signal(SIGSEGV, on_crash);
        signal(SIGABRT, on_crash);
        setvbuf(stdout, nullptr, _IOLBF, 0);
    }
} crash_report;

void write_header(FILE* f, int w, int h, int c, int n) { fprintf(f, "VSRAW1\n%d %d %d %d\n", w, h, c, n); }

}  // namespace

namespace cv {

Mat imread(const std::string& filename, int flags)
{
    FILE* f = fopen(filename.c_str(), "rb");
    if (!f) return Mat();
    RawHeader hd;
    Mat img;
    if (read_header(f, hd)) {
        Mat raw(hd.h, hd.w, hd.c == 3 ? CV_8UC3 : CV_8UC1);
        if (fread(raw.data, 1, (size_t)hd.w * hd.h * hd.c, f) == (size_t)hd.w * hd.h * hd.c) {
            if (flags == IMREAD_COLOR && hd.c == 1) {       // replicate gray into BGR, as imread does
                img = Mat(hd.h, hd.w, CV_8UC3);
                for (int y = 0; y < hd.h; y++)
                    for (int x = 0; x < hd.w; x++)
                        img.ptr(y)[3 * x] = img.ptr(y)[3 * x + 1] = img.ptr(y)[3 * x + 2] = raw.ptr(y)[x];
            } else {
                img = raw;
            }
        }
    }
    fclose(f);
    return img;
}

bool imwrite(const std::string& filename, const Mat& img)
{
    if (img.empty() || img.depth() != CV_8U) return false;
    FILE* f = fopen(filename.c_str(), "wb");
    if (!f) return false;
    write_header(f, img.cols, img.rows, img.channels(), 1);
    for (int y = 0; y < img.rows; y++) fwrite(img.ptr(y), 1, (size_t)img.cols * img.channels(), f);
    fclose(f);
    return true;
}

struct VideoCapture::Impl {
    FILE* f = nullptr;
    RawHeader hd;
    int next = 0;
    ~Impl() { if (f) fclose(f); }
};

VideoCapture::VideoCapture() {}
VideoCapture::VideoCapture(const std::string& filename) { open(filename); }
VideoCapture::~VideoCapture() {}
bool VideoCapture::open(const std::string& filename)
{
    impl_.reset(new Impl());
    impl_->f = fopen(filename.c_str(), "rb");
    if (!impl_->f || !read_header(impl_->f, impl_->hd) || impl_->hd.c != 3) { impl_.reset(); return false; }
    return true;
}
bool VideoCapture::isOpened() const { return (bool)impl_; }
double VideoCapture::get(int prop) const
{
    if (!impl_) return 0;
    switch (prop) {
    case CAP_PROP_FPS: return 30.0;
    case CAP_PROP_FRAME_WIDTH: return impl_->hd.w;
    case CAP_PROP_FRAME_HEIGHT: return impl_->hd.h;
    case CAP_PROP_FRAME_COUNT: return impl_->hd.n;
    default: return 0;
    }
}
bool VideoCapture::read(Mat& frame)
{
    if (!impl_ || impl_->next >= impl_->hd.n) return false;
    frame.create(impl_->hd.h, impl_->hd.w, CV_8UC3);      // a decoder reuses its buffer too
    const size_t bytes = (size_t)impl_->hd.w * impl_->hd.h * 3;
    if (fread(frame.data, 1, bytes, impl_->f) != bytes) return false;
    impl_->next++;
    return true;
}
void VideoCapture::release() { impl_.reset(); }

struct VideoWriter::Impl {
    std::string path;
    FILE* f = nullptr;
    Size size;
    int frames = 0, empty = 0;
    ~Impl() { close(); }
    void close()
    {
        if (!f) return;
        fclose(f);
        f = nullptr;
        // the frame count is only known now: rewrite the header in place (fixed-width count field)
        FILE* g = fopen(path.c_str(), "r+b");
        if (g) {
            fprintf(g, "VSRAW1\n%d %d %d %09d\n", size.width, size.height, 3, frames);
            fclose(g);
        }
        printf("[cv_io] %s: %d frames written, %d empty frames skipped\n", path.c_str(), frames, empty);
    }
};

VideoWriter::VideoWriter() {}
VideoWriter::~VideoWriter() {}
bool VideoWriter::open(const std::string& filename, int, double, Size frameSize, bool)
{
    impl_.reset(new Impl());
    impl_->path = filename;
    impl_->size = frameSize;
    impl_->f = fopen(filename.c_str(), "wb");
    if (!impl_->f) { impl_.reset(); return false; }
    fprintf(impl_->f, "VSRAW1\n%d %d %d %09d\n", frameSize.width, frameSize.height, 3, 0);
    return true;
}
bool VideoWriter::isOpened() const { return (bool)impl_; }
bool VideoWriter::set(int, double) { return true; }
void VideoWriter::write(const Mat& frame)
{
    if (!impl_ || !impl_->f) return;
    // video_test.cpp:109 writes the (empty) cv::Mat of the first `lag` frames too; OpenCV drops those
    if (frame.empty()) { impl_->empty++; return; }
    if (frame.cols != impl_->size.width || frame.rows != impl_->size.height || frame.type() != CV_8UC3) return;
    for (int y = 0; y < frame.rows; y++) fwrite(frame.ptr(y), 1, (size_t)frame.cols * 3, impl_->f);
    impl_->frames++;
}
void VideoWriter::release() { if (impl_) impl_->close(); impl_.reset(); }

}  // namespace cv
