// Test infrastructure (oracle/): what the DETECTION build of the compiled reference needs on top of oracle/ref_stubs/ (included at the end of its commons.pc.h
// when REF_DETECT is defined; oracle/build_ref.py builds oracle/_ref/libref_detect.so with it):
//   commons/common/processing/RawProcessing.cpp      RawProcessing::generate_binary
//   tracker/python/BackgroundSubtraction.cpp         BackgroundSubtraction::apply / Data::set
// Those two files are OpenCV call sequences.  OpenCV itself is third party and has no headers in this image, but the REAL library is here as Python's
// cv2 (4.13): every cv:: function of the detection path forwards to a callback that the test installs (tests/test_oracle_ref_detect.py) and that runs the
// real cv2 function on the same bytes.  So the reference's control flow runs as compiled, and the pixel arithmetic is OpenCV's own -- nothing in between is
// restated.  Constants carry OpenCV's numeric values so that they can be handed to cv2 unchanged.  The tag branch of RawProcessing.cpp (contours, drawing)
// only has to compile; its functions abort.  Never linked into the product.
#pragma once
#include <optional>

namespace cv {
using UMat = Mat;
struct Point { int x = 0, y = 0; Point() = default; template<typename A, typename B> Point(A x, B y) : x(int(x)), y(int(y)) {} };
template<typename T, int N> struct VecN { T v[N] = {}; T& operator[](int i) { return v[i]; } const T& operator[](int i) const { return v[i]; } };
using Vec2i = VecN<int, 2>;
using Vec4i = VecN<int, 4>;
struct Range { int start = 0, end = 0; };
struct ParallelLoopBody { virtual ~ParallelLoopBody() {} virtual void operator()(const Range&) const = 0; };
template<typename T, typename S> inline T saturate_cast(S v) { return T(std::clamp<double>(double(v), 0., 255.)); }

enum { COLOR_BGRA2BGR = 1, COLOR_BGRA2RGBA = 5, COLOR_GRAY2BGRA = 9, COLOR_BGRA2GRAY = 10 };
enum { THRESH_BINARY = 0, THRESH_TOZERO = 3, ADAPTIVE_THRESH_MEAN_C = 0, MORPH_ELLIPSE = 2, RETR_CCOMP = 2, CHAIN_APPROX_SIMPLE = 2, FONT_HERSHEY_PLAIN = 1 };

// ---- the bridge to the real OpenCV ----
struct BridgeMat { int32_t rows, cols, type, channels; int64_t step; unsigned char *data; };
// op: the function's name; a, b: inputs (b may be null); out: an opaque cv::Mat* the callback sizes with ref_cv_out(out, rows, cols, type) and then fills;
// params: the call's scalar arguments in order
using bridge_fn_t = void (*)(const char *op, const BridgeMat *a, const BridgeMat *b, void *out, const double *params, int32_t n_params);
inline bridge_fn_t& bridge() { static bridge_fn_t f = nullptr; return f; }
inline BridgeMat describe(const Mat& m) { return BridgeMat{m.rows, m.cols, m._type, m.channels(), (int64_t)m.step.p[0], m.data}; }
inline void forward(const char *op, const Mat *a, const Mat *b, const Mat& out, std::initializer_list<double> params)
{
    if (!bridge()) { std::fprintf(stderr, "cv::%s: no OpenCV bridge installed\n", op); std::abort(); }
    BridgeMat da{}, db{};
    if (a) da = describe(*a);
    if (b) db = describe(*b);
    std::vector<double> p(params);
    bridge()(op, a ? &da : nullptr, b ? &db : nullptr, const_cast<Mat *>(&out), p.data(), (int32_t)p.size());
}

// outputs are OpenCV OutputArrays: they bind to temporaries and const references as well
inline void cvtColor(const Mat& src, const Mat& dst, int code) { forward("cvtColor", &src, nullptr, dst, {double(code)}); }
inline void absdiff(const Mat& a, const Mat& b, const Mat& dst) { forward("absdiff", &a, &b, dst, {}); }
inline void subtract(const Mat& a, const Mat& b, const Mat& dst) { forward("subtract", &a, &b, dst, {}); }
inline void subtract(double a, const Mat& b, const Mat& dst) { forward("subtract_scalar_mat", &b, nullptr, dst, {a}); }
inline double threshold(const Mat& src, const Mat& dst, double thresh, double maxval, int type) { forward("threshold", &src, nullptr, dst, {thresh, maxval, double(type)}); return thresh; }
inline void blur(const Mat& src, const Mat& dst, Size k) { forward("blur", &src, nullptr, dst, {double(k.width), double(k.height)}); }
inline void inRange(const Mat& src, double lo, double hi, const Mat& dst) { forward("inRange", &src, nullptr, dst, {lo, hi}); }
inline void adaptiveThreshold(const Mat& src, const Mat& dst, double maxval, int method, int type, int block, double c) { forward("adaptiveThreshold", &src, nullptr, dst, {maxval, double(method), double(type), double(block), c}); }
inline void dilate(const Mat& src, const Mat& dst, const Mat& kernel) { forward("dilate", &src, &kernel, dst, {}); }
inline void erode(const Mat& src, const Mat& dst, const Mat& kernel) { forward("erode", &src, &kernel, dst, {}); }
inline void bitwise_and(const Mat& a, const Mat& b, const Mat& dst) { forward("bitwise_and", &a, &b, dst, {}); }
inline void bitwise_or(const Mat& a, const Mat& b, const Mat& dst) { forward("bitwise_or", &a, &b, dst, {}); }
inline void add(const Mat& a, const Mat& b, const Mat& dst) { forward("add", &a, &b, dst, {}); }
inline void max(const Mat& a, const Mat& b, const Mat& dst) { forward("max", &a, &b, dst, {}); }
inline void min(const Mat& a, const Mat& b, const Mat& dst) { forward("min", &a, &b, dst, {}); }
inline void divide(const Mat& a, double b, const Mat& dst) { forward("divide_mat_scalar", &a, nullptr, dst, {b}); }      // a 1 x 1 double divides every channel
inline void Mat::convert_32f_to_8u(const Mat& src, Mat& dst) { forward("convertTo_8u", &src, nullptr, dst, {}); }
inline void equalizeHist(const Mat& src, const Mat& dst) { forward("equalizeHist", &src, nullptr, dst, {}); }
inline Mat getStructuringElement(int shape, Size k, Point anchor) { Mat m; forward("getStructuringElement", nullptr, nullptr, m, {double(shape), double(k.width), double(k.height), double(anchor.x), double(anchor.y)}); return m; }
// pure data movement: done here
inline void split(const Mat& src, Mat *planes)
{
    const int n = src.channels();
    for (int c = 0; c < n; ++c) {
        Mat m(src.rows, src.cols, CV_8UC1);
        for (int y = 0; y < src.rows; ++y) for (int x = 0; x < src.cols; ++x) m.ptr(y)[x] = src.ptr(y)[x * n + c];
        planes[c] = m;
    }
}
inline void split(const Mat& src, std::vector<Mat>& planes) { planes.resize((size_t)src.channels()); split(src, planes.data()); }
inline void merge(const std::vector<Mat>& planes, Mat& dst)
{
    const int n = (int)planes.size();
    Mat m(planes[0].rows, planes[0].cols, (n - 1) << 3);
    for (int c = 0; c < n; ++c) for (int y = 0; y < m.rows; ++y) for (int x = 0; x < m.cols; ++x) m.ptr(y)[x * n + c] = planes[(size_t)c].ptr(y)[x];
    dst = m;
}
// the tag branch of RawProcessing.cpp (tags_enable; out of scope): has to compile, is never run
#define REF_CV_UNUSED(NAME, RET) template<typename... A> inline RET NAME(A&&...) { std::fprintf(stderr, "cv::" #NAME " stand-in used\n"); std::abort(); }
REF_CV_UNUSED(drawContours, void)
REF_CV_UNUSED(arcLength, double)
REF_CV_UNUSED(approxPolyDP, void)
REF_CV_UNUSED(putText, void)
REF_CV_UNUSED(contourArea, double)
REF_CV_UNUSED(line, void)
REF_CV_UNUSED(getRotationMatrix2D, Mat)
REF_CV_UNUSED(findContours, void)
#undef REF_CV_UNUSED
}

typedef cv::Mat gpuMat;                                         // commons.pc.h:426-430 (cv::UMat with OpenCL, cv::Mat without)
#define CV_32FC(n) (5 + (((n) - 1) << 3))

#include <misc/FormatColor.h>
namespace cmn {
template<FormatterType, typename... A> inline std::string format(const A&...) { return std::string(); }
namespace gui {
struct Color { uint8_t r = 0, g = 0, b = 0, a = 255; constexpr Color() = default; constexpr Color(uint8_t r, uint8_t g, uint8_t b, uint8_t a = 255) : r(r), g(g), b(b), a(a) {} };
constexpr Color Red{255, 0, 0}, Cyan{0, 255, 255};
}
template<typename T> inline std::string hex(T) { return std::string(); }
// misc/detail.h:533-559: every pixel through vec_to_r3g3b2 (restated in misc/detail_bits.h)
template<int channels> inline void convert_to_r3g3b2(const cv::Mat& input, cv::Mat& output)
{
    if (output.rows != input.rows || output.cols != input.cols || output.type() != CV_8UC1) output = cv::Mat::zeros(input.rows, input.cols, CV_8UC1);
    for (int y = 0; y < input.rows; ++y) for (int x = 0; x < input.cols; ++x) output.ptr(y)[x] = vec_to_r3g3b2(input.ptr(y) + x * channels);
}
}
