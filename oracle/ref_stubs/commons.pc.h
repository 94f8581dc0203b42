// Test infrastructure (oracle/): a MINIMAL stand-in for TRex's precompiled header `commons.pc.h`, so that a few of the reference's own source files
// compile here, unmodified and from where they lie under /root/reference, without OpenCV / glaze / cnpy (oracle/build_ref.py):
//   commons/common/misc/CircularGraph.cpp   eft, ieft, curvature, differentiate, find_peaks and its fast::cos polynomial
//   commons/common/misc/curve_discussion.cpp, commons/common/gui/Transform.cpp
//   tracker/tracking/Outline.cpp            Outline::resample / smooth / offset_to_middle / calculate_midline, Midline::post_process / normalize / fix_length
// Only the declarations those files touch are provided, written for this purpose.  The arithmetic is NOT restated: the vector types and the scalar maths are
// the reference's own headers, included from the checkout --
//   commons/common/misc/vec2.h (+ vec2.cpp)   Vec2 / Size2 / Bounds with every operator, length / normalize / atan2, sqdistance, euclidean_distance
//   commons/common/misc/math.h                cmn::sqrt / sin / cos / atan2 (float arguments call the f-suffixed C functions), fast_atan2, abs, isnan,
//                                             t_circle_line, crosses_zero, infinity<T>
// -- like misc/IllegalVector.h, misc/EnumClass.h, misc/matharray.h, misc/bid.h, misc/Median.h, processing/encoding.h and every header the compiled .cpp
// files bring themselves (Outline.h, CircularGraph.h, PixelTree.h, Background.h, PVBlob.h ...).  Restated here, citing their lines:
//   SQR, DEGREE, RADIANS, GETTER*    commons.pc.h:444-457
//   cmn::min / cmn::max              commons.pc.h:462-528 (mixed arithmetic types: computed in the wider of the two, the second on a tie)
//   narrow_cast                      commons.pc.h (value-preserving static_cast);  saturate  misc/detail.h:225-229;  Range / arange  misc/ranges.h (misc/ranges.h of this directory)
// Printing, timing, exceptions, drawing and OpenCV types are inert.  Nothing under trex_b200/ includes this.
#pragma once
#include <algorithm>
#include <array>
#include <atomic>
#include <cassert>
#include <climits>
#include <condition_variable>
#include <cfloat>
#include <cmath>
#include <concepts>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <expected>
#include <functional>
#include <limits>
#include <map>
#include <memory>
#include <mutex>
#include <numeric>
#include <optional>
#include <set>
#include <shared_mutex>
#include <stdexcept>
#include <string>
#include <string_view>
#include <tuple>
#include <unordered_map>
#include <vector>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#define SQR(X) ((X)*(X))
#define DEGREE(radians) ((radians) * (cmn::ScalarType(1.0) / cmn::ScalarType(M_PI) * cmn::ScalarType(180)))
#define RADIANS(degree) ((degree) * (cmn::ScalarType(1.0) / cmn::ScalarType(180) * cmn::ScalarType(M_PI)))
#define GETTER(TYPE, VAR) public: [[nodiscard]] const TYPE& VAR() const { return _##VAR; } protected: TYPE _##VAR
#define GETTER_NCONST(TYPE, VAR) public: [[nodiscard]] inline const TYPE& VAR() const { return _##VAR; } [[nodiscard]] inline TYPE& VAR() { return _##VAR; } protected: TYPE _##VAR
#define UNUSED(X) (void)(X)
#define GETTER_I(TYPE, VAR, INIT) public: [[nodiscard]] const TYPE& VAR() const { return _##VAR; } protected: TYPE _##VAR = INIT
#define GETTER_SETTER(TYPE, VAR) public: [[nodiscard]] inline const TYPE& VAR() const { return _##VAR; } inline void set_##VAR(const TYPE& value) { _##VAR = value; } protected: TYPE _##VAR
#define GETTER_SETTER_I(TYPE, VAR, INIT) public: [[nodiscard]] inline const TYPE& VAR() const { return _##VAR; } inline void set_##VAR(const TYPE& value) { _##VAR = value; } protected: TYPE _##VAR = INIT

using long_t = int32_t;

namespace cmn {
using Float2_t = float;
typedef float ScalarType;

template<typename To, typename From> constexpr To narrow_cast(From&& v) { return static_cast<To>(v); }
template<typename To, typename From> constexpr To sign_cast(From&& v) { return static_cast<To>(v); }

}
// the REFERENCE'S OWN commons/common/misc/math.h from the checkout: cmn::cos / sin / sqrt / atan2 (float arguments call the f-suffixed C functions), fast_atan /
// fast_atan2, abs, isnan / isinf, sqdistance / euclidean_distance, t_circle_line, crosses_zero, infinity<T>, next_pow2 ...  It needs only these OpenCV look-alikes
// (three of its templates mention them; none is instantiated by the compiled files)
namespace cv {
struct Point2f { float x = 0, y = 0; Point2f() = default; Point2f(float x, float y) : x(x), y(y) {} float dot(const Point2f& o) const { return x * o.x + y * o.y; } };
template<typename T, int m, int n> struct Matx { T v[m * n] = {}; T operator()(int i) const { return v[i]; } };
template<typename T> struct Mat_ {
    int rows = 0, cols = 0; std::vector<T> v;
    Mat_() = default;
    Mat_(int r, int c) : rows(r), cols(c), v((size_t)r * c) {}
    T& operator()(int r, int c) { return v[(size_t)r * cols + c]; }
    Mat_ operator*(const Mat_& o) const { Mat_ out(rows, o.cols); for (int i = 0; i < rows; ++i) for (int j = 0; j < o.cols; ++j) { T a{}; for (int k = 0; k < cols; ++k) a += v[(size_t)i * cols + k] * o.v[(size_t)k * o.cols + j]; out(i, j) = a; } return out; }
};
}
namespace cmn {
// cmn::min / cmn::max (commons.pc.h:462-528) -- math.h calls them
template<typename A, typename B> requires (std::is_arithmetic_v<std::remove_cvref_t<A>> && std::is_arithmetic_v<std::remove_cvref_t<B>>)
constexpr auto min(A&& a, B&& b) { using A_ = std::remove_cvref_t<A>; using B_ = std::remove_cvref_t<B>; using R = std::conditional_t<(sizeof(A_) > sizeof(B_)), A_, B_>; return std::min(R(a), R(b)); }
template<typename A, typename B> requires (std::is_arithmetic_v<std::remove_cvref_t<A>> && std::is_arithmetic_v<std::remove_cvref_t<B>>)
constexpr auto max(A&& a, B&& b) { using A_ = std::remove_cvref_t<A>; using B_ = std::remove_cvref_t<B>; using R = std::conditional_t<(sizeof(A_) > sizeof(B_)), A_, B_>; return std::max(R(a), R(b)); }
template<typename A, typename B, typename Cc> requires (std::is_arithmetic_v<A> && std::is_arithmetic_v<B> && std::is_arithmetic_v<Cc>)
constexpr auto min(const A& x, const B& y, const Cc& z) -> decltype(x + y + z) { using R = decltype(x + y + z); return std::min(R(x), std::min(R(y), R(z))); }
template<typename A, typename B, typename Cc> requires (std::is_arithmetic_v<A> && std::is_arithmetic_v<B> && std::is_arithmetic_v<Cc>)
constexpr auto max(const A& x, const B& y, const Cc& z) -> decltype(x + y + z) { using R = decltype(x + y + z); return std::max(R(x), std::max(R(y), R(z))); }

namespace check_abs_detail {            // misc/useful_concepts.h:186-196
template<typename T> concept has_coordinates = requires(T t) { { t.x } -> std::convertible_to<float>; };
template<typename T> concept has_get = requires(T t) { { t.get() } -> std::convertible_to<float>; };
}
}
#include <misc/math.h>
namespace cmn {
}
// ---- the REFERENCE'S OWN commons/common/misc/vec2.h from the checkout: Vec2, Size2, Bounds with every operator, length / normalize / atan2, sqdistance ... ----
// (commons/common/misc/vec2.cpp is compiled with it: the explicit instantiations, Bounds::restrict_to / distance.)  What it expects from TRex's precompiled
// header before it: OpenCV's Point_ / Size_ / Rect_ templates and cv::Mat (below), the is_numeric concept (misc/useful_concepts.h:175), saturate
// (misc/detail.h:225-229), and glaze / Meta names that only its string and JSON helpers touch (inert here).
using uint = unsigned int;
namespace cv {
template<typename T> struct Point_ { T x{}, y{}; Point_() = default; template<typename A, typename B> Point_(A x_, B y_) : x(T(x_)), y(T(y_)) {} };
template<typename T> struct Size_ { T width{}, height{}; Size_() = default; template<typename A, typename B> Size_(A w, B h) : width(T(w)), height(T(h)) {} };
template<typename T> struct Rect_ { T x{}, y{}, width{}, height{}; Rect_() = default; template<typename A, typename B, typename C_, typename D> Rect_(A x_, B y_, C_ w, D h) : x(T(x_)), y(T(y_)), width(T(w)), height(T(h)) {} };
using Size = Size_<int>;
using Rect2i = Rect_<int>;
}
namespace cmn { class Bounds; }
// the handful of OpenCV names the compiled files mention.  cv::Mat is a plain row-major byte image (rows, cols, type = CV_8UC1 / CV_8UC3, step.p =
// {bytes per row, bytes per pixel}, ptr(row)) -- enough for Source::extract_lines, which only compares bytes with zero; Transform::toCV, the debug drawing of
// offset_to_middle and the outline_use_dft branch of find_tail are never reached (they abort if they are)
#define CV_8UC1 0
#define CV_8UC3 16
#define CV_32FC1 5
#define CV_64F 6
#define CV_8UC4 24
namespace cv {
struct Scalar { Scalar(double = 0, double = 0, double = 0, double = 0) {} };
struct Mat {
    int rows = 0, cols = 0, _type = CV_8UC1;
    std::shared_ptr<std::vector<unsigned char>> store;      // copies share the pixels, like cv::Mat
    unsigned char *data = nullptr;
    struct Step { size_t p[2] = {0, 0}; } step;
    Mat() {}
    Mat(int r, int c, int t) : rows(r), cols(c), _type(t) { alloc(); }
    Mat(int r, int c, int t, const Scalar&) : Mat(r, c, t) {}
    Mat(int r, int c, int t, void *d) : rows(r), cols(c), _type(t), data((unsigned char *)d) { step.p[1] = elemSize(); step.p[0] = (size_t)c * step.p[1]; }
    template<typename T> explicit Mat(const std::vector<T>&) { std::fprintf(stderr, "cv::Mat(std::vector) stand-in used\n"); std::abort(); }      // RawProcessing.cpp's tag branch only
    static Mat ones(int r, int c, int t) { Mat m(r, c, t); std::memset(m.data, 1, (size_t)r * m.step.p[0]); return m; }                              // 8-bit only
    void release() { store.reset(); data = nullptr; rows = cols = 0; }
    void setTo(int v) { for (int y = 0; y < rows; ++y) std::memset(ptr(y), v, (size_t)cols * step.p[1]); }                   // 8-bit only
    void convertTo(const Mat& dst_, int t, double alpha = 1.0) const                           // 8U -> 8U (a copy) and 8U -> 32F (scaled) here; 32F -> 8U is OpenCV's rounding: bridged
    {
        Mat& dst = const_cast<Mat&>(dst_);
        Mat src = *this;
#ifdef REF_DETECT
        if ((t & 7) == 0 && src.depth() == 5) { convert_32f_to_8u(src, dst); return; }
#endif
        if ((t & 7) == 0 && src.depth() == 0) { src.copyTo(dst); return; }
        if ((t & 7) != 5 || src.depth() != 0) { std::fprintf(stderr, "cv::Mat::convertTo stand-in: unsupported types\n"); std::abort(); }
        Mat out(src.rows, src.cols, 5 + ((src.channels() - 1) << 3));
        for (int y = 0; y < src.rows; ++y) for (int x = 0; x < src.cols * src.channels(); ++x) reinterpret_cast<float *>(out.ptr(y))[x] = float(src.ptr(y)[x] * alpha);
        dst = out;
    }
    static Mat zeros(int r, int c, int t) { return Mat(r, c, t); }
    size_t elemSize() const { return (size_t)channels() * (depth() == 6 ? 8 : (depth() == 5 || depth() == 4 ? 4 : (depth() == 2 || depth() == 3 ? 2 : 1))); }
    void alloc() { step.p[1] = elemSize(); step.p[0] = (size_t)cols * step.p[1]; store = std::make_shared<std::vector<unsigned char>>((size_t)rows * step.p[0] + 64, 0); data = store->data(); }
    int type() const { return _type; }
    static void convert_32f_to_8u(const Mat& src, Mat& dst);          // detection build: forwarded to cv2 (oracle/ref_stubs_detect/cv_detect.h)
    int channels() const { return (_type >> 3) + 1; }                                      // OpenCV's type code: depth in the low three bits, channels - 1 above
    int depth() const { return _type & 7; }
    bool isContinuous() const { return true; }
    const unsigned char *ptr(int r = 0) const { return data + (size_t)r * step.p[0]; }
    unsigned char *ptr(int r = 0) { return data + (size_t)r * step.p[0]; }
    const unsigned char *ptr(int r, int c) const { return data + (size_t)r * step.p[0] + (size_t)c * step.p[1]; }
    unsigned char *ptr(int r, int c) { return data + (size_t)r * step.p[0] + (size_t)c * step.p[1]; }
    template<typename T> T *ptr(int r, int = 0) { (void)r; std::fprintf(stderr, "cv::Mat::ptr<T> stand-in used\n"); std::abort(); return nullptr; }
    template<typename T> T& at(int r, int c) { return *reinterpret_cast<T *>(data + (size_t)r * step.p[0] + (size_t)c * sizeof(T)); }
    bool empty() const { return data == nullptr; }
    Mat& operator=(const Scalar&) { if (data) std::memset(data, 0, (size_t)rows * step.p[0]); return *this; }
    // what tracker/tracking/FilterCache.cpp uses on top: size(), copyTo (plain and masked: a destination of another size / type is re-created, zero-filled
    // for the masked form), the ROI view padded(Bounds) (the rectangle's members truncated to int: vec2.h:433)
    Size size() const { return Size(cols, rows); }
    void create(int r, int c, int t) { if (r != rows || c != cols || t != _type || !data) { rows = r; cols = c; _type = t; alloc(); } }
    void copyTo(const Mat& dst_) const                 // OpenCV's OutputArray binds to const cv::Mat& as well (FilterCache.cpp:65 copies a const image onto itself)
    {
        Mat& dst = const_cast<Mat&>(dst_);
        if (dst.data == data && dst.rows == rows && dst.cols == cols) return;
        Mat src = *this;                                   // keeps the pixels alive when dst is the parent of this view
        dst.create(src.rows, src.cols, src._type);
        for (int y = 0; y < src.rows; ++y) std::memcpy(dst.ptr(y), src.ptr(y), (size_t)src.cols * src.step.p[1]);
    }
    void copyTo(const Mat& dst_, const Mat& mask) const
    {
        Mat& dst = const_cast<Mat&>(dst_);
        if (dst.data == data && dst.rows == rows && dst.cols == cols) return;
        Mat src = *this;
        dst.create(src.rows, src.cols, src._type);
        const size_t es = src.step.p[1];
        for (int y = 0; y < src.rows; ++y) for (int x = 0; x < src.cols; ++x) if (mask.ptr(y)[x]) std::memcpy(dst.ptr(y) + x * es, src.ptr(y) + x * es, es);
    }
    Mat operator()(const cmn::Bounds& b) const;          // defined below, after vec2.h
};
enum { INTER_NEAREST = 0, INTER_LINEAR = 1, BORDER_CONSTANT = 0 };
// cv::warpAffine is OpenCV's (third party): the test installs the oracle's bit-exact restatement of its 8-bit INTER_LINEAR / BORDER_CONSTANT path
// (oracle/trex_oracle.c to_warp_affine_u8, pinned on cv2 4.13 by tests/test_oracle_moments.py) through ref_filtercache_set_warp
using warp_fn_t = void (*)(const unsigned char *src, int sw, int sh, const double *M, unsigned char *dst, int dw, int dh);
inline warp_fn_t& warp_hook() { static warp_fn_t f = nullptr; return f; }
inline void warpAffine(const Mat& src, Mat& dst, const Mat& M, Size dsize, int flags = INTER_LINEAR, int = BORDER_CONSTANT)
{
    if (!warp_hook() || flags != INTER_LINEAR || src.channels() != 1 || src.step.p[0] != (size_t)src.cols) { std::fprintf(stderr, "cv::warpAffine stand-in: unsupported call\n"); std::abort(); }
    double m[6];
    for (int r = 0; r < 2; ++r) for (int c = 0; c < 3; ++c) m[r * 3 + c] = const_cast<Mat&>(M).at<double>(r, c);
    Mat out(dsize.height, dsize.width, src._type);
    warp_hook()(src.data, src.cols, src.rows, m, out.data, dsize.width, dsize.height);
    dst = out;
}
inline void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int, int)
{
    Mat in = src, out(in.rows + top + bottom, in.cols + left + right, in._type);
    for (int y = 0; y < in.rows; ++y) std::memcpy(out.ptr(y + top) + (size_t)left * in.step.p[1], in.ptr(y), (size_t)in.cols * in.step.p[1]);
    dst = out;
}
// cv::resize(src, dst, Size(), f, f, INTER_NEAREST) (OpenCV imgproc/resize.cpp): dsize = cvRound(n * f) (round half to even), source index =
// min(cvFloor(d * (1 / f)), n - 1); the oracle's resize_nearest restates the same and is checked against cv2 in tests/test_oracle_golden.py
// the test can install the real cv2.resize instead (ref_filtercache_set_resize): (src, sw, sh, channels, fx, fy, dst, dw, dh)
using resize_fn_t = void (*)(const unsigned char *src, int sw, int sh, int channels, double fx, double fy, unsigned char *dst, int dw, int dh);
inline resize_fn_t& resize_hook() { static resize_fn_t f = nullptr; return f; }
inline void resize(const Mat& src, Mat& dst, Size, double fx, double fy, int flags)
{
    if (flags != INTER_NEAREST) { std::fprintf(stderr, "cv::resize stand-in: only INTER_NEAREST\n"); std::abort(); }
    Mat in = src, out((int)std::lrint(in.rows * fy), (int)std::lrint(in.cols * fx), in._type);
    if (resize_hook() && in.step.p[0] == (size_t)in.cols * in.step.p[1]) { resize_hook()(in.data, in.cols, in.rows, in.channels(), fx, fy, out.data, out.cols, out.rows); dst = out; return; }
    const double ifx = 1. / fx, ify = 1. / fy; const size_t es = in.step.p[1];
    for (int y = 0; y < out.rows; ++y) {
        const int sy = std::min((int)std::floor(y * ify), in.rows - 1);
        for (int x = 0; x < out.cols; ++x) { const int sx = std::min((int)std::floor(x * ifx), in.cols - 1); std::memcpy(out.ptr(y) + x * es, in.ptr(sy) + sx * es, es); }
    }
    dst = out;
}
enum { DFT_INVERSE = 1, DFT_SCALE = 2 };
inline void dft(const Mat&, Mat&, int = 0) { std::fprintf(stderr, "cv::dft stand-in used\n"); std::abort(); }
}

namespace glz {
struct json_t { json_t() = default; template<typename... T> json_t(T&&...) {} bool contains(const char *) const { return false; } json_t operator[](const char *) const { return {}; } double get_number() const { return 0; } };
enum class error_code { none };
template<typename... A> inline error_code read_json(A&&...) { return error_code::none; }
template<typename... A> inline std::string format_error(A&&...) { return std::string(); }
template<typename T> struct meta;
template<auto Read, auto Write> inline constexpr int custom = 0;
template<typename... A> constexpr int array(A...) { return 0; }
}
namespace cmn {
struct Meta {
    template<typename T> static std::string toStr(const T&) { return std::string(); }
    template<typename T> static std::string name() { return std::string(); }
    template<typename T, typename S> static T fromStr(S&&) { return T{}; }
};
template<typename Str> concept StringLike = std::is_same_v<std::remove_cvref_t<Str>, std::string> || std::is_same_v<std::remove_cvref_t<Str>, const char*> ||
                                            std::is_same_v<std::remove_cvref_t<Str>, std::string_view> || std::is_array_v<std::remove_cvref_t<Str>>;
template<typename T> concept is_numeric = (!std::is_same_v<std::remove_cvref_t<T>, bool>) && (std::floating_point<T> || std::integral<T>);      // misc/useful_concepts.h:175
template<typename K, typename T = K> constexpr inline T saturate(K val, T min = 0, T max = 255) { return std::clamp(T(val), min, max); }          // misc/detail.h:225-229
}
#include <misc/vec2.h>
namespace cv {
inline Mat Mat::operator()(const cmn::Bounds& b) const          // the ROI view: the rectangle's members truncated to int (vec2.h:433)
{
    Mat v = *this;
    v.data = data + (size_t)(int)b.y * step.p[0] + (size_t)(int)b.x * step.p[1]; v.rows = (int)b.height; v.cols = (int)b.width;
    return v;
}
}
namespace cmn {
class Minimizable { public: virtual void minimize_memory() = 0; virtual ~Minimizable() {} };

struct Frame_t {
    int32_t _frame = -1;
    constexpr Frame_t() = default;
    explicit constexpr Frame_t(int32_t f) : _frame(f) {}
    constexpr bool valid() const { return _frame >= 0; }
    constexpr int32_t get() const { return _frame; }
    constexpr bool operator==(const Frame_t&) const = default;
    constexpr auto operator<=>(const Frame_t& o) const { return _frame <=> o._frame; }
    constexpr Frame_t operator-(const Frame_t& o) const { return Frame_t(_frame - o._frame); }
    constexpr Frame_t operator+(const Frame_t& o) const { return Frame_t(_frame + o._frame); }
};
constexpr Frame_t operator""_f(unsigned long long v) { return Frame_t(int32_t(v)); }
// LOGGED_MUTEX / LOGGED_LOCK without the logging (commons.pc.h:1399-1400)
struct LoggedMutexStandIn : std::mutex { LoggedMutexStandIn(const char *) {} };
#define LOGGED_MUTEX(NAME) cmn::LoggedMutexStandIn(NAME)
#define LOGGED_LOCK(MUTEX) std::unique_lock<std::mutex>{ MUTEX }

struct glz_json_placeholder {};
template<typename T> inline auto cvt2json(const T&) { return glz_json_placeholder{}; }
template<int N> struct dec { template<typename T> dec(T) {} std::string toStr() const { return std::string(); } };
template<typename... A> inline void Print(const A&...) {}
template<typename... A> inline void FormatWarning(const A&...) {}
template<typename... A> inline void FormatError(const A&...) {}
template<typename... A> inline void FormatExcept(const A&...) {}
template<typename... A> inline std::runtime_error U_EXCEPTION(const char *msg, const A&...) { return std::runtime_error(msg); }
template<typename... A> inline std::runtime_error RuntimeError(const char *msg, const A&...) { return std::runtime_error(msg); }
}

namespace cmn::utils {
inline bool lowercase_equal_to(std::string_view a, std::string_view b)
{
    if (a.size() != b.size()) return false;
    for (size_t i = 0; i < a.size(); ++i) if (std::tolower((unsigned char)a[i]) != std::tolower((unsigned char)b[i])) return false;
    return true;
}
}
#include <misc/base_types.h>
#include <misc/EnumClass.h>
#include <misc/detail_bits.h>
using namespace cmn;
#ifdef REF_DETECT
#include <cv_detect.h>
#endif
