// Test infrastructure (oracle/): a MINIMAL stand-in for TRex's precompiled header `commons.pc.h`, so that a few of the reference's own source files
// compile here, unmodified and from where they lie under /root/reference, without OpenCV / glaze / cnpy (oracle/build_ref.py):
//   commons/common/misc/CircularGraph.cpp   eft, ieft, curvature, differentiate, find_peaks and its fast::cos polynomial
//   commons/common/misc/curve_discussion.cpp, commons/common/gui/Transform.cpp
//   tracker/tracking/Outline.cpp            Outline::resample / smooth / offset_to_middle / calculate_midline, Midline::post_process / normalize / fix_length
// Only the declarations those files touch are provided, written for this purpose; the ones that carry arithmetic follow the reference's
// definitions operation for operation and cite them:
//   Vec2 / Size2      commons/common/misc/vec2.h:20-215   (Vector2D<float>: component-wise float operators; length() = std::sqrt(x*x + y*y);
//                     normalize() = (L != 0) * (v / ((L == 0) + L)); member atan2() = std::atan2(y, x) in float; free atan2(v) = ::atan2 in DOUBLE, :371)
//   SQR, DEGREE, RADIANS, GETTER*    commons.pc.h:444-457
//   cmn::min / cmn::max              commons.pc.h:462-528 (mixed arithmetic types: computed in the wider of the two, the second on a tie)
//   narrow_cast                      commons.pc.h (value-preserving static_cast)
// The scalar maths is NOT restated: commons/common/misc/math.h is the reference's own file, included from the checkout (cmn::sqrt / sin / cos / atan2, fast_atan2,
// abs, isnan, sqdistance, euclidean_distance, t_circle_line, crosses_zero, infinity<T>) -- like misc/IllegalVector.h, misc/EnumClass.h, misc/matharray.h,
// misc/bid.h, misc/Median.h, processing/encoding.h and every header the compiled .cpp files bring themselves (Outline.h, CircularGraph.h, PixelTree.h, Background.h ...).
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
constexpr Float2_t operator""_F(long double v) { return Float2_t(v); }
constexpr Float2_t operator""_F(unsigned long long v) { return Float2_t(v); }

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
namespace cv { struct Mat; struct Size { int width = 0, height = 0; Size() = default; template<typename A, typename B> Size(A w, B h) : width(int(w)), height(int(h)) {} }; }
namespace cmn {
template<bool IsVec>
struct Vector2D {
    union { Float2_t x; Float2_t width; };      // vec2.h: Size2 names its two members width / height (FilterCache.cpp reads them)
    union { Float2_t y; Float2_t height; };
    Vector2D(const cv::Size& s) noexcept : x(Float2_t(s.width)), y(Float2_t(s.height)) {}                 // vec2.h:30-33
    operator cv::Size() const { return cv::Size(x, y); }
    explicit Vector2D(const cv::Mat& m) noexcept;                                                         // (cols, rows): vec2.h Size2(const cv::Mat&)

    constexpr Vector2D() noexcept : x(0), y(0) {}
    constexpr Vector2D(const Vector2D& o) noexcept : x(o.x), y(o.y) {}
    constexpr Vector2D& operator=(const Vector2D& o) noexcept { x = o.x; y = o.y; return *this; }
    template<typename S> requires std::is_arithmetic_v<S>
    constexpr Vector2D(S v) noexcept : x(Float2_t(v)), y(Float2_t(v)) {}
    template<typename S0, typename S1> requires (std::is_arithmetic_v<S0> && std::is_arithmetic_v<S1>)
    constexpr Vector2D(S0 a, S1 b) noexcept : x(Float2_t(a)), y(Float2_t(b)) {}
    template<bool K> constexpr Vector2D(const Vector2D<K>& o) noexcept : x(o.x), y(o.y) {}
    constexpr Float2_t A() const { return x; }
    constexpr Float2_t B() const { return y; }
    constexpr Vector2D& operator+=(const Vector2D& o) { x += o.x; y += o.y; return *this; }
    constexpr Vector2D& operator-=(const Vector2D& o) { x -= o.x; y -= o.y; return *this; }
    constexpr Vector2D& operator+=(Float2_t o) { x += o; y += o; return *this; }
    constexpr Vector2D& operator-=(Float2_t o) { x -= o; y -= o; return *this; }
    constexpr Vector2D& operator*=(Float2_t o) { x *= o; y *= o; return *this; }
    constexpr Vector2D& operator/=(Float2_t o) { x /= o; y /= o; return *this; }
    template<typename S> requires std::is_arithmetic_v<S> constexpr Vector2D operator/(S o) const { return Vector2D{x / Float2_t(o), y / Float2_t(o)}; }
    template<typename S> requires std::is_arithmetic_v<S> constexpr Vector2D operator*(S o) const { return Vector2D{x * Float2_t(o), y * Float2_t(o)}; }
    template<typename S> requires std::is_arithmetic_v<S> constexpr Vector2D operator-(S o) const { return Vector2D{x - Float2_t(o), y - Float2_t(o)}; }
    template<typename S> requires std::is_arithmetic_v<S> constexpr Vector2D operator+(S o) const { return Vector2D{x + Float2_t(o), y + Float2_t(o)}; }
    constexpr Vector2D operator+(Vector2D o) const { return Vector2D{x + o.x, y + o.y}; }
    constexpr Vector2D operator-(Vector2D o) const { return Vector2D{x - o.x, y - o.y}; }
    constexpr Vector2D operator-() const { return Vector2D{-x, -y}; }
    constexpr Vector2D mul(const Vector2D& o) const { return Vector2D{x * o.x, y * o.y}; }
    constexpr Vector2D div(const Vector2D& o) const { return Vector2D{x / o.x, y / o.y}; }
    constexpr Vector2D perp() const { return Vector2D{y, -x}; }
    constexpr Vector2D T() const { return Vector2D{y, x}; }
    constexpr Float2_t dot(const Vector2D& o) const { return x * o.x + y * o.y; }
    constexpr Float2_t sqlength() const { return x * x + y * y; }
    Float2_t length() const { return std::sqrt(sqlength()); }
    Vector2D normalize() const { auto L = length(); return Float2_t(L != 0) * (*this / (Float2_t(L == 0) + L)); }
    Float2_t atan2() const { return std::atan2(y, x); }
    Vector2D abs() const { return Vector2D{std::abs(x), std::abs(y)}; }
    constexpr Float2_t max() const { return std::max(x, y); }
    constexpr Float2_t min() const { return std::min(x, y); }
    constexpr Float2_t mean() const { return (x + y) * 0.5f; }
    constexpr bool empty() const { return x == 0 && y == 0; }
    constexpr bool operator==(const Vector2D& o) const { return x == o.x && y == o.y; }
    constexpr bool operator!=(const Vector2D& o) const { return x != o.x || y != o.y; }
    constexpr bool operator<(const Vector2D& o) const { return o.y < y || (o.y == y && o.x < x); }      // vec2.h:99-102
    friend constexpr Vector2D operator*(Float2_t s, const Vector2D& v) { return Vector2D{v.x * s, v.y * s}; }
};
using Vec2 = Vector2D<true>;
using Size2 = Vector2D<false>;
inline Float2_t length(const Vec2& v) { return cmn::sqrt(v.x * v.x + v.y * v.y); }      // = math.h:219-222 (this stand-in's width / height aliases would make math.h's member-detecting overloads ambiguous)
inline auto atan2(const Vec2& v) { return ::atan2(v.y, v.x); }           // double: vec2.h:370-373
inline Vec2 abs(const Vec2& v) { return Vec2(cmn::abs(v.x), cmn::abs(v.y)); }
struct Bounds {
    Float2_t x, y, width, height;
    constexpr Bounds(Float2_t x = 0, Float2_t y = 0, Float2_t w = 0, Float2_t h = 0) : x(x), y(y), width(w), height(h) {}
    Bounds(const Vec2& p, const Size2& s) : x(p.x), y(p.y), width(s.x), height(s.y) {}
    // only used by Posture.cpp's pose-based outline (never run here): the union of two boxes and shifts by a vector
    void combine(const Bounds& o) { const Float2_t x1 = std::max(x + width, o.x + o.width), y1 = std::max(y + height, o.y + o.height); x = std::min(x, o.x); y = std::min(y, o.y); width = x1 - x; height = y1 - y; }
    template<bool K> Bounds operator-(const Vector2D<K>& v) const { return Bounds(x - v.x, y - v.y, width, height); }
    template<bool K> Bounds operator+(const Vector2D<K>& v) const { return Bounds(x + v.x, y + v.y, width, height); }
    Vec2 pos() const { return Vec2(x, y); }
    Size2 size() const { return Size2(width, height); }
    void restrict_to(const Bounds& b)                                   // only RawProcessing.cpp's tag branch (never run here): clamp to a surrounding box
    {
        const Float2_t x1 = std::min(x + width, b.x + b.width), y1 = std::min(y + height, b.y + b.height);
        x = std::max(x, b.x); y = std::max(y, b.y); width = x1 - x; height = y1 - y;
    }
    void operator<<(const Size2& s) { width = s.x; height = s.y; }      // vec2.h:437-444
    void operator<<(const Vec2& p) { x = p.x; y = p.y; }
};

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
struct Meta {
    template<typename T> static std::string toStr(const T&) { return std::string(); }
    template<typename T> static std::string name() { return std::string(); }
    template<typename T, typename S> static T fromStr(S&&) { return T{}; }
};
template<typename T> inline auto cvt2json(const T&) { return glz_json_placeholder{}; }
template<int N> struct dec { template<typename T> dec(T) {} std::string toStr() const { return std::string(); } };
template<typename... A> inline void Print(const A&...) {}
template<typename... A> inline void FormatWarning(const A&...) {}
template<typename... A> inline void FormatError(const A&...) {}
template<typename... A> inline void FormatExcept(const A&...) {}
template<typename... A> inline std::runtime_error U_EXCEPTION(const char *msg, const A&...) { return std::runtime_error(msg); }
template<typename... A> inline std::runtime_error RuntimeError(const char *msg, const A&...) { return std::runtime_error(msg); }
}

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
    Mat operator()(const cmn::Bounds& b) const
    {
        Mat v = *this;
        v.data = data + (size_t)(int)b.y * step.p[0] + (size_t)(int)b.x * step.p[1]; v.rows = (int)b.height; v.cols = (int)b.width;
        return v;
    }
};
}
namespace cmn { template<bool K> inline Vector2D<K>::Vector2D(const cv::Mat& m) noexcept : x(Float2_t(m.cols)), y(Float2_t(m.rows)) {} }
namespace cv {
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
namespace glz { struct json_t { json_t() = default; template<typename T> json_t(T&&) {} }; }
namespace cmn::utils {
inline bool lowercase_equal_to(std::string_view a, std::string_view b)
{
    if (a.size() != b.size()) return false;
    for (size_t i = 0; i < a.size(); ++i) if (std::tolower((unsigned char)a[i]) != std::tolower((unsigned char)b[i])) return false;
    return true;
}
}
namespace cmn {
template<typename Str> concept StringLike = std::is_same_v<std::remove_cvref_t<Str>, std::string> || std::is_same_v<std::remove_cvref_t<Str>, const char*> ||
                                            std::is_same_v<std::remove_cvref_t<Str>, std::string_view> || std::is_array_v<std::remove_cvref_t<Str>>;
}
#include <misc/base_types.h>
#include <misc/EnumClass.h>
#include <misc/detail_bits.h>
using namespace cmn;
#ifdef REF_DETECT
#include <cv_detect.h>
#endif
