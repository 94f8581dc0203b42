// Test infrastructure (oracle/): a MINIMAL stand-in for TRex's precompiled header `commons.pc.h`, so that the reference's own
// commons/common/misc/CircularGraph.cpp (eft, ieft, curvature, differentiate, find_peaks and its fast::cos polynomial) compiles here, unmodified
// and from where it lies under /root/reference, without OpenCV / glaze / cnpy (oracle/build_ref.py).  Only the declarations that file touches
// are mirrored; the arithmetic-bearing ones restate the reference 1:1:
//   Vec2            commons/common/misc/vec2.h:20-215  (Vector2D<float, true>: component-wise float operators, length() = std::sqrt(x*x + y*y))
//   sqdistance      commons/common/misc/vec2.h:380-383
//   SQR             commons/common/commons.pc.h:446
//   cmn::sqrt       commons/common/misc/math.h:23-32   (the float specialisation calls ::sqrtf: EFT::dt's unqualified `sqrt(...)` resolves to it)
// Everything else (printing, timing, exceptions) is inert.  Nothing under trex_b200/ includes this.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <concepts>
#include <cstdint>
#include <cstdio>
#include <memory>
#include <set>
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
#define GETTER(TYPE, VAR) protected: TYPE _##VAR; public: const TYPE& VAR() const { return _##VAR; } protected:

using long_t = int32_t;

namespace cmn {
using Float2_t = float;
constexpr Float2_t operator""_F(long double v) { return Float2_t(v); }
constexpr Float2_t operator""_F(unsigned long long v) { return Float2_t(v); }

template<typename T = double> inline T sqrt(const T& s) { return ::sqrt(s); }
template<> inline float sqrt(const float& s) { return ::sqrtf(s); }

struct Vec2 {
    Float2_t x, y;
    constexpr Vec2() noexcept : x(0), y(0) {}
    template<typename S> requires std::is_arithmetic_v<S>
    constexpr Vec2(S v) noexcept : x(Float2_t(v)), y(Float2_t(v)) {}
    template<typename S0, typename S1> requires (std::is_arithmetic_v<S0> && std::is_arithmetic_v<S1>)
    constexpr Vec2(S0 a, S1 b) noexcept : x(Float2_t(a)), y(Float2_t(b)) {}
    constexpr Float2_t A() const { return x; }
    constexpr Float2_t B() const { return y; }
    constexpr Vec2& operator+=(const Vec2& o) { x += o.x; y += o.y; return *this; }
    constexpr Vec2& operator-=(const Vec2& o) { x -= o.x; y -= o.y; return *this; }
    constexpr Vec2& operator+=(Float2_t o) { x += o; y += o; return *this; }
    constexpr Vec2& operator-=(Float2_t o) { x -= o; y -= o; return *this; }
    constexpr Vec2& operator*=(Float2_t o) { x *= o; y *= o; return *this; }
    constexpr Vec2& operator/=(Float2_t o) { x /= o; y /= o; return *this; }
    template<typename S> requires std::is_arithmetic_v<S> constexpr Vec2 operator/(S o) const { return Vec2{x / Float2_t(o), y / Float2_t(o)}; }
    template<typename S> requires std::is_arithmetic_v<S> constexpr Vec2 operator*(S o) const { return Vec2{x * Float2_t(o), y * Float2_t(o)}; }
    template<typename S> requires std::is_arithmetic_v<S> constexpr Vec2 operator-(S o) const { return Vec2{x - Float2_t(o), y - Float2_t(o)}; }
    template<typename S> requires std::is_arithmetic_v<S> constexpr Vec2 operator+(S o) const { return Vec2{x + Float2_t(o), y + Float2_t(o)}; }
    constexpr Vec2 operator+(Vec2 o) const { return Vec2{x + o.x, y + o.y}; }
    constexpr Vec2 operator-(Vec2 o) const { return Vec2{x - o.x, y - o.y}; }
    constexpr Vec2 operator-() const { return Vec2{-x, -y}; }
    constexpr Float2_t sqlength() const { return x * x + y * y; }
    Float2_t length() const { return std::sqrt(sqlength()); }
    constexpr bool operator==(const Vec2& o) const { return x == o.x && y == o.y; }
    constexpr bool operator<(const Vec2& o) const { return o.y < y || (o.y == y && o.x < x); }      // vec2.h:99-102
};
template<typename S> requires std::is_arithmetic_v<S> constexpr Vec2 operator*(S s, const Vec2& v) { return Vec2{Float2_t(s) * v.x, Float2_t(s) * v.y}; }
inline Float2_t sqdistance(const Vec2& p0, const Vec2& p1) { return SQR(p1.A() - p0.A()) + SQR(p1.B() - p0.B()); }

struct Meta {
    template<typename T> static std::string toStr(const T&) { return std::string(); }
};
template<typename... A> inline void Print(const A&...) {}
template<typename... A> inline void FormatWarning(const A&...) {}
template<typename... A> inline std::runtime_error U_EXCEPTION(const char *msg, const A&...) { return std::runtime_error(msg); }
}
using namespace cmn;
