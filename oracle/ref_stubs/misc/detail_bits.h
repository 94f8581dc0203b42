// Test infrastructure: the pieces of commons/common/misc/detail.{h,cpp}, misc/Image.h and misc/GlobalSettings.h that processing/Background.{h,cpp} use.
// Arithmetic-bearing ones restate the reference:
//   saturate(val, min, max) = std::clamp(T(val), min, max) with T deduced from `min`            misc/detail.h:225-229
//   vec_to_r3g3b2 / r3g3b2_to_vec                                                                 misc/detail.h:508-531
//   lines_dimensions: bounding box of the runs through float min / max                            misc/detail.cpp:440-458
//   cv::cvtColor(BGR2GRAY) = (B * 1868 + G * 9617 + R * 4899 + 8192) >> 14 (OpenCV's 8-bit path; the oracle pins the same formula against cv2),
//   cv::cvtColor(GRAY2BGR) = three copies
#pragma once
#include <future>
#include <misc/matharray.h>
namespace cv {
struct Vec3b { unsigned char v[3]; Vec3b() : v{0, 0, 0} {} Vec3b(unsigned char a, unsigned char b, unsigned char c) : v{a, b, c} {} unsigned char& operator[](int i) { return v[i]; } const unsigned char& operator[](int i) const { return v[i]; } };
enum { COLOR_GRAY2BGR = 8, COLOR_BGR2GRAY = 6 };
#ifndef REF_DETECT
inline void cvtColor(const Mat& src, Mat dst, int code)
{
    for (int y = 0; y < src.rows; ++y) {
        const unsigned char *s = src.ptr(y); unsigned char *d = dst.ptr(y);
        for (int x = 0; x < src.cols; ++x) {
            if (code == COLOR_GRAY2BGR) { d[3 * x] = d[3 * x + 1] = d[3 * x + 2] = s[x]; }
            else d[x] = (unsigned char)((s[3 * x] * 1868 + s[3 * x + 1] * 9617 + s[3 * x + 2] * 4899 + 8192) >> 14);
        }
    }
}
#endif
}
#define CV_8UC(n) ((n) == 3 ? CV_8UC3 : ((n) == 4 ? CV_8UC4 : CV_8UC1))
namespace cmn {
enum class ImageMode { GRAY, RGB, R3G3B2, RGBA };
template<typename Vec> constexpr uint8_t vec_to_r3g3b2(const Vec& bgr) { return (uint8_t(bgr[0] / 64) << 6) | (uint8_t(bgr[1] / 32) << 3) | (uint8_t(bgr[2] / 32) << 0); }
template<uint8_t channels = 3> constexpr auto r3g3b2_to_vec(const uint8_t& c, const uint8_t = 255)
{
    return RGBArray{static_cast<unsigned char>(((uint8_t(c) >> 6) & 3) * 64), static_cast<unsigned char>(((uint8_t(c) >> 3) & 7) * 32), static_cast<unsigned char>((uint8_t(c) & 7) * 32)};
}
template<typename T> inline void resize_image(T& mat, double factor, int flags = cv::INTER_NEAREST) { cv::resize(mat, mat, cv::Size(), factor, factor, flags); }      // misc/detail.h:465-469
inline void convert_from_r3g3b2(const cv::Mat&, cv::Mat&) { std::fprintf(stderr, "convert_from_r3g3b2 stand-in used\n"); std::abort(); }
inline cv::Rect2i lines_dimensions(const std::vector<HorizontalLine>& lines)
{
    float mx = FLT_MAX, my = FLT_MAX, px = -FLT_MAX, py = -FLT_MAX;
    for (auto &l : lines) { if (mx > l.x0) mx = l.x0; if (my > l.y) my = l.y; if (px < l.x1) px = l.x1; if (py < l.y) py = l.y; }
    float w = px - mx + 1, h = py - my + 1;
    return cv::Rect2i(mx, my, w, h);
}
template<typename... A> inline std::invalid_argument InvalidArgumentException(const A&...) { return std::invalid_argument("invalid argument"); }
struct CallbackFuture { bool set = false; operator bool() const { return set; } };

class Image {
public:
    using Ptr = std::unique_ptr<Image>;
    using SPtr = std::shared_ptr<Image>;
    uint32_t cols = 0, rows = 0, dims = 1;
    cv::Mat mat;
    Image(uint32_t r, uint32_t c, uint32_t d) : cols(c), rows(r), dims(d), mat((int)r, (int)c, (int)((d - 1) << 3)) {}      // CV_8UC(d)
    static Ptr Make(uint32_t r, uint32_t c, uint32_t d) { return std::make_unique<Image>(r, c, d); }
    Image() {}
    static Ptr Make() { return std::make_unique<Image>(); }
    template<typename A, typename B> requires (std::is_arithmetic_v<A> && std::is_arithmetic_v<B>) static Ptr Make(A r, B c) { return std::make_unique<Image>((uint32_t)r, (uint32_t)c, 1u); }
    void create(uint32_t r, uint32_t c, uint32_t d) { rows = r; cols = c; dims = d; mat = cv::Mat((int)r, (int)c, (int)((d - 1) << 3)); }
    uchar *ptr(uint32_t y, uint32_t x) { return mat.ptr((int)y, (int)x); }
    const uchar *ptr(uint32_t y, uint32_t x) const { return mat.ptr((int)y, (int)x); }
    explicit Image(const cv::Mat& m) : Image((uint32_t)m.rows, (uint32_t)m.cols, (uint32_t)m.channels()) { m.copyTo(mat); }      // Image::Make(cv::Mat): a copy of the pixels
    static Ptr Make(const cv::Mat& m) { return std::make_unique<Image>(m); }
    const uchar *data() const { return mat.data; }
    uchar *data() { return mat.data; }
    ptr_safe_t channels() const { return dims; }
    size_t size() const { return (size_t)rows * cols * dims; }           // bytes (misc/Image.h)
    cv::Mat get() const { return mat; }
    Bounds bounds() const { return Bounds(0, 0, (float)cols, (float)rows); }
};

// settings: a three-entry table the test wrapper fills; register_callbacks runs the callback once for every name (as TRex does on registration)
struct RefSettings {
    bool track_threshold_is_absolute = true, track_background_subtraction = true; int meta_encoding = 0; std::function<void(std::string_view)> cb;
    std::map<std::string, double> num;                                   // every other (numeric / boolean) setting by name: the detection build (RawProcessing.cpp, BackgroundSubtraction.cpp)
    std::vector<std::function<void(std::string_view)>> cbs;              // every callback registered so far (cb = the last one)
};
inline RefSettings& ref_settings() { static RefSettings s; return s; }
inline bool bool_setting_config(const char *name)
{
    if (std::string_view(name) == "track_threshold_is_absolute") return ref_settings().track_threshold_is_absolute;
    if (std::string_view(name) == "track_background_subtraction") return ref_settings().track_background_subtraction;
    return ref_settings().num[name] != 0;
}
struct NoType {};
struct SettingValueStandIn { double v; template<typename T> T value() const { return T(v); } };
struct GlobalSettings {
    template<typename Str, typename F> static CallbackFuture register_callbacks(std::initializer_list<Str> names, F&& fn)
    {
        ref_settings().cb = fn;
        ref_settings().cbs.push_back(fn);
        for (auto &n : names) fn(std::string_view(n));
        return CallbackFuture{true};
    }
    static void unregister_callbacks(CallbackFuture&&) {}
    static bool is_runtime_quiet() { return true; }
    template<typename T> static T read_value_with_default(const char *, T fallback) { return fallback; }
    template<typename> static SettingValueStandIn read_value(std::string_view key) { return SettingValueStandIn{ref_settings().num[std::string(key)]}; }
};
}
namespace cmn {
template<typename T> inline T read_setting_config(const char *name)
{
    if constexpr (requires { T{}.value(); }) return T{static_cast<std::remove_cvref_t<decltype(T{}.value())>>(ref_settings().meta_encoding)};      // meta_encoding_t
    else if constexpr (std::is_arithmetic_v<T>) return T(ref_settings().num[name]);
    else return T{};
}
template<typename T> inline T read_setting_config_or(const char *name, T fallback) { auto it = ref_settings().num.find(name); return it == ref_settings().num.end() ? fallback : T(it->second); }
}
#define READ_SETTING_WITH_DEFAULT(NAME, DEFAULT) (cmn::read_setting_config_or( #NAME, DEFAULT ))
#define BOOL_SETTING(NAME) (cmn::bool_setting_config(#NAME))
#define READ_SETTING(NAME, ...) (cmn::read_setting_config< __VA_ARGS__ >( #NAME ))
