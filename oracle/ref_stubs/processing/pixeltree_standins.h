// Test infrastructure: what commons/common/processing/PixelTree.{h,cpp} needs around pixel::find_outer_points -- the ONLY function of that file the
// tests call (oracle/ref_pixeltree.cpp).  The file also defines threshold_blob & co., which drag in pv::Blob, Background, CPULabeling and the
// image-mode dispatch; for those the declarations below only have to COMPILE (their bodies are never run: CPULabeling::run returns nothing,
// call_image_mode_function calls nothing).  What find_outer_points itself touches is restated from the reference:
//   HorizontalLine {x0, x1, y}, coord_t = uint16_t, ptr_safe_t = uint64_t      commons/common/misc/detail.h:71-116
//   pv::Blob::hor_lines() / bounds()  (bounds = bounding box of the lines: x, y, width = max x1 - min x0 + 1 ...)   processing/PVBlob.cpp
//   contains(vector, value) = std::find                                        commons.pc.h
#pragma once
#include <commons.pc.h>
#include <misc/ranges.h>
#include <span>
#include <deque>

using uchar = unsigned char;
namespace cmn {
using coord_t = uint16_t;
using ptr_safe_t = uint64_t;
struct HorizontalLine {
    coord_t x0, x1;
    coord_t y, padding;
    constexpr HorizontalLine() noexcept = default;
    constexpr HorizontalLine(coord_t y_, coord_t x0_, coord_t x1_) noexcept : x0(x0_), x1(x1_), y(y_), padding(0) {}
    constexpr ptr_safe_t length() const noexcept { return ptr_safe_t(x1) - ptr_safe_t(x0) + 1; }
};
using PixelArray_t = std::vector<uchar>;
using Rangel = Range<long_t>;
template<typename T, typename... A> constexpr bool is_in(const T& v, const A&... a) { return ((v == T(a)) || ...); }
template<typename Cont, typename V> inline bool contains(const Cont& c, const V& v) { return std::find(c.begin(), c.end(), v) != c.end(); }
// sorted-vector helpers of pixel::Tree (commons.pc.h:693-716): upper_bound insertion, lower_bound look-up with the caller's comparator
template<class T, typename Cmp> inline auto insert_sorted(std::vector<T>& v, T&& e, Cmp&& cmp) { return v.insert(std::upper_bound(v.begin(), v.end(), e, std::forward<Cmp>(cmp)), std::move(e)); }
template<class T, typename Cmp, class K = T> inline auto find_sorted(const std::vector<T>& v, const K& e, Cmp&& cmp)
{
    auto it = std::lower_bound(v.begin(), v.end(), e, cmp);
    return (it != v.end() && !cmp(e, *it)) ? it : v.end();
}

struct Image { using Ptr = std::unique_ptr<Image>; };
enum class meta_encoding_t { gray, r3g3b2, rgb8, binary };
struct InputInfo { uint8_t channels = 1; meta_encoding_t encoding = meta_encoding_t::gray; constexpr bool is_r3g3b2() const { return encoding == meta_encoding_t::r3g3b2; } };
struct OutputInfo {
    uint8_t channels = 1; meta_encoding_t encoding = meta_encoding_t::gray;
    constexpr bool is_r3g3b2() const { return encoding == meta_encoding_t::r3g3b2; }
    constexpr OutputInfo& operator=(const InputInfo& i) { channels = i.channels; encoding = i.encoding; return *this; }
};
namespace DifferenceMethod_t { enum Class { absolute, sign, none }; }
using DifferenceMethod = DifferenceMethod_t::Class;
constexpr OutputInfo DIFFERENCE_OUTPUT_FORMAT{};
template<InputInfo, OutputInfo> using PixelOutput_t = uchar;
template<InputInfo, OutputInfo> inline uchar grey_diffable_pixel_value(const uchar *p) { return *p; }
template<typename F> inline void call_image_mode_function(InputInfo, OutputInfo, F&&) {}

struct BackgroundInfo { const uchar *data = nullptr; ptr_safe_t width = 0; };
class Background {
public:
    template<OutputInfo, DifferenceMethod, typename> BackgroundInfo info() const { return {}; }
    template<OutputInfo, DifferenceMethod> bool is_different(coord_t, coord_t, uchar, int, const BackgroundInfo&) const { return false; }
};

namespace blob {
using lines_t = std::vector<HorizontalLine>;
using line_ptr_t = std::unique_ptr<lines_t>;
using pixel_ptr_t = std::unique_ptr<PixelArray_t>;
struct Prediction {};
struct Pair {
    line_ptr_t lines; pixel_ptr_t pixels; uint8_t extra_flags = 0; Prediction pred;
    Pair() = default;
    Pair(line_ptr_t&& l, pixel_ptr_t&& p, uint8_t f = 0, Prediction&& pr = {}) : lines(std::move(l)), pixels(std::move(p)), extra_flags(f), pred(pr) {}
};
}
using blobs_t = std::vector<blob::Pair>;

namespace CPULabeling {
struct ListCache_t {};
inline blobs_t run(const std::vector<HorizontalLine>&, const PixelArray_t&, ListCache_t&, uint8_t) { return {}; }
inline blobs_t run(const std::vector<HorizontalLine>&, std::span<uchar>, ListCache_t&, uint8_t) { return {}; }
}
}

namespace pv {
class Blob;
using BlobPtr = std::unique_ptr<Blob>;
class Blob {
    cmn::blob::line_ptr_t _lines;
    cmn::blob::pixel_ptr_t _pixels;
    uint8_t _flags = 0;
    cmn::Bounds _bounds;
public:
    Blob(cmn::blob::line_ptr_t&& l, cmn::blob::pixel_ptr_t&& p, uint8_t flags = 0, cmn::blob::Prediction&& = {}) : _lines(std::move(l)), _pixels(std::move(p)), _flags(flags)
    {
        // bounding box of the lines (pv::Blob::init / calculate bounds, processing/PVBlob.cpp): x = min x0, y = first y, width = max x1 - x + 1, height = last y - y + 1
        if (_lines && !_lines->empty()) {
            int x0 = 1 << 30, x1 = -1;
            for (auto &h : *_lines) { x0 = std::min<int>(x0, h.x0); x1 = std::max<int>(x1, h.x1); }
            _bounds = cmn::Bounds((float)x0, (float)_lines->front().y, (float)(x1 - x0 + 1), (float)(_lines->back().y - _lines->front().y + 1));
        }
    }
    Blob(const Blob& o) : Blob(std::make_unique<cmn::blob::lines_t>(*o._lines), o._pixels ? std::make_unique<cmn::PixelArray_t>(*o._pixels) : nullptr, o._flags) {}
    template<typename... A> static BlobPtr Make(A&&... a) { return std::make_unique<Blob>(std::forward<A>(a)...); }
    const std::vector<cmn::HorizontalLine>& hor_lines() const { return *_lines; }
    const cmn::blob::line_ptr_t& lines() const { return _lines; }
    const cmn::blob::pixel_ptr_t& pixels() const { return _pixels; }
    cmn::blob::line_ptr_t&& steal_lines() { return std::move(_lines); }
    const cmn::Bounds& bounds() const { return _bounds; }
    uint8_t flags() const { return _flags; }
    uint32_t blob_id() const { return 0; }
    cmn::blob::Prediction prediction() const { return {}; }
    uint8_t channels() const { return 1; }
    cmn::InputInfo input_info() const { return {}; }
    bool is_binary() const { return false; }
    bool is_rgb() const { return false; }
    bool is_r3g3b2() const { return false; }
    size_t num_pixels() const { size_t n = 0; for (auto &h : *_lines) n += h.length(); return n; }
    static uint8_t copy_flags(const Blob& b) { return b._flags; }
};
}
