// Test infrastructure: what PixelTree.h's threshold templates need from commons/common/processing/Background.h in order to COMPILE.  They are never
// instantiated with a background here (the tests call find_outer_points and CPULabeling::run only), so nothing below carries reference arithmetic.
#pragma once
#include <commons.pc.h>
#include <misc/ranges.h>
#include <processing/PVBlob.h>
namespace cmn {
using Rangel = Range<long_t>;
struct Image { using Ptr = std::unique_ptr<Image>; };
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
}
