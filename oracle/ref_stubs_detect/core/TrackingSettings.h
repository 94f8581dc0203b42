// Test infrastructure: of tracker/core/TrackingSettings.h BackgroundSubtraction.cpp needs the type of cm_per_pixel (Float2_t, TrackingSettings.h:82) and
// SizeFilters -- the REFERENCE'S OWN tracker/core/SizeFilters.{h,cpp}, compiled from the checkout.  The settings of apply() come from the test's table.
#pragma once
#include <commons.pc.h>
#include <core/SizeFilters.h>
namespace track { namespace Settings { using cm_per_pixel_t = cmn::Float2_t; } }
namespace cmn {
struct DetectSettingsStandIn { std::vector<Range<double>> size_filter; std::optional<uint8_t> color_channel; };
inline DetectSettingsStandIn& detect_settings() { static DetectSettingsStandIn s; return s; }
template<> inline SizeFilters read_setting_config<SizeFilters>(const char *) { return SizeFilters(detect_settings().size_filter); }
template<> inline std::optional<uint8_t> read_setting_config<std::optional<uint8_t>>(const char *) { return detect_settings().color_channel; }
}
