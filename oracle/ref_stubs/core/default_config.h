// Test infrastructure: the two setting enumerations of tracker/core/default_config.h that Outline.cpp names (peak_mode, :default_config.cpp
// "peak_mode"; individual_image_normalization), as plain scoped-enum look-alikes of TRex's ENUM_CLASS objects.
#pragma once
#include <commons.pc.h>
namespace default_config {
namespace peak_mode_t {
enum Class { pointy, broad };
struct Named { Class v; constexpr Class value() const { return v; } constexpr operator Class() const { return v; } };
constexpr Named pointy_{Class::pointy}, broad_{Class::broad};
}
namespace individual_image_normalization_t {
enum Class { none, moments, posture, legacy };
}
}
