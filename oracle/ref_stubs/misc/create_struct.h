// Test infrastructure: stands in for TRex's settings cache generator (commons/common/misc/create_struct.h).  Outline.cpp declares
// `CREATE_STRUCT(Settings, (type, name)...)` in namespace outline and reads `outline::Settings::copy<outline::Settings::name>()`; here the struct is
// a plain table that the test wrapper fills (ref_outline_settings), and the macro expands to nothing.
#pragma once
#include <commons.pc.h>
#include <core/default_config.h>
namespace outline {
struct Settings {
    enum Variables { outline_curvature_range_ratio, outline_use_dft, peak_mode, midline_walk_offset, outline_approximate, outline_smooth_samples,
                     midline_start_with_head, midline_stiff_percentage, midline_resolution, posture_closing_steps, posture_closing_size, outline_resample };
    struct Values {
        float outline_curvature_range_ratio = 0.03f; bool outline_use_dft = false; default_config::peak_mode_t::Class peak_mode = default_config::peak_mode_t::pointy;
        float midline_walk_offset = 0.025f; uint8_t outline_approximate = 3; uint8_t outline_smooth_samples = 4; bool midline_start_with_head = false;
        float midline_stiff_percentage = 0.15f; uint32_t midline_resolution = 25; uint8_t posture_closing_steps = 0; uint8_t posture_closing_size = 2;
        float outline_resample = 1.f;
        long_t outline_smooth_step = 1; bool midline_invert = false;          // the two FAST_SETTINGs Outline.cpp reads
        int track_posture_threshold = 0; float outline_compression = 0.f;      // + what Posture.cpp reads (with posture_closing_*, outline_resample above)
        float individual_image_scale = 1.f; bool calculate_posture = true;     // + what FilterCache.cpp reads
    };
    static Values& values() { static Values v; return v; }
    static void init() {}
    template<Variables V> static auto copy() {
        auto &v = values();
        if constexpr (V == outline_curvature_range_ratio) return v.outline_curvature_range_ratio;
        else if constexpr (V == outline_use_dft) return v.outline_use_dft;
        else if constexpr (V == peak_mode) return v.peak_mode;
        else if constexpr (V == midline_walk_offset) return v.midline_walk_offset;
        else if constexpr (V == outline_approximate) return v.outline_approximate;
        else if constexpr (V == outline_smooth_samples) return v.outline_smooth_samples;
        else if constexpr (V == midline_start_with_head) return v.midline_start_with_head;
        else if constexpr (V == midline_stiff_percentage) return v.midline_stiff_percentage;
        else if constexpr (V == midline_resolution) return v.midline_resolution;
        else if constexpr (V == posture_closing_steps) return v.posture_closing_steps;
        else if constexpr (V == posture_closing_size) return v.posture_closing_size;
        else return v.outline_resample;
    }
};
}
#ifdef REF_REAL_PVBLOB
namespace pv {            // CREATE_STRUCT(PVSettings, (Float2_t, cm_per_pixel), (bool, correct_illegal_lines)) of processing/PVBlob.cpp:58-61 as a plain table
struct PVSettings {
    enum Variables { cm_per_pixel, correct_illegal_lines };
    static float& cm() { static float v = 1.f; return v; }
    static void init() {}
    template<Variables V> static auto get() { if constexpr (V == cm_per_pixel) return cm(); else return false; }
};
}
#endif
#define CREATE_STRUCT(NAME, ...) static_assert(true, "settings table provided by the stand-in");
#define FAST_SETTING(NAME) (outline::Settings::values().NAME)
