// Test infrastructure (oracle/): a C ABI around the REFERENCE's own posture::calculate_posture(Frame_t, pv::BlobWeakPtr) (tracker/tracking/Posture.cpp:305-400):
// the threshold loop over pixel::threshold_get_biggest_blob -> find_outer_points -> longest outline -> Outline::resample -> calculate_midline, the +2
// retries and the first-outline fall-back -- every part of it the reference's code, compiled unmodified (oracle/build_ref.py).  The stand-ins supply the
// blob (pv::Blob look-alike), the Background object Tracker::background() returns, and the settings.  Never linked into the product.
#include <tracking/Posture.h>
#include <tracking/Tracker.h>
#include <processing/PVBlob.h>

using namespace cmn;

extern "C" {

void ref_posture_settings(int track_posture_threshold, float outline_resample)
{
    auto &v = outline::Settings::values();
    v.track_posture_threshold = track_posture_threshold; v.outline_resample = outline_resample; v.posture_closing_steps = 0; v.outline_compression = 0;
}

// Returns the number of midline segments (>= 0), -1 when the result carries an outline but no midline (the fall-back), -2 for std::unexpected.
// pts: the result's outline (n_pts points; blob-relative coordinates like the reference's)
int64_t ref_calculate_posture(const uint16_t *in_lines, int64_t n, const uint8_t *in_px, int64_t n_px, int channels, const uint8_t *bg, int w, int h, int bg_channels, int rgb8,
                              float *pts, int64_t cap_pts, int64_t *n_pts, float *segments, int64_t cap_seg, int64_t *tail, int64_t *head)
{
    auto img = Image::Make((uint32_t)h, (uint32_t)w, (uint32_t)bg_channels);
    std::memcpy(img->data(), bg, (size_t)w * h * bg_channels);
    Background background(std::move(img), rgb8 ? meta_encoding_t::rgb8 : meta_encoding_t::gray);
    track::Tracker::background_slot() = &background;
    auto l = std::make_unique<blob::lines_t>((size_t)n);
    for (int64_t i = 0; i < n; ++i) (*l)[(size_t)i] = HorizontalLine(in_lines[4 * i + 2], in_lines[4 * i], in_lines[4 * i + 1]);
    pv::Blob blob(std::move(l), std::make_unique<PixelArray_t>(in_px, in_px + n_px), pv::Blob::get_only_flag(pv::Blob::Flags::is_rgb, channels == 3));
    auto r = track::posture::calculate_posture(Frame_t{}, &blob);
    track::Tracker::background_slot() = nullptr;
    if (!r) return -2;
    auto &res = r.value();
    *n_pts = (int64_t)res.outline.size();
    for (int64_t i = 0; i < *n_pts && i < cap_pts; ++i) { pts[2 * i] = res.outline[(size_t)i].x; pts[2 * i + 1] = res.outline[(size_t)i].y; }
    if (!res.midline) return -1;
    *tail = res.midline->tail_index(); *head = res.midline->head_index();
    const auto &s = res.midline->segments();
    if ((int64_t)s.size() > cap_seg) return -4;
    for (size_t i = 0; i < s.size(); ++i) { segments[4 * i] = s[i].pos.x; segments[4 * i + 1] = s[i].pos.y; segments[4 * i + 2] = s[i].height; segments[4 * i + 3] = s[i].l_length; }
    return (int64_t)s.size();
}

}
