// Test infrastructure (oracle/): a C ABI around the REFERENCE's own crop code -- tracker/tracking/FilterCache.cpp, compiled unmodified (oracle/build_ref.py):
// constraints::diff_image with its four normalisations (:266-299), i.e. image::calculate_diff_image (:157-238: the masked blob image, individual_image_scale,
// the centre pad / centre cut to the output size and the position it reports) and image::calculate_normalized_diff_image -> normalize_image (:21-131: the
// transform handed to cv::warpAffine, the pad / cut that follows, the reported point), over the reference's imageFromLines (processing/Background.cpp),
// gui::Transform (commons/common/gui/Transform.cpp) and Midline::transform (tracker/tracking/Outline.cpp).  constraints::local_midline_length (:321-410, the
// median midline length the `posture` crops are scaled by) runs over a list of frames.  OpenCV's own functions are stand-ins (oracle/ref_stubs/commons.pc.h):
// cv::warpAffine is the oracle's cv2-pinned restatement, installed by the test; pv::Blob is the look-alike of the other wrappers (orientation handed in).
// Never linked into the product.
#include <tracking/FilterCache.h>
#include <tracking/Outline.h>
#include <tracking/Stuffs.h>
#include <tracking/Tracker.h>
#include <processing/PVBlob.h>
#include <gui/Transform.h>
#include <misc/create_struct.h>

using namespace cmn;

extern "C" {

void ref_filtercache_set_warp(cv::warp_fn_t f) { cv::warp_hook() = f; }
void ref_filtercache_set_resize(cv::resize_fn_t f) { cv::resize_hook() = f; }
void ref_filtercache_settings(float individual_image_scale) { outline::Settings::values().individual_image_scale = individual_image_scale; }

// mode: 0 none, 1 moments, 2 posture, 3 legacy (default_config::individual_image_normalization_t).  The midline of modes 2 / 3 carries what a normalised
// midline carries (angle, offset, front = (0, 0)); mode 1 takes the blob's orientation.  with_background = 0 hands a null Background (grey values instead of
// differences).  Returns 1 and fills out (dims = {rows, cols, channels}) and pos, 0 when the reference returns no image, -1 when it throws.
int ref_diff_image(int mode, const uint16_t *in_lines, int64_t n, const uint8_t *in_px, int64_t n_px, int channels, const uint8_t *bg, int w, int h, int bg_channels,
                   int rgb8, int with_background, float orientation, float mid_angle, float mid_offx, float mid_offy, float midline_length, int out_w, int out_h,
                   uint8_t *out, int64_t cap, int32_t *dims, float *pos)
{
    auto img = Image::Make((uint32_t)h, (uint32_t)w, (uint32_t)bg_channels);
    std::memcpy(img->data(), bg, (size_t)w * h * bg_channels);
    Background background(std::move(img), rgb8 ? meta_encoding_t::rgb8 : meta_encoding_t::gray);
    auto l = std::make_unique<blob::lines_t>((size_t)n);
    for (int64_t i = 0; i < n; ++i) (*l)[(size_t)i] = HorizontalLine(in_lines[4 * i + 2], in_lines[4 * i], in_lines[4 * i + 1]);
    pv::Blob blob(std::move(l), std::make_unique<PixelArray_t>(in_px, in_px + n_px), pv::Blob::get_only_flag(pv::Blob::Flags::is_rgb, channels == 3));
    blob._orientation = orientation;
    track::Midline m;
    m.segments().resize(1);
    m.angle() = mid_angle; m.offset() = Vec2(mid_offx, mid_offy);
    using namespace default_config::individual_image_normalization_t;
    const Class type = mode == 1 ? moments : (mode == 2 ? posture : (mode == 3 ? legacy : none));
    gui::Transform tr = mode >= 2 ? m.transform(type) : gui::Transform();
    try {
        auto [image, p] = track::constraints::diff_image(type, &blob, tr, midline_length, Size2((float)out_w, (float)out_h), with_background ? &background : nullptr);
        if (!image) return 0;
        dims[0] = (int32_t)image->rows; dims[1] = (int32_t)image->cols; dims[2] = (int32_t)image->dims;
        const int64_t bytes = (int64_t)image->rows * image->cols * image->dims;
        if (bytes > cap) return -2;
        std::memcpy(out, image->data(), (size_t)bytes);
        pos[0] = p.x; pos[1] = p.y;
        return 1;
    } catch (const std::exception&) {
        return -1;
    }
}

// constraints::local_midline_length over n consecutive frames: per frame the cached midline length / angle (has[i] = 0: posture without a cached midline),
// the outline's point count (0: none), whether the blob was split.  out = {median_midline_length_px, median_number_outline_pts, midline_length_px_std,
// outline_pts_std, median_angle_diff}
void ref_local_midline_length(uint32_t identity, int32_t first_frame, int64_t n, const float *length, const float *angle, const uint8_t *has, const uint32_t *outline_pts,
                              const uint8_t *split, int calculate_std, float *out)
{
    track::Individual fish;
    fish.id = track::Idx_t(identity);
    fish.range = Range<Frame_t>(Frame_t(first_frame), Frame_t(first_frame + (int32_t)n - 1));
    fish.frames.resize((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        auto &[b, p] = fish.frames[(size_t)i];
        b.blob.b = Bounds(10, 10, 5, 5); b.blob.is_split = split[i] != 0;
        p.frame = Frame_t(first_frame + (int32_t)i); p.outline.n = outline_pts[i];
        if (has[i]) { p.midline_length.v = length[i]; p.midline_angle.v = angle[i]; }
    }
    track::constraints::FilterCache::clear();
    auto c = track::constraints::local_midline_length(&fish, fish.range, calculate_std != 0);
    out[0] = c->median_midline_length_px; out[1] = c->median_number_outline_pts; out[2] = c->midline_length_px_std; out[3] = c->outline_pts_std; out[4] = c->median_angle_diff;
}

}
