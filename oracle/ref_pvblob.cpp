// Test infrastructure (oracle/): a C ABI around the REFERENCE's own pv::Blob -- commons/common/processing/PVBlob.{h,cpp} with processing/BlobIdentity.cpp and
// misc/bid.h, compiled unmodified (oracle/build_ref.py build_pvblob, -DREF_REAL_PVBLOB; Grid / ProximityGrid / DataFormat / Buffers are placeholders in
// oracle/ref_stubs_pvblob/, the rest comes from oracle/ref_stubs/): the constructor's init() -> calculate_properties (bounds, centre, pixel count) and the blob
// id, calculate_moments -> orientation (the angle the `moments` crop normalisation rotates by, FilterCache.cpp:329-341), recount(threshold, background) =
// raw_recount * SQR(cm_per_pixel) with its threshold-0 shortcut and its cache, and threshold(value, background).  Never linked into the product.
#include <processing/PVBlob.h>
#include <processing/Background.h>
#include <misc/create_struct.h>

using namespace cmn;

static std::unique_ptr<pv::Blob> make(const uint16_t *in_lines, int64_t n, const uint8_t *in_px, int64_t n_px, int channels)
{
    auto l = std::make_unique<blob::lines_t>((size_t)n);
    for (int64_t i = 0; i < n; ++i) (*l)[(size_t)i] = HorizontalLine(in_lines[4 * i + 2], in_lines[4 * i], in_lines[4 * i + 1]);
    blob::pixel_ptr_t px = in_px ? std::make_unique<PixelArray_t>(in_px, in_px + n_px) : nullptr;
    return std::make_unique<pv::Blob>(std::move(l), std::move(px), pv::Blob::get_only_flag(pv::Blob::Flags::is_rgb, channels == 3), blob::Prediction{});
}

extern "C" {

// track_threshold_is_absolute, track_background_subtraction, meta_encoding (0 gray, 2 rgb8): the settings Background.cpp caches through its callbacks (as in oracle/ref_background.cpp)
void ref_background_settings(int absolute, int subtraction, int meta_encoding)
{
    auto &s = ref_settings();
    s.track_threshold_is_absolute = absolute != 0; s.track_background_subtraction = subtraction != 0; s.meta_encoding = meta_encoding;
    (void)Background::track_threshold_is_absolute();
    for (auto &cb : s.cbs) for (const char *n : {"track_threshold_is_absolute", "track_background_subtraction", "meta_encoding"}) cb(n);
}

// out = {orientation, centre.x, centre.y (after calculate_moments), bounds.x, .y, .width, .height, num_pixels, centre.x, centre.y (before: calculate_properties)}; returns the blob id
uint32_t ref_pvblob_properties(const uint16_t *in_lines, int64_t n, float *out)
{
    auto b = make(in_lines, n, nullptr, 0, 1);
    out[8] = b->center().x; out[9] = b->center().y;
    out[3] = b->bounds().x; out[4] = b->bounds().y; out[5] = b->bounds().width; out[6] = b->bounds().height; out[7] = (float)b->num_pixels();
    b->calculate_moments();
    out[0] = b->orientation(); out[1] = b->center().x; out[2] = b->center().y;
    return (uint32_t)b->blob_id();
}

// pv::Blob::recount(threshold, background): twice, to go through the cache (PVBlob.cpp:941-951) -- both results are returned
void ref_pvblob_recount(const uint16_t *in_lines, int64_t n, const uint8_t *in_px, int64_t n_px, int channels, const uint8_t *bg, int w, int h, int bg_channels, int rgb8,
                        int threshold, float cm_per_pixel, float *out)
{
    auto img = Image::Make((uint32_t)h, (uint32_t)w, (uint32_t)bg_channels);
    std::memcpy(img->data(), bg, (size_t)w * h * bg_channels);
    Background background(std::move(img), rgb8 ? meta_encoding_t::rgb8 : meta_encoding_t::gray);
    pv::PVSettings::cm() = cm_per_pixel;
    auto b = make(in_lines, n, in_px, n_px, channels);
    out[0] = b->recount(threshold, background);
    out[1] = b->recount(threshold, background);
    out[2] = (float)b->last_recount_threshold();
}

}
