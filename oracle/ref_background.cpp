// Test infrastructure (oracle/): a C ABI around the REFERENCE's own Background code -- commons/common/processing/Background.{h,cpp} with processing/encoding.h,
// misc/EnumClass.h and misc/matharray.h (per-pixel difference / is_different / count_above_threshold, the tracker's bgr2gray, imageFromLines) and, on
// top of it, the Background overload of pixel::threshold_blob (PixelTree.cpp:186-356) -- compiled unmodified from the reference checkout
// (oracle/build_ref.py; cmn::Image, cv::Mat, the settings callbacks and pv::Blob are stand-ins in oracle/ref_stubs/).  Never linked into the product.
#include <processing/Background.h>
#include <processing/PixelTree.h>
#include <processing/PVBlob.h>

using namespace cmn;

static std::unique_ptr<Background> make_background(const uint8_t *bg, int w, int h, int bg_channels, int rgb8)
{
    auto img = Image::Make((uint32_t)h, (uint32_t)w, (uint32_t)bg_channels);
    std::memcpy(img->data(), bg, (size_t)w * h * bg_channels);
    return std::make_unique<Background>(std::move(img), rgb8 ? meta_encoding_t::rgb8 : meta_encoding_t::gray);
}

static std::unique_ptr<pv::Blob> make_blob(const uint16_t *lines, int64_t n, const uint8_t *px, int64_t n_px, int channels)
{
    auto l = std::make_unique<blob::lines_t>((size_t)n);
    for (int64_t i = 0; i < n; ++i) (*l)[(size_t)i] = HorizontalLine(lines[4 * i + 2], lines[4 * i], lines[4 * i + 1]);
    auto p = std::make_unique<PixelArray_t>(px, px + n_px);
    return std::make_unique<pv::Blob>(std::move(l), std::move(p), pv::Blob::get_only_flag(pv::Blob::Flags::is_rgb, channels == 3));
}

extern "C" {

// track_threshold_is_absolute, track_background_subtraction, meta_encoding (0 gray, 2 rgb8): the three settings Background.cpp caches through its callbacks
void ref_background_settings(int absolute, int subtraction, int meta_encoding)
{
    auto &s = ref_settings();
    s.track_threshold_is_absolute = absolute != 0; s.track_background_subtraction = subtraction != 0; s.meta_encoding = meta_encoding;
    (void)Background::track_threshold_is_absolute();                       // registers the callbacks on first use ...
    if (s.cb) for (const char *n : {"track_threshold_is_absolute", "track_background_subtraction", "meta_encoding"}) s.cb(n);      // ... and they fire on every change
}

// pixel::threshold_blob(cache, blob, threshold, background)  -- the call of Tracker.cpp:833 (size_range = Rangel(-1, -1)).
// bg: h x w x bg_channels (1 for gray encoding, 3 = B,G,R for rgb8).  Output like ref_label_image.
int64_t ref_threshold_blob_bg(const uint16_t *in_lines, int64_t n, const uint8_t *in_px, int64_t n_px, int channels, const uint8_t *bg, int w, int h, int bg_channels,
                              int rgb8, int threshold, uint16_t *lines, int64_t cap_lines, uint8_t *pixels, int64_t cap_px, int64_t *line_off, int64_t *px_off,
                              uint8_t *flags, int64_t cap_blobs)
{
    auto background = make_background(bg, w, h, bg_channels, rgb8);
    auto blob = make_blob(in_lines, n, in_px, n_px, channels);
    CPULabeling::ListCache_t cache;
    auto out = pixel::threshold_blob(cache, blob.get(), threshold, background.get());
    int64_t k = 0, nl = 0, np = 0;
    line_off[0] = 0; px_off[0] = 0;
    for (auto &b : out) {
        if (k >= cap_blobs) return -4;
        for (auto &hl : b->hor_lines()) {
            if (nl >= cap_lines) return -4;
            lines[4 * nl] = hl.x0; lines[4 * nl + 1] = hl.x1; lines[4 * nl + 2] = hl.y; lines[4 * nl + 3] = 0; ++nl;
        }
        if (b->pixels()) {
            if (np + (int64_t)b->pixels()->size() > cap_px) return -4;
            std::memcpy(pixels + np, b->pixels()->data(), b->pixels()->size());
            np += (int64_t)b->pixels()->size();
        }
        flags[k] = b->flags();
        ++k;
        line_off[k] = nl; px_off[k] = np;
    }
    return k;
}

// pv::Blob::raw_recount(threshold, background) for threshold > 0 (processing/PVBlob.cpp:953-1017): the same dispatch (call_image_mode_function with the
// known gray output) and the same Background::count_above_threshold per run, summed in a float
float ref_raw_recount(const uint16_t *in_lines, int64_t n, const uint8_t *in_px, int64_t n_px, int channels, const uint8_t *bg, int w, int h, int bg_channels, int rgb8, int threshold)
{
    auto background_p = make_background(bg, w, h, bg_channels, rgb8);
    const Background &background = *background_p;
    auto blob = make_blob(in_lines, n, in_px, n_px, channels);
    float recount = 0;
    auto ptr = blob->pixels()->data();
    InputInfo input = blob->input_info();
    constexpr OutputInfo output{.channels = 1u, .encoding = meta_encoding_t::gray};
    call_image_mode_function<output>(input, KnownOutputType{}, [&]<InputInfo in, OutputInfo out, DifferenceMethod method>() {
        for (auto &line : blob->hor_lines()) {
            if constexpr (in.channels == 0) {
                recount += ptr_safe_t(line.x1) - ptr_safe_t(line.x0) + 1;
            } else {
                const auto L = (ptr_safe_t(line.x1) - ptr_safe_t(line.x0) + 1) * in.channels;
                recount += background.count_above_threshold<method, in>(line.x0, line.x1, line.y, std::span(ptr, ptr + L), threshold);
                ptr += L;
            }
        }
    });
    return recount;
}

// imageFromLines(input, lines, &mask, &grey, &difference, pixels, base_threshold, background, 0)  (Background.cpp:113-299): the three images of a blob's bounding
// box (mask: 1 byte per pixel; grey / difference: out_channels = 1 (gray input) or 3 bytes per pixel).  rect = {x, y, width, height}; returns the recount.
int64_t ref_image_from_lines(const uint16_t *in_lines, int64_t n, const uint8_t *in_px, int64_t n_px, int channels, const uint8_t *bg, int w, int h, int bg_channels, int rgb8,
                             int base_threshold, int with_background, int32_t *rect, uint8_t *mask, uint8_t *grey, uint8_t *diff)
{
    auto background = make_background(bg, w, h, bg_channels, rgb8);
    auto blob = make_blob(in_lines, n, in_px, n_px, channels);
    cv::Mat m_mask, m_grey, m_diff;
    auto [r, cnt] = imageFromLines(blob->input_info(), blob->hor_lines(), &m_mask, &m_grey, with_background ? &m_diff : nullptr, blob->pixels().get(), base_threshold,
                                   with_background ? background.get() : nullptr, 0);
    rect[0] = r.x; rect[1] = r.y; rect[2] = r.width; rect[3] = r.height;
    const int oc = channels == 3 ? 3 : 1;
    for (int y = 0; y < r.height; ++y) {
        std::memcpy(mask + (size_t)y * r.width, m_mask.ptr(y), (size_t)r.width);
        std::memcpy(grey + (size_t)y * r.width * oc, m_grey.ptr(y), (size_t)r.width * oc);
        if (with_background) std::memcpy(diff + (size_t)y * r.width * oc, m_diff.ptr(y), (size_t)r.width * oc);
    }
    return (int64_t)cnt;
}

}
