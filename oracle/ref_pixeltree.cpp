// Test infrastructure (oracle/): a C ABI around the REFERENCE's own pixel::find_outer_points (commons/common/processing/PixelTree.cpp:497-651 with
// pixel::Tree::add / generate_edges / walk, :652-1133), compiled unmodified from the reference checkout (oracle/build_ref.py; stand-ins in
// oracle/ref_stubs/processing/pixeltree_standins.h).  Called like posture::calculate_posture does (tracker/tracking/Posture.cpp:330-348).
// Never linked into the product.
#include <processing/PixelTree.h>
#include <processing/PVBlob.h>

extern "C" {

// lines: n x {x0, x1, y, pad} (the memory layout of HorizontalLine).  pts: all outlines back to back, off[k] .. off[k + 1] = outline k.
// Returns the number of outlines, -4 when a capacity is too small.
int64_t ref_find_outer_points(const uint16_t *lines, int64_t n, float *pts, int64_t cap_pts, int64_t *off, int64_t cap_off)
{
    auto l = std::make_unique<cmn::blob::lines_t>((size_t)n);
    for (int64_t i = 0; i < n; ++i) (*l)[(size_t)i] = cmn::HorizontalLine(lines[4 * i + 2], lines[4 * i], lines[4 * i + 1]);
    pv::Blob blob(std::move(l), nullptr);
    auto outlines = cmn::pixel::find_outer_points(&blob, 0);
    int64_t k = 0, total = 0;
    off[0] = 0;
    for (auto &o : outlines) {
        if (k + 1 > cap_off || total + (int64_t)o->size() > cap_pts) return -4;
        for (auto &p : *o) { pts[2 * total] = p.x; pts[2 * total + 1] = p.y; ++total; }
        off[++k] = total;
    }
    return k;
}


// pixel::threshold_blob(cache, blob, difference_cache, threshold, Rangel(-1, -1))  (PixelTree.cpp:362-374 over _threshold_blob :90-184): the overload that
// takes the per-pixel difference values ready-made (tracker/tracking/SplitBlob.cpp:164) -- the same run cutting, relabeling (CPULabeling::run, the
// reference's own, compiled) and `pixels->size() > 1` rule as the Background overload the tracker calls (:344-356), without needing a Background.
// lines: n x {x0, x1, y, pad}; pixels: channels bytes per pixel; diff: one byte per pixel.  Output like ref_label_image.
int64_t ref_threshold_blob_cache(const uint16_t *in_lines, int64_t n, const uint8_t *in_px, int64_t n_px, int channels, const uint8_t *diff, int threshold,
                                 uint16_t *lines, int64_t cap_lines, uint8_t *pixels, int64_t cap_px, int64_t *line_off, int64_t *px_off, uint8_t *flags, int64_t cap_blobs)
{
    auto l = std::make_unique<cmn::blob::lines_t>((size_t)n);
    for (int64_t i = 0; i < n; ++i) (*l)[(size_t)i] = cmn::HorizontalLine(in_lines[4 * i + 2], in_lines[4 * i], in_lines[4 * i + 1]);
    auto px = std::make_unique<cmn::PixelArray_t>(in_px, in_px + n_px);
    pv::Blob blob(std::move(l), std::move(px), pv::Blob::get_only_flag(pv::Blob::Flags::is_rgb, channels == 3));
    cmn::PixelArray_t cache(diff, diff + n_px / channels);
    cmn::CPULabeling::ListCache_t lc;
    auto out = cmn::pixel::threshold_blob(lc, &blob, cache, threshold, cmn::Rangel(-1, -1));
    int64_t k = 0, nl = 0, np = 0;
    line_off[0] = 0; px_off[0] = 0;
    for (auto &b : out) {
        if (k >= cap_blobs) return -4;
        for (auto &h : b->hor_lines()) {
            if (nl >= cap_lines) return -4;
            lines[4 * nl] = h.x0; lines[4 * nl + 1] = h.x1; lines[4 * nl + 2] = h.y; lines[4 * nl + 3] = 0; ++nl;
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

}
