// trexb200.hpp -- header-only C++ host layer above the C ABI (include/trexb200.h), mirroring the
// reference's C++ interfaces for this path so TRex code can call it with its own vocabulary:
//   trexb200::HorizontalLine        == cmn::HorizontalLine        (C/misc/detail.h:71-131)
//   trexb200::Pair                  ~  cmn::blob::Pair             (C/misc/types.h:591-604)
//   trexb200::BackgroundSubtraction ~  track::BackgroundSubtraction (T/python/BackgroundSubtraction.h:10-27)
//   trexb200::labeling_run          ~  cmn::CPULabeling::run       (C/processing/CPULabeling.h:16)
//   trexb200::VINetwork             ~  Python::VINetwork           (T/ml/VisualIdentification.h:104-133)
// Errors surface as std::runtime_error carrying tb_last_error(), which is what the reference's
// callers already handle (promise->set_exception, BackgroundSubtraction.cpp:322-326; SoftException).
#pragma once
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "trexb200.h"

namespace trexb200 {

using HorizontalLine = tb_line;                        // {u16 x0, x1, y, padding}
static_assert(sizeof(HorizontalLine) == 8, "layout of cmn::HorizontalLine");

struct Pair {                                          // blob::Pair: lines + pixels + flags
    std::unique_ptr<std::vector<HorizontalLine>> lines;
    std::unique_ptr<std::vector<uint8_t>> pixels;
    uint8_t extra_flags = 0;
    uint32_t bid = 0;                                  // pv::bid::from_data of the first line
};
using blobs_t = std::vector<Pair>;

inline void check(int rc, const char *what)
{
    if (rc != TB_OK) throw std::runtime_error(std::string(what) + ": " + tb_last_error());
}

// meta_encoding_t (T/core/default_config.cpp:59): the encodings this library builds
enum class meta_encoding_t { gray = 0, rgb8 = 1, r3g3b2 = 2 };

class BackgroundSubtraction {
public:
    // channels = Image::dims of the frames apply() receives (TileImage.images[0]): 1 gray, 3 BGR, 4 BGRA.
    // Like BackgroundSubtraction::apply (T/python/BackgroundSubtraction.cpp:177-181), rgb8 refuses gray frames
    // ("Invalid number of channels"): here at construction instead of per image.
    BackgroundSubtraction(int width, int height, int max_batch = 1, int max_individuals = 0, int device = 0,
                          int channels = 1, meta_encoding_t encoding = meta_encoding_t::gray)
    {
        tb_seg_config cfg{};
        cfg.device = device; cfg.width = width; cfg.height = height; cfg.max_batch = max_batch;
        cfg.max_crops_per_frame = max_individuals; cfg.crop_width = 80; cfg.crop_height = 80; cfg.crop_method = 1;
        cfg.channels = channels; cfg.encoding = (int)encoding;
        check(tb_seg_create(&cfg, &_h), "tb_seg_create");
        tb_seg_default_params(&_p);
        _w = width; _hgt = height; _opx = encoding == meta_encoding_t::rgb8 ? 3 : 1; _r3 = encoding == meta_encoding_t::r3g3b2;
    }
    ~BackgroundSubtraction() { tb_seg_destroy(_h); }
    BackgroundSubtraction(const BackgroundSubtraction &) = delete;
    BackgroundSubtraction &operator=(const BackgroundSubtraction &) = delete;

    tb_seg_params &settings() { return _p; }           // detect_threshold, detect_size_filter, cm_per_pixel ...
    void update_settings() { check(tb_seg_set_params(_h, &_p), "tb_seg_set_params"); }

    // set_background(Image::Ptr&&): un-pauses the pipeline (BackgroundSubtraction.cpp:86-90)
    // the average image has 1 channel for gray and 3 (B,G,R) for rgb8 (RawProcessing.cpp:343)
    void set_background(const uint8_t *average, int64_t stride = 0)
    {
        check(tb_seg_set_background_c(_h, average, _w, _hgt, _opx, stride), "tb_seg_set_background");
    }

    // apply(std::vector<TileImage>&&): one blobs_t per image (BackgroundSubtraction.cpp:126-347)
    std::vector<blobs_t> apply(const std::vector<const uint8_t *> &images, int64_t stride = 0)
    {
        check(tb_seg_submit(_h, images.data(), (int)images.size(), stride, 1), "tb_seg_submit");
        check(tb_seg_wait(_h), "tb_seg_wait");
        std::vector<blobs_t> out(images.size());
        for (size_t i = 0; i < images.size(); ++i) {
            tb_blob_view v;
            check(tb_seg_result(_h, (int)i, &v), "tb_seg_result");
            out[i].reserve(v.info.n_blobs);
            for (uint32_t k = 0; k < v.info.n_blobs; ++k) {
                const tb_blob_rec &r = v.recs[k];
                Pair p;
                const tb_line *l = v.lines + (r.line_off - v.info.line_begin);
                const uint8_t *px = v.pixels + (r.px_off - v.info.px_begin);
                p.lines = std::make_unique<std::vector<HorizontalLine>>(l, l + r.n_lines);
                p.pixels = std::make_unique<std::vector<uint8_t>>(px, px + (size_t)r.n_pixels * _opx);
                p.extra_flags = _opx == 3 ? (1u << 5) : (_r3 ? (1u << 6) : 0);     // pv::Blob::Flags::is_rgb = bit 5 (CPULabeling.cpp:193), is_r3g3b2 = bit 6 (BackgroundSubtraction.cpp:221; PVBlob.h:138-169)
                p.bid = r.bid;
                out[i].emplace_back(std::move(p));
            }
        }
        return out;
    }
    // Posture's first stage for every blob of the last apply(), in apply()'s order (frame by frame): the points
    // pixel::find_outer_points + the longest-outline choice + Outline::resample(outline_resample) leave in
    // posture::Result::outline before calculate_midline (T/tracking/Posture.cpp:326-352); x,y pairs relative to the blob's bounds.
    std::vector<std::vector<float>> outlines(float outline_resample = 1.f)
    {
        check(tb_seg_outlines(_h, outline_resample), "tb_seg_outlines");
        const tb_outline_rec *recs = nullptr; const float *raw = nullptr, *pts = nullptr; uint32_t n = 0;
        check(tb_seg_outline_result(_h, &recs, &raw, &pts, &n), "tb_seg_outline_result");
        std::vector<std::vector<float>> out(n);
        for (uint32_t k = 0; k < n; ++k) out[k].assign(pts + 2 * (size_t)recs[k].res_off, pts + 2 * ((size_t)recs[k].res_off + recs[k].n_res));
        return out;
    }
    // Posture's second stage, after outlines(): Outline::calculate_midline per blob (T/tracking/Outline.cpp:768-868; peak_mode
    // pointy).  One entry per blob of the last apply(): the midline segments {pos.x, pos.y, height, l_length} from the tail on
    // (empty where the reference reports too few segments) and the tail / head indices into the walked outline.
    struct Midline { std::vector<float> segments; int tail_index = -1, head_index = -1; };
    std::vector<Midline> midlines(float outline_resample = 1.f)
    {
        check(tb_seg_outlines(_h, outline_resample), "tb_seg_outlines");
        tb_posture_params pp; tb_posture_default_params(&pp);
        check(tb_seg_midlines(_h, &pp), "tb_seg_midlines");
        const tb_midline_rec *recs = nullptr; const float *pts = nullptr, *segs = nullptr; uint32_t n = 0;
        check(tb_seg_midline_result(_h, &recs, &pts, &segs, &n), "tb_seg_midline_result");
        std::vector<Midline> out(n);
        for (uint32_t k = 0; k < n; ++k) {
            out[k].segments.assign(segs + 4 * (size_t)recs[k].seg_off, segs + 4 * ((size_t)recs[k].seg_off + recs[k].n_seg));
            out[k].tail_index = recs[k].tail; out[k].head_index = recs[k].head;
        }
        return out;
    }
    // The whole posture chain of the last apply() in one call: Individual::calculate_midline_for's result per blob
    // (T/tracking/Individual.cpp:1348-1383 = Outline::calculate_midline + Midline::post_process + Midline::normalize): the normalised
    // midline (midline_resolution segments {pos.x, pos.y, height, l_length}; empty where the reference has none) with Midline::len(),
    // angle(), offset() -- what Midline::transform(type) needs (T/tracking/Outline.cpp:1238-1256).
    struct NormalizedMidline { std::vector<float> segments; float len = 0, angle = 0, offset_x = 0, offset_y = 0; int tail_index = -1, head_index = -1; bool inverted_because_previous = false; };
    std::vector<NormalizedMidline> posture(float outline_resample = 1.f, const tb_posture_params *params = nullptr, float median_midline_length_px = 0.f)
    {
        tb_posture_request q; tb_posture_default_request(&q);
        if (params) q.params = *params;
        q.outline_resample = outline_resample; q.normalize = 1; q.fetch = 1; q.median_midline_length_px = median_midline_length_px;
        check(tb_seg_posture(_h, &q), "tb_seg_posture");
        check(tb_seg_posture_wait(_h), "tb_seg_posture_wait");
        tb_posture_view v;
        check(tb_seg_posture_result(_h, &v), "tb_seg_posture_result");
        std::vector<NormalizedMidline> out(v.n_blobs);
        for (uint32_t k = 0; k < v.n_blobs; ++k) {
            const tb_midline_norm &n = v.normalized[k];
            NormalizedMidline &m = out[k];
            m.len = n.len; m.angle = n.angle; m.offset_x = n.offx; m.offset_y = n.offy; m.tail_index = n.tail; m.head_index = n.head;
            m.inverted_because_previous = (n.flags & 1u) != 0;
            if (n.n_points) m.segments.assign(v.norm_points + 4 * (size_t)k * v.midline_resolution, v.norm_points + 4 * ((size_t)k + 1) * v.midline_resolution);
        }
        return out;
    }
    // pv::Blob::recount(threshold, background) for every blob of the last apply() (C/processing/PVBlob.cpp:934-1027)
    std::vector<float> recount(int threshold)
    {
        uint32_t t[4];
        check(tb_seg_totals(_h, t), "tb_seg_totals");
        std::vector<float> out(t[0] ? t[0] : 1);
        check(tb_seg_recount(_h, threshold, out.data(), (uint32_t)out.size()), "tb_seg_recount");
        out.resize(t[0]);
        return out;
    }
    tb_seg *handle() { return _h; }

private:
    tb_seg *_h = nullptr;
    tb_seg_params _p{};
    int _w = 0, _hgt = 0, _opx = 1;
    bool _r3 = false;
};

// CPULabeling::run(const cv::Mat&, ...): label an already-binary image (any non-zero pixel is foreground)
inline blobs_t labeling_run(const uint8_t *image, int width, int height)
{
    BackgroundSubtraction bs(width, height, 1, 0);
    bs.settings().detect_threshold = 0; bs.settings().enable_difference = 0; bs.settings().n_size_ranges = 0;
    bs.update_settings();
    std::vector<uint8_t> zero((size_t)width * height, 0);
    bs.set_background(zero.data());
    return std::move(bs.apply({image})[0]);
}

class VINetwork {
public:
    // version: visual_identification_version as tb_vi_config.arch (0 v118_3, 1 v100, 2 v110, 3 v119, 4 v200)
    VINetwork(int num_classes, int max_images = 4096, int device = 0, int channels = 1, int version = 0) : _m(num_classes), _c(channels)
    {
        tb_vi_config cfg{};
        cfg.device = device; cfg.width = 80; cfg.height = 80; cfg.channels = channels;
        cfg.num_classes = num_classes; cfg.max_images = max_images;
        cfg.arch = version;
        cfg.precision = version <= 1 ? 3 : 0;   // v118_3, v100: "fp16c" on tensor cores (fp16 + e5m2 correction terms, ~2e-5 of fp32 on O(1) logits); the others: fp32
        check(tb_vi_create(&cfg, &_h), "tb_vi_create");
    }
    ~VINetwork() { tb_vi_destroy(_h); }
    VINetwork(const VINetwork &) = delete;
    VINetwork &operator=(const VINetwork &) = delete;

    // load_weights(VIWeights&&): tensors under their state_dict names, then commit
    void set_tensor(const std::string &name, const std::vector<float> &v) { check(tb_vi_set_tensor(_h, name.c_str(), v.data(), (int64_t)v.size()), "tb_vi_set_tensor"); }
    void commit() { check(tb_vi_commit(_h), "tb_vi_commit"); }

    // probabilities(std::vector<Image::Ptr>&&) (sync form): N x M softmax rows, flat
    std::vector<float> probabilities(const std::vector<const uint8_t *> &images)
    {
        const size_t ib = (size_t)6400 * _c;
        std::vector<uint8_t> packed(images.size() * ib);
        for (size_t i = 0; i < images.size(); ++i) std::memcpy(packed.data() + i * ib, images[i], ib);
        std::vector<float> probs(images.size() * (size_t)_m);
        if (!images.empty()) check(tb_vi_predict(_h, packed.data(), (int)images.size(), probs.data(), nullptr), "tb_vi_predict");
        return probs;
    }

    // paverages(ids, images) (T/ml/VisualIdentification.h:146-181): the mean probability row of the images of every id -- rows are
    // added in image order in float and divided by float(samples), like the reference's std::transform chain
    struct Average { size_t samples = 0; std::vector<float> values; };
    std::map<uint32_t, Average> paverages(const std::vector<uint32_t> &ids, const std::vector<const uint8_t *> &images)
    {
        if (ids.size() != images.size()) throw std::runtime_error("paverages: ids and images differ in length");
        const std::vector<float> probs = probabilities(images);
        std::map<uint32_t, Average> averages;
        for (size_t i = 0; i < ids.size(); ++i) {
            Average &a = averages[ids[i]];
            if (a.values.empty()) { a.values.assign((size_t)_m, 0.f); a.samples = 0; }
            ++a.samples;
            for (int k = 0; k < _m; ++k) a.values[(size_t)k] = probs[i * (size_t)_m + (size_t)k] + a.values[(size_t)k];
        }
        for (auto &kv : averages) {
            const float N = float(kv.second.samples);
            for (float &v : kv.second.values) v = v / N;
        }
        return averages;
    }

private:
    tb_vi *_h = nullptr;
    int _m, _c;
};

// Page-locked frame buffer for the pool BackgroundSubtraction::apply reads from (buffers::TileBuffers / ImageMaker,
// T/core/TileBuffers.h:9-22): tb_seg_submit copies from it at the full PCIe rate and asynchronously.
class HostFrame {
public:
    explicit HostFrame(size_t bytes) : _n(bytes) { void *p = nullptr; check(tb_host_alloc(bytes, &p), "tb_host_alloc"); _p = (uint8_t *)p; }
    ~HostFrame() { tb_host_free(_p); }
    HostFrame(const HostFrame &) = delete;
    HostFrame &operator=(const HostFrame &) = delete;
    uint8_t *data() { return _p; }
    size_t size() const { return _n; }
private:
    uint8_t *_p = nullptr;
    size_t _n = 0;
};

}  // namespace trexb200
