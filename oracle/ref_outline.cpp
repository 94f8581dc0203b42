// Test infrastructure (oracle/): a C ABI around the REFERENCE's own posture-chain code -- tracker/tracking/Outline.cpp (Outline::resample, smooth,
// offset_to_middle, calculate_midline; Midline::post_process, normalize, fix_length), compiled unmodified from the reference checkout together with
// commons/common/misc/{CircularGraph,curve_discussion}.cpp and commons/common/gui/Transform.cpp (oracle/build_ref.py; stand-ins for TRex's precompiled
// header, settings cache and unrelated headers in oracle/ref_stubs/).  The calls follow tracker/tracking/Posture.cpp:226-302 and
// Individual.cpp:507-522, so that tests/test_oracle_ref_outline.py can hold oracle/trex_oracle.c's restatements to the real code.
// Never linked into the product.
#include <tracking/Outline.h>
#include <misc/create_struct.h>
#include <gui/Transform.h>

namespace track {
// Outline declares `friend class DebugDrawing` (Outline.h): the name opens the protected members to this wrapper
class DebugDrawing {
public:
    static std::vector<Vec2>& points(Outline& o) { return *o._points; }
};
}

using namespace track;

static std::unique_ptr<std::vector<Vec2>> to_points(const float *p, int64_t n)
{
    auto v = std::make_unique<std::vector<Vec2>>((size_t)n);
    for (int64_t i = 0; i < n; ++i) (*v)[(size_t)i] = Vec2(p[2 * i], p[2 * i + 1]);
    return v;
}

static void fill_midline(Midline& m, const float *seg, int64_t n, int64_t tail, int64_t head)
{
    m.segments().resize((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        auto &s = m.segments()[(size_t)i];
        s.pos = Vec2(seg[4 * i], seg[4 * i + 1]); s.height = seg[4 * i + 2]; s.l_length = seg[4 * i + 3];
    }
    m.tail_index() = (long_t)tail; m.head_index() = (long_t)head;
}

static void store_segments(const std::vector<MidlineSegment>& s, float *out)
{
    for (size_t i = 0; i < s.size(); ++i) { out[4 * i] = s[i].pos.x; out[4 * i + 1] = s[i].pos.y; out[4 * i + 2] = s[i].height; out[4 * i + 3] = s[i].l_length; }
}

extern "C" {

// the settings Outline.cpp reads (outline::Settings cache + the two FAST_SETTINGs)
void ref_outline_settings(float curvature_range_ratio, int peak_mode_broad, float midline_walk_offset, int outline_approximate, int outline_smooth_samples,
                          int outline_smooth_step, int midline_start_with_head, int midline_invert, float midline_stiff_percentage, uint32_t midline_resolution)
{
    auto &v = outline::Settings::values();
    v.outline_curvature_range_ratio = curvature_range_ratio;
    v.peak_mode = peak_mode_broad ? default_config::peak_mode_t::broad : default_config::peak_mode_t::pointy;
    v.midline_walk_offset = midline_walk_offset; v.outline_approximate = (uint8_t)outline_approximate; v.outline_smooth_samples = (uint8_t)outline_smooth_samples;
    v.outline_smooth_step = outline_smooth_step; v.midline_start_with_head = midline_start_with_head != 0; v.midline_invert = midline_invert != 0;
    v.midline_stiff_percentage = midline_stiff_percentage; v.midline_resolution = midline_resolution; v.outline_use_dft = false;
}

// Outline(points).resample(distance)  (Posture.cpp:271); returns the new number of points (out: up to cap)
int64_t ref_outline_resample(const float *pts, int64_t n, float distance, float *out, int64_t cap)
{
    Outline o(to_points(pts, n));
    o.resample(distance);
    const int64_t m = (int64_t)o.size();
    for (int64_t i = 0; i < m && i < cap; ++i) { out[2 * i] = o[(size_t)i].x; out[2 * i + 1] = o[(size_t)i].y; }
    return m;
}

// Outline(points).calculate_midline({})  (Posture.cpp:230).  pts: the (resampled) outline, overwritten with the outline as the call left it
// (n_after points: the reference works on the Outline in place).  Returns the number of midline segments, or -1 for std::unexpected.
int64_t ref_calculate_midline(float *pts, int64_t n, int64_t *n_after, float *segments, int64_t cap, int64_t *tail, int64_t *head)
{
    Outline o(to_points(pts, n));
    auto r = o.calculate_midline({});
    *n_after = (int64_t)o.size();
    for (int64_t i = 0; i < *n_after && i < n; ++i) { pts[2 * i] = o[(size_t)i].x; pts[2 * i + 1] = o[(size_t)i].y; }
    if (!r) return -1;
    auto &m = *r.value();
    *tail = m.tail_index(); *head = m.head_index();
    if ((int64_t)m.segments().size() > cap) return -4;
    store_segments(m.segments(), segments);
    return (int64_t)m.segments().size();
}

// Midline::post_process(movement, {})  (Individual.cpp:518); move_dir: MovementInformation::direction ((0, 0): none).  Returns 0 / 1 =
// inverted_because_previous, -5 when the reference throws
int ref_midline_post_process(float *seg, int64_t n, const float *move_dir, int64_t *tail, int64_t *head)
{
    Midline m;
    fill_midline(m, seg, n, *tail, *head);
    MovementInformation mv;
    if (move_dir) mv.direction = Vec2(move_dir[0], move_dir[1]);
    try { m.post_process(mv, DebugInfo{}); } catch (...) { return -5; }
    store_segments(m.segments(), seg);
    *tail = m.tail_index(); *head = m.head_index();
    return m.inverted_because_previous() ? 1 : 0;
}

// Midline::normalize(fix_length)  (Individual.cpp:519, 1372); info = {len, angle, offset.x, offset.y}.  Returns the number of segments, 0 for nullptr
int64_t ref_midline_normalize(const float *seg, int64_t n, int64_t tail, int64_t head, float fix_length, float *out, int64_t cap, float *info)
{
    Midline m;
    fill_midline(m, seg, n, tail, head);
    auto r = m.normalize(fix_length);
    if (!r) return 0;
    if ((int64_t)r->segments().size() > cap) return -4;
    store_segments(r->segments(), out);
    info[0] = r->len(); info[1] = r->angle(); info[2] = r->offset().x; info[3] = r->offset().y;
    return (int64_t)r->segments().size();
}


// The 2 x 3 matrix normalize_image hands to cv::warpAffine (tracker/tracking/FilterCache.cpp:47-63): translate(size / 2) . scale(image_scale) .
// translate(legacy ? (-len / 2, 0) : (0.4 len, 0.4 len)) . Midline::transform(type) -- the last factor is the reference's compiled
// Midline::transform (Outline.cpp:1237-1256), the composition uses the reference's compiled gui::Transform (commons/common/gui/Transform.cpp);
// toCV()'s element order (Transform.h:66-79).  The midline carries what a normalised one carries: angle, offset, front = (0, 0).
void ref_posture_matrix(float angle, float offx, float offy, float midline_length, float image_scale, int legacy, int out_w, int out_h, double *M)
{
    Midline m;
    m.segments().resize(1);
    m.angle() = angle; m.offset() = Vec2(offx, offy);
    const auto type = legacy ? default_config::individual_image_normalization_t::legacy : default_config::individual_image_normalization_t::posture;
    gui::Transform midline_transform = m.transform(type);
    const Size2 size((float)out_w, (float)out_h);
    gui::Transform tr;
    if (legacy) {
        tr.translate(size * 0.5);
        tr.scale(image_scale);
        tr.translate(Vec2(-midline_length * 0.5, 0));
    } else {
        tr.translate(size * 0.5);
        tr.scale(image_scale);
        tr.translate(Vec2(midline_length * 0.4));
    }
    tr.combine(midline_transform);
    const auto *g = tr.getMatrix();
    M[0] = g[0]; M[1] = g[4]; M[2] = g[12]; M[3] = g[1]; M[4] = g[5]; M[5] = g[13];
}

}
