// Test infrastructure (oracle/): a C ABI around the REFERENCE's own periodic:: functions, compiled from
// /root/reference/Application/src/commons/common/misc/CircularGraph.cpp where it lies (oracle/build_ref.py; stubs for its precompiled header in
// oracle/ref_stubs/).  Each entry calls the reference exactly as Outline::offset_to_middle does (tracker/tracking/Outline.cpp:493-528), so that
// tests/test_oracle_ref_circular_graph.py can pin oracle/trex_oracle.c's to_periodic_curvature / to_orientation_sum / to_eft / to_ieft /
// to_find_peaks on the real code.  Never linked into the product.
#include <misc/CircularGraph.h>

using namespace cmn;
using namespace cmn::periodic;

static points_t::element_type to_points(const float *p, int64_t n)
{
    points_t::element_type v((size_t)n);
    for (int64_t i = 0; i < n; ++i) v[(size_t)i] = Vec2(p[2 * i], p[2 * i + 1]);
    return v;
}

extern "C" {

// periodic::curvature(points, r, absolute)  (Outline.cpp:515)
void ref_curvature(const float *p, int64_t n, int r, int absolute, float *out)
{
    auto c = curvature(to_points(p, n), r, absolute != 0);
    for (size_t i = 0; i < c->size(); ++i) out[i] = (*c)[i];
}

// differentiate_and_test_clockwise(points)  (Outline.cpp:493): the orientation sum
float ref_orientation_sum(const float *p, int64_t n)
{
    auto && [sum, d] = differentiate_and_test_clockwise(to_points(p, n));
    return sum;
}

// periodic::eft(points, order)  (Outline.cpp:507); coeffs: order x {x, y, z, w}
void ref_eft(const float *p, int64_t n, int order, float *coeffs)
{
    auto c = eft(to_points(p, n), (size_t)order);
    for (size_t i = 0; i < c->size(); ++i) { coeffs[4 * i] = (*c)[i].x; coeffs[4 * i + 1] = (*c)[i].y; coeffs[4 * i + 2] = (*c)[i].z; coeffs[4 * i + 3] = (*c)[i].w; }
}

// periodic::ieft(coeffs, coeffs.size(), n_points, center, false).front()  (Outline.cpp:509)
void ref_ieft(const float *coeffs, int order, int64_t n_points, float offx, float offy, float *out)
{
    coeff_t::element_type c((size_t)order);
    for (int i = 0; i < order; ++i) c[(size_t)i] = Vec4{coeffs[4 * i], coeffs[4 * i + 1], coeffs[4 * i + 2], coeffs[4 * i + 3]};
    auto pts = std::move(ieft(c, c.size(), (size_t)n_points, Vec2(offx, offy), false).front());
    for (size_t i = 0; i < pts->size(); ++i) { out[2 * i] = (*pts)[i].x; out[2 * i + 1] = (*pts)[i].y; }
}

// diffs = periodic::differentiate(curv, 2); find_peaks(curv, 0, diffs, mode)  (Outline.cpp:516-528).  Returns the number of maxima;
// rec: cap x {x, y, width, integral, range.start, range.end, max_y_extrema, max_y, number of points}
int64_t ref_find_peaks(const float *v, int64_t n, int broad, float *rec, int64_t cap)
{
    auto curv = std::make_unique<scalars_t::element_type>(v, v + n);
    auto diffs = differentiate(*curv, 2);
    auto && [maxima, minima] = find_peaks(curv, 0, diffs, broad ? PeakMode::FIND_BROAD : PeakMode::FIND_POINTY);
    int64_t k = 0;
    for (auto &pk : *maxima) {
        if (k >= cap) break;
        float *r = rec + 9 * k++;
        r[0] = pk.position.x; r[1] = pk.position.y; r[2] = pk.width; r[3] = pk.integral; r[4] = pk.range.start; r[5] = pk.range.end;
        r[6] = pk.max_y_extrema; r[7] = pk.max_y; r[8] = (float)pk.points.size();
    }
    return (int64_t)maxima->size();
}

}
