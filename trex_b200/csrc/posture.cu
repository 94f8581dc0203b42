// "Next" row N4, second and third stage: from the resampled outline of every blob (outline.cu) to the raw midline, the
// post-processed / normalised midline and the posture-normalised crop.
//   Outline::smooth / smooth_outline         T/tracking/Outline.cpp:330-452
//   Outline::offset_to_middle                T/tracking/Outline.cpp:454-718   (peak_mode pointy and broad)
//   periodic::differentiate_and_test_clockwise, eft, ieft, curvature, find_peaks, fast::cos
//                                            C/misc/CircularGraph.cpp:12-606
//   Outline::calculate_midline               T/tracking/Outline.cpp:768-868
//   Midline::post_process / normalize / fix_length / calculate_angle / transform
//                                            T/tracking/Outline.cpp:870-1085,1113-1456
//   image::normalize_image (posture, legacy) T/tracking/FilterCache.cpp:21-115,133-154,266-276
//
// One WARP per blob, the blob's work arrays in shared memory (a slice of the CTA's pool; outlines that do not fit take a slice of
// a global arena and run the same code).  Everything that is independent per outline point -- smoothing, the first differences,
// the harmonics' cos / sin tables, the inverse transform, the curvature, the extremum test -- runs lane-parallel over the points.
// The reference's float ACCUMULATIONS (orientation sum, centre, cumulative arc length, the 4 x order Fourier sums) fix the order
// of the additions, so each of them is one sequential chain -- but the chains are independent: every chain gets its own lane and
// all of them advance together (7 chains in the first pass: the orientation sum plus the centre and arc-length sums for BOTH
// orientations, so the pass does not have to wait for the orientation test; 12 chains = 3 harmonics x {a, b, c, d} in the second).
// The maxima of the curvature need no state machine: find_peaks' `sign` always equals the previous sample's sign, so "i is an
// extremum" is a test on diff[i - 1], diff[i]; tail and head are warp arg-max reductions with the reference's tie rules.  The
// pairing walk is sequential over the midline segments; the warp evaluates the max_offset candidates of a step in parallel
// (redux.min on the distances' bit patterns + ballot for the first minimum).  Broad tails (find_peaks' ranges / integrals) and
// Midline::post_process / normalize are short scalar programs over a handful of peaks / <= N / 2 segments: lane 0 runs them on
// the arrays the warp left in shared memory.
// Compiled with -fmad=false: the arithmetic must round like the reference's scalar float code; double promotions are the
// reference's.  libm calls of the reference (atan2f, cos, sin, acos) are evaluated in double and rounded to float: the correctly
// rounded value (what glibc >= 2.41's atan2f returns; older libms are within one ulp of it), the same on the host and the device
// up to double-precision ties.
#include "common.h"

#include <math.h>

#include <algorithm>

namespace tb {

constexpr int ML_WARPS = 8;                      // blobs per CTA pass
constexpr int ML_POOL = 18400;                   // floats of shared memory per CTA (73 600 B: three CTAs per SM)
constexpr float ML_FLT_MAX = 3.402823466e+38f;

__device__ __forceinline__ float ml_fast_cos(float x)
{
    const float tp = (float)(1. / (2. * 3.14159265358979323846264338327950288));
    x *= tp;
    x -= 0.25f + floorf(x + 0.25f);
    x *= 16.f * (fabsf(x) - 0.5f);
    x += 0.225f * x * (fabsf(x) - 1.f);
    return x;
}
__device__ __forceinline__ float ml_fast_sin(float x) { return ml_fast_cos(x - (float)1.57079632679489661923132169163975144); }

__device__ __forceinline__ float ml_atan2(float y, float x) { return (float)atan2((double)y, (double)x); }
__device__ __forceinline__ void ml_vnormalize(float x, float y, float &ox, float &oy)      // Vector2D::normalize (C/misc/vec2.h:159-162)
{
    const float L = sqrtf(x * x + y * y);
    const float s = (float)(L != 0), d = (float)(L == 0) + L;
    ox = s * (x / d); oy = s * (y / d);
}

struct MlPeak { float x, y, integral, r0, r1, max_y; };        // Peak::width and max_y_extrema are never read by offset_to_middle

// ---- lane 0: periodic::find_peaks' ranges and integrals + the broad tail index (CircularGraph.cpp:185-402, Outline.cpp:621-650).
// curv[N]; ext_i / ext_m: the extrema in index order; mx: the maxima in index order (x, y set); order: scratch of n_max ints (the
// ranges found so far are those of mx[order[0 .. oi)], in the order the reference pushes them).
__device__ float ml_broad_tail(const float *curv, int N, const float *ext_i, const int *ext_m, int n_ext, MlPeak *mx, int n_max, int *order)
{
    float min_y = ML_FLT_MAX;
    for (int i = 0; i < N; ++i) if (curv[i] < min_y) min_y = curv[i];
    for (int i = 0; i < n_max; ++i) order[i] = i;
    for (int i = 1; i < n_max; ++i) {                      // descending (y, index): std::set<tuple<y, idx>, greater<>>
        const int k = order[i]; int j = i - 1;
        while (j >= 0 && (mx[order[j]].y < mx[k].y || (mx[order[j]].y == mx[k].y && order[j] < k))) { order[j + 1] = order[j]; --j; }
        order[j + 1] = k;
    }
    for (int oi = 0; oi < n_max; ++oi) {
        MlPeak *peak = &mx[order[oi]];
        int after = 0, prev = n_ext - 1;
        for (; after != n_ext; ++after) {
            if (ext_i[after] == peak->x) { ++after; if (after == n_ext) after = 0; break; }
            prev = after;
        }
        if (after == n_ext) after = 0;
        float minimum_left = ML_FLT_MAX, index_left = ext_i[prev], minimum_right = ML_FLT_MAX, index_right = ext_i[after];
        float left_border = 0, right_border = (float)N;
        for (int k = 0; k < oi; ++k) {
            const float rs = mx[order[k]].r0, re = mx[order[k]].r1;
            if (re > left_border && re < peak->x) left_border = re;
            if (rs < right_border && rs > peak->x) right_border = rs;
        }
        int cl = prev;
        float last_y = peak->y, offset = 0;
        while (ext_i[cl] + offset >= left_border) {
            const float x = ext_i[cl], y = curv[(size_t)x];
            if (ext_m[cl]) { if ((double)y > (double)last_y * 1.05) break; last_y = y; }
            if (y < minimum_left) { minimum_left = y; index_left = x; }
            if (cl == 0) { offset = -(float)N; cl = n_ext - 1; }
            if (cl == after) break;
            --cl;
            if (cl < 0) break;
        }
        offset = 0; cl = after; last_y = peak->y;
        while (ext_i[cl] + offset <= right_border) {
            const float x = ext_i[cl], y = curv[(size_t)x];
            if (ext_m[cl]) { if ((double)y > (double)last_y * 1.05) break; last_y = y; }
            if (y < minimum_right) { minimum_right = y; index_right = x; }
            if (++cl == n_ext) { offset = (float)N; cl = 0; }
            if (cl == prev) break;
        }
        while (index_left > peak->x) index_left -= (float)N;
        while (index_right < peak->x) index_right += (float)N;
        peak->r0 = index_left; peak->r1 = index_right;
    }
    for (int k = 0; k < n_max; ++k) {                      // points above half height, max_y and the integral (broad: factor 1)
        MlPeak *peak = &mx[k];
        float chk[3][2]; int nchk = 1;
        chk[0][0] = peak->r0; chk[0][1] = peak->r1;
        float x0 = peak->r0, x1 = peak->r1;
        if (x0 < 0) { x0 += (float)N; chk[0][0] = 0; chk[nchk][0] = x0; chk[nchk][1] = (float)(N - 1); ++nchk; }
        if (x1 >= (float)N) { x1 -= (float)N; chk[0][1] = (float)(N - 1); chk[nchk][0] = 0; chk[nchk][1] = x1; ++nchk; }
        float max_y = 0;
        for (int c = 0; c < nchk; ++c) {
            const float first = chk[c][0], last = chk[c][1];
            if (!(last - first >= 0)) continue;
            const size_t steps = (size_t)((last - first) / 1.f);
            for (size_t s = 0; s < steps; ++s) {
                const float i = first + (float)s * 1.f;
                const float y = curv[(size_t)i] - min_y;
                if ((double)(y / (peak->y - min_y)) >= 0.5 && y > max_y) max_y = y;
            }
        }
        peak->max_y = max_y;
        const double median = (double)max_y * 0.5;
        float integral = 0;
        for (int c = 0; c < nchk; ++c) {
            const float first = chk[c][0], last = chk[c][1];
            if (!(last - first >= 0)) continue;
            const size_t steps = (size_t)((last - first) / 1.f);
            for (size_t s = 0; s < steps; ++s) {
                const float i = first + (float)s * 1.f;
                const float y = curv[(size_t)i] - min_y;
                if ((double)(y / (peak->y - min_y)) >= 0.5) integral = (float)((double)integral + ((double)y - median) * (double)1.f);
            }
        }
        peak->integral = integral;
    }
    // Outline.cpp:621-650
    float max_int = -1;
    for (int k = 0; k < n_max; ++k) if (mx[k].integral > max_int) max_int = mx[k].integral;
    float m0 = 0, m1 = 0, max_y = 0; bool first = true;
    for (int k = 0; k < n_max; ++k) {
        if (!((double)fabsf(mx[k].integral - max_int) <= 1e-5)) continue;
        if (first) { m0 = mx[k].r0; m1 = mx[k].r1; first = false; }
        else { if (mx[k].r0 < m0) m0 = mx[k].r0; if (mx[k].r1 > m1) m1 = mx[k].r1; }
        if (mx[k].max_y > max_y) max_y = mx[k].max_y;
    }
    float start = m1, end = m0;
    const float period = (float)(size_t)N;
    for (int k = 0; k < n_max; ++k) {
        if (!((double)fabsf(mx[k].integral - max_int) <= 1e-5)) continue;
        float chk[3][2]; int nchk = 1;
        chk[0][0] = mx[k].r0; chk[0][1] = mx[k].r1;
        float x0 = mx[k].r0, x1 = mx[k].r1;
        if (x0 < 0) { x0 += (float)N; chk[0][0] = 0; chk[nchk][0] = x0; chk[nchk][1] = (float)(N - 1); ++nchk; }
        if (x1 >= (float)N) { x1 -= (float)N; chk[0][1] = (float)(N - 1); chk[nchk][0] = 0; chk[nchk][1] = x1; ++nchk; }
        for (int c = 0; c < nchk; ++c) {
            if (!(chk[c][1] - chk[c][0] >= 0)) continue;
            const size_t steps = (size_t)(chk[c][1] - chk[c][0]);
            for (size_t s2 = 0; s2 < steps; ++s2) {
                const float i = chk[c][0] + (float)s2;
                const float y = curv[(size_t)i] - min_y;
                if (!((double)(y / (mx[k].y - min_y)) >= 0.5)) continue;
                float cx; bool in_range;                   // is_in_periodic_range (Outline.cpp:442-452), Range::contains half open
                if (m0 < 0) { if (i - period >= m0) { cx = i - period; in_range = true; } else { cx = i; in_range = i <= m1; } }
                else if (m1 >= period) { if (i + period <= m1) { cx = i + period; in_range = true; } else { cx = i; in_range = i >= m0; } }
                else { cx = i; in_range = i >= m0 && i < m1; }
                if ((double)y >= (double)max_y * 0.9 && in_range) { if (start > cx) start = cx; if (end < cx) end = cx; }
            }
        }
    }
    float idx = (float)round((double)start + (double)(end - start) * 0.5);
    if (idx < 0) idx += period;
    if (idx >= period) idx -= period;
    return idx;
}

__device__ __forceinline__ void ml_reverse_segments(float4 *seg, int n)
{
    for (int i = 0, j = n - 1; i < j; ++i, --j) { const float4 t = seg[i]; seg[i] = seg[j]; seg[j] = t; }
}

// ---- lane 0: Midline::post_process (Outline.cpp:895-1062) in place.  Returns 0 / 1 (inverted because of the movement), -5 where
// the reference's segments().at(i + 1) throws.
__device__ int ml_post_process(float4 *seg, int n, const tb_posture_params &P, float mdx, float mdy, int &tail, int &head, float4 *copy)
{
    if (n <= 2) return 0;
    float dx = 0, dy = 0;
    {   // midline_direction (:870-887)
        const int samples = (int)fmaxf(1.f, (float)(size_t)n * P.midline_stiff_percentage);
        int counted = 0;
        for (int i = 0; i < samples && i + 1 < n; i++, counted++) { dx += seg[i + 1].x - seg[i].x; dy += seg[i + 1].y - seg[i].y; }
        if (counted > 0) { dx /= (float)counted; dy /= (float)counted; ml_vnormalize(dx, dy, dx, dy); }
    }
    bool needs_invert = !P.midline_invert; int inverted_prev = 0;
    if (!needs_invert) { dx = -dx; dy = -dy; }
    if (mdx != 0 || mdy != 0) {
        const float a = (-dx) * mdx + (-dy) * mdy, b = dx * mdx + dy * mdy;
        if (acos((double)a) < acos((double)b)) { needs_invert = !needs_invert; inverted_prev = 1; const int t = tail; tail = head; head = t; }
    }
    if (needs_invert) { if (!P.midline_start_with_head) ml_reverse_segments(seg, n); }
    else if (P.midline_start_with_head) ml_reverse_segments(seg, n);
    const float stiff = P.midline_stiff_percentage;
    if (stiff > 0) {
        const size_t center = (size_t)fminf((float)(size_t)n - 1, roundf((float)(size_t)n * stiff) + 1);
        const float cpx = seg[center].x, cpy = seg[center].y;
        float ax = 0, ay = 0; uint32_t count = 0;
        const size_t extra = (size_t)fmin((double)n, (double)center + fmax(0.0, (double)(size_t)n * 0.1));
        for (size_t i = center; i < extra; ++i) {
            if (i + 1 >= (size_t)n) return -5;
            float nx, ny;
            ml_vnormalize(seg[i].x - seg[i + 1].x, seg[i].y - seg[i + 1].y, nx, ny);
            ax += nx; ay += ny; ++count;
        }
        if (count > 0) { ax /= (float)count; ay /= (float)count; }
        for (size_t i = 0; i <= center; ++i) copy[i] = seg[i];
        for (size_t i = center; i > 0; --i) {
            const float p1x = seg[i].x, p1y = seg[i].y;
            const float lx = copy[i].x - copy[i - 1].x, ly = copy[i].y - copy[i - 1].y;
            const float L = sqrtf(lx * lx + ly * ly);
            float dcx, dcy, tx, ty;
            ml_vnormalize(seg[i - 1].x - cpx, seg[i - 1].y - cpy, dcx, dcy);
            ml_vnormalize((dcx + ax) * 0.5f, (dcy + ay) * 0.5f, tx, ty);
            seg[i - 1].x = p1x + L * tx; seg[i - 1].y = p1y + L * ty;
        }
    }
    ml_reverse_segments(seg, n);
    return inverted_prev;
}

__device__ float ml_calculate_angle(const float4 *seg, int n, float stiff)     // Midline::calculate_angle (:1113-1123)
{
    if (n < 2) return 0;
    const float center = fmaxf(0.f, (float)((size_t)n - 2) - (float)(size_t)n * stiff);
    const size_t start = (size_t)center;
    const float rest = center - (float)start;
    const float lx = seg[n - 1].x - (seg[start].x * (1 - rest) + seg[start + 1].x * rest);
    const float ly = seg[n - 1].y - (seg[start].y * (1 - rest) + seg[start + 1].y * rest);
    return ml_atan2(ly, lx);
}

__device__ __forceinline__ void ml_t_circle_line(float x0, float y0, float x1, float y1, float h, float k, float r, float &t0, float &t1)
{
    const float a = (x1 - x0) * (x1 - x0) + (y1 - y0) * (y1 - y0);
    const float b = 2 * (x1 - x0) * (x0 - h) + 2 * (y1 - y0) * (y0 - k);
    const float c = (x0 - h) * (x0 - h) + (y0 - k) * (y0 - k) - r * r;
    float disc = b * b - 4 * a * c;
    if (disc < 0) { t0 = -1; t1 = -1; return; }
    disc = sqrtf(disc);
    t0 = (-b + disc) / (2 * a); t1 = (-b - disc) / (2 * a);
}

// Midline::fix_length (:1125-1236): pts (n_pts, reversed order) -> out (<= resolution + 1 entries); returns the count
__device__ int ml_fix_length(float len, const float4 *pts, int n_pts, uint32_t resolution, float4 *out)
{
    const float step = len / (float)resolution;
    int n_out = 0;
    float4 seg = pts[0];
    out[n_out++] = seg;
    uint32_t j = 1;
    float last_t = -1;
    for (uint32_t i = 1; i < resolution; i++) {
        bool found = false;
        float mx = 0, my = 0;
        for (; j < resolution && j < (uint32_t)n_pts; j++) {
            const float4 v0 = pts[j - 1], v1 = pts[j];
            float t0, t1;
            ml_t_circle_line(v0.x, v0.y, v1.x, v1.y, seg.x, seg.y, step, t0, t1);
            if (t0 >= 0 && t0 <= 1 && t0 > last_t) {
                found = true; mx = v0.x + (v1.x - v0.x) * t0; my = v0.y + (v1.y - v0.y) * t0;
                seg.z = t0 * v1.z + (1 - t0) * v0.z; last_t = t0;
                break;
            } else if (t1 >= 0 && t1 <= 1 && t1 > last_t) {
                found = true; mx = v0.x + (v1.x - v0.x) * t1; my = v0.y + (v1.y - v0.y) * t1;
                seg.z = t1 * v1.z + (1 - t1) * v0.z; last_t = t1;
                break;
            }
            last_t = -1;
        }
        if (found) { seg.x = mx; seg.y = my; out[n_out++] = seg; }
        else if (j >= resolution) {
            if (n_pts >= 3) {
                const float lx = pts[n_pts - 1].x - pts[n_pts - 2].x, ly = pts[n_pts - 1].y - pts[n_pts - 2].y;
                const float l1x = pts[n_pts - 2].x - pts[n_pts - 3].x, l1y = pts[n_pts - 2].y - pts[n_pts - 3].y;
                const float angle0 = ml_atan2(ly, lx), angle1 = ml_atan2(l1y, l1x);
                const float change = angle0 - angle1;
                float angle = angle0;
                while ((uint32_t)n_out < resolution) {
                    seg.x += (float)cos((double)angle) * step; seg.y += (float)sin((double)angle) * step;
                    angle += change;
                    seg.z *= 0.5f;
                    out[n_out++] = seg;
                }
            }
            break;
        }
    }
    return n_out;
}

// ---- lane 0: Midline::normalize (:1268-1456).  seg: n post-processed segments; red / tmp: scratch of n + resolution + 4 and
// resolution + 2 entries; out: `resolution` normalised segments.  Returns resolution or 0 (nullptr).
__device__ int ml_normalize(const float4 *seg, int n, const tb_posture_params &P, float fix_length, float4 *red, float4 *tmp,
                            float4 *out, float info[4])
{
    if (n < 2) return 0;
    double len = 0.0;
    for (int i = 1; i < n; i++) { const float lx = seg[i].x - seg[i - 1].x, ly = seg[i].y - seg[i - 1].y; len += (double)sqrtf(lx * lx + ly * ly); }
    if (len == 0.0) return 0;
    const uint32_t resolution = (uint32_t)P.midline_resolution;
    const int max_segments = (int)(resolution - 1);
    const double step = len / (double)max_segments;
    if (step < 0) return 0;
    size_t index = 0;
    int nr = 0;
    red[nr++] = seg[0];
    double last_pt_distance = 0.0, distance;
    for (distance = 0.0; distance <= len && index < (size_t)n - 1;) {
        while (distance - last_pt_distance < step && index < (size_t)n - 1) {
            const float lx = seg[index + 1].x - seg[index].x, ly = seg[index + 1].y - seg[index].y;
            distance += (double)sqrtf(lx * lx + ly * ly);
            index++;
        }
        float off = (float)(distance - last_pt_distance);
        if ((double)off < step) break;
        while ((double)off >= step) {
            off = (float)((double)off - step);
            if (index > 0) {
                const float4 s0 = seg[index - 1], s1 = seg[index];
                const float lx = s1.x - s0.x, ly = s1.y - s0.y;
                const float local_d = sqrtf(lx * lx + ly * ly);
                float percent = off;
                if (local_d > 0) percent /= local_d;
                percent = 1.f - percent;
                float4 o;
                o.x = s0.x + lx * percent; o.y = s0.y + ly * percent;
                o.z = (float)((double)(s0.z * percent) + (double)s1.z * (1.0 - (double)percent));
                o.w = s0.w < s1.w ? s1.w : s0.w;
                red[nr++] = o;
                const float q = (float)(1.0 - (double)percent);
                const float qx = lx * q, qy = ly * q;
                last_pt_distance = distance - (double)sqrtf(qx * qx + qy * qy);
            } else {
                float4 o = seg[index]; o.w = 0;
                red[nr++] = o;
                last_pt_distance = distance;
            }
            if ((uint32_t)nr > (uint32_t)n + resolution) return 0;
        }
    }
    {
        const float lx = red[nr - 1].x - seg[n - 1].x, ly = red[nr - 1].y - seg[n - 1].y;
        if ((double)sqrtf(lx * lx + ly * ly) >= 0.01) red[nr++] = seg[n - 1];
    }
    if ((uint32_t)nr != resolution) return 0;
    {
        const float lx = red[1].x - red[0].x, ly = red[1].y - red[0].y;
        float percent = sqrtf(lx * lx + ly * ly);
        if (len > 0) percent = (float)((double)percent / len);
        red[0].z = (float)((double)(red[1].z * percent) + (double)red[0].z * (1.0 - (double)percent));
    }
    if (fix_length > 0) {
        ml_reverse_segments(red, nr);
        const int m = ml_fix_length(fix_length, red, nr, resolution, tmp);
        for (int i = 0; i < m; ++i) red[i] = tmp[i];
        nr = m;
        ml_reverse_segments(red, nr);
    }
    len = 0.0;
    for (int i = 1; i < nr; i++) { const float lx = red[i].x - red[i - 1].x, ly = red[i].y - red[i - 1].y; len += (double)sqrtf(lx * lx + ly * ly); }
    const float ang = ml_calculate_angle(red, nr, P.midline_stiff_percentage);
    const float angle = (float)((double)(-ang) + 3.14159265358979323846);
    const float offx = red[nr - 1].x, offy = red[nr - 1].y;
    const float deg = angle * (1.0f / 3.14159274f * 180.0f);
    const double rad = (double)deg * 3.141592654 / 180.0;
    const double c = cos(rad), s = sin(rad);
    const double tx = (double)(-offx), ty = (double)(-offy);
    const double m0 = c, m4 = -s, m1 = s, m5 = c;
    const double m12 = m0 * tx + m4 * ty + 0.0, m13 = m1 * tx + m5 * ty + 0.0;
    for (int i = nr - 1, k = 0; i >= 0; i--, k++) {
        const double x = red[i].x, y = red[i].y;
        out[k] = make_float4((float)(m0 * x + m4 * y + m12), (float)(m1 * x + m5 * y + m13), red[i].z, red[i].w);
    }
    const float fx = out[0].x, fy = out[0].y;
    if (fx != 0 || fy != 0) for (int k = 0; k < nr; ++k) { out[k].x -= fx; out[k].y -= fy; }
    info[0] = (float)len; info[1] = ang; info[2] = offx; info[3] = offy;
    return nr;
}

// work space of one outline, in floats: [p | t: 4 NP][extra: 4 (RES + 8)][a0..a3: 4 NP][xs: 3 NP][normalize's out + tmp: 8 (RES + 8)]; NP >= N + 2, a multiple of 4.
// The tables of three harmonics (6 N) live in t, a1 and xs while p, the tangents and phi are in use; later xs holds the midline segments.
__device__ __forceinline__ int ml_np(int N) { return (N + 5) & ~3; }      // >= N + 2, a multiple of 4: every array starts on 16 bytes
__device__ __forceinline__ int ml_slice_floats(int N, int res) { return 11 * ml_np(N) + 12 * (res + 8); }

// One warp, one outline: everything from the resampled points to the records (see the file header).  w_: the work space.
__device__ __forceinline__ void ml_process_outline(const tb_outline_rec o, const int N, const uint32_t q, float *w_, const int lane, const int RES,
                                                   const float *__restrict__ res, const tb_posture_params &P, const int do_norm,
                                                   const float *__restrict__ move_dir, const float *__restrict__ fix_len,
                                                   float *__restrict__ pts_out, float4 *__restrict__ segs, tb_midline_rec *__restrict__ mrecs,
                                                   tb_midline_norm *__restrict__ nrecs, float4 *__restrict__ norm_pts)
{
    tb_midline_rec mr; mr.seg_off = o.res_off; mr.n_seg = 0; mr.tail = -1; mr.head = -1;
    tb_midline_norm nr{};
    const int NP = ml_np(N), EX = 4 * (RES + 8);
    float *p = w_, *t = w_ + 2 * NP, *a0 = w_ + 4 * NP + EX, *a1 = a0 + NP, *a2 = a0 + 2 * NP, *a3 = a0 + 3 * NP, *xs = a0 + 4 * NP;
    float4 *nscr = (float4 *)(xs + 3 * NP);                    // 2 * (RES + 8) float4: normalize's out + tmp
    {
        const float2 *src = (const float2 *)res + o.res_off;
        for (int i = lane; i < N; i += 32) { const float2 v = src[i]; p[2 * i] = v.x; p[2 * i + 1] = v.y; }
    }
    __syncwarp();
    // ---- Outline::smooth (:380-389, smooth_outline :330-378): the normalised tap weights once per outline (a0 .. a3 are free: 4 NP >= 2 N + 1
    // taps), then one weighted sum per point, taps in the reference's order
    if (P.outline_smooth_samples > 0 && (float)N > (float)P.outline_smooth_samples) {
        const float range = (float)P.outline_smooth_samples;
        const int step = P.outline_smooth_step;
        const float step_row = range * (float)step;
        const int k0 = (int)(-step_row);
        int nw = 0;
        float wsum = 0;
        for (int i = k0; (float)i <= step_row; i += step) { wsum += (step_row - (float)abs(i)) / step_row; ++nw; }
        float *wt = a0;
        const int nwc = min(nw, 4 * NP);
        for (int k = lane; k < nwc; k += 32) wt[k] = ((step_row - (float)abs(k0 + k * step)) / step_row) / wsum;
        __syncwarp();
        for (int i = lane; i < N; i += 32) {
            float px = 0, py = 0;
            int k = 0;
            for (int j = (int)((float)i - step_row); (float)j <= (float)i + step_row; j += step, ++k) {
                int idx = j;
                if (idx < 0) { idx += N; while (idx < 0) idx += N; }
                else if (idx >= N) { idx -= N; while (idx >= N) idx -= N; }
                const float wgt = k < nwc ? wt[k] : ((step_row - (float)abs(k0 + k * step)) / step_row) / wsum;
                px += p[2 * idx] * wgt; py += p[2 * idx + 1] * wgt;
            }
            t[2 * i] = px; t[2 * i + 1] = py;
        }
        __syncwarp();
        float *sw = p; p = t; t = sw;
    }
    // ---- offset_to_middle (:454-718).  Pass 1: orientation terms and first differences for the sequential chains
    const int nd = N - 1;
    const bool approx = P.outline_approximate > 0;
    for (int i = lane; i < N; i += 32) {
        const int j = i + 1 < N ? i + 1 : 0;
        const float x0 = p[2 * i], y0 = p[2 * i + 1], x1 = p[2 * j], y1 = p[2 * j + 1];
        // _differentiate<true> (CircularGraph.cpp:409-463): the wrap-around term enters with the operands swapped
        a0[i] = i + 1 < N ? x0 * y1 - x1 * y0 : x1 * y0 - x0 * y1;
        if (approx && i < nd) { const float dx = x1 - x0, dy = y1 - y0; a1[i] = (float)((double)sqrtf(dx * dx + dy * dy) + 1e-10); }
    }
    __syncwarp();
    // the sequential float sums, one lane each, as TWO code paths (divergent paths of a warp issue one after the other): lanes 0 - 4 sum an
    // array front to back or back to front (orientation terms; x and y of the centre in both orders), lanes 5 - 6 the arc-length prefix sums
    float acc = 0;
    if (lane < 5 && (lane == 0 || approx)) {
        const float *src = lane == 0 ? a0 : p + ((lane - 1) & 1);
        const int stride = lane == 0 ? 1 : 2;
        const bool back = lane >= 3;
        const float *q0 = back ? src + (N - 1) * stride : src;
        const int inc = back ? -stride : stride;
        for (int i = 0; i < N; ++i, q0 += inc) acc += *q0;
    } else if (approx && lane < 7) {
        const bool back = lane == 6;
        float *dst = back ? a3 : a2;
        const float *q0 = back ? a1 + (nd - 1) : a1;
        const int inc = back ? -1 : 1;
        dst[0] = 0;
        for (int i = 0; i < nd; ++i, q0 += inc) { acc += *q0; dst[i + 1] = acc; }
    }
    __syncwarp();
    const bool reversed = __shfl_sync(0xffffffffu, acc, 0) < 0;
    if (reversed) {
        for (int i = lane; i < N / 2; i += 32) {
            const int j = N - 1 - i;
            const float tx = p[2 * i], ty = p[2 * i + 1];
            p[2 * i] = p[2 * j]; p[2 * i + 1] = p[2 * j + 1]; p[2 * j] = tx; p[2 * j + 1] = ty;
        }
        __syncwarp();
    }
    if (approx) {
        const float ccx = __shfl_sync(0xffffffffu, acc, reversed ? 3 : 1) / (float)(size_t)N;
        const float ccy = __shfl_sync(0xffffffffu, acc, reversed ? 4 : 2) / (float)(size_t)N;
        float *cum = reversed ? a3 : a2, *cyv = reversed ? a2 : a3, *cxv = a0, *dt = a1;
        const float T = cum[nd];
        __syncwarp();
        // eft (:484-561): dt, phi, the unit tangents
        for (int i = lane; i < N; i += 32) {
            const float cumi = cum[i];
            float dti = 0, x = 0, y = 0;
            if (i < nd) {
                x = p[2 * (i + 1)] - p[2 * i]; y = p[2 * (i + 1) + 1] - p[2 * i + 1];
                dti = (float)((double)sqrtf(x * x + y * y) + 1e-10);
            }
            cum[i] = i == 0 ? 0.f : (float)(2 * 3.14159265358979323846 * (double)cumi);        // phi, in place
            if (i < nd) { dt[i] = dti; cxv[i] = x / dti; cyv[i] = y / dti; }
        }
        __syncwarp();
        const float *phi = cum;
        const float norm_base = (float)((double)T / (2 * (3.14159265358979323846 * 3.14159265358979323846)));
        const int order_n = min(P.outline_approximate, 8);
        float coef[8][4];
#pragma unroll
        for (int g = 0; g < 8; ++g) { coef[g][0] = coef[g][1] = coef[g][2] = coef[g][3] = 0.f; }
#pragma unroll
        for (int g0 = 0; g0 < 8; g0 += 3) {
            if (g0 >= order_n) break;
            const int ng = min(3, order_n - g0);
            // cos / sin tables of up to three harmonics in the arrays that are free now: t (2 NP), dt, and the three extra arrays
            float *const tab[6] = {t, t + NP, dt, xs, xs + NP, xs + 2 * NP};
            for (int e = lane; e < ng * N; e += 32) {
                const int h = e / N, i = e - h * N;
                const float phi_n = phi[i] * (float)(g0 + h + 1) / T;
                float *tc = h == 0 ? tab[0] : (h == 1 ? tab[2] : tab[4]), *ts = h == 0 ? tab[1] : (h == 1 ? tab[3] : tab[5]);
                tc[i] = ml_fast_cos(phi_n); ts[i] = ml_fast_sin(phi_n);
            }
            __syncwarp();
            float sum = 0;
            if (lane < 4 * ng) {                                // chain (h, k): k = 0 cnx, 1 cny, 2 snx, 3 sny
                const int h = lane >> 2, k = lane & 3;
                const int ti = 2 * h + (k >> 1);
                const float *v = (k & 1) ? cyv : cxv;
                const float *tr = ti == 0 ? tab[0] : ti == 1 ? tab[1] : ti == 2 ? tab[2] : ti == 3 ? tab[3] : ti == 4 ? tab[4] : tab[5];
                float prev = tr[0];
                for (int i = 0; i < nd; ++i) { const float nx = tr[i + 1]; sum += v[i] * (nx - prev); prev = nx; }
                const int n = g0 + h + 1;
                sum *= norm_base / (float)((size_t)n * (size_t)n);
            }
            __syncwarp();
#pragma unroll
            for (int h = 0; h < 3; ++h) {
                if (g0 + h < 8) {
                    coef[(g0 + h) & 7][0] = __shfl_sync(0xffffffffu, sum, 4 * h + 0);       // a = cnx
                    coef[(g0 + h) & 7][1] = __shfl_sync(0xffffffffu, sum, 4 * h + 2);       // b = snx
                    coef[(g0 + h) & 7][2] = __shfl_sync(0xffffffffu, sum, 4 * h + 1);       // c = cny
                    coef[(g0 + h) & 7][3] = __shfl_sync(0xffffffffu, sum, 4 * h + 3);       // d = sny
                }
            }
        }
        // ieft (:563-606)
        for (int j = lane; j < N; j += 32) {
            float x = ccx, y = ccy;
            const float tt = (float)((double)j / (double)(N - 1) * 3.14159265358979323846 * 2.0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (i < order_n) {
                    const float ct = ml_fast_cos(tt * (float)(i + 1)), st = ml_fast_sin(tt * (float)(i + 1));
                    x += 1.f * (coef[i][0] * ct + coef[i][1] * st);
                    y += 1.f * (coef[i][2] * ct + coef[i][3] * st);
                }
            }
            p[2 * j] = x; p[2 * j + 1] = y;
        }
        __syncwarp();
    }
    // curvature (CircularGraph.cpp:49-113) -> a0, its first difference -> a1
    float *curv = a0, *diff = a1;
    {
        float rf = P.outline_curvature_range_ratio * (float)(size_t)N;
        if (rf < 1.f) rf = 1.f;
        const int r = (int)rf;
        for (int i = lane; i < N; i += 32) {
            const int i1 = ((i - r) % N + N) % N, i3 = (i + r) % N;
            const float x1 = p[2 * i1], y1 = p[2 * i1 + 1], x2 = p[2 * i], y2 = p[2 * i + 1], x3 = p[2 * i3], y3 = p[2 * i3 + 1];
            const bool e12 = x1 == x2 && y1 == y2, e13 = x1 == x3 && y1 == y3, e23 = x2 == x3 && y2 == y3;
            float v = 0.f;
            if (!e12 && !e13 && !e23) {
                const float cross = (x2 - x1) * (y3 - y2) - (y2 - y1) * (x3 - x2);
                const float d12 = (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1), d23 = (x3 - x2) * (x3 - x2) + (y3 - y2) * (y3 - y2),
                            d13 = (x3 - x1) * (x3 - x1) + (y3 - y1) * (y3 - y1);
                v = 2.f * (approx ? fabsf(cross) : cross) / sqrtf(d12 * d23 * d13);
            }
            curv[i] = v;
        }
        __syncwarp();
        for (int i = lane; i < N; i += 32) diff[i] = curv[i + 1 < N ? i + 1 : 0] - curv[i];
        __syncwarp();
    }
    // find_peaks (:115-183): sample i is an extremum iff the sign of diff changes between i - 1 and i (and diff[i] != 0); it is a
    // maximum iff diff[i - 1] >= 0.  Pointy tail (:617-619): the highest maximum (> -1), the first among equals.
    float idxf = 0;
    if (P.peak_mode == 0) {
        float by = -1.f; int bi = 0x7fffffff;
        for (int i = lane; i < N; i += 32) {
            const bool cprev = diff[i == 0 ? N - 1 : i - 1] < 0, c = diff[i] < 0;
            if (c != cprev && diff[i] != 0 && !cprev) {
                const float y = curv[i];
                if (y > by) { by = y; bi = i; }
            }
        }
#pragma unroll
        for (int d = 16; d; d >>= 1) {
            const float oy = __shfl_xor_sync(0xffffffffu, by, d); const int oi = __shfl_xor_sync(0xffffffffu, bi, d);
            if (oy > by || (oy == by && oi < bi)) { by = oy; bi = oi; }
        }
        idxf = bi == 0x7fffffff ? 0.f : (float)bi;
    } else {
        // broad tails: compact the extrema / maxima in index order, then lane 0 runs the range bookkeeping.  Free arrays: the extremum
        // indices -> a2, the is-maximum flags -> a3, order -> t, the peaks (6 floats each; maxima alternate with minima, so there are at
        // most N / 2) -> the three extra arrays
        float *ext_i = a2; int *ext_m = (int *)a3; int *order = (int *)t; MlPeak *mx = (MlPeak *)xs;
        int n_ext = 0, n_max = 0;
        for (int i0 = 0; i0 < N; i0 += 32) {
            const int i = i0 + lane;
            bool is_ext = false, is_max = false;
            if (i < N) {
                const bool cprev = diff[i == 0 ? N - 1 : i - 1] < 0, c = diff[i] < 0;
                is_ext = c != cprev && diff[i] != 0; is_max = is_ext && !cprev;
            }
            const unsigned be = __ballot_sync(0xffffffffu, is_ext), bm = __ballot_sync(0xffffffffu, is_max);
            const unsigned below = (1u << lane) - 1u;
            if (is_ext) { const int k = n_ext + __popc(be & below); ext_i[k] = (float)i; ext_m[k] = is_max ? 1 : 0; }
            if (is_max) {
                const int k = n_max + __popc(bm & below);
                MlPeak pk; pk.x = (float)i; pk.y = curv[i]; pk.integral = 0; pk.r0 = -1; pk.r1 = -1; pk.max_y = 0; mx[k] = pk;
            }
            n_ext += __popc(be); n_max += __popc(bm);
        }
        __syncwarp();
        if (lane == 0) idxf = n_max ? ml_broad_tail(curv, N, ext_i, ext_m, n_ext, mx, n_max, order) : 0.f;
        idxf = __shfl_sync(0xffffffffu, idxf, 0);
        __syncwarp();
    }
    // head (:652-682): the maximum farthest (periodically) from the tail, the first among equals; d must exceed 0
    int tail = (int)idxf, head = -1;
    {
        float bd = 0.f; int bi = 0x7fffffff;
        const float sz = (float)(size_t)N;
        for (int i = lane; i < N; i += 32) {
            const bool cprev = diff[i == 0 ? N - 1 : i - 1] < 0, c = diff[i] < 0;
            if (c != cprev && diff[i] != 0 && !cprev) {
                const float px = (float)i;
                float d;
                if (px >= idxf) { const float a = fabsf(px - idxf), b = fabsf(px - idxf - sz); d = a < b ? a : b; }
                else { const float a = fabsf(idxf - px), b = fabsf(idxf - px - sz); d = a < b ? a : b; }
                if (d > bd) { bd = d; bi = i; }
            }
        }
#pragma unroll
        for (int d = 16; d; d >>= 1) {
            const float od = __shfl_xor_sync(0xffffffffu, bd, d); const int oi = __shfl_xor_sync(0xffffffffu, bi, d);
            if (od > bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
        }
        if (bi != 0x7fffffff) head = bi;
    }
    int rot;
    if (P.midline_start_with_head && head != -1) {
        if (tail != -1) { tail -= head; if (tail < 0) tail += N; }
        rot = head; head = 0;
    } else {
        if (head != -1) { head -= tail; if (head < 0) head += N; }
        rot = tail; tail = 0;
    }
    __syncwarp();
    if (rot > 0 && rot < N) {                                  // std::rotate(begin, begin + rot, end)
        for (int i = lane; i < N; i += 32) { const int s = i + rot < N ? i + rot : i + rot - N; t[2 * i] = p[2 * s]; t[2 * i + 1] = p[2 * s + 1]; }
        __syncwarp();
        float *sw = p; p = t; t = sw;
    }
    if (P.midline_invert) { const int tt = tail; tail = head; head = tt; }
    mr.tail = tail; mr.head = head;
    {
        float2 *dst = (float2 *)pts_out + o.res_off;
        for (int i = lane; i < N; i += 32) dst[i] = make_float2(p[2 * i], p[2 * i + 1]);
    }
    // ---- calculate_midline: the pairing walk (:786-868); segments -> shared (the extra arrays, as float4) and global
    float4 *sseg = (float4 *)xs;                               // 3 NP floats: room for 0.75 NP segments, the walk makes < N / 2
    uint32_t ns = 0;
    if (N > 1) {
        const int L = N;
        int idx_r = 1, idx_l = -1;
        float mo = P.midline_walk_offset * (float)L;
        if (mo < 3.f) mo = 3.f;
        const int max_offset = (int)mo;
        float4 *so = segs + o.res_off;
        while (idx_r < L + idx_l) {
            float prx = 0, pry = 0, plx = p[2 * (L + idx_l)], ply = p[2 * (L + idx_l) + 1];
            float min_d = ML_FLT_MAX; int min_idx = -1;
            for (int b = 0; b < max_offset && idx_r + b < L; b += 32) {
                const int i = b + lane;
                unsigned bits = 0xffffffffu;
                if (i < max_offset && idx_r + i < L) {
                    const float dx = p[2 * (idx_r + i)] - plx, dy = p[2 * (idx_r + i) + 1] - ply;
                    bits = __float_as_uint(sqrtf(dx * dx + dy * dy));
                }
                const unsigned m = __reduce_min_sync(0xffffffffu, bits);
                if (m != 0xffffffffu && __uint_as_float(m) < min_d) {
                    min_d = __uint_as_float(m);
                    min_idx = idx_r + b + (__ffs(__ballot_sync(0xffffffffu, bits == m)) - 1);
                }
            }
            if (min_idx != -1) { prx = p[2 * min_idx]; pry = p[2 * min_idx + 1]; idx_r = min_idx; }
            min_d = ML_FLT_MAX; min_idx = 1;
            for (int b = 0; b < max_offset && idx_l - b > -L; b += 32) {
                const int i = b + lane;
                unsigned bits = 0xffffffffu;
                if (i < max_offset && idx_l - i > -L) {
                    const float dx = prx - p[2 * (L + idx_l - i)], dy = pry - p[2 * (L + idx_l - i) + 1];
                    bits = __float_as_uint(sqrtf(dx * dx + dy * dy));
                }
                const unsigned m = __reduce_min_sync(0xffffffffu, bits);
                if (m != 0xffffffffu && __uint_as_float(m) < min_d) {
                    min_d = __uint_as_float(m);
                    min_idx = idx_l - b - (__ffs(__ballot_sync(0xffffffffu, bits == m)) - 1);
                }
            }
            if (min_idx != 1) { plx = p[2 * (L + min_idx)]; ply = p[2 * (L + min_idx) + 1]; idx_l = min_idx; }
            const float lx = prx - plx, ly = pry - ply;
            const float mx_ = plx + lx * 0.5f, my_ = ply + ly * 0.5f;
            if (ns < o.n_res && 4 * ns + 4 <= 3u * (uint32_t)NP && lane == 0) {
                const float4 sg = make_float4(mx_, my_, sqrtf((prx - plx) * (prx - plx) + (pry - ply) * (pry - ply)),
                                              sqrtf((plx - mx_) * (plx - mx_) + (ply - my_) * (ply - my_)));
                so[ns] = sg;
                sseg[ns] = sg;
            }
            ++ns;
            idx_r++; idx_l--;
        }
        mr.n_seg = ns > 2 ? min(ns, o.n_res) : 0;             // "Too few midline segments calculated." (:863-866)
    }
    __syncwarp();
    // ---- Midline::post_process + normalize (lane 0; Individual::calculate_midline_for, T/tracking/Individual.cpp:1348-1383)
    if (do_norm) {
        nr.n_points = 0; nr.flags = 0;
        if (lane == 0 && mr.n_seg > 2) {
            const int n = (int)mr.n_seg;
            float4 *red = (float4 *)w_;                        // p | t | extra are free now: NP + RES + 8 entries >= n + RES + 4
            int tl = mr.tail, hd = mr.head;
            const float mdx = move_dir ? move_dir[2 * q] : 0.f, mdy = move_dir ? move_dir[2 * q + 1] : 0.f;
            const int rc = ml_post_process(sseg, n, P, mdx, mdy, tl, hd, red);
            nr.tail = tl; nr.head = hd;
            if (rc < 0) nr.flags = 2u;                         // the reference throws std::out_of_range here
            else {
                float info[4];
                float4 *out = nscr, *tmp = nscr + (RES + 8);
                const int m = ml_normalize(sseg, n, P, fix_len ? fix_len[q] : -1.f, red, tmp, out, info);
                if (m == RES) {
                    nr.len = info[0]; nr.angle = info[1]; nr.offx = info[2]; nr.offy = info[3];
                    nr.n_points = (uint32_t)m; nr.flags = (uint32_t)rc;
                    for (int k = 0; k < m; ++k) norm_pts[(size_t)q * RES + k] = out[k];
                } else nr.flags = (uint32_t)rc | 4u;           // normalize() returned nullptr
            }
        }
        if (lane == 0) nrecs[q] = nr;
    }
    if (lane == 0) mrecs[q] = mr;
}

__global__ void __launch_bounds__(ML_WARPS * 32, 3)
midline_warp_kernel(const tb_outline_rec *__restrict__ orecs, const uint32_t *__restrict__ nb_dev, uint32_t nb_max,
                    const float *__restrict__ res, uint32_t cap_pts, tb_posture_params P, int do_norm,
                    const float *__restrict__ move_dir, const float *__restrict__ fix_len,
                    float *__restrict__ pts_out, float4 *__restrict__ segs, tb_midline_rec *__restrict__ mrecs,
                    tb_midline_norm *__restrict__ nrecs, float4 *__restrict__ norm_pts,
                    float *__restrict__ arena, unsigned long long arena_floats, unsigned long long *__restrict__ arena_used,
                    uint32_t *__restrict__ status, const uint8_t *__restrict__ skip)
{
    extern __shared__ __align__(16) float s_pool[];
    __shared__ uint32_t s_need[ML_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t nb = min(nb_dev ? *nb_dev : nb_max, nb_max);
    const int RES = max(P.midline_resolution, 2);
    for (uint32_t base = blockIdx.x * ML_WARPS; base < nb; base += gridDim.x * ML_WARPS) {
        const uint32_t q = base + warp;
        tb_outline_rec o{};
        int N = 0;
        const bool live = q < nb && !(skip && skip[q]);
        if (live) { o = orecs[q]; N = (int)o.n_res; }
        if (N <= 0 || (unsigned long long)o.res_off + o.n_res > cap_pts) N = 0;
        const uint32_t need = N > 0 ? (uint32_t)((ml_slice_floats(N, RES) + 3) & ~3) : 0u;
        bool todo = need != 0;
        if (live && !todo && lane == 0) {                          // no outline: empty records
            tb_midline_rec mr; mr.seg_off = o.res_off; mr.n_seg = 0; mr.tail = -1; mr.head = -1;
            mrecs[q] = mr;
            if (do_norm) nrecs[q] = tb_midline_norm{};
        }
        // The 8 outlines of a pass share the pool: those that fit (in warp order) run now, the others in a further round of the same
        // pass -- shared memory is an order of magnitude faster for this kernel than the global arena, which only takes outlines that
        // exceed the whole pool.
        for (;;) {
            __syncthreads();                                       // the previous round is done with the pool
            if (lane == 0) s_need[warp] = todo ? need : 0u;
            __syncthreads();
            uint32_t off = 0, any = 0;
            for (int w = 0; w < ML_WARPS; ++w) {
                const uint32_t nw = s_need[w];
                any |= nw;
                if (w < warp && nw <= (uint32_t)ML_POOL) off += nw;
            }
            if (!any) break;
            float *w_ = nullptr;
            bool run = false;
            if (todo) {
                if (need <= (uint32_t)ML_POOL) { if (off + need <= (uint32_t)ML_POOL) { w_ = s_pool + off; run = true; } }
                else {
                    unsigned long long a = 0;
                    if (lane == 0) a = atomicAdd(arena_used, (unsigned long long)need);
                    a = __shfl_sync(0xffffffffu, a, 0);
                    run = true;
                    if (a + need <= arena_floats) w_ = arena + a;
                    else if (lane == 0) {                          // the global arena is too small for this batch's long outlines
                        atomicOr(status, 1u);
                        tb_midline_rec mr; mr.seg_off = o.res_off; mr.n_seg = 0; mr.tail = -1; mr.head = -1;
                        mrecs[q] = mr;
                        if (do_norm) nrecs[q] = tb_midline_norm{};
                    }
                }
            }
            if (run) {
                if (w_) ml_process_outline(o, N, q, w_, lane, RES, res, P, do_norm, move_dir, fix_len, pts_out, segs, mrecs, nrecs, norm_pts);
                todo = false;
            }
        }
    }
}

// The affine map of the `posture` / `legacy` crop per crop (FilterCache.cpp:47-62 + Midline::transform, Outline.cpp:1238-1256),
// inverted like cv::warpAffine does, into coef[6] for crop_warp_kernel.  A crop whose blob has no normalised midline, or whose
// midline length is negative, gets the all-zero map of a failed image (crop_warp_kernel then renders zeros: D = 0 -> coordinates 0).
__global__ void posture_coef_kernel(const uint32_t *__restrict__ totals, const uint32_t *__restrict__ crop_blob,
                                    const tb_midline_norm *__restrict__ nrecs, const float *__restrict__ median_len, float median_len_all,
                                    float image_scale, int legacy, int out_w, int out_h, double *__restrict__ coef, uint8_t *__restrict__ valid)
{
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= totals[3]) return;
    const uint32_t b = crop_blob[q];
    const tb_midline_norm nr = nrecs[b];
    const float ml = median_len ? median_len[b] : median_len_all;
    double M[6] = {0, 0, 0, 0, 0, 0};
    const bool ok = nr.n_points > 0 && !(ml < 0);
    if (ok) {
        double mt[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, tr[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        auto combine = [](double *a, const double *bb) {
            double r[9];
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) r[3 * i + j] = a[3 * i] * bb[j] + a[3 * i + 1] * bb[3 + j] + a[3 * i + 2] * bb[6 + j];
            for (int i = 0; i < 9; ++i) a[i] = r[i];
        };
        const float angle = legacy ? (float)((double)(-nr.angle) + 3.14159265358979323846) : (float)((double)(-nr.angle) + 3.14159265358979323846 * (double)0.25f);
        const double t0[9] = {1, 0, (double)(-0.f), 0, 1, (double)(-0.f), 0, 0, 1};
        combine(mt, t0);
        const float deg = angle * (1.0f / 3.14159274f * 180.0f);
        const double rad = (double)deg * 3.141592654 / 180.0, c = cos(rad), s = sin(rad);
        const double rot[9] = {c, -s, 0, s, c, 0, 0, 0, 1};
        combine(mt, rot);
        const double t1[9] = {1, 0, (double)(-nr.offx), 0, 1, (double)(-nr.offy), 0, 0, 1};
        combine(mt, t1);
        const double c0[9] = {1, 0, (double)((float)out_w * 0.5f), 0, 1, (double)((float)out_h * 0.5f), 0, 0, 1};
        combine(tr, c0);
        const double sc[9] = {(double)image_scale, 0, 0, 0, (double)image_scale, 0, 0, 0, 1};
        combine(tr, sc);
        if (legacy) { const double t2[9] = {1, 0, (double)(-ml * 0.5f), 0, 1, 0.0, 0, 0, 1}; combine(tr, t2); }
        else { const float v = (float)((double)ml * 0.4); const double t2[9] = {1, 0, (double)v, 0, 1, (double)v, 0, 0, 1}; combine(tr, t2); }
        combine(tr, mt);
        M[0] = tr[0]; M[1] = tr[1]; M[2] = tr[2]; M[3] = tr[3]; M[4] = tr[4]; M[5] = tr[5];
        double D = M[0] * M[4] - M[1] * M[3];
        D = D != 0 ? 1. / D : 0;
        const double A11 = M[4] * D, A22 = M[0] * D;
        M[0] = A11; M[1] *= -D; M[3] *= -D; M[4] = A22;
        const double b1 = -M[0] * M[2] - M[1] * M[5], b2 = -M[3] * M[2] - M[4] * M[5];
        M[2] = b1; M[5] = b2;
    } else { M[2] = -1e6; M[5] = -1e6; }                           // every tap falls outside the blob image: a zero crop
    for (int k = 0; k < 6; ++k) coef[(size_t)q * 6 + k] = M[k];
    if (valid) valid[q] = ok ? 1 : 0;
}

// ---- posture::calculate_posture's threshold loop (T/tracking/Posture.cpp:305-400), one round.
// state per parent: 0 active, 1 done (midline found), 2 finished without a midline.
// (a) every sub-blob finds its parent -- the blob of the same frame that holds the first pixel of its first line -- and competes for
//     "biggest sub-blob" (pixel::threshold_get_biggest_blob, C/processing/PixelTree.cpp:297-340: most pixels, the first among equals)
__global__ void posture_parent_kernel(PostureRound R, uint32_t max_subs)
{
    const uint32_t ns = min(*R.n_subs, max_subs);
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < ns; s += gridDim.x * blockDim.x) {
        const tb_blob_rec r = R.sub_recs[s];
        const tb_line first = R.sub_lines[r.line_off];
        const tb_frame_info fi = R.parent_infos[r.frame];
        for (uint32_t p = fi.blob_begin; p < fi.blob_begin + fi.n_blobs; ++p) {
            const tb_blob_rec pr = R.parent_recs[p];
            if (first.y < pr.y0 || first.y > pr.y1 || first.x0 < pr.x0 || first.x0 > pr.x1) continue;
            const tb_line *L = R.parent_lines + pr.line_off;
            uint32_t lo = 0, hi = pr.n_lines;
            while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (L[mid].y < first.y) lo = mid + 1; else hi = mid; }
            bool inside = false;
            for (uint32_t j = lo; j < pr.n_lines && L[j].y == first.y && L[j].x0 <= first.x0; ++j) inside |= first.x0 <= L[j].x1;
            if (inside) {
                if (R.state[p] == 0) atomicMax(R.best + p, ((unsigned long long)r.n_pixels << 32) | (unsigned long long)(0xFFFFFFFFu - s));
                break;
            }
        }
    }
}
// (b) per parent: the chosen sub-blob of this round
__global__ void posture_pick_kernel(PostureRound R, uint32_t max_parents)
{
    const uint32_t np = min(*R.n_parents, max_parents);
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < np; p += gridDim.x * blockDim.x) {
        if (R.state[p]) continue;
        const unsigned long long b = R.best[p];
        R.index[p] = b ? 0xFFFFFFFFu - (uint32_t)(b & 0xFFFFFFFFull) : 0xFFFFFFFFu;
        R.sub_npx[p] = (uint32_t)(b >> 32);
        R.best[p] = 0;
    }
}
// (c) after the round's outlines and midlines: found -> done; else remember the first resampled outline (:358-366) and go on while the
//     sub-blob kept at least max(1, pixels / 10) pixels (:373-378; the caller bounds the threshold by start + 100); parents that stop
//     without a midline get their first outline back (:381-397)
__global__ void posture_round_end_kernel(PostureRound R, uint32_t max_parents, tb_outline_rec *orecs, tb_midline_rec *mrecs, tb_midline_norm *nrecs,
                                         int do_norm, int last_round)
{
    const uint32_t np = min(*R.n_parents, max_parents);
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < np; p += gridDim.x * blockDim.x) {
        if (R.state[p]) continue;
        if (mrecs[p].n_seg > 2) { R.state[p] = 1; continue; }
        if (R.first_outline[p].n_res == 0 && orecs[p].n_res != 0) R.first_outline[p] = orecs[p];
        const uint32_t npx = R.parent_recs[p].n_pixels, minimum = max(1u, npx / 10u);
        if (R.sub_npx[p] < minimum || last_round) {
            R.state[p] = 2;
            orecs[p] = R.first_outline[p];
            tb_midline_rec mr; mr.seg_off = orecs[p].res_off; mr.n_seg = 0; mr.tail = -1; mr.head = -1;
            mrecs[p] = mr;
            if (do_norm) nrecs[p] = tb_midline_norm{};
        } else atomicAdd(R.remaining, 1u);
    }
}

int launch_posture_parents(const PostureRound &R, uint32_t max_parents, uint32_t max_subs, int sms, cudaStream_t s)
{
    if (max_subs) posture_parent_kernel<<<std::max(1, sms) * 2, 256, 0, s>>>(R, max_subs);
    if (max_parents) posture_pick_kernel<<<std::max(1, sms) * 2, 256, 0, s>>>(R, max_parents);
    TB_CUDA(cudaGetLastError());
    return TB_OK;
}

int launch_posture_round_end(const PostureRound &R, uint32_t max_parents, tb_outline_rec *orecs, tb_midline_rec *mrecs, tb_midline_norm *nrecs,
                             int do_norm, int last_round, int sms, cudaStream_t s)
{
    if (max_parents) posture_round_end_kernel<<<std::max(1, sms) * 2, 256, 0, s>>>(R, max_parents, orecs, mrecs, nrecs, do_norm, last_round);
    TB_CUDA(cudaGetLastError());
    return TB_OK;
}

int launch_midlines(const tb_outline_rec *orecs, const uint32_t *nb_dev, uint32_t nb_max, const float *res, uint32_t cap_pts,
                    const tb_posture_params *P, int do_norm, const float *move_dir, const float *fix_len,
                    float *pts_out, float *segs, tb_midline_rec *mrecs, tb_midline_norm *nrecs, float *norm_pts,
                    float *arena, unsigned long long arena_floats, unsigned long long *arena_used, uint32_t *status, int sms, cudaStream_t s,
                    const uint8_t *skip)
{
    if (nb_max == 0) return TB_OK;
    static DeviceOnce once;
    if (once.need()) {
        TB_CUDA(cudaFuncSetAttribute(midline_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ML_POOL * (int)sizeof(float)));
        once.done();
    }
    TB_CUDA(cudaMemsetAsync(arena_used, 0, sizeof(unsigned long long), s));
    const unsigned grid = (unsigned)std::min<uint64_t>(((uint64_t)nb_max + ML_WARPS - 1) / ML_WARPS, (uint64_t)sms * 3);
    midline_warp_kernel<<<grid, ML_WARPS * 32, ML_POOL * sizeof(float), s>>>(orecs, nb_dev, nb_max, res, cap_pts, *P, do_norm, move_dir, fix_len,
                                                                                 pts_out, (float4 *)segs, mrecs, nrecs, (float4 *)norm_pts,
                                                                                 arena, arena_floats, arena_used, status, skip);
    TB_CUDA(cudaGetLastError());
    return TB_OK;
}

int launch_posture_crops(const tb_blob_rec *recs, const uint32_t *totals, const uint32_t *crop_blob, const tb_line *lines,
                         const uint32_t *line_px, const uint8_t *pixels, const uint8_t *bg, int W, int crop_method,
                         int out_w, int out_h, const tb_midline_norm *nrecs, const float *median_len, float median_len_all,
                         float image_scale, int legacy, uint8_t *crops, double *coef, uint8_t *valid, int max_crops_total, cudaStream_t s)
{
    if (max_crops_total <= 0) return TB_OK;
    posture_coef_kernel<<<(max_crops_total + 127) / 128, 128, 0, s>>>(totals, crop_blob, nrecs, median_len, median_len_all, image_scale, legacy,
                                                                        out_w, out_h, coef, valid);
    TB_CUDA(cudaGetLastError());
    return launch_crop_warp(recs, totals, crop_blob, lines, line_px, pixels, bg, W, crop_method, out_w, out_h, coef, crops, max_crops_total, s);
}

}  // namespace tb
