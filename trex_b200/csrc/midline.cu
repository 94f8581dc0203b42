// "Next" row N4, second stage: from the resampled outline of every blob (tb_seg_outlines) to the raw midline.
//   Outline::smooth / smooth_outline         T/tracking/Outline.cpp:330-452
//   Outline::offset_to_middle                T/tracking/Outline.cpp:454-718   (peak_mode pointy, the default)
//   periodic::differentiate_and_test_clockwise, eft, ieft, curvature, find_peaks, fast::cos
//                                            C/misc/CircularGraph.cpp:12-606
//   Outline::calculate_midline               T/tracking/Outline.cpp:768-868
// One thread per blob; the per-blob work arrays live in a global scratch arena (32 floats per outline point).  The arithmetic is
// the reference's scalar float code with its double promotions, compiled with -fmad=false so that it rounds the same way.
#include "common.h"

namespace tb {

struct MlPeak { float x, y, width, integral, r0, r1, max_y_extrema, max_y; };

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

// _differentiate<true> (CircularGraph.cpp:409-463): cyclic first difference + the orientation sum as the reference accumulates it
__device__ float ml_differentiate(const float *p, int N, float *dxy)
{
    float sum = 0;
    for (int i = 1; i < N; ++i) {
        if (dxy) { dxy[2 * (i - 1)] = p[2 * i] - p[2 * (i - 1)]; dxy[2 * (i - 1) + 1] = p[2 * i + 1] - p[2 * (i - 1) + 1]; }
        sum += p[2 * (i - 1)] * p[2 * i + 1] - p[2 * i] * p[2 * (i - 1) + 1];
    }
    if (dxy) { dxy[2 * (N - 1)] = p[0] - p[2 * (N - 1)]; dxy[2 * (N - 1) + 1] = p[1] - p[2 * (N - 1) + 1]; }
    sum += p[0] * p[2 * (N - 1) + 1] - p[2 * (N - 1)] * p[1];
    return sum;
}

__global__ void __launch_bounds__(64)
midline_kernel(const tb_outline_rec *__restrict__ orecs, uint32_t nb, const float *__restrict__ res, uint32_t cap_pts,
               tb_posture_params P, float *__restrict__ pts_out, float4 *__restrict__ segs, tb_midline_rec *__restrict__ mrecs,
               float *__restrict__ scratch)
{
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nb) return;
    const tb_outline_rec o = orecs[q];
    tb_midline_rec mr; mr.seg_off = o.res_off; mr.n_seg = 0; mr.tail = -1; mr.head = -1;
    const int N = (int)o.n_res;
    if (N <= 0 || (unsigned long long)o.res_off + o.n_res > cap_pts) { mr.tail = -1; mrecs[q] = mr; return; }
    float *p = pts_out + 2 * (size_t)o.res_off;
    float *w = scratch + 32 * ((size_t)o.res_off + 2 * (size_t)q);           // 32 * (N + 2) floats of work space
    for (int i = 0; i < 2 * N; ++i) p[i] = res[2 * (size_t)o.res_off + i];
    float *sm = w;                 // 2N
    float *dxy = w + 2 * (N + 2);  // 2N
    float *dt = w + 4 * (N + 2), *phi = w + 5 * (N + 2), *cx = w + 6 * (N + 2), *cy = w + 7 * (N + 2), *cum = w + 8 * (N + 2);
    float *cs = w + 9 * (N + 2);   // 2N
    float *curv = w + 11 * (N + 2), *diff = w + 12 * (N + 2);
    float *ext_i = w + 13 * (N + 2); int *ext_m = (int *)(w + 14 * (N + 2));
    MlPeak *mx = (MlPeak *)(w + 15 * (N + 2));                                // 8 floats each, up to w + 23 (N + 2)

    // ---- Outline::smooth (:380-389, smooth_outline :330-378)
    if (P.outline_smooth_samples > 0) {
        const float range = (float)P.outline_smooth_samples;
        if ((float)N > range) {
            const int step = P.outline_smooth_step;
            const float step_row = range * (float)step;
            float wsum = 0; int nw = 0;
            for (int i = (int)(-step_row); (float)i <= step_row; i += step) { wsum += (step_row - (float)abs(i)) / step_row; ++nw; }
            for (int i = 0; i < N; ++i) {
                float px = 0, py = 0;
                int k = (int)(-step_row);
                for (int j = (int)((float)i - step_row); (float)j <= (float)i + step_row; j += step, k += step) {
                    int idx = j;
                    while (idx < 0) idx += N;
                    while (idx >= N) idx -= N;
                    const float wgt = ((step_row - (float)abs(k)) / step_row) / wsum;
                    px += p[2 * idx] * wgt; py += p[2 * idx + 1] * wgt;
                }
                sm[2 * i] = px; sm[2 * i + 1] = py;
            }
            (void)nw;
            for (int i = 0; i < 2 * N; ++i) p[i] = sm[i];
        }
    }
    // ---- offset_to_middle (:454-718)
    if (ml_differentiate(p, N, nullptr) < 0)
        for (int i = 0, j = N - 1; i < j; ++i, --j) {
            const float tx = p[2 * i], ty = p[2 * i + 1];
            p[2 * i] = p[2 * j]; p[2 * i + 1] = p[2 * j + 1]; p[2 * j] = tx; p[2 * j + 1] = ty;
        }
    if (P.outline_approximate > 0) {
        float ccx = 0, ccy = 0;
        for (int i = 0; i < N; ++i) { ccx += p[2 * i]; ccy += p[2 * i + 1]; }
        ccx /= (float)N; ccy /= (float)N;
        // eft (:484-561)
        ml_differentiate(p, N, dxy);
        const int nd = N - 1;
        float sum = 0;
        cum[0] = 0; phi[0] = 0;
        for (int i = 0; i < nd; ++i) {
            const float x = dxy[2 * i], y = dxy[2 * i + 1];
            dt[i] = (float)((double)sqrtf(x * x + y * y) + 1e-10);
            sum += dt[i];
            cum[i + 1] = sum;
            phi[i + 1] = (float)(2 * 3.14159265358979323846 * (double)cum[i + 1]);
            cx[i] = x / dt[i]; cy[i] = y / dt[i];
        }
        const float T = cum[nd];
        const float norm_base = (float)((double)T / (2 * (3.14159265358979323846 * 3.14159265358979323846)));
        const int np = nd + 1;
        float coef[8][4];
        const int order_n = min(P.outline_approximate, 8);
        for (int n = 1; n < order_n + 1; ++n) {
            const float norm = norm_base / (float)((size_t)n * (size_t)n);
            float cnx = 0, cny = 0, snx = 0, sny = 0;
            for (int i = 0; i < np; ++i) {
                const float phi_n = phi[i] * (float)n / T;
                cs[2 * i] = ml_fast_cos(phi_n); cs[2 * i + 1] = ml_fast_sin(phi_n);
            }
            for (int i = 0; i < np - 1; ++i) {
                const float dc = cs[2 * (i + 1)] - cs[2 * i], ds = cs[2 * (i + 1) + 1] - cs[2 * i + 1];
                cnx += cx[i] * dc; cny += cy[i] * dc;
                snx += cx[i] * ds; sny += cy[i] * ds;
            }
            cnx *= norm; cny *= norm; snx *= norm; sny *= norm;
            coef[n - 1][0] = cnx; coef[n - 1][1] = snx; coef[n - 1][2] = cny; coef[n - 1][3] = sny;
        }
        // ieft (:563-606)
        for (int j = 0; j < N; ++j) { p[2 * j] = ccx; p[2 * j + 1] = ccy; }
        for (int i = 0; i < order_n; ++i)
            for (int j = 0; j < N; ++j) {
                const float t = (float)((double)j / (double)(N - 1) * 3.14159265358979323846 * 2.0);
                const float ct = ml_fast_cos(t * (float)(i + 1)), st = ml_fast_sin(t * (float)(i + 1));
                p[2 * j] += 1.f * (coef[i][0] * ct + coef[i][1] * st);
                p[2 * j + 1] += 1.f * (coef[i][2] * ct + coef[i][3] * st);
            }
    }
    // curvature (:49-113)
    {
        float rf = P.outline_curvature_range_ratio * (float)N;
        if (rf < 1.f) rf = 1.f;
        const int r = (int)rf;
        const bool absolute = P.outline_approximate > 0;
        for (int i = 0; i < N; ++i) {
            const int i1 = ((i - r) % N + N) % N, i3 = (i + r) % N;
            const float x1 = p[2 * i1], y1 = p[2 * i1 + 1], x2 = p[2 * i], y2 = p[2 * i + 1], x3 = p[2 * i3], y3 = p[2 * i3 + 1];
            const bool e12 = x1 == x2 && y1 == y2, e13 = x1 == x3 && y1 == y3, e23 = x2 == x3 && y2 == y3;
            float v = 0.f;
            if (!e12 && !e13 && !e23) {
                const float cross = (x2 - x1) * (y3 - y2) - (y2 - y1) * (x3 - x2);
                const float d12 = (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1), d23 = (x3 - x2) * (x3 - x2) + (y3 - y2) * (y3 - y2),
                            d13 = (x3 - x1) * (x3 - x1) + (y3 - y1) * (y3 - y1);
                v = 2.f * (absolute ? fabsf(cross) : cross) / sqrtf(d12 * d23 * d13);
            }
            curv[i] = v;
        }
    }
    // find_peaks (:115-407), FIND_POINTY; only the maxima's positions and heights decide tail and head in this mode, but the
    // ranges are computed like the reference does (they bound each other)
    int n_ext = 0, n_max = 0;
    {
        for (int i = 0; i + 1 < N; ++i) diff[i] = curv[i + 1] - curv[i];
        diff[N - 1] = curv[0] - curv[N - 1];
        int sign = diff[N - 1] < 0;
        for (int i = 0; i < N; ++i) {
            const int c = diff[i] < 0;
            if (c != sign) {
                if (diff[i] != 0) {
                    if (!sign) { MlPeak pk; pk.x = (float)i; pk.y = curv[i]; pk.width = 0; pk.integral = 0; pk.r0 = -1; pk.r1 = -1; pk.max_y_extrema = 0; pk.max_y = 0; mx[n_max++] = pk; }
                    ext_i[n_ext] = (float)i; ext_m[n_ext] = !sign; ++n_ext;
                }
                sign = c;
            }
        }
    }
    float max_y = -1, max_y_idx = 0;
    for (int k = 0; k < n_max; ++k)
        if (mx[k].y > max_y) { max_y = mx[k].y; max_y_idx = mx[k].x; }
    const float idx = max_y_idx;
    int tail = (int)idx, head = -1;
    float max_d = 0;
    for (int k = 0; k < n_max; ++k) {
        float d;
        const float px = mx[k].x, sz = (float)N;
        if (px >= idx) { const float a = fabsf(px - idx), b = fabsf(px - idx - sz); d = a < b ? a : b; }
        else { const float a = fabsf(idx - px), b = fabsf(idx - px - sz); d = a < b ? a : b; }
        if (d > max_d) { max_d = d; head = (int)px; }
    }
    int rot;
    if (P.midline_start_with_head && head != -1) {
        if (tail != -1) { tail -= head; if (tail < 0) tail += N; }
        rot = head; head = 0;
    } else {
        if (head != -1) { head -= tail; if (head < 0) head += N; }
        rot = tail; tail = 0;
    }
    if (rot > 0 && rot < N) {                       // std::rotate(begin, begin + rot, end) through the smoothing buffer
        for (int i = 0; i < N; ++i) { const int s = (i + rot) % N; sm[2 * i] = p[2 * s]; sm[2 * i + 1] = p[2 * s + 1]; }
        for (int i = 0; i < 2 * N; ++i) p[i] = sm[i];
    }
    if (P.midline_invert) { const int t = tail; tail = head; head = t; }
    mr.tail = tail; mr.head = head;
    // ---- calculate_midline: the pairing walk (:786-868)
    if (N > 1) {
        const int L = N;
        int idx_r = 1, idx_l = -1;
        float mo = P.midline_walk_offset * (float)L;
        if (mo < 3.f) mo = 3.f;
        const int max_offset = (int)mo;
        uint32_t ns = 0;
        float4 *so = segs + o.res_off;
        while (idx_r < L + idx_l) {
            float prx = 0, pry = 0, plx = p[2 * (L + idx_l)], ply = p[2 * (L + idx_l) + 1];
            float min_d = 3.402823466e+38f; int min_idx = -1;
            for (int i = 0; i < max_offset; ++i) {
                if (idx_r + i >= L) break;
                const float dx = p[2 * (idx_r + i)] - plx, dy = p[2 * (idx_r + i) + 1] - ply;
                const float len = sqrtf(dx * dx + dy * dy);
                if (len < min_d) { min_d = len; min_idx = idx_r + i; }
            }
            if (min_idx != -1) { prx = p[2 * min_idx]; pry = p[2 * min_idx + 1]; idx_r = min_idx; }
            min_d = 3.402823466e+38f; min_idx = 1;
            for (int i = 0; i < max_offset; ++i) {
                if (idx_l - i <= -L) break;
                const float dx = prx - p[2 * (L + idx_l - i)], dy = pry - p[2 * (L + idx_l - i) + 1];
                const float len = sqrtf(dx * dx + dy * dy);
                if (len < min_d) { min_d = len; min_idx = idx_l - i; }
            }
            if (min_idx != 1) { plx = p[2 * (L + min_idx)]; ply = p[2 * (L + min_idx) + 1]; idx_l = min_idx; }
            const float lx = prx - plx, ly = pry - ply;
            const float mx_ = plx + lx * 0.5f, my_ = ply + ly * 0.5f;
            if (ns < o.n_res)
                so[ns] = make_float4(mx_, my_, sqrtf((prx - plx) * (prx - plx) + (pry - ply) * (pry - ply)),
                                     sqrtf((plx - mx_) * (plx - mx_) + (ply - my_) * (ply - my_)));
            ++ns;
            idx_r++; idx_l--;
        }
        mr.n_seg = ns > 2 ? min(ns, o.n_res) : 0;         // "Too few midline segments calculated." (:863-866)
    }
    mrecs[q] = mr;
}

int launch_midlines(const tb_outline_rec *orecs, uint32_t nb, const float *res, uint32_t cap_pts, const tb_posture_params *P,
                    float *pts_out, float *segs, tb_midline_rec *mrecs, float *scratch, cudaStream_t s)
{
    if (nb == 0) return TB_OK;
    midline_kernel<<<(nb + 63) / 64, 64, 0, s>>>(orecs, nb, res, cap_pts, *P, pts_out, (float4 *)segs, mrecs, scratch);
    TB_CUDA(cudaGetLastError());
    return TB_OK;
}

}  // namespace tb
