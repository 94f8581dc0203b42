// "Next" row N4, first stage: the outline of every blob of a batch.
//   pixel::find_outer_points   C/processing/PixelTree.cpp:497-651 (Tree::generate_edges / add_edge / walk :657-1130)
//   the outline calculate_posture selects (the first of maximal size, T/tracking/Posture.cpp:341-348)
//   Outline::resample          T/tracking/Outline.cpp:724-766
// The reference builds a graph of side midpoints with two vectors of sub-nodes and walks it with a deque.  What that
// bookkeeping produces has a closed form, which is what runs here (the CPU restatement used by the parity tests emulates
// the bookkeeping literally; tests/test_gpu_outline.py compares the two):
//   * every missing 4-neighbour side s of a blob pixel has exactly one successor succ(s) (generate_edges :876-965: the
//     diagonal pixel's perpendicular side, else the straight neighbour's same side, else the pixel's own next side) and one
//     predecessor; the sides form closed loops,
//   * edges are emitted pixel by pixel in leaf_index order (x major, then y), sides in the order TOP, LEFT, RIGHT, BOTTOM;
//     a sub-node enters `_sides` when its second edge arrives, i.e. at time max(t(s), t(pred(s))), ties inside one add_edge
//     call going to the sub-node created first, min(t(s), t(pred(s))),
//   * walk() starts at the first sub-node of `_sides` that is not walked yet and, because edges[1] (or the only edge) of a
//     sub-node always points to its successor, emits s, succ(s), succ(succ(s)), ... once around the loop.
// One thread per blob: loops are traced through binary searches in the blob's line list, no per-blob image is built.
// Compiled with -fmad=false: the resampling arithmetic must round like the reference's scalar float code.
#include "common.h"

namespace tb {

__constant__ int c_vx[8] = {0, 1, 1, 1, 0, -1, -1, -1};
__constant__ int c_vy[8] = {-1, -1, 0, 1, 1, 1, 0, -1};

struct BlobLines {
    const tb_line *l; int n;
    // index of the line holding pixel (x, y), or -1
    __device__ int find(int x, int y) const
    {
        if (x < 0 || y < 0 || x > 65535 || y > 65535) return -1;
        const uint32_t key = ((uint32_t)y << 16) | (uint32_t)x;
        int lo = 0, hi = n;                     // last line with (y, x0) <= (y, x)
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            const tb_line t = l[mid];
            if ((((uint32_t)t.y << 16) | t.x0) <= key) lo = mid + 1; else hi = mid;
        }
        if (lo == 0) return -1;
        const tb_line t = l[lo - 1];
        return (t.y == y && x <= (int)t.x1) ? lo - 1 : -1;
    }
};

struct Side { int x, y, b, li; };               // pixel, Direction of the missing neighbour (0 TOP, 2 RIGHT, 4 BOTTOM, 6 LEFT), line index

__device__ __forceinline__ int side_order(int b) { return b == 0 ? 0 : (b == 6 ? 1 : (b == 2 ? 2 : 3)); }     // direction_from_bool
__device__ __forceinline__ unsigned long long side_key(const Side &s, int bx0, int by0)
{
    return ((((unsigned long long)(s.x - bx0) << 20) | (unsigned long long)(s.y - by0)) << 2) | (unsigned long long)side_order(s.b);
}
__device__ __forceinline__ Side side_succ(const BlobLines &L, const Side &s)
{
    const int l = (s.b + 7) & 7, ll = (s.b + 6) & 7;
    int li = L.find(s.x + c_vx[l], s.y + c_vy[l]);
    if (li >= 0) return Side{s.x + c_vx[l], s.y + c_vy[l], (s.b + 2) & 7, li};
    li = L.find(s.x + c_vx[ll], s.y + c_vy[ll]);
    if (li >= 0) return Side{s.x + c_vx[ll], s.y + c_vy[ll], s.b, li};
    return Side{s.x, s.y, ll, s.li};
}
__device__ __forceinline__ void side_pos(const Side &s, int bx0, int by0, float &px, float &py)
{
    px = ((float)(s.x - bx0) + 0.5f) + (float)c_vx[s.b] * 0.5f;
    py = ((float)(s.y - by0) + 0.5f) + (float)c_vy[s.b] * 0.5f;
}

// Outline::resample as a state machine over the point stream
struct Resampler {
    float rd, walked; uint32_t n;
    __device__ void init(float r) { rd = r; walked = 0.f; n = 0; }
    __device__ void segment(float x0, float y0, float x1, float y1, float *out)
    {
        const float lx = x1 - x0, ly = y1 - y0;
        const float len = sqrtf(lx * lx + ly * ly);
        walked += len;
        const float percent = len / rd;
        float wp = walked / rd;
        int offset = 0;
        while ((double)wp >= 1.0) {
            const float f = (float)((double)offset * 1.0 / (double)percent);
            if (out) { out[2 * n] = x0 + lx * f; out[2 * n + 1] = y0 + ly * f; }
            ++n; ++offset;
            walked -= rd;
            wp = (float)((double)wp - 1.0);
        }
    }
};

// pass 1: choose the outline (longest loop, the earliest in `_sides` among equals), count its raw and resampled points
__global__ void outline_select_kernel(const tb_blob_rec *__restrict__ recs, uint32_t nb, const tb_line *__restrict__ lines,
                                      const uint32_t *__restrict__ line_px, int opx, uint8_t *__restrict__ visited, float rd,
                                      int4 *__restrict__ sel, tb_outline_rec *__restrict__ orecs)
{
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nb) return;
    const tb_blob_rec r = recs[q];
    const BlobLines L{lines + r.line_off, (int)r.n_lines};
    const int bx0 = r.x0, by0 = r.y0;
    const int FB[4] = {0, 6, 2, 4};
    uint32_t best_n = 0; unsigned long long best_c = 0, best_cr = 0; Side best{0, 0, 0, 0};
    for (int li = 0; li < L.n; ++li) {
        const tb_line ln = L.l[li];
        const uint32_t vbase = line_px[r.line_off + li] / (uint32_t)opx;
        for (int x = ln.x0; x <= (int)ln.x1; ++x) {
            // interior pixels of a line can only miss their TOP / BOTTOM neighbours
            for (int bi = 0; bi < 4; ++bi) {
                const int b = FB[bi];
                if (b == 6 && x > (int)ln.x0) continue;
                if (b == 2 && x < (int)ln.x1) continue;
                if (visited[vbase + (x - ln.x0)] & (1u << bi)) continue;
                if ((b == 0 || b == 4) && L.find(x, (int)ln.y + c_vy[b]) >= 0) continue;
                // a new loop: once around, tracking the sub-node that enters `_sides` first
                const Side start{x, (int)ln.y, b, li};
                Side cur = start;
                const uint32_t max_n = 4u * r.n_pixels + 4u;          // a loop cannot hold more sides than the blob has
                uint32_t n = 0;
                unsigned long long prev_key = 0, first_key = 0, lc = ~0ull, lcr = ~0ull; Side ls = start;
                do {
                    const uint32_t vb = line_px[r.line_off + cur.li] / (uint32_t)opx + (uint32_t)(cur.x - (int)L.l[cur.li].x0);
                    visited[vb] |= (uint8_t)(1u << side_order(cur.b));
                    const unsigned long long k = side_key(cur, bx0, by0);
                    if (n == 0) first_key = k;
                    else {
                        const unsigned long long c = max(k, prev_key), cr = min(k, prev_key);
                        if (c < lc || (c == lc && cr < lcr)) { lc = c; lcr = cr; ls = cur; }
                    }
                    prev_key = k;
                    cur = side_succ(L, cur);
                    ++n;
                } while (!(cur.x == start.x && cur.y == start.y && cur.b == start.b) && n < max_n);
                {   // the start's predecessor is the last sub-node of the loop
                    const unsigned long long c = max(first_key, prev_key), cr = min(first_key, prev_key);
                    if (c < lc || (c == lc && cr < lcr)) { lc = c; lcr = cr; ls = start; }
                }
                if (n > best_n || (n == best_n && (lc < best_c || (lc == best_c && lcr < best_cr)))) { best_n = n; best_c = lc; best_cr = lcr; best = ls; }
            }
        }
    }
    uint32_t n_res = best_n;
    if (best_n > 1 && rd > 0.f) {                 // dry run of the resampling from the chosen start
        Resampler rs; rs.init(rd);
        Side cur = best;
        float x0, y0, fx, fy; side_pos(cur, bx0, by0, x0, y0); fx = x0; fy = y0;
        for (uint32_t i = 0; i < best_n; ++i) {
            float x1, y1;
            if (i + 1 < best_n) { cur = side_succ(L, cur); side_pos(cur, bx0, by0, x1, y1); } else { x1 = fx; y1 = fy; }
            rs.segment(x0, y0, x1, y1, nullptr);
            x0 = x1; y0 = y1;
        }
        n_res = rs.n;
    }
    sel[q] = make_int4(best.x, best.y, best.b, best.li);
    tb_outline_rec o; o.raw_off = 0; o.n_raw = best_n; o.res_off = 0; o.n_res = n_res;
    orecs[q] = o;
}

// arena offsets: exclusive prefix sums of n_raw / n_res over the blobs (one CTA); totals[0..1] = sums
__global__ void outline_scan_kernel(tb_outline_rec *__restrict__ orecs, uint32_t nb, uint32_t *__restrict__ totals)
{
    __shared__ uint32_t ws[33];
    uint32_t base_raw = 0, base_res = 0;
    for (uint32_t i0 = 0; i0 < nb; i0 += blockDim.x) {
        const uint32_t i = i0 + threadIdx.x;
        const uint32_t a = i < nb ? orecs[i].n_raw : 0u, b = i < nb ? orecs[i].n_res : 0u;
        uint32_t ta, tb_;
        const uint32_t ea = block_excl_scan(a, ws, ta);
        const uint32_t eb = block_excl_scan(b, ws, tb_);
        if (i < nb) { orecs[i].raw_off = base_raw + ea; orecs[i].res_off = base_res + eb; }
        base_raw += ta; base_res += tb_;
    }
    if (threadIdx.x == 0) { totals[0] = base_raw; totals[1] = base_res; }
}

// pass 2: write the raw outline and its resampled version
__global__ void outline_emit_kernel(const tb_blob_rec *__restrict__ recs, uint32_t nb, const tb_line *__restrict__ lines,
                                    const int4 *__restrict__ sel, const tb_outline_rec *__restrict__ orecs, float rd,
                                    float *__restrict__ raw, float *__restrict__ res, uint32_t cap_pts)
{
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nb) return;
    const tb_outline_rec o = orecs[q];
    if (o.n_raw == 0 || o.raw_off + o.n_raw > cap_pts || o.res_off + o.n_res > cap_pts) return;
    const tb_blob_rec r = recs[q];
    const BlobLines L{lines + r.line_off, (int)r.n_lines};
    const int bx0 = r.x0, by0 = r.y0;
    const int4 s4 = sel[q];
    Side cur{s4.x, s4.y, s4.z, s4.w};
    float *rp = raw + 2 * (size_t)o.raw_off, *sp = res + 2 * (size_t)o.res_off;
    const bool resample = o.n_raw > 1 && rd > 0.f;
    Resampler rs; rs.init(rd);
    float x0, y0, fx, fy; side_pos(cur, bx0, by0, x0, y0); fx = x0; fy = y0;
    for (uint32_t i = 0; i < o.n_raw; ++i) {
        rp[2 * i] = x0; rp[2 * i + 1] = y0;
        if (!resample) { sp[2 * i] = x0; sp[2 * i + 1] = y0; }
        float x1, y1;
        if (i + 1 < o.n_raw) { cur = side_succ(L, cur); side_pos(cur, bx0, by0, x1, y1); } else { x1 = fx; y1 = fy; }
        if (resample) rs.segment(x0, y0, x1, y1, sp);
        x0 = x1; y0 = y1;
    }
}

int launch_outlines(const tb_blob_rec *recs, uint32_t nb, const tb_line *lines, const uint32_t *line_px, int opx,
                    uint8_t *visited, size_t visited_bytes, float rd, int4 *sel, tb_outline_rec *orecs, uint32_t *totals,
                    float *raw, float *res, uint32_t cap_pts, cudaStream_t s)
{
    if (nb == 0) { TB_CUDA(cudaMemsetAsync(totals, 0, 8, s)); return TB_OK; }
    TB_CUDA(cudaMemsetAsync(visited, 0, visited_bytes, s));
    outline_select_kernel<<<(nb + 63) / 64, 64, 0, s>>>(recs, nb, lines, line_px, opx, visited, rd, sel, orecs);
    outline_scan_kernel<<<1, 1024, 0, s>>>(orecs, nb, totals);
    outline_emit_kernel<<<(nb + 63) / 64, 64, 0, s>>>(recs, nb, lines, sel, orecs, rd, raw, res, cap_pts);
    TB_CUDA(cudaGetLastError());
    return TB_OK;
}

}  // namespace tb
