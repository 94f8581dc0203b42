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
// One warp per blob: the warp rasterises the blob's bounding box into a bit image in shared memory (boxes that do not fit
// fall back to a row table into the blob's line list) and its lanes trace the loops.
// Compiled with -fmad=false: the resampling arithmetic must round like the reference's scalar float code.
#include "common.h"

namespace tb {

__constant__ int c_vx[8] = {0, 1, 1, 1, 0, -1, -1, -1};
__constant__ int c_vy[8] = {-1, -1, 0, 1, 1, 1, 0, -1};

// "Is pixel (x, y) of the blob set": either a bit image of the bounding box (+ 1 pixel border) in shared memory, built by
// the blob's warp, or -- for blobs whose box does not fit -- a row table into the blob's line list in global memory.
constexpr int OL_WARPS = 8, OL_BM_WORDS = 512;           // 2 KB of bit image per warp: boxes up to ~ 126 x 126 / 94 x 168 ...

struct Occupancy {
    const uint32_t *bm; int pw;                 // bit image (nullptr: use the lines), words per row
    const tb_line *l; int n; const uint32_t *row_first;
    int bx0, by0, bx1, by1;
    __device__ __forceinline__ bool set(int x, int y) const
    {
        if (bm) {
            const int rx = x - bx0 + 1, ry = y - by0 + 1;
            return (bm[ry * pw + (rx >> 5)] >> (rx & 31)) & 1u;
        }
        if (x < bx0 || x > bx1 || y < by0 || y > by1) return false;
        for (int li = (int)row_first[y - by0]; li < n; ++li) {
            const tb_line t = l[li];
            if ((int)t.y != y || (int)t.x0 > x) return false;
            if (x <= (int)t.x1) return true;
        }
        return false;
    }
};

// the warp's view of blob r: builds the bit image (all lanes) or the row table (fallback); bm_store = this warp's 2 KB
__device__ __forceinline__ Occupancy make_occupancy(const tb_blob_rec &r, const tb_line *bl, uint32_t *bm_store, uint32_t *rf, bool build_rf, int lane)
{
    Occupancy O;
    O.l = bl; O.n = (int)r.n_lines; O.row_first = rf;
    O.bx0 = r.x0; O.by0 = r.y0; O.bx1 = r.x1; O.by1 = r.y1;
    const int w = O.bx1 - O.bx0 + 1, h = O.by1 - O.by0 + 1;
    O.pw = (w + 2 + 31) >> 5;
    if (O.pw * (h + 2) <= OL_BM_WORDS) {
        for (int i = lane; i < O.pw * (h + 2); i += 32) bm_store[i] = 0u;
        __syncwarp();
        for (int li = lane; li < O.n; li += 32) {
            const tb_line t = bl[li];
            const int a = (int)t.x0 - O.bx0 + 1, b = (int)t.x1 - O.bx0 + 1;
            uint32_t *row = bm_store + ((int)t.y - O.by0 + 1) * O.pw;
            for (int wd = a >> 5; wd <= (b >> 5); ++wd) {
                const int lo = max(a, wd << 5) & 31, hi = min(b, (wd << 5) + 31) & 31;
                atomicOr(row + wd, (0xFFFFFFFFu >> (31 - hi)) & (0xFFFFFFFFu << lo));
            }
        }
        __syncwarp();
        O.bm = bm_store;
    } else {
        O.bm = nullptr;
        if (build_rf) {
            for (int li = lane; li < O.n; li += 32)
                if (li == 0 || bl[li - 1].y != bl[li].y) rf[(int)bl[li].y - O.by0] = (uint32_t)li;
            __syncwarp();
            __threadfence_block();
        }
    }
    return O;
}

struct Side { int x, y, b; };                   // pixel, Direction of the missing neighbour (0 TOP, 2 RIGHT, 4 BOTTOM, 6 LEFT)

__device__ __forceinline__ int side_order(int b) { return b == 0 ? 0 : (b == 6 ? 1 : (b == 2 ? 2 : 3)); }     // direction_from_bool
__device__ __forceinline__ unsigned long long side_key(const Side &s, int bx0, int by0)
{
    return ((((unsigned long long)(s.x - bx0) << 20) | (unsigned long long)(s.y - by0)) << 2) | (unsigned long long)side_order(s.b);
}
__device__ __forceinline__ Side side_succ(const Occupancy &O, const Side &s)
{
    const int l = (s.b + 7) & 7, ll = (s.b + 6) & 7;
    if (O.set(s.x + c_vx[l], s.y + c_vy[l])) return Side{s.x + c_vx[l], s.y + c_vy[l], (s.b + 2) & 7};
    if (O.set(s.x + c_vx[ll], s.y + c_vy[ll])) return Side{s.x + c_vx[ll], s.y + c_vy[ll], s.b};
    return Side{s.x, s.y, ll};
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
    __device__ void segment(float x0, float y0, float x1, float y1, float *out, uint32_t cap)
    {
        const float lx = x1 - x0, ly = y1 - y0;
        const float len = sqrtf(lx * lx + ly * ly);
        walked += len;
        const float percent = len / rd;
        float wp = walked / rd;
        int offset = 0;
        while ((double)wp >= 1.0) {
            const float f = (float)((double)offset * 1.0 / (double)percent);
            if (n < cap) { out[2 * n] = x0 + lx * f; out[2 * n + 1] = y0 + ly * f; }
            ++n; ++offset;
            walked -= rd;
            wp = (float)((double)wp - 1.0);
        }
    }
};

// pass 1, one warp per blob: the outline to take -- the longest loop, the earliest in `_sides` among equals.
// Every loop holds at least one maximal horizontal run of TOP sides (a closed curve has sides facing up) and such a run lies
// in one loop (TOP(x) -> TOP(x - 1) while the left neighbour exists and has no pixel above it).  The left ends of the runs are
// the start candidates; they are dealt to the lanes line by line, every lane traces the loops of its candidates, and a trace
// stops as soon as it meets a run end that precedes its own start in (y, x) order -- so each loop is completed by exactly one
// lane, the one that holds its first run end.  No visited flags, no atomics; the lanes' best loops are then reduced.
__global__ void __launch_bounds__(OL_WARPS * 32)
outline_select_kernel(const tb_blob_rec *__restrict__ recs, uint32_t nb, const tb_line *__restrict__ lines, float rd,
                      uint32_t *__restrict__ row_first, int4 *__restrict__ sel, tb_outline_rec *__restrict__ orecs)
{
    __shared__ uint32_t s_bm[OL_WARPS][OL_BM_WORDS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t q = blockIdx.x * OL_WARPS + warp;
    if (q >= nb) return;
    const tb_blob_rec r = recs[q];
    const tb_line *bl = lines + r.line_off;
    const Occupancy O = make_occupancy(r, bl, s_bm[warp], row_first + r.line_off, true, lane);
    const int bx0 = r.x0, by0 = r.y0;
    const uint32_t max_n = 4u * r.n_pixels + 4u;            // a loop cannot hold more sides than the blob has
    uint32_t best_n = 0; unsigned long long best_c = ~0ull, best_cr = ~0ull; Side best{0, 0, 0};
    for (int li = lane; li < O.n; li += 32) {
        const tb_line ln = bl[li];
        const int y = ln.y;
        for (int x = ln.x0; x <= (int)ln.x1; ++x) {
            if (O.set(x, y - 1) || !(x == (int)ln.x0 || O.set(x - 1, y - 1))) continue;       // not the left end of a run without pixels above
            const Side start{x, y, 0};
            Side cur = start;
            uint32_t n = 0;
            bool mine = true;
            unsigned long long prev_key = 0, first_key = 0, lc = ~0ull, lcr = ~0ull; Side ls = start;
            do {
                const Side nx = side_succ(O, cur);
                if (cur.b == 0 && !(nx.b == 0 && nx.y == cur.y) && (cur.y < start.y || (cur.y == start.y && cur.x < start.x))) { mine = false; break; }
                const unsigned long long k = side_key(cur, bx0, by0);
                if (n == 0) first_key = k;
                else {
                    const unsigned long long c = max(k, prev_key), cr = min(k, prev_key);
                    if (c < lc || (c == lc && cr < lcr)) { lc = c; lcr = cr; ls = cur; }
                }
                prev_key = k;
                cur = nx;
                ++n;
            } while (!(cur.x == start.x && cur.y == start.y && cur.b == start.b) && n < max_n);
            if (!mine) continue;
            {   // the start's predecessor is the last sub-node of the loop
                const unsigned long long c = max(first_key, prev_key), cr = min(first_key, prev_key);
                if (c < lc || (c == lc && cr < lcr)) { lc = c; lcr = cr; ls = start; }
            }
            if (n > best_n || (n == best_n && (lc < best_c || (lc == best_c && lcr < best_cr)))) { best_n = n; best_c = lc; best_cr = lcr; best = ls; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {          // the warp's best: most sides, then the smallest (entry time, creation time)
        const uint32_t on = __shfl_xor_sync(0xffffffffu, best_n, o);
        const unsigned long long oc = __shfl_xor_sync(0xffffffffu, best_c, o), ocr = __shfl_xor_sync(0xffffffffu, best_cr, o);
        const int ox = __shfl_xor_sync(0xffffffffu, best.x, o), oy = __shfl_xor_sync(0xffffffffu, best.y, o), ob = __shfl_xor_sync(0xffffffffu, best.b, o);
        if (on > best_n || (on == best_n && (oc < best_c || (oc == best_c && ocr < best_cr)))) { best_n = on; best_c = oc; best_cr = ocr; best = Side{ox, oy, ob}; }
    }
    if (lane == 0) {
        sel[q] = make_int4(best.x, best.y, best.b, 0);
        tb_outline_rec o; o.raw_off = 0; o.n_raw = best_n; o.res_off = 0;
        // room for the resampled outline: the perimeter is at most n_raw (steps of 1 or sqrt(1/2)), one point per outline_resample walked
        o.n_res = (best_n > 1 && rd > 0.f) ? (uint32_t)fminf((float)best_n / rd + 2.f, 4.0e9f) : best_n;
        orecs[q] = o;
    }
}

// arena offsets: exclusive prefix sums of n_raw / the n_res bounds over the blobs (one CTA); totals[0..1] = sums
__global__ void outline_scan_kernel(tb_outline_rec *__restrict__ orecs, uint32_t nb, uint32_t *__restrict__ totals)
{
    __shared__ uint32_t ws[33];
    unsigned long long base_raw = 0, base_res = 0;
    for (uint32_t i0 = 0; i0 < nb; i0 += blockDim.x) {
        const uint32_t i = i0 + threadIdx.x;
        const uint32_t a = i < nb ? orecs[i].n_raw : 0u, b = i < nb ? orecs[i].n_res : 0u;
        uint32_t ta, tb_;
        const uint32_t ea = block_excl_scan(a, ws, ta);
        const uint32_t eb = block_excl_scan(b, ws, tb_);
        if (i < nb) {
            orecs[i].raw_off = (uint32_t)min(base_raw + ea, 0xFFFFFFFFull); orecs[i].res_off = (uint32_t)min(base_res + eb, 0xFFFFFFFFull);
        }
        base_raw += ta; base_res += tb_;
    }
    if (threadIdx.x == 0) { totals[0] = (uint32_t)min(base_raw, 0xFFFFFFFFull); totals[1] = (uint32_t)min(base_res, 0xFFFFFFFFull); }
}

// pass 2, one warp per blob (the warp builds the bit image, lane 0 walks): the raw outline and its resampled version;
// n_res becomes the number of resampled points
__global__ void __launch_bounds__(OL_WARPS * 32)
outline_emit_kernel(const tb_blob_rec *__restrict__ recs, uint32_t nb, const tb_line *__restrict__ lines,
                    uint32_t *__restrict__ row_first, const int4 *__restrict__ sel,
                    tb_outline_rec *__restrict__ orecs, float rd, float *__restrict__ raw, float *__restrict__ res, uint32_t cap_pts)
{
    __shared__ uint32_t s_bm[OL_WARPS][OL_BM_WORDS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t q = blockIdx.x * OL_WARPS + warp;
    if (q >= nb) return;
    const tb_outline_rec o = orecs[q];
    if (o.n_raw == 0 || (unsigned long long)o.raw_off + o.n_raw > cap_pts || (unsigned long long)o.res_off + o.n_res > cap_pts) return;
    const tb_blob_rec r = recs[q];
    const Occupancy O = make_occupancy(r, lines + r.line_off, s_bm[warp], row_first + r.line_off, false, lane);
    if (lane != 0) return;
    const int bx0 = r.x0, by0 = r.y0;
    const int4 s4 = sel[q];
    Side cur{s4.x, s4.y, s4.z};
    float *rp = raw + 2 * (size_t)o.raw_off, *sp = res + 2 * (size_t)o.res_off;
    const bool resample = o.n_raw > 1 && rd > 0.f;
    Resampler rs; rs.init(rd);
    float x0, y0, fx, fy; side_pos(cur, bx0, by0, x0, y0); fx = x0; fy = y0;
    for (uint32_t i = 0; i < o.n_raw; ++i) {
        rp[2 * i] = x0; rp[2 * i + 1] = y0;
        if (!resample) { sp[2 * i] = x0; sp[2 * i + 1] = y0; }
        float x1, y1;
        if (i + 1 < o.n_raw) { cur = side_succ(O, cur); side_pos(cur, bx0, by0, x1, y1); } else { x1 = fx; y1 = fy; }
        if (resample) rs.segment(x0, y0, x1, y1, sp, o.n_res);
        x0 = x1; y0 = y1;
    }
    if (resample) orecs[q].n_res = min(rs.n, o.n_res);
}

int launch_outlines(const tb_blob_rec *recs, uint32_t nb, const tb_line *lines, float rd, uint32_t *row_first, int4 *sel,
                    tb_outline_rec *orecs, uint32_t *totals, float *raw, float *res, uint32_t cap_pts, cudaStream_t s)
{
    if (nb == 0) { TB_CUDA(cudaMemsetAsync(totals, 0, 8, s)); return TB_OK; }
    const unsigned grid = (nb + OL_WARPS - 1) / OL_WARPS;
    outline_select_kernel<<<grid, OL_WARPS * 32, 0, s>>>(recs, nb, lines, rd, row_first, sel, orecs);
    outline_scan_kernel<<<1, 1024, 0, s>>>(orecs, nb, totals);
    outline_emit_kernel<<<grid, OL_WARPS * 32, 0, s>>>(recs, nb, lines, row_first, sel, orecs, rd, raw, res, cap_pts);
    TB_CUDA(cudaGetLastError());
    return TB_OK;
}

}  // namespace tb
