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
// One thread per blob, 64 blobs per CTA: a thread rasterises its blob's bounding box into a bit image in a slice of the
// CTA's shared-memory pool (boxes that do not fit fall back to a row table into the blob's line list) and traces its loops.
// Compiled with -fmad=false: the resampling arithmetic must round like the reference's scalar float code.
#include "common.h"

#include <algorithm>

namespace tb {

// Direction d = 0 .. 7 (TOP, TOPR, RIGHT, BOTTOMR, BOTTOM, BOTTOML, LEFT, TOPL): offsets {0,1,1,1,0,-1,-1,-1} / {-1,-1,0,1,1,1,0,-1} from 2-bit
// tables in registers (a walk step is a dependent chain: no constant-memory round trip inside it)
__device__ __forceinline__ int dir_vx(int d) { return (int)((0x01A9u >> (2 * d)) & 3u) - 1; }        // (vx + 1): 1,2,2,2,1,0,0,0
__device__ __forceinline__ int dir_vy(int d) { return (int)((0x1A90u >> (2 * d)) & 3u) - 1; }        // (vy + 1): 0,0,1,2,2,2,1,0

// "Is pixel (x, y) of the blob set": either a bit image of the bounding box (+ 1 pixel border) in shared memory, built by
// the blob's thread from a pool its CTA shares, or -- for blobs whose box does not fit -- a row table into the blob's line
// list in global memory.
// 18 KB of bit images (blob + visited run ends) per CTA of 16 blobs, 12 CTAs per SM.  A walk is one dependent chain per blob: the chip has
// far more warp slots than blobs / 32, so only every OL_SPREAD-th lane owns a blob -- 4 x the warps in flight for the same number of walks,
// and at most 8 lanes of a warp hit the shared-memory banks with their (random) bit-image words per step instead of 32
constexpr int OL_NT = 64, OL_SPREAD = 4, OL_BLOBS = OL_NT / OL_SPREAD, OL_POOL_WORDS = 4608, OL_MAX_WORDS = 2048, OL_CTAS_PER_SM = 12;

struct Occupancy {
    const uint32_t *bm; int pw;                 // bit image (nullptr: use the lines), words per row
    uint32_t *vis;                              // select pass: visited run ends, a second bit image (bm != nullptr) ...
    uint8_t *gvis; const uint32_t *line_px; int opx;   // ... or one byte per blob pixel in global memory (row-table path)
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
    // run end (x, y) seen before?  marks it
    __device__ __forceinline__ bool visit(int x, int y) const
    {
        if (bm) {
            const int rx = x - bx0 + 1, ry = y - by0 + 1;
            uint32_t &wd = vis[ry * pw + (rx >> 5)];
            const uint32_t bit = 1u << (rx & 31);
            const bool seen = wd & bit;
            wd |= bit;
            return seen;
        }
        for (int li = (int)row_first[y - by0]; li < n; ++li) {
            const tb_line t = l[li];
            if (x <= (int)t.x1) {
                uint8_t &f = gvis[line_px[li] / (uint32_t)opx + (uint32_t)(x - (int)t.x0)];
                const bool seen = f != 0;
                f = 1;
                return seen;
            }
        }
        return true;
    }
};

// One thread per blob.  All threads of the CTA call this (it contains block barriers); `valid` = the thread has a blob.
// The blob's thread rasterises its lines into its slice of the pool; blobs that do not get a slice build / use the row table.
// with_vis (select pass): the slice also holds the visited image.
__device__ __forceinline__ Occupancy make_occupancy(bool valid, const tb_blob_rec &r, const tb_line *bl, uint32_t *pool, uint32_t *ws,
                                                    uint32_t *rf, bool build_rf, bool with_vis)
{
    Occupancy O;
    O.vis = nullptr; O.gvis = nullptr; O.line_px = nullptr; O.opx = 1;
    O.l = bl; O.n = valid ? (int)r.n_lines : 0; O.row_first = rf;
    O.bx0 = r.x0; O.by0 = r.y0; O.bx1 = r.x1; O.by1 = r.y1;
    const int w = O.bx1 - O.bx0 + 1, h = O.by1 - O.by0 + 1;
    O.pw = (w + 2 + 31) >> 5;
    const uint32_t words = (valid && (long long)O.pw * (h + 2) <= OL_MAX_WORDS) ? (uint32_t)(O.pw * (h + 2)) : 0u;
    const uint32_t need = with_vis ? 2u * words : words;
    uint32_t total;
    const uint32_t off = block_excl_scan(need, ws, total);
    O.bm = nullptr;
    if (need && off + need <= OL_POOL_WORDS) {
        uint32_t *bm = pool + off;
        for (uint32_t i = 0; i < need; ++i) bm[i] = 0u;
        if (with_vis) O.vis = bm + words;
        for (int li = 0; li < O.n; ++li) {
            const tb_line t = bl[li];
            const int a = (int)t.x0 - O.bx0 + 1, b = (int)t.x1 - O.bx0 + 1;
            uint32_t *row = bm + ((int)t.y - O.by0 + 1) * O.pw;
            for (int wd = a >> 5; wd <= (b >> 5); ++wd) {
                const int lo = max(a, wd << 5) & 31, hi = min(b, (wd << 5) + 31) & 31;
                row[wd] |= (0xFFFFFFFFu >> (31 - hi)) & (0xFFFFFFFFu << lo);
            }
        }
        O.bm = bm;
    } else if (valid && build_rf) {
        for (int li = 0; li < O.n; ++li)
            if (li == 0 || bl[li - 1].y != bl[li].y) rf[(int)bl[li].y - O.by0] = (uint32_t)li;
    }
    return O;
}

struct Side { int x, y, b; };                   // pixel, Direction of the missing neighbour (0 TOP, 2 RIGHT, 4 BOTTOM, 6 LEFT)

__device__ __forceinline__ int side_order(int b) { return (int)((0x1320u >> (2 * b)) & 3u); }     // direction_from_bool: TOP 0, LEFT 1, RIGHT 2, BOTTOM 3
// emission order of a sub-node: pixel x-major, then y, then the side; 14 + 16 + 2 bits (frames are at most 16384 x 65534)
__device__ __forceinline__ uint32_t side_key(const Side &s, int bx0, int by0)
{
    return ((((uint32_t)(s.x - bx0) << 16) | (uint32_t)(s.y - by0)) << 2) | (uint32_t)side_order(s.b);
}
__device__ __forceinline__ Side side_succ(const Occupancy &O, const Side &s)
{
    const int l = (s.b + 7) & 7, ll = (s.b + 6) & 7;
    if (O.set(s.x + dir_vx(l), s.y + dir_vy(l))) return Side{s.x + dir_vx(l), s.y + dir_vy(l), (s.b + 2) & 7};
    if (O.set(s.x + dir_vx(ll), s.y + dir_vy(ll))) return Side{s.x + dir_vx(ll), s.y + dir_vy(ll), s.b};
    return Side{s.x, s.y, ll};
}
__device__ __forceinline__ void side_pos(const Side &s, int bx0, int by0, float &px, float &py)
{
    px = ((float)(s.x - bx0) + 0.5f) + (float)dir_vx(s.b) * 0.5f;
    py = ((float)(s.y - by0) + 0.5f) + (float)dir_vy(s.b) * 0.5f;
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
        const float percent = rd == 1.f ? len : len / rd;          // x / 1 is exact: the default outline_resample skips two divisions per step
        float wp = rd == 1.f ? walked : walked / rd;
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

// pass 1, one thread per blob: the outline to take -- the longest loop, the earliest in `_sides` among equals.
// Every loop holds at least one maximal horizontal run of TOP sides (a closed curve has sides facing up) and such a run lies
// in one loop (TOP(x) -> TOP(x - 1) while the left neighbour exists and has no pixel above it).  The left ends of the runs are
// the start candidates; a trace flags the run ends it passes (a second bit image, or a byte per blob pixel on the row-table
// path), so every loop is walked exactly once, from its first run end in (y, x) order.
struct LoopPick {
    uint32_t best_n = 0, best_c = ~0u, best_cr = ~0u; int bx, by, bb;
    __device__ void trace(const Occupancy &O, int x, int y, uint32_t max_n)
    {
        const Side start{x, y, 0};
        Side cur = start;
        uint32_t n = 0;
        uint32_t prev_key = 0, first_key = 0, lc = ~0u, lcr = ~0u; Side ls = start;
        do {
            const Side nx = side_succ(O, cur);
            if (n && cur.b == 0 && !(nx.b == 0 && nx.y == cur.y)) O.visit(cur.x, cur.y);       // a run end of this loop: not a start any more
            const uint32_t k = side_key(cur, O.bx0, O.by0);
            if (n == 0) first_key = k;
            else {
                const uint32_t c = max(k, prev_key), cr = min(k, prev_key);
                if (c < lc || (c == lc && cr < lcr)) { lc = c; lcr = cr; ls = cur; }
            }
            prev_key = k;
            cur = nx;
            ++n;
        } while (!(cur.x == start.x && cur.y == start.y && cur.b == start.b) && n < max_n);
        {   // the start's predecessor is the last sub-node of the loop
            const uint32_t c = max(first_key, prev_key), cr = min(first_key, prev_key);
            if (c < lc || (c == lc && cr < lcr)) { lc = c; lcr = cr; ls = start; }
        }
        if (n > best_n || (n == best_n && (lc < best_c || (lc == best_c && lcr < best_cr)))) { best_n = n; best_c = lc; best_cr = lcr; bx = ls.x; by = ls.y; bb = ls.b; }
    }
};

__global__ void __launch_bounds__(OL_NT)
outline_select_kernel(const tb_blob_rec *__restrict__ recs, const uint32_t *__restrict__ nb_dev, uint32_t nb_max, const tb_line *__restrict__ lines,
                      const uint32_t *__restrict__ line_px, int opx, uint8_t *__restrict__ visited, float rd,
                      uint32_t *__restrict__ row_first, int4 *__restrict__ sel, tb_outline_rec *__restrict__ orecs, OutlineMap map)
{
    __shared__ uint32_t s_pool[OL_POOL_WORDS];
    __shared__ uint32_t s_ws[33];
    const uint32_t nb = min(nb_dev ? *nb_dev : nb_max, nb_max);
    for (uint32_t base = blockIdx.x * OL_BLOBS; base < nb; base += gridDim.x * OL_BLOBS) {
    const uint32_t q = base + threadIdx.x / OL_SPREAD;
    const bool live = (threadIdx.x % OL_SPREAD) == 0 && q < nb && !(map.skip && map.skip[q]);
    const uint32_t qi = live ? (map.index ? map.index[q] : q) : 0xFFFFFFFFu;       // record q's outline comes from blob qi (posture of a sub-blob)
    const bool valid = qi != 0xFFFFFFFFu;
    tb_blob_rec r{};
    if (valid) r = recs[qi];
    const tb_line *bl = lines + r.line_off;
    __syncthreads();                                         // the previous group is done with the pool
    Occupancy O = make_occupancy(valid, r, bl, s_pool, s_ws, row_first + r.line_off, true, true);
    if (!valid) { if (live) orecs[q] = tb_outline_rec{0, 0, 0, 0}; continue; }
    O.gvis = visited; O.line_px = line_px + r.line_off; O.opx = opx;
    const uint32_t max_n = 4u * r.n_pixels + 4u;            // a loop cannot hold more sides than the blob has
    LoopPick pick; pick.bx = pick.by = pick.bb = 0;
    if (O.bm) {
        // run ends from the bit image: pixels without a pixel above whose left neighbour is not one of them
        const int h = O.by1 - O.by0 + 1;
        for (int ry = 1; ry <= h; ++ry) {
            uint32_t carry = 0;                              // "no pixel above" flag of the previous word's last pixel
            for (int wd = 0; wd < O.pw; ++wd) {
                const uint32_t tm = O.bm[ry * O.pw + wd] & ~O.bm[(ry - 1) * O.pw + wd];
                uint32_t ends = tm & ~((tm << 1) | carry);
                carry = tm >> 31;
                while (ends) {
                    const int bit = __ffs(ends) - 1;
                    ends &= ends - 1;
                    const int x = (wd << 5) + bit - 1 + O.bx0, y = ry - 1 + O.by0;
                    if (!O.visit(x, y)) pick.trace(O, x, y, max_n);
                }
            }
        }
    } else {
        for (int li = 0; li < O.n; ++li) {
            const tb_line ln = bl[li];
            for (int x = ln.x0; x <= (int)ln.x1; ++x)
                if (!O.set(x, (int)ln.y - 1) && (x == (int)ln.x0 || O.set(x - 1, (int)ln.y - 1)) && !O.visit(x, ln.y)) pick.trace(O, x, ln.y, max_n);
        }
    }
    sel[q] = make_int4(pick.bx, pick.by, pick.bb, 0);
    tb_outline_rec o; o.raw_off = 0; o.n_raw = pick.best_n; o.res_off = 0;
    // room for the resampled outline: the perimeter is at most n_raw (steps of 1 or sqrt(1/2)), one point per outline_resample walked
    o.n_res = (pick.best_n > 1 && rd > 0.f) ? (uint32_t)fminf((float)pick.best_n / rd + 2.f, 4.0e9f) : pick.best_n;
    orecs[q] = o;
    }
}

// arena offsets: exclusive prefix sums of n_raw / the n_res bounds over the blobs (one CTA); totals[0..1] = sums
__global__ void outline_scan_kernel(tb_outline_rec *__restrict__ orecs, const uint32_t *__restrict__ nb_dev, uint32_t nb_max, uint32_t *__restrict__ totals,
                                    const uint8_t *__restrict__ skip, int append)
{
    __shared__ uint32_t ws[33];
    const uint32_t nb = min(nb_dev ? *nb_dev : nb_max, nb_max);
    unsigned long long base_raw = append ? totals[0] : 0, base_res = append ? totals[1] : 0;      // append: behind the points of earlier rounds
    __syncthreads();
    for (uint32_t i0 = 0; i0 < nb; i0 += blockDim.x) {
        const uint32_t i = i0 + threadIdx.x;
        const bool live = i < nb && !(skip && skip[i]);
        const uint32_t a = live ? orecs[i].n_raw : 0u, b = live ? orecs[i].n_res : 0u;
        uint32_t ta, tb_;
        const uint32_t ea = block_excl_scan(a, ws, ta);
        const uint32_t eb = block_excl_scan(b, ws, tb_);
        if (live) {
            orecs[i].raw_off = (uint32_t)min(base_raw + ea, 0xFFFFFFFFull); orecs[i].res_off = (uint32_t)min(base_res + eb, 0xFFFFFFFFull);
        }
        base_raw += ta; base_res += tb_;
    }
    if (threadIdx.x == 0) { totals[0] = (uint32_t)min(base_raw, 0xFFFFFFFFull); totals[1] = (uint32_t)min(base_res, 0xFFFFFFFFull); }
}

// pass 2, one thread per blob: the raw outline and its resampled version; n_res becomes the number of resampled points
__global__ void __launch_bounds__(OL_NT)
outline_emit_kernel(const tb_blob_rec *__restrict__ recs, const uint32_t *__restrict__ nb_dev, uint32_t nb_max, const tb_line *__restrict__ lines,
                    uint32_t *__restrict__ row_first, const int4 *__restrict__ sel,
                    tb_outline_rec *__restrict__ orecs, float rd, float *__restrict__ raw, float *__restrict__ res, uint32_t cap_pts, OutlineMap map)
{
    __shared__ uint32_t s_pool[OL_POOL_WORDS];
    __shared__ uint32_t s_ws[33];
    const uint32_t nb = min(nb_dev ? *nb_dev : nb_max, nb_max);
    for (uint32_t base = blockIdx.x * OL_BLOBS; base < nb; base += gridDim.x * OL_BLOBS) {
    const uint32_t q = base + threadIdx.x / OL_SPREAD;
    const bool live = (threadIdx.x % OL_SPREAD) == 0 && q < nb && !(map.skip && map.skip[q]);
    tb_outline_rec o{};
    if (live) o = orecs[q];
    const bool valid = live && o.n_raw != 0 && (unsigned long long)o.raw_off + o.n_raw <= cap_pts && (unsigned long long)o.res_off + o.n_res <= cap_pts;
    tb_blob_rec r{};
    if (valid) r = recs[map.index ? map.index[q] : q];
    __syncthreads();                                         // the previous group is done with the pool
    const Occupancy O = make_occupancy(valid, r, lines + r.line_off, s_pool, s_ws, row_first + r.line_off, false, false);
    if (!valid) { if (live && o.n_raw != 0) orecs[q].n_res = 0; continue; }       // arena overflow: no points (the totals report it)
    // points are relative to the blob's own bounds, or -- posture of a sub-blob -- to its parent's (Posture.cpp:337)
    const int bx0 = map.origin ? (int)map.origin[q].x0 : (int)r.x0, by0 = map.origin ? (int)map.origin[q].y0 : (int)r.y0;
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
}

int launch_outlines(const tb_blob_rec *recs, const uint32_t *nb_dev, uint32_t nb_max, const tb_line *lines, const uint32_t *line_px, int opx,
                    uint8_t *visited, size_t visited_bytes, float rd, uint32_t *row_first, int4 *sel,
                    tb_outline_rec *orecs, uint32_t *totals, float *raw, float *res, uint32_t cap_pts, int sms, cudaStream_t s, const OutlineMap *map)
{
    const OutlineMap m = map ? *map : OutlineMap{nullptr, nullptr, nullptr, 0};
    if (nb_max == 0) { if (!m.append) TB_CUDA(cudaMemsetAsync(totals, 0, 8, s)); return TB_OK; }
    TB_CUDA(cudaMemsetAsync(visited, 0, visited_bytes, s));
    // persistent over groups of OL_BLOBS blobs
    const unsigned grid = (unsigned)std::min<uint64_t>(((uint64_t)nb_max + OL_BLOBS - 1) / OL_BLOBS, (uint64_t)std::max(1, sms) * OL_CTAS_PER_SM);
    outline_select_kernel<<<grid, OL_NT, 0, s>>>(recs, nb_dev, nb_max, lines, line_px, opx, visited, rd, row_first, sel, orecs, m);
    outline_scan_kernel<<<1, 1024, 0, s>>>(orecs, nb_dev, nb_max, totals, m.skip, m.append);
    outline_emit_kernel<<<grid, OL_NT, 0, s>>>(recs, nb_dev, nb_max, lines, row_first, sel, orecs, rd, raw, res, cap_pts, m);
    TB_CUDA(cudaGetLastError());
    return TB_OK;
}

}  // namespace tb
