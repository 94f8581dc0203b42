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
// One thread per blob: loops are traced through a per-blob row table into the blob's line list, no per-blob image is built.
// Compiled with -fmad=false: the resampling arithmetic must round like the reference's scalar float code.
#include "common.h"

namespace tb {

__constant__ int c_vx[8] = {0, 1, 1, 1, 0, -1, -1, -1};
__constant__ int c_vy[8] = {-1, -1, 0, 1, 1, 1, 0, -1};

struct BlobLines {
    const tb_line *l; int n;
    const uint32_t *row_first;                  // first line index of every row of the bounding box (a connected blob has no empty row)
    int bx0, by0, bx1, by1;
    // index of the line holding pixel (x, y), or -1
    __device__ int find(int x, int y) const
    {
        if (x < bx0 || x > bx1 || y < by0 || y > by1) return -1;
        for (int li = (int)row_first[y - by0]; li < n; ++li) {
            const tb_line t = l[li];
            if ((int)t.y != y || (int)t.x0 > x) return -1;
            if (x <= (int)t.x1) return li;
        }
        return -1;
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

// pass 1: the row table of the blob, then the outline to take: the longest loop, the earliest in `_sides` among equals.
// Every loop holds at least one maximal horizontal run of TOP sides (a closed curve has sides facing up), and such a run
// lies in one loop (TOP(x) -> TOP(x - 1) while the left neighbour exists and has no pixel above it), so the left ends of
// the runs are the only start candidates that need a visited flag: one flag per blob pixel.
__global__ void outline_select_kernel(const tb_blob_rec *__restrict__ recs, uint32_t nb, const tb_line *__restrict__ lines,
                                      const uint32_t *__restrict__ line_px, int opx, uint8_t *__restrict__ visited, float rd,
                                      uint32_t *__restrict__ row_first, int4 *__restrict__ sel, tb_outline_rec *__restrict__ orecs)
{
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nb) return;
    const tb_blob_rec r = recs[q];
    uint32_t *rf = row_first + r.line_off;
    const tb_line *bl = lines + r.line_off;
    for (int li = 0; li < (int)r.n_lines; ++li)
        if (li == 0 || bl[li - 1].y != bl[li].y) rf[(int)bl[li].y - (int)r.y0] = (uint32_t)li;
    const BlobLines L{bl, (int)r.n_lines, rf, (int)r.x0, (int)r.y0, (int)r.x1, (int)r.y1};
    const int bx0 = r.x0, by0 = r.y0;
    const uint32_t max_n = 4u * r.n_pixels + 4u;            // a loop cannot hold more sides than the blob has
    uint32_t best_n = 0; unsigned long long best_c = 0, best_cr = 0; Side best{0, 0, 0, 0};
    for (int li = 0; li < L.n; ++li) {
        const tb_line ln = bl[li];
        const uint32_t vbase = line_px[r.line_off + li] / (uint32_t)opx;
        // runs of pixels of this line without a pixel above: the line minus the lines of the row above
        int s0 = ln.x0;
        int ui = ((int)ln.y > by0) ? (int)rf[(int)ln.y - 1 - by0] : L.n;
        while (s0 <= (int)ln.x1) {
            while (ui < L.n && (int)bl[ui].y == (int)ln.y - 1 && (int)bl[ui].x1 < s0) ++ui;        // upper lines left of s0
            const bool up = ui < L.n && (int)bl[ui].y == (int)ln.y - 1;
            if (up && (int)bl[ui].x0 <= s0) { s0 = (int)bl[ui].x1 + 1; continue; }                 // s0 has a pixel above: jump past that line
            const int e0 = (up && (int)bl[ui].x0 <= (int)ln.x1) ? (int)bl[ui].x0 - 1 : (int)ln.x1;  // the run [s0, e0] has no pixel above
            if (!(visited[vbase + (uint32_t)(s0 - (int)ln.x0)] & 1u)) {
                // a new loop: once around from the left end of the run, tracking the sub-node that enters `_sides` first
                const Side start{s0, (int)ln.y, 0, li};
                Side cur = start;
                uint32_t n = 0;
                unsigned long long prev_key = 0, first_key = 0, lc = ~0ull, lcr = ~0ull; Side ls = start;
                do {
                    const Side nx = side_succ(L, cur);
                    if (cur.b == 0 && !(nx.b == 0 && nx.y == cur.y)) {        // left end of a run of TOP sides
                        const uint32_t vb = line_px[r.line_off + cur.li] / (uint32_t)opx + (uint32_t)(cur.x - (int)bl[cur.li].x0);
                        visited[vb] |= 1u;
                    }
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
                {   // the start's predecessor is the last sub-node of the loop
                    const unsigned long long c = max(first_key, prev_key), cr = min(first_key, prev_key);
                    if (c < lc || (c == lc && cr < lcr)) { lc = c; lcr = cr; ls = start; }
                }
                if (n > best_n || (n == best_n && (lc < best_c || (lc == best_c && lcr < best_cr)))) { best_n = n; best_c = lc; best_cr = lcr; best = ls; }
            }
            s0 = e0 + 1;
        }
    }
    sel[q] = make_int4(best.x, best.y, best.b, best.li);
    tb_outline_rec o; o.raw_off = 0; o.n_raw = best_n; o.res_off = 0;
    // room for the resampled outline: the perimeter is at most n_raw (steps of 1 or sqrt(1/2)), one point per outline_resample walked
    o.n_res = (best_n > 1 && rd > 0.f) ? (uint32_t)fminf((float)best_n / rd + 2.f, 4.0e9f) : best_n;
    orecs[q] = o;
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

// pass 2: write the raw outline and its resampled version; n_res becomes the number of resampled points
__global__ void outline_emit_kernel(const tb_blob_rec *__restrict__ recs, uint32_t nb, const tb_line *__restrict__ lines,
                                    const uint32_t *__restrict__ row_first, const int4 *__restrict__ sel,
                                    tb_outline_rec *__restrict__ orecs, float rd, float *__restrict__ raw, float *__restrict__ res, uint32_t cap_pts)
{
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nb) return;
    const tb_outline_rec o = orecs[q];
    if (o.n_raw == 0 || (unsigned long long)o.raw_off + o.n_raw > cap_pts || (unsigned long long)o.res_off + o.n_res > cap_pts) return;
    const tb_blob_rec r = recs[q];
    const BlobLines L{lines + r.line_off, (int)r.n_lines, row_first + r.line_off, (int)r.x0, (int)r.y0, (int)r.x1, (int)r.y1};
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
        if (resample) rs.segment(x0, y0, x1, y1, sp, o.n_res);
        x0 = x1; y0 = y1;
    }
    if (resample) orecs[q].n_res = min(rs.n, o.n_res);
}

int launch_outlines(const tb_blob_rec *recs, uint32_t nb, const tb_line *lines, const uint32_t *line_px, int opx,
                    uint8_t *visited, size_t visited_bytes, float rd, uint32_t *row_first, int4 *sel, tb_outline_rec *orecs, uint32_t *totals,
                    float *raw, float *res, uint32_t cap_pts, cudaStream_t s)
{
    if (nb == 0) { TB_CUDA(cudaMemsetAsync(totals, 0, 8, s)); return TB_OK; }
    TB_CUDA(cudaMemsetAsync(visited, 0, visited_bytes, s));
    outline_select_kernel<<<(nb + 31) / 32, 32, 0, s>>>(recs, nb, lines, line_px, opx, visited, rd, row_first, sel, orecs);
    outline_scan_kernel<<<1, 1024, 0, s>>>(orecs, nb, totals);
    outline_emit_kernel<<<(nb + 31) / 32, 32, 0, s>>>(recs, nb, lines, row_first, sel, orecs, rd, raw, res, cap_pts);
    TB_CUDA(cudaGetLastError());
    return TB_OK;
}

}  // namespace tb
