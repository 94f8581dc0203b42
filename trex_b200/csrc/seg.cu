// Segmentation hot path on sm_100a:  |frame - background| -> threshold -> (mask & frame) -> row RLE
// -> 8-connected run labeling -> per-blob sorted line lists + pixel bytes -> size filter ->
// individual crops.  Replaces, behind the C ABI of include/trexb200.h,
//   RawProcessing::generate_binary   C/processing/RawProcessing.cpp:263-600 (default branch family)
//   Source::extract_lines            C/processing/Source.cpp:156-255
//   merge_lines / run_fast           C/processing/CPULabeling.cpp:44-343
//   BackgroundSubtraction::apply     T/python/BackgroundSubtraction.cpp:126-347 (size filter :259, :306)
//   image::calculate_diff_image      T/tracking/FilterCache.cpp:158-235
// Not a translation of those: the reference walks pixels serially and merges blobs through an
// object graph; here a frame batch is streamed once from HBM (K1), runs are labelled with a
// lock-free union-find whose root is the raster-first run (K2), and blobs are materialised by
// warps walking their bounding boxes (K3).  Data layout and rooflines: DESIGN.md.
#include "common.h"
#include "umma.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace tb {

// ------------------------------------------------------------------------------------------------
// device-side description of one seg handle
// ------------------------------------------------------------------------------------------------
struct SegK {                  // threshold configuration, bytes replicated x4
    uint32_t flags;            // bit0 enable_difference, bit1 absolute, bit2 image_invert,
                               // bit3 invert mask (T<0), bit4 inRange
    uint32_t t4, lo4, hi4;
};
enum { F_DIFF = 1, F_ABS = 2, F_INV = 4, F_INVMASK = 8, F_RANGE = 16, F_PREMASK = 32, F_GE = 64 };

struct SegDev {
    int W, H, B, cpr, rpt, n_bands, aligned;
    int wpr;                   // 64-bit words per row of K1's row-padded bit image (ceil(cpr / 4) + 1 zero word)
    uint32_t rcap;             // runs per frame (scratch)
    uint32_t lines_cap, px_cap, blobs_cap, crops_cap;   // batch arenas
    uint32_t px_frame_cap, max_crops;
    int crop_w, crop_h, crop_method;
    float crop_scale;          // individual_image_scale; != 1: crops are rendered by crop_scale_kernel after K3 (K3 only assigns the crop slots)
    int crop_norm;             // 1: `moments` normalisation, rendered by crop_norm.cu after K3 (K3 only assigns the crop slots)
    float sqcm; int n_ranges; double lo[4], hi[4];
    const uint8_t *bg;
    size_t bg_stride;          // 0: one background for all frames; else `bg` holds one mask image per frame (morphology path)
    const uint8_t *keep_mask;  // optional [B][H][W] 0xFF/0 image ANDed with the foreground (tracker-side re-threshold), or null
    uint32_t min_payload;      // tracker-side re-threshold: blobs whose payload is not MORE than this many bytes are dropped (pixel::threshold_blob,
                               // C/processing/PixelTree.cpp:350-352: `pixels->size() > 1` -- a lone grey pixel goes, a lone rgb8 pixel stays); 0 elsewhere
    // colour inputs (T/python/BackgroundSubtraction.cpp:151-188): frames are [B][H][W][CN] interleaved B,G,R(,A)
    int CN;                    // bytes per pixel of the submitted frames: 1, 3, 4
    int enc;                   // 0 gray (1 byte per blob pixel), 1 rgb8 (B,G,R per blob pixel)
    int cc;                    // color_channel (gray encoding): -1 = cvtColor, else the plane to take
    int opx;                   // bytes per blob pixel: enc ? 3 : 1
    int cpx;                   // bytes per crop pixel: 3 for rgb8 and r3g3b2 (imageFromLines renders r3g3b2 blobs as B,G,R), else 1
    int r3;                    // meta_encoding r3g3b2: frames become 1-byte codes (convert_to_r3g3b2, cc = CC_R3G3B2); enc stays 0
    const uint8_t *bg3;        // rgb8: the 3-channel background (crops difference against it); `bg` is its grey image
    const uint8_t *nz_plane;   // rgb8 on the plane path: [B][H][W] "any of B,G,R != 0" bytes (0xFF / 0), or null
    // K1 outputs
    uint32_t *run_count;       // [B]
    uint32_t *band_base, *band_cnt;   // [B][n_bands]
    tb_line *runs_raw;         // [B][rcap]
    // K2 scratch
    tb_line *runs;             // [B][rcap] raster order
    uint32_t *parent;          // [B][rcap] -> final label (root run index)
    uint32_t *bidx;            // [B][rcap] blob index of a root run
    uint32_t *row_start, *row_end;    // [B][H]
    uint32_t *b_npx, *b_nl, *b_xmin, *b_xmax, *b_ymax, *b_root, *b_loff, *b_poff, *kept; // [B][rcap]
    uint32_t *frame_tot;       // [B][4]: kept blobs, lines, pixels, status
    // K3 outputs
    tb_frame_info *infos;      // [B]
    tb_blob_rec *recs;         // [blobs_cap]
    tb_line *lines;            // [lines_cap]
    uint32_t *line_px;         // [lines_cap] pixel offset of each line
    uint8_t *pixels;           // [px_cap]
    uint8_t *crops;            // [crops_cap][crop_h*crop_w]
    uint32_t *crop_blob;       // [crops_cap]
    uint32_t *totals;          // [4]: blobs, lines, pixels, crops
};

// ------------------------------------------------------------------------------------------------
// K1: fused difference / threshold / mask / run-length extraction.  HBM-bound: every frame byte is
// read exactly once with coalesced 16-byte loads; the background tile is loaded once per CTA and
// reused for `fpc` frames.  A CTA owns a band of whole rows (<= 1024 16-pixel chunks).  Each thread
// turns 4 chunks into 16-bit foreground masks (striped, for coalescing), the masks are transposed
// through shared memory so that every thread then owns 64 consecutive pixels as one 64-bit word, in
// which run starts / ends are single-bit events.  They are ranked with one warp scan + one block scan
// and written straight into the frame's run array (one atomicAdd per band and frame).
// ------------------------------------------------------------------------------------------------
constexpr int K1_NT = 256, K1_KPT = 4, K1_CHUNKS = K1_NT * K1_KPT;

// the threshold mask of 4 pixels (0xFF / 0 bytes), before it is ANDed with the input
template <bool GENERIC>
__device__ __forceinline__ uint32_t mask4(uint32_t f, uint32_t b, const SegK &p)
{
    if (!GENERIC) return __vcmpgtu4(__vabsdiffu4(f, b), p.t4);      // default settings: |f-b| > T
    if (p.flags & F_PREMASK) return b;                              // b = mask bytes after morphology
    uint32_t in = (p.flags & F_INV) ? ~f : f;                       // 255 - x
    uint32_t d = in;
    if (p.flags & F_DIFF) d = (p.flags & F_ABS) ? __vabsdiffu4(in, b) : __vsubus4(b, in);
    uint32_t m = (p.flags & F_RANGE) ? (__vcmpgeu4(d, p.lo4) & __vcmpleu4(d, p.hi4))
               : (p.flags & F_GE) ? __vcmpgeu4(d, p.t4) : __vcmpgtu4(d, p.t4);     // F_GE: tracker-side comparison >=
    if (p.flags & F_INVMASK) m = ~m;
    return m;
}
template <bool GENERIC>
__device__ __forceinline__ uint32_t fg4(uint32_t f, uint32_t b, const SegK &p)
{
    return mask4<GENERIC>(f, b, p) & __vcmpne4(f, 0u);              // (mask & input) != 0
}
__device__ __forceinline__ uint32_t pack4(uint32_t m)               // 4 byte masks -> 4 bits
{
    return ((m & 0x08040201u) * 0x01010101u) >> 24;
}
template <bool GENERIC>
__device__ __forceinline__ uint32_t fg16(const uint4 &f, const uint4 &b, const SegK &p)
{
    return pack4(fg4<GENERIC>(f.x, b.x, p)) | (pack4(fg4<GENERIC>(f.y, b.y, p)) << 4) |
           (pack4(fg4<GENERIC>(f.z, b.z, p)) << 8) | (pack4(fg4<GENERIC>(f.w, b.w, p)) << 12);
}
template <bool GENERIC>
__device__ __forceinline__ uint32_t mask16(const uint4 &f, const uint4 &b, const SegK &p)
{
    return pack4(mask4<GENERIC>(f.x, b.x, p)) | (pack4(mask4<GENERIC>(f.y, b.y, p)) << 4) |
           (pack4(mask4<GENERIC>(f.z, b.z, p)) << 8) | (pack4(mask4<GENERIC>(f.w, b.w, p)) << 12);
}
// Fast path (default settings and |T| <= 127): per-byte flags in bit 7 by SWAR arithmetic, then the 16
// flags are gathered with four dp4a (bytes 0x80 times weights 1,2,4,8 / 16,32,64,128).
__device__ __forceinline__ uint32_t fgflag_fast(uint32_t f, uint32_t b, uint32_t kadd)
{
    const uint32_t d = __vabsdiffu4(f, b);
    const uint32_t t = (d & 0x7f7f7f7fu) + kadd;               // bit 7: (d & 127) > T
    const uint32_t u = (f & 0x7f7f7f7fu) + 0x7f7f7f7fu;        // bit 7: (f & 127) != 0
    return (t | d) & (u | f) & 0x80808080u;
}
__device__ __forceinline__ uint32_t fg16_fast(const uint4 &f, const uint4 &b, uint32_t kadd)
{
    uint32_t lo = __dp4a(fgflag_fast(f.x, b.x, kadd), 0x08040201u, 0u);
    lo = __dp4a(fgflag_fast(f.y, b.y, kadd), 0x80402010u, lo);
    uint32_t hi = __dp4a(fgflag_fast(f.z, b.z, kadd), 0x08040201u, 0u);
    hi = __dp4a(fgflag_fast(f.w, b.w, kadd), 0x80402010u, hi);
    return (lo >> 7) | ((hi >> 7) << 8);
}
// the same without the "grey != 0" test (rgb8: the non-zero test is on B|G|R)
__device__ __forceinline__ uint32_t fgflag_fast_nonz(uint32_t f, uint32_t b, uint32_t kadd)
{
    const uint32_t d = __vabsdiffu4(f, b);
    return (((d & 0x7f7f7f7fu) + kadd) | d) & 0x80808080u;
}
__device__ __forceinline__ uint32_t fg16_fast_nonz(const uint4 &f, const uint4 &b, uint32_t kadd)
{
    uint32_t lo = __dp4a(fgflag_fast_nonz(f.x, b.x, kadd), 0x08040201u, 0u);
    lo = __dp4a(fgflag_fast_nonz(f.y, b.y, kadd), 0x80402010u, lo);
    uint32_t hi = __dp4a(fgflag_fast_nonz(f.z, b.z, kadd), 0x08040201u, 0u);
    hi = __dp4a(fgflag_fast_nonz(f.w, b.w, kadd), 0x80402010u, hi);
    return (lo >> 7) | ((hi >> 7) << 8);
}
__device__ __forceinline__ uint4 ld_stream(const uint4 *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
// 16 pixels starting at column col*16 of a row that may be shorter / unaligned (generic widths)
__device__ __forceinline__ uint4 ld_edge(const uint8_t *row, int col, int W)
{
    uint32_t v[4] = {0, 0, 0, 0};
    int x0 = col * 16, n = min(16, W - x0);
    for (int i = 0; i < n; ++i) v[i >> 2] |= (uint32_t)row[x0 + i] << (8 * (i & 3));
    return make_uint4(v[0], v[1], v[2], v[3]);
}

// cv::cvtColor(BGR[A]2GRAY) on 8-bit data (OpenCV's fixed point, T/python/BackgroundSubtraction.cpp:165-169,
// C/processing/RawProcessing.cpp:358): Y = (3735 B + 19235 G + 9798 R + 16384) >> 15.  p = B | G << 8 | R << 16 | x << 24;
// the 16-bit weights are split into byte halves so that two dp4a do the three multiplications (byte 3 has weight 0).
__device__ __forceinline__ uint32_t gray_of(uint32_t p)
{
    const uint32_t t = __dp4a(p, 0x00462397u, 16384u);          // low bytes  151, 35, 70
    const uint32_t s = __dp4a(p, 0x00264B0Eu, 0u);              // high bytes  14, 75, 38
    return ((s << 8) + t) >> 15;
}
// cmn::bgr2gray (C/processing/Background.h:76-81), the tracker side's grey value of a B,G,R triple:
// saturate(float(B) * 0.114 + float(G) * 0.587 + float(R) * 0.299 + 0.5, 0, 255), evaluated in double without
// contraction (explicit round-to-nearest multiplies and adds), truncated.  Differs from cvtColor on ~0.14 % of triples.
__device__ __forceinline__ uint32_t gray_px_tracker(const uint8_t *q)
{
    double v = __dmul_rn((double)q[0], 0.114);
    v = __dadd_rn(v, __dmul_rn((double)q[1], 0.587));
    v = __dadd_rn(v, __dmul_rn((double)q[2], 0.299));
    v = __dadd_rn(v, 0.5);
    return (uint32_t)fmin(fmax(v, 0.0), 255.0);
}
constexpr int CC_R3G3B2 = 8;      // SegDev.cc value: the "plane" of a colour frame is its r3g3b2 code image
// vec_to_r3g3b2 / r3g3b2_to_vec, C/misc/detail.h:508-531
__device__ __forceinline__ uint32_t to_r3g3b2(uint32_t b, uint32_t g, uint32_t r) { return ((b >> 6) << 6) | ((g >> 5) << 3) | (r >> 5); }
__device__ __forceinline__ int r3g3b2_channel(uint32_t code, int k) { return k == 0 ? (int)((code >> 6) & 3u) * 64 : (k == 1 ? (int)((code >> 3) & 7u) * 32 : (int)(code & 7u) * 32); }
__device__ __forceinline__ uint32_t gray_px(const uint8_t *q, int CN, int cc)
{
    if (CN == 1) return q[0];
    if (cc == CC_R3G3B2) return to_r3g3b2(q[0], q[1], q[2]);
    if (cc >= 0) return q[cc];
    return gray_of((uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16));
}
// 4 consecutive pixels of an interleaved colour row -> 4 grey bytes; a, b, c = the 12 (CN 3) or a, b, c, e = the 16
// (CN 4) bytes.  nz4 (RGB only): bit i set when any of B,G,R of pixel i is non-zero (Source.cpp:200-206).
template <int CN, bool RGB>
__device__ __forceinline__ uint32_t gray4(uint32_t a, uint32_t b, uint32_t c, uint32_t e, uint32_t &nz4)
{
    uint32_t p0, p1, p2, p3;
    if (CN == 3) { p0 = a; p1 = __byte_perm(a, b, 0x0543); p2 = __byte_perm(b, c, 0x0432); p3 = c >> 8; }
    else { p0 = a; p1 = b; p2 = c; p3 = e; }
    if (RGB) {
        nz4 = (((p0 & 0xFFFFFFu) + 0xFFFFFFu) >> 24) | ((((p1 & 0xFFFFFFu) + 0xFFFFFFu) >> 24) << 1) |
              ((((p2 & 0xFFFFFFu) + 0xFFFFFFu) >> 24) << 2) | ((((p3 & 0xFFFFFFu) + 0xFFFFFFu) >> 24) << 3);
    }
    return gray_of(p0) | (gray_of(p1) << 8) | (gray_of(p2) << 16) | (gray_of(p3) << 24);
}

template <bool GENERIC>
__global__ void __launch_bounds__(K1_NT, 4)
seg_rle_kernel(const uint8_t *__restrict__ frames, SegDev d, SegK p, int fpc)
{
    __shared__ __align__(16) uint16_t s_mask[K1_CHUNKS + 8];
    __shared__ uint32_t s_wtot[K1_NT / 32];
    __shared__ uint32_t s_base;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int band = blockIdx.x;
    const int row0 = band * d.rpt;
    const int tile_rows = min(d.rpt, d.H - row0);
    const int tile_chunks = tile_rows * d.cpr;
    const size_t frame_bytes = (size_t)d.W * d.H;

    // load role: chunks warp*128 + k*32 + lane
    uint4 bgc[K1_KPT];
#pragma unroll
    for (int k = 0; k < K1_KPT; ++k) {
        const int c = warp * (32 * K1_KPT) + k * 32 + lane;
        bgc[k] = make_uint4(0, 0, 0, 0);
        if (c < tile_chunks) {
            if (d.aligned) bgc[k] = *reinterpret_cast<const uint4 *>(d.bg + (size_t)row0 * d.W + (size_t)c * 16);
            else bgc[k] = ld_edge(d.bg + (size_t)(row0 + c / d.cpr) * d.W, c % d.cpr, d.W);
        }
    }
    // run role: chunks 4*tid .. 4*tid+3 = 64 consecutive pixels; row boundaries inside the word
    const int g0 = 4 * tid;
    const int gr0 = g0 / d.cpr, gc0 = g0 % d.cpr;
    uint64_t RS = 0, RE = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int col = (g0 + i) % d.cpr;
        if (col == 0) RS |= 1ull << (16 * i);
        if (col == d.cpr - 1) RE |= 1ull << (16 * i + 15);
    }
    if (tid == 0) for (int i = 0; i < 8; ++i) s_mask[K1_CHUNKS + i] = 0;

    const int f0 = blockIdx.y * fpc, f1 = min(d.B, f0 + fpc);
    for (int f = f0; f < f1; ++f) {
        const uint8_t *fb = frames + (size_t)f * frame_bytes;
        if (d.bg_stride) {                                      // per-frame mask image instead of the background
            const uint8_t *mb = d.bg + (size_t)f * d.bg_stride;
#pragma unroll
            for (int k = 0; k < K1_KPT; ++k) {
                const int c = warp * (32 * K1_KPT) + k * 32 + lane;
                if (c < tile_chunks) {
                    if (d.aligned) bgc[k] = *reinterpret_cast<const uint4 *>(mb + (size_t)row0 * d.W + (size_t)c * 16);
                    else bgc[k] = ld_edge(mb + (size_t)(row0 + c / d.cpr) * d.W, c % d.cpr, d.W);
                }
            }
        }
        uint4 cur[K1_KPT];
#pragma unroll
        for (int k = 0; k < K1_KPT; ++k) {
            const int c = warp * (32 * K1_KPT) + k * 32 + lane;
            cur[k] = make_uint4(0, 0, 0, 0);
            if (c < tile_chunks) {
                if (d.aligned) cur[k] = ld_stream(reinterpret_cast<const uint4 *>(fb + (size_t)row0 * d.W) + c);
                else cur[k] = ld_edge(fb + (size_t)(row0 + c / d.cpr) * d.W, c % d.cpr, d.W);
            }
        }
#pragma unroll
        for (int k = 0; k < K1_KPT; ++k) {
            const int c = warp * (32 * K1_KPT) + k * 32 + lane;
            uint32_t m16 = (c < tile_chunks) ? fg16<GENERIC>(cur[k], bgc[k], p) : 0u;
            if (d.nz_plane && c < tile_chunks) {                // rgb8: the pixel is set when any of B,G,R is, not the grey value
                const uint8_t *nb = d.nz_plane + (size_t)f * frame_bytes;
                const uint4 nz = d.aligned ? *reinterpret_cast<const uint4 *>(nb + (size_t)row0 * d.W + (size_t)c * 16)
                                           : ld_edge(nb + (size_t)(row0 + c / d.cpr) * d.W, c % d.cpr, d.W);
                m16 = mask16<GENERIC>(cur[k], bgc[k], p) & (pack4(nz.x) | (pack4(nz.y) << 4) | (pack4(nz.z) << 8) | (pack4(nz.w) << 12));
            }
            if (d.keep_mask && c < tile_chunks) {               // pixels outside the kept detection blobs are never foreground
                const uint8_t *kb = d.keep_mask + (size_t)f * frame_bytes;
                const uint4 mk = d.aligned ? *reinterpret_cast<const uint4 *>(kb + (size_t)row0 * d.W + (size_t)c * 16)
                                           : ld_edge(kb + (size_t)(row0 + c / d.cpr) * d.W, c % d.cpr, d.W);
                m16 &= pack4(mk.x) | (pack4(mk.y) << 4) | (pack4(mk.z) << 8) | (pack4(mk.w) << 12);
            }
            s_mask[c] = (uint16_t)m16;
        }
        __syncthreads();
        const uint64_t M = *reinterpret_cast<const uint64_t *>(s_mask + g0);
        const uint64_t prev = g0 > 0 ? (uint64_t)(s_mask[g0 - 1] >> 15) : 0ull;
        const uint64_t next = (uint64_t)(s_mask[g0 + 4] & 1u);
        uint64_t st = M & ~(((M << 1) | prev) & ~RS);
        uint64_t en = M & ~(((M >> 1) | (next << 63)) & ~RE);
        uint32_t run, ex = warp_excl_scan((uint32_t)__popcll(st) | ((uint32_t)__popcll(en) << 16), run);
        if (lane == 0) s_wtot[warp] = run;
        __syncthreads();
        uint32_t wbase = 0, total = 0;
#pragma unroll
        for (int w = 0; w < K1_NT / 32; ++w) {
            uint32_t t = s_wtot[w];
            if (w < warp) wbase += t;
            total += t;
        }
        if (tid == 0) {
            uint32_t ns = total & 0xFFFFu;
            uint32_t base = ns ? atomicAdd(&d.run_count[f], ns) : 0u;
            s_base = base;
            d.band_base[(size_t)f * d.n_bands + band] = base;
            d.band_cnt[(size_t)f * d.n_bands + band] = ns;
        }
        __syncthreads();
        if (st | en) {
            uint16_t *out = reinterpret_cast<uint16_t *>(d.runs_raw + (size_t)f * d.rcap);
            uint32_t os = s_base + ((wbase + ex) & 0xFFFFu), oe = s_base + ((wbase + ex) >> 16);
            while (st) {
                const int b = __ffsll((long long)st) - 1; st &= st - 1;
                int col = gc0 + (b >> 4), row = gr0;
                while (col >= d.cpr) { col -= d.cpr; ++row; }
                if (os < d.rcap) {
                    out[(size_t)os * 4 + 0] = (uint16_t)(col * 16 + (b & 15));
                    *reinterpret_cast<uint32_t *>(out + (size_t)os * 4 + 2) = (uint32_t)(row0 + row);   // y, pad = 0
                }
                ++os;
            }
            while (en) {
                const int b = __ffsll((long long)en) - 1; en &= en - 1;
                int col = gc0 + (b >> 4);
                while (col >= d.cpr) col -= d.cpr;
                if (oe < d.rcap) out[(size_t)oe * 4 + 1] = (uint16_t)(col * 16 + (b & 15));
                ++oe;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K1, bulk-copy variant (frame widths that are multiples of 16): the band's tile of every frame is a
// contiguous <= 16 KB range, so a producer warp streams it into a 3-stage shared-memory ring with
// cp.async.bulk (TMA engine, mbarrier completion) while 8 consumer warps pull their chunks out of the
// ring into registers, release the stage at once and run the same mask / 64-bit run-word / ranking code.
// The memory pipeline depth no longer depends on registers or occupancy.
// ------------------------------------------------------------------------------------------------
constexpr int K1T_STAGES = 3, K1T_NT = K1_NT + 32, K1T_SMEM = K1T_STAGES * K1_CHUNKS * 16;

template <bool GENERIC>
__global__ void __launch_bounds__(K1T_NT, 4)
seg_rle_tma_kernel(const uint8_t *__restrict__ frames, SegDev d, SegK p, int fpc)
{
    extern __shared__ __align__(128) uint8_t k1t_dsm[];
    uint4 (*s_tile)[K1_CHUNKS] = reinterpret_cast<uint4 (*)[K1_CHUNKS]>(k1t_dsm);
    __shared__ __align__(16) uint16_t s_mask[K1_CHUNKS + 8];
    __shared__ uint32_t s_wtot[K1_NT / 32];
    __shared__ uint32_t s_base;
    __shared__ uint64_t bar_full[K1T_STAGES], bar_empty[K1T_STAGES];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int band = blockIdx.x;
    const int row0 = band * d.rpt;
    const int tile_rows = min(d.rpt, d.H - row0);
    const int tile_chunks = tile_rows * d.cpr;
    const size_t frame_bytes = (size_t)d.W * d.H;
    const int f0 = blockIdx.y * fpc, f1 = min(d.B, f0 + fpc);

    if (tid == 0) {
        for (int i = 0; i < K1T_STAGES; ++i) { umma::mbar_init(&bar_full[i], 1); umma::mbar_init(&bar_empty[i], K1_NT / 32); }
        umma::fence_mbar_init();
        for (int i = 0; i < 8; ++i) s_mask[K1_CHUNKS + i] = 0;
    }
    __syncthreads();

    if (warp == K1_NT / 32) {                      // producer warp
        if (lane == 0) {
            const uint32_t bytes = (uint32_t)tile_chunks * 16u;
            for (int f = f0, it = 0; f < f1; ++f, ++it) {
                const int s = it % K1T_STAGES;
                umma::mbar_wait(&bar_empty[s], ((it / K1T_STAGES) & 1) ^ 1);
                umma::mbar_expect_tx(&bar_full[s], bytes);
                umma::bulk_g2s(&s_tile[s][0], frames + (size_t)f * frame_bytes + (size_t)row0 * d.W, bytes, &bar_full[s]);
            }
        }
        return;
    }

    uint4 bgc[K1_KPT];
#pragma unroll
    for (int k = 0; k < K1_KPT; ++k) {
        const int c = warp * (32 * K1_KPT) + k * 32 + lane;
        bgc[k] = make_uint4(0, 0, 0, 0);
        if (c < tile_chunks) bgc[k] = *reinterpret_cast<const uint4 *>(d.bg + (size_t)row0 * d.W + (size_t)c * 16);
    }
    const int g0 = 4 * tid;
    const int gr0 = g0 / d.cpr, gc0 = g0 % d.cpr;
    uint64_t RS = 0, RE = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int col = (g0 + i) % d.cpr;
        if (col == 0) RS |= 1ull << (16 * i);
        if (col == d.cpr - 1) RE |= 1ull << (16 * i + 15);
    }

    // pixel coordinates of the 64-pixel word: at most one row break when a row holds >= 4 chunks
    const bool single_break = d.cpr >= 4;
    const int x_first = gc0 * 16, y_first = row0 + gr0;
    const int brk = min(64, (d.cpr - gc0) * 16);                 // bit index at which the next row starts
    for (int f = f0, it = 0; f < f1; ++f, ++it) {
        const int s = it % K1T_STAGES;
        umma::mbar_wait(&bar_full[s], (it / K1T_STAGES) & 1);
        uint4 cur[K1_KPT];
#pragma unroll
        for (int k = 0; k < K1_KPT; ++k) {
            const int c = warp * (32 * K1_KPT) + k * 32 + lane;
            cur[k] = (c < tile_chunks) ? s_tile[s][c] : make_uint4(0, 0, 0, 0);
        }
        __syncwarp();
        if (lane == 0) umma::mbar_arrive(&bar_empty[s]);          // stage is in registers: refill it
#pragma unroll
        for (int k = 0; k < K1_KPT; ++k) {
            const int c = warp * (32 * K1_KPT) + k * 32 + lane;
            uint32_t m16 = 0;
            if (c < tile_chunks) m16 = GENERIC ? fg16<true>(cur[k], bgc[k], p) : fg16_fast(cur[k], bgc[k], p.lo4);
            s_mask[c] = (uint16_t)m16;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const uint64_t M = *reinterpret_cast<const uint64_t *>(s_mask + g0);
        const uint64_t prev = g0 > 0 ? (uint64_t)(s_mask[g0 - 1] >> 15) : 0ull;
        const uint64_t next = (uint64_t)(s_mask[g0 + 4] & 1u);
        uint64_t st = M & ~(((M << 1) | prev) & ~RS);
        uint64_t en = M & ~(((M >> 1) | (next << 63)) & ~RE);
        uint32_t run, ex = warp_excl_scan((uint32_t)__popcll(st) | ((uint32_t)__popcll(en) << 16), run);
        if (lane == 0) s_wtot[warp] = run;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        uint32_t wbase = 0, total = 0;
#pragma unroll
        for (int w = 0; w < K1_NT / 32; ++w) {
            uint32_t t = s_wtot[w];
            if (w < warp) wbase += t;
            total += t;
        }
        if (tid == 0) {
            uint32_t ns = total & 0xFFFFu;
            uint32_t base = ns ? atomicAdd(&d.run_count[f], ns) : 0u;
            s_base = base;
            d.band_base[(size_t)f * d.n_bands + band] = base;
            d.band_cnt[(size_t)f * d.n_bands + band] = ns;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (st | en) {
            uint16_t *out = reinterpret_cast<uint16_t *>(d.runs_raw + (size_t)f * d.rcap);
            uint32_t os = s_base + ((wbase + ex) & 0xFFFFu), oe = s_base + ((wbase + ex) >> 16);
            while (st) {
                const int b = __ffsll((long long)st) - 1; st &= st - 1;
                int x, y;
                if (single_break) { const bool nx = b >= brk; x = nx ? b - brk : x_first + b; y = y_first + (nx ? 1 : 0); }
                else { int col = gc0 + (b >> 4), row = gr0; while (col >= d.cpr) { col -= d.cpr; ++row; } x = col * 16 + (b & 15); y = row0 + row; }
                if (os < d.rcap) {
                    out[(size_t)os * 4 + 0] = (uint16_t)x;
                    *reinterpret_cast<uint32_t *>(out + (size_t)os * 4 + 2) = (uint32_t)y;
                }
                ++os;
            }
            while (en) {
                const int b = __ffsll((long long)en) - 1; en &= en - 1;
                int x;
                if (single_break) x = b >= brk ? b - brk : x_first + b;
                else { int col = gc0 + (b >> 4); while (col >= d.cpr) col -= d.cpr; x = col * 16 + (b & 15); }
                if (oe < d.rcap) out[(size_t)oe * 4 + 1] = (uint16_t)x;
                ++oe;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K1, warp-specialised persistent variant (the default for widths that are multiples of 16).
// The batch is cut into units (band of rows, frame) in band-major order.  A grid of (SMs x resident CTAs)
// persistent CTAs first works through a static contiguous share of the units (so the background tile,
// kept in registers, changes only a few times per CTA) and then takes the remaining units one at a time
// from a global counter, which absorbs the speed differences between SMs.  Roles inside a CTA:
//   * 1 producer warp: picks the units and streams their tiles into a shared-memory ring with
//     cp.async.bulk (TMA engine); the unit's (band, frame) travels with the stage;
//   * 8 mask warps: 16 pixels -> 16 foreground bits (SWAR + dp4a), written into a row-padded bit image
//     (every image row starts on a 64-bit word and is followed by at least one zero word, so a run never
//     crosses a word boundary between rows and no row-break masks are needed);
//   * EW extraction warps, taking units round-robin: each lane owns 8 consecutive 64-bit words of the
//     bit image (512 pixels), counts run starts with 8 popcounts, ranks them with one warp scan, reserves
//     space in the frame's run array with one atomicAdd per unit and writes x0,y / x1.
// Mask warps never wait for the extraction of the same unit (ring of EW bit images), all waits are
// hardware-suspended mbarrier waits, and per byte the kernel issues about half the instructions of the
// variant above, where every thread ranks its own word behind three block barriers.
// ------------------------------------------------------------------------------------------------
constexpr int K1W_MW = 8, K1W_STAGES = 3;
constexpr int K1W_STAGE_BYTES = K1_CHUNKS * 16;
constexpr int K1W_WORDS = 256;                              // 64-bit words of the padded bit image (32 lanes x 8)
constexpr int K1W_MASK_BYTES = (K1W_WORDS / 8 + 1) * 80 + 128;   // per lane 64 B (8 words) + 16 B pad (bank spread); 1 spare zero block; 128 B dump area
constexpr uint32_t K1W_DONE = 0xFFFFFFFFu;
constexpr int k1w_smem(int ew, int cn = 1) { return K1W_STAGES * (cn == 1 ? K1W_STAGE_BYTES : 512 * 16 * cn) + ew * K1W_MASK_BYTES; }

__device__ __forceinline__ uint4 lds_v4(uint32_t a)
{
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a));
    return r;
}
__device__ __forceinline__ uint64_t lds_u64(uint32_t a) { uint64_t r; asm volatile("ld.shared.u64 %0, [%1];" : "=l"(r) : "r"(a)); return r; }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t r; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(a)); return r; }
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" :: "r"(a), "h"((unsigned short)v) : "memory"); }

// Colour frames (CN = 3 BGR / 4 BGRA, fused cvtColor): a unit holds half as many pixels (KPT = 2 chunks per mask
// thread) so that three stages of CN x 8 KB still leave room for 2-3 CTAs per SM; the mask warps turn every 4 pixels
// into 4 grey bytes with dp4a (gray4) and continue as for grey frames.  RGB (rgb8 encoding): a pixel is foreground
// when the grey difference exceeds T and any of B,G,R is non-zero.
template <bool GENERIC, int EW, int CN = 1, bool RGB = false>
__global__ void __launch_bounds__((K1W_MW + EW + 1) * 32, CN == 1 ? 3 : 2)
seg_rle_ws_kernel(const uint8_t *__restrict__ frames, SegDev d, SegK p, uint32_t static_units, unsigned long long *dbg)
{
    static_assert(CN == 1 || !GENERIC, "colour frames with non-default settings take the plane path");
    constexpr int NT = (K1W_MW + EW + 1) * 32;
    constexpr int KPT = CN == 1 ? K1_KPT : 2;                       // chunks per mask thread and unit
    constexpr int STAGE_BYTES = K1W_MW * 32 * KPT * 16 * CN;
    extern __shared__ __align__(128) uint8_t k1w_dsm[];
    __shared__ uint64_t s_bar[2 * K1W_STAGES + 2 * EW];
    __shared__ uint32_t s_meta[K1W_STAGES + EW];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tile_a = umma::smem_u32(k1w_dsm), mask_a = tile_a + K1W_STAGES * STAGE_BYTES;
    const uint32_t full_a = umma::smem_u32(s_bar), empty_a = full_a + 8 * K1W_STAGES;
    const uint32_t mkf_a = empty_a + 8 * K1W_STAGES, mke_a = mkf_a + 8 * EW;
    const uint32_t meta_a = umma::smem_u32(s_meta), mmeta_a = meta_a + 4 * K1W_STAGES;

    if (tid == 0) {
        for (int i = 0; i < K1W_STAGES; ++i) { umma::mbar_init(&s_bar[i], 1); umma::mbar_init(&s_bar[K1W_STAGES + i], K1W_MW); }
        for (int i = 0; i < EW; ++i) { umma::mbar_init(&s_bar[2 * K1W_STAGES + i], K1W_MW); umma::mbar_init(&s_bar[2 * K1W_STAGES + EW + i], 1); }
        umma::fence_mbar_init();
    }
    for (int i = tid; i < EW * K1W_MASK_BYTES / 4; i += NT) sts_u32(mask_a + 4 * i, 0u);
    __syncthreads();

    if (warp == K1W_MW + EW) {                     // ---- producer warp ----
        if (lane != 0) return;
        if (dbg) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); dbg[blockIdx.x * 4 + 0] = t; }
        const uint32_t B = (uint32_t)d.B, U = (uint32_t)d.n_bands * B;
        const size_t frame_bytes = (size_t)d.W * d.H * CN;
        uint32_t s = 0, ph = 0;
        auto issue = [&](uint32_t band, uint32_t f) {
            const int row0 = (int)band * d.rpt;
            const uint32_t bytes = (uint32_t)min(d.rpt, d.H - row0) * (uint32_t)d.W * CN;
            umma::mbar_wait_suspend_a(empty_a + 8 * s, ph ^ 1);
            sts_u32(meta_a + 4 * s, (band << 16) | f);
            umma::mbar_expect_tx_a(full_a + 8 * s, bytes);
            umma::bulk_g2s_a(tile_a + s * STAGE_BYTES, frames + (size_t)f * frame_bytes + (size_t)row0 * d.W * CN, bytes, full_a + 8 * s);
            if (++s == K1W_STAGES) { s = 0; ph ^= 1; }
        };
        {                                          // static share: units [i * static_units, (i + 1) * static_units)
            const uint32_t u0 = blockIdx.x * static_units;
            uint32_t band = u0 / B, f = u0 % B;
            for (uint32_t i = 0; i < static_units; ++i) {
                issue(band, f);
                if (++f == B) { f = 0; ++band; }
            }
        }
        uint32_t *counter = d.run_count + d.B;     // zeroed with run_count before every launch
        const uint32_t dyn0 = gridDim.x * static_units;
        for (;;) {                                 // a unit is taken only once a stage is free for it, so that no CTA sits on
            umma::mbar_wait_suspend_a(empty_a + 8 * s, ph ^ 1);      // queued work while others run dry at the end
            const uint32_t u = dyn0 + atomicAdd(counter, 1u);
            if (u >= U) break;
            issue(u / B, u % B);
        }
        umma::mbar_wait_suspend_a(empty_a + 8 * s, ph ^ 1);
        sts_u32(meta_a + 4 * s, K1W_DONE);
        umma::mbar_arrive_a(full_a + 8 * s);
        if (dbg) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); dbg[blockIdx.x * 4 + 1] = t; }
        return;
    }

    if (warp < K1W_MW) {                           // ---- mask warps ----
        const int full_chunks = d.rpt * d.cpr;
        const int cb = warp * (32 * KPT) + lane;                      // chunks cb + 32 k
        uint32_t mpos[KPT];                                           // byte offset of the chunk's 16 flags inside a bit image
#pragma unroll
        for (int k = 0; k < KPT; ++k) {
            const int c = cb + 32 * k;
            mpos[k] = (uint32_t)(K1W_MASK_BYTES - 128 + 2 * lane);    // chunks beyond the band: dump area
            if (c < full_chunks) {
                const int q = (c / d.cpr) * d.wpr * 4 + c % d.cpr;    // chunk position in the row-padded bit image
                mpos[k] = (uint32_t)((q >> 5) * 80 + (q & 31) * 2);
            }
        }
        const uint32_t my_tile = tile_a + (uint32_t)cb * (16u * CN);
        uint4 bgc[KPT];
        uint32_t cur_band = 0xFFFFFFFFu, s = 0, ph = 0, b = 0, bph = 0;
        int tile_chunks = 0;
        for (;;) {
            umma::mbar_wait_suspend_a(full_a + 8 * s, ph);
            const uint32_t meta = lds_u32(meta_a + 4 * s);
            if (meta == K1W_DONE) break;
            const uint32_t band = meta >> 16;
            if (band != cur_band) {                                   // new band: background tile into registers (L2-resident)
                cur_band = band;
                const int row0 = (int)band * d.rpt;
                tile_chunks = min(d.rpt, d.H - row0) * d.cpr;
                const uint4 *bgp = reinterpret_cast<const uint4 *>(d.bg + (size_t)row0 * d.W) + cb;
#pragma unroll
                for (int k = 0; k < KPT; ++k) bgc[k] = (cb + 32 * k < tile_chunks) ? bgp[32 * k] : make_uint4(0, 0, 0, 0);
            }
            uint32_t m16[KPT];
            if (CN == 1) {
                uint4 cur[KPT];
#pragma unroll
                for (int k = 0; k < KPT; ++k) cur[k] = lds_v4(my_tile + s * STAGE_BYTES + k * 512);
                __syncwarp();
                if (lane == 0) umma::mbar_arrive_a(empty_a + 8 * s);  // stage is in registers: refill it
#pragma unroll
                for (int k = 0; k < KPT; ++k) {
                    const uint32_t m = GENERIC ? fg16<true>(cur[k], bgc[k], p) : fg16_fast(cur[k], bgc[k], p.lo4);
                    m16[k] = (cb + 32 * k < tile_chunks) ? m : 0u;    // rows below the image (last band) are stale shared memory
                }
            } else {
                uint4 cur[KPT][CN];
#pragma unroll
                for (int k = 0; k < KPT; ++k)
#pragma unroll
                    for (int j = 0; j < CN; ++j) cur[k][j] = lds_v4(my_tile + s * STAGE_BYTES + k * (512 * CN) + j * 16);
                __syncwarp();
                if (lane == 0) umma::mbar_arrive_a(empty_a + 8 * s);
#pragma unroll
                for (int k = 0; k < KPT; ++k) {
                    const uint32_t *w = reinterpret_cast<const uint32_t *>(&cur[k][0]);
                    uint4 g; uint32_t nz = 0, n4 = 0;
                    if (CN == 3) {
                        g.x = gray4<3, RGB>(w[0], w[1], w[2], 0u, n4);  nz |= n4;
                        g.y = gray4<3, RGB>(w[3], w[4], w[5], 0u, n4);  nz |= n4 << 4;
                        g.z = gray4<3, RGB>(w[6], w[7], w[8], 0u, n4);  nz |= n4 << 8;
                        g.w = gray4<3, RGB>(w[9], w[10], w[11], 0u, n4); nz |= n4 << 12;
                    } else {
                        g.x = gray4<4, RGB>(w[0], w[1], w[2], w[3], n4);    nz |= n4;
                        g.y = gray4<4, RGB>(w[4], w[5], w[6], w[7], n4);    nz |= n4 << 4;
                        g.z = gray4<4, RGB>(w[8], w[9], w[10], w[11], n4);  nz |= n4 << 8;
                        g.w = gray4<4, RGB>(w[12], w[13], w[14], w[15], n4); nz |= n4 << 12;
                    }
                    const uint32_t m = RGB ? (fg16_fast_nonz(g, bgc[k], p.lo4) & nz) : fg16_fast(g, bgc[k], p.lo4);
                    m16[k] = (cb + 32 * k < tile_chunks) ? m : 0u;
                }
            }
            umma::mbar_wait_suspend_a(mke_a + 8 * b, bph ^ 1);
            const uint32_t mb = mask_a + b * K1W_MASK_BYTES;
#pragma unroll
            for (int k = 0; k < KPT; ++k) sts_u16(mb + mpos[k], m16[k]);
            if (tid == 0) sts_u32(mmeta_a + 4 * b, meta);
            __syncwarp();
            if (lane == 0) umma::mbar_arrive_a(mkf_a + 8 * b);
            if (++s == K1W_STAGES) { s = 0; ph ^= 1; }
            if (++b == EW) { b = 0; bph ^= 1; }
        }
        for (int i = 0; i < EW; ++i) {                                // tell every extraction warp that the CTA is done
            umma::mbar_wait_suspend_a(mke_a + 8 * b, bph ^ 1);
            if (tid == 0) sts_u32(mmeta_a + 4 * b, K1W_DONE);
            __syncwarp();
            if (lane == 0) umma::mbar_arrive_a(mkf_a + 8 * b);
            if (++b == EW) { b = 0; bph ^= 1; }
        }
        return;
    }

    // ---- extraction warps ----
    const int e = warp - K1W_MW;
    const uint32_t mb = mask_a + e * K1W_MASK_BYTES;
    const uint32_t wpr = (uint32_t)d.wpr;
    const uint32_t r0 = (uint32_t)(lane * 8) / wpr, c0 = (uint32_t)(lane * 8) % wpr;
    for (uint32_t ph = 0;; ph ^= 1) {
        umma::mbar_wait_suspend_a(mkf_a + 8 * e, ph);
        const uint32_t meta = lds_u32(mmeta_a + 4 * e);
        if (meta == K1W_DONE) break;
        const uint32_t band = meta >> 16, f = meta & 0xFFFFu;
        uint64_t M[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint4 q = lds_v4(mb + lane * 80 + j * 16);
            M[2 * j] = (uint64_t)q.x | ((uint64_t)q.y << 32);
            M[2 * j + 1] = (uint64_t)q.z | ((uint64_t)q.w << 32);
        }
        // starts / ends inside this lane's 512 pixels
        uint32_t up = __shfl_up_sync(0xffffffffu, (uint32_t)(M[7] >> 63), 1);
        uint32_t dn = __shfl_down_sync(0xffffffffu, (uint32_t)M[0] & 1u, 1);
        if (lane == 0) up = 0;
        if (lane == 31) dn = 0;
        uint32_t cs = 0, nz = 0;
        uint64_t prev = up;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            cs += (uint32_t)__popcll(M[j] & ~((M[j] << 1) | prev));
            prev = M[j] >> 63;
            nz |= (M[j] != 0ull ? 1u : 0u) << j;
        }
        const uint32_t ce = cs + (up & (uint32_t)M[0] & 1u) - ((uint32_t)(M[7] >> 63) & dn);
        uint32_t total;
        const uint32_t ex = warp_excl_scan(cs | (ce << 16), total);
        const uint32_t ns = total & 0xFFFFu;
        uint32_t base = 0;
        if (lane == 0) {
            if (ns) base = atomicAdd(&d.run_count[f], ns);
            d.band_base[(size_t)f * d.n_bands + band] = base;
            d.band_cnt[(size_t)f * d.n_bands + band] = ns;
        }
        if (ns) {
            base = __shfl_sync(0xffffffffu, base, 0);
            uint32_t os = base + (ex & 0xFFFFu), oe = base + (ex >> 16);
            uint16_t *out = reinterpret_cast<uint16_t *>(d.runs_raw + (size_t)f * d.rcap);
            const uint32_t y0 = band * (uint32_t)d.rpt;
            while (nz) {
                const uint32_t j = (uint32_t)__ffs((int)nz) - 1u; nz &= nz - 1u;
                const uint32_t w = (uint32_t)lane * 8u + j;
                const uint64_t Mw = lds_u64(mb + lane * 80 + j * 8);
                const uint64_t pw = w ? lds_u64(mb + ((w - 1) >> 3) * 80 + ((w - 1) & 7) * 8) : 0ull;
                const uint64_t nw = lds_u64(mb + ((w + 1) >> 3) * 80 + ((w + 1) & 7) * 8);
                uint64_t st = Mw & ~((Mw << 1) | (pw >> 63));
                uint64_t en = Mw & ~((Mw >> 1) | (nw << 63));
                uint32_t c = c0 + j, r = r0;
                while (c >= wpr) { c -= wpr; ++r; }
                const uint32_t xb = c * 64u, y = y0 + r;
                while (st) {
                    const uint32_t bit = (uint32_t)__ffsll((long long)st) - 1u; st &= st - 1;
                    if (os < d.rcap) {
                        out[(size_t)os * 4 + 0] = (uint16_t)(xb + bit);
                        *reinterpret_cast<uint32_t *>(out + (size_t)os * 4 + 2) = y;                 // y, pad = 0
                    }
                    ++os;
                }
                while (en) {
                    const uint32_t bit = (uint32_t)__ffsll((long long)en) - 1u; en &= en - 1;
                    if (oe < d.rcap) out[(size_t)oe * 4 + 1] = (uint16_t)(xb + bit);
                    ++oe;
                }
            }
        }
        __syncwarp();
        if (lane == 0) umma::mbar_arrive_a(mke_a + 8 * e);
    }
    if (dbg && e == 0 && lane == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); dbg[blockIdx.x * 4 + 2] = t; }
}

// Colour frames on the plane path (non-default settings, widths that are not multiples of 16, tracker-side
// re-threshold): the grey plane the threshold runs on (cvtColor or color_channel) and, for rgb8, the plane of
// "any of B,G,R != 0" bytes.  One thread per 4 pixels.
__global__ void to_gray_kernel(const uint8_t *__restrict__ frames, uint8_t *__restrict__ gray, uint8_t *__restrict__ nz,
                               size_t total, int CN, int cc, int tracker_formula = 0)
{
    for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < total; i += (size_t)gridDim.x * blockDim.x * 4) {
        uint32_t g = 0, z = 0;
        const int n = (int)min((size_t)4, total - i);
        for (int k = 0; k < n; ++k) {
            const uint8_t *q = frames + (i + k) * CN;
            g |= (tracker_formula ? gray_px_tracker(q) : gray_px(q, CN, cc)) << (8 * k);
            if (nz) z |= ((q[0] | q[1] | q[2]) ? 0xFFu : 0u) << (8 * k);
        }
        if (n == 4 && (i & 3) == 0) {
            *reinterpret_cast<uint32_t *>(gray + i) = g;
            if (nz) *reinterpret_cast<uint32_t *>(nz + i) = z;
        } else {
            for (int k = 0; k < n; ++k) { gray[i + k] = (uint8_t)(g >> (8 * k)); if (nz) nz[i + k] = (uint8_t)(z >> (8 * k)); }
        }
    }
}

// generate_binary's output for a colour frame, parity tests only: mask & grey (gray encoding) or mask & each of
// B,G,R (rgb8, RawProcessing.cpp:581-589).  mask: optional 0xFF/0 plane after morphology, else thresholded here.
__global__ void binary_image_color_kernel(const uint8_t *__restrict__ frame, const uint8_t *__restrict__ gray, const uint8_t *__restrict__ bg,
                                          const uint8_t *__restrict__ mask, uint8_t *__restrict__ out, size_t n, SegK p, int CN, int enc)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t g = gray[i];
        const uint32_t m = mask ? mask[i] : (mask4<true>(g, bg[i], p) & 0xFFu);
        if (!enc) out[i] = m ? (uint8_t)g : (uint8_t)0;
        else for (int k = 0; k < 3; ++k) out[i * 3 + k] = m ? frame[i * CN + k] : (uint8_t)0;
    }
}

// generate_binary's output image, for parity tests only (RawProcessing.cpp:597-600)
__global__ void binary_image_kernel(const uint8_t *__restrict__ frame, const uint8_t *__restrict__ bg,
                                    uint8_t *__restrict__ out, size_t n, SegK p)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t f = frame[i];
        uint32_t m = fg4<true>(f, bg[i], p) & 0xFFu;
        out[i] = m ? (uint8_t)f : (uint8_t)0;
    }
}

// ------------------------------------------------------------------------------------------------
// Optional morphology of generate_binary (default off in the reference): threshold mask ->
// closing (dilate, erode with the elliptical (2k+1)^2 element, RawProcessing.cpp:438-448,504-505) ->
// dilation_size > 0: dilate ones(n,n); < 0: erode, keep pixels whose difference still exceeds |T|,
// closing again (:541-550).  Pixels outside the image are ignored (OpenCV's default morphology border).
// These are plain per-pixel kernels over a mask image; K1 then runs on (mask, frame) with F_PREMASK.
// ------------------------------------------------------------------------------------------------
struct MorphEl { uint32_t rows[15]; int kw, kh, ax, ay; };

__global__ void thresh_mask_kernel(const uint8_t *__restrict__ frames, const uint8_t *__restrict__ bg, uint8_t *__restrict__ mask,
                                   uint8_t *__restrict__ diff, size_t frame_px, size_t total, SegK p)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t f = frames[i], b = bg[i % frame_px];
        const uint32_t in = (p.flags & F_INV) ? 255u - f : f;
        uint32_t d = in;
        if (p.flags & F_DIFF) d = (p.flags & F_ABS) ? (in > b ? in - b : b - in) : (b > in ? b - in : 0u);
        const uint32_t t = p.t4 & 0xFFu, lo = p.lo4 & 0xFFu, hi = p.hi4 & 0xFFu;
        bool m = (p.flags & F_RANGE) ? (d >= lo && d <= hi) : (d > t);
        if (p.flags & F_INVMASK) m = !m;
        mask[i] = m ? 255 : 0;
        if (diff) diff[i] = (uint8_t)d;
    }
}

// blur_difference (RawProcessing.cpp:371-380): the difference (always taken, absolute or signed) with values <= |T| zeroed
// (cv::threshold THRESH_TOZERO); the 25x25 blur and the second threshold follow
__global__ void diff_tozero_kernel(const uint8_t *__restrict__ frames, const uint8_t *__restrict__ bg, uint8_t *__restrict__ out,
                                   size_t frame_px, size_t total, SegK p)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t f = frames[i], b = bg[i % frame_px];
        const uint32_t in = (p.flags & F_INV) ? 255u - f : f;
        const uint32_t d = (p.flags & F_ABS) ? (in > b ? in - b : b - in) : (b > in ? b - in : 0u);
        out[i] = d > (p.t4 & 0xFFu) ? (uint8_t)d : (uint8_t)0;
    }
}
__global__ void gt_mask_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ mask, size_t total, uint32_t t)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
        mask[i] = src[i] > t ? 255 : 0;
}
// cv::adaptiveThreshold(MEAN_C, THRESH_BINARY, n, -T) on the difference image: 255 where diff - mean > T; inverted for T < 0
// (RawProcessing.cpp:487,498 / :526,537)
__global__ void adaptive_mask_kernel(const uint8_t *__restrict__ diff, const uint8_t *__restrict__ mean, uint8_t *__restrict__ mask, size_t total, int T)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        bool m = (int)diff[i] - (int)mean[i] > T;
        if (T < 0) m = !m;
        mask[i] = m ? 255 : 0;
    }
}

__global__ void morph_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int W, int H, size_t total, MorphEl el, int dilate)
{
    const size_t frame_px = (size_t)W * H;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t fo = i / frame_px * frame_px;
        const int y = (int)((i - fo) / W), x = (int)((i - fo) % W);
        int v = dilate ? 0 : 255;
        for (int j = 0; j < el.kh; ++j) {
            const int yy = y + j - el.ay;
            if (yy < 0 || yy >= H) continue;
            const uint32_t bits = el.rows[j];
            for (int k = 0; k < el.kw; ++k) {
                if (!((bits >> k) & 1u)) continue;
                const int xx = x + k - el.ax;
                if (xx < 0 || xx >= W) continue;
                const int sv = src[fo + (size_t)yy * W + xx];
                v = dilate ? max(v, sv) : min(v, sv);
            }
        }
        dst[i] = (uint8_t)v;
    }
}

// dilation_size < 0: keep the eroded mask where the difference still exceeds |T|
__global__ void remask_kernel(const uint8_t *__restrict__ eroded, const uint8_t *__restrict__ diff, uint8_t *__restrict__ mask, size_t total, uint32_t t)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
        mask[i] = (eroded[i] && diff[i] > t) ? 255 : 0;
}

__global__ void mask_and_kernel(const uint8_t *__restrict__ mask, const uint8_t *__restrict__ frame, uint8_t *__restrict__ out, size_t total)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
        out[i] = mask[i] & frame[i];
}

// Tracker-side re-threshold ("next" row N3a, pixel::threshold_blob, C/processing/PixelTree.cpp:186-291):
// the pixels of the kept detection blobs of a batch are painted into a per-frame mask; K1 then runs with the
// tracker's comparison (difference >= track_threshold) ANDed with that mask, and K2/K3 relabel.  Blobs of one
// frame are never 8-adjacent, so relabelling the frame equals relabelling every blob on its own.
__global__ void paint_blobs_kernel(const tb_blob_rec *__restrict__ recs, const uint32_t *__restrict__ totals, const tb_line *__restrict__ lines,
                                   uint8_t *__restrict__ mask, int W, int H)
{
    const uint32_t nb = totals[0];
    const int lane = threadIdx.x & 31;
    for (uint32_t b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); b < nb; b += gridDim.x * (blockDim.x >> 5)) {
        const tb_blob_rec r = recs[b];
        uint8_t *m = mask + (size_t)r.frame * W * H;
        for (uint32_t l = 0; l < r.n_lines; ++l) {
            const tb_line ln = lines[r.line_off + l];
            for (uint32_t x = ln.x0 + lane; x <= ln.x1; x += 32) m[(size_t)ln.y * W + x] = 0xFF;
        }
    }
}

// pv::Blob::recount (C/processing/PVBlob.cpp:934-1027) with Background::count_above_threshold (C/processing/Background.h:430-489) for every
// blob of a batch: the number of blob pixels whose difference to the background is >= threshold -- method 0 none (the value itself),
// 1 absolute |bg - v|, 2 sign bg - v (int32, may be negative) -- times SQR(cm_per_pixel).  rgb8 blobs compare cmn::bgr2gray of the pixel
// with the background's grey image (diffable_pixel_value<rgb8 -> gray>).  One warp per blob, lanes over the blob's lines.
__global__ void recount_kernel(const tb_blob_rec *__restrict__ recs, const uint32_t *__restrict__ totals, const tb_line *__restrict__ lines,
                               const uint32_t *__restrict__ line_px, const uint8_t *__restrict__ pixels, int opx, const uint8_t *__restrict__ bg, int W,
                               int method, int threshold, float sqcm, float *__restrict__ out)
{
    const uint32_t nb = totals[0];
    const int lane = threadIdx.x & 31;
    for (uint32_t b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); b < nb; b += gridDim.x * (blockDim.x >> 5)) {
        const tb_blob_rec r = recs[b];
        uint32_t cnt = 0;
        if (threshold == 0) cnt = lane == 0 ? r.n_pixels : 0u;            // :942-949: num_pixels()
        else
            for (uint32_t l = lane; l < r.n_lines; l += 32) {
                const tb_line ln = lines[r.line_off + l];
                const uint8_t *px = pixels + line_px[r.line_off + l];
                const uint8_t *bgr = bg + (size_t)ln.y * W + ln.x0;
                const int n = (int)ln.x1 - (int)ln.x0 + 1;
                for (int i = 0; i < n; ++i) {
                    const int v = opx == 3 ? (int)gray_px_tracker(px + 3 * i) : (int)px[i];
                    const int d = method == 0 ? v : (method == 1 ? abs((int)bgr[i] - v) : (int)bgr[i] - v);
                    cnt += d >= threshold;
                }
            }
#pragma unroll
        for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (lane == 0) out[b] = (float)cnt * sqcm;
    }
}

// ------------------------------------------------------------------------------------------------
// K2: one CTA per frame.  Orders the band chunks into raster order, labels runs with a union-find
// (8-connectivity between vertically adjacent rows, HLine.h:90-92 / CPULabeling.cpp:60-62,91; the
// root of a set is its smallest run index = the blob's first line), gathers per-blob statistics and
// applies the size filter.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t K2_SMEM_RUNS = 8192;      // union-find lives in shared memory up to this many runs

__device__ __forceinline__ uint32_t uf_find(uint32_t *par, uint32_t i)
{
    uint32_t q;
    while ((q = *(volatile uint32_t *)(par + i)) != i) i = q;
    return i;
}
__device__ __forceinline__ void uf_union(uint32_t *par, uint32_t a, uint32_t b)
{
    for (;;) {
        a = uf_find(par, a); b = uf_find(par, b);
        if (a == b) return;
        if (a < b) { uint32_t t = a; a = b; b = t; }        // a > b: hook the larger root under the smaller
        uint32_t old = atomicMin(par + a, b);
        if (old == a) return;
        a = old;                                            // lost a race: keep merging old's set with b
    }
}

// General path: any run count up to rcap, any height; scratch in global memory (the union-find in shared memory up to
// K2_SMEM_RUNS runs).  s_par: K2_SMEM_RUNS words of shared memory, ws: 34 words.
__device__ void ccl_label_body(const SegDev &d, uint32_t *s_par, uint32_t *ws)
{
    const int K2_NT = (int)blockDim.x;
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t R = d.run_count[f];
    uint32_t *tot = d.frame_tot + (size_t)f * 4;
    if (R > d.rcap) {                                      // capacity exceeded: drop the frame, flag it
        if (tid == 0) { tot[0] = 0; tot[1] = 0; tot[2] = 0; tot[3] = 1u; }
        return;
    }
    tb_line *runs = d.runs + (size_t)f * d.rcap;
    const tb_line *raw = d.runs_raw + (size_t)f * d.rcap;
    uint32_t *gpar = d.parent + (size_t)f * d.rcap;
    uint32_t *par = (R <= K2_SMEM_RUNS) ? s_par : gpar;
    uint32_t *bidx = d.bidx + (size_t)f * d.rcap;
    uint32_t *rs = d.row_start + (size_t)f * d.H, *re = d.row_end + (size_t)f * d.H;
    const uint32_t *bbase = d.band_base + (size_t)f * d.n_bands, *bcnt = d.band_cnt + (size_t)f * d.n_bands;

    // 1. raster order: exclusive scan of the band counts, then copy band chunks (a warp per band)
    for (int y = tid; y < d.H; y += K2_NT) { rs[y] = 0; re[y] = 0; }
    uint32_t carry = 0;
    for (int b0 = 0; b0 < d.n_bands; b0 += K2_NT) {
        int b = b0 + tid;
        uint32_t cnt = b < d.n_bands ? bcnt[b] : 0u, total;
        uint32_t ex = block_excl_scan(cnt, ws, total);
        if (b < d.n_bands) bidx[b] = carry + ex;          // bidx reused as the band offset table (n_bands <= rcap)
        carry += total;
    }
    __syncthreads();
    for (int b = warp; b < d.n_bands; b += K2_NT / 32) {
        const uint32_t src = bbase[b], dst = bidx[b], n = bcnt[b];
        for (uint32_t i = lane; i < n; i += 32) runs[dst + i] = raw[src + i];
    }
    __syncthreads();
    // 2. init union-find, row table
    for (uint32_t i = tid; i < R; i += K2_NT) {
        par[i] = i;
        const uint32_t y = runs[i].y;
        if (i == 0 || runs[i - 1].y != y) rs[y] = i;
        if (i + 1 == R || runs[i + 1].y != y) re[y] = i + 1;
    }
    __syncthreads();
    // 3. unions with the runs of row y-1
    for (uint32_t i = tid; i < R; i += K2_NT) {
        const tb_line r = runs[i];
        if (r.y == 0) continue;
        uint32_t s = rs[r.y - 1], e = re[r.y - 1];
        if (s >= e) continue;
        uint32_t lo = s, hi = e;                           // first j with x1[j] + 1 >= x0
        while (lo < hi) {
            uint32_t mid = (lo + hi) >> 1;
            if ((uint32_t)runs[mid].x1 + 1u < (uint32_t)r.x0) lo = mid + 1; else hi = mid;
        }
        for (uint32_t j = lo; j < e && (uint32_t)runs[j].x0 <= (uint32_t)r.x1 + 1u; ++j) uf_union(par, i, j);
    }
    __syncthreads();
    // 4. flatten; roots get dense blob indices in raster order of their first run
    uint32_t kbase = 0;
    for (uint32_t i0 = 0; i0 < R; i0 += K2_NT) {
        uint32_t i = i0 + tid, root = 0, isroot = 0;
        if (i < R) { root = uf_find(par, i); isroot = root == i; }
        uint32_t total, ex = block_excl_scan(isroot, ws, total);
        if (i < R) {
            gpar[i] = root;                                // final label, global for K3
            if (isroot) { bidx[i] = kbase + ex; d.b_root[(size_t)f * d.rcap + kbase + ex] = i; }
        }
        kbase += total;
    }
    const uint32_t K = kbase;
    uint32_t *npx = d.b_npx + (size_t)f * d.rcap, *nl = d.b_nl + (size_t)f * d.rcap;
    uint32_t *xmin = d.b_xmin + (size_t)f * d.rcap, *xmax = d.b_xmax + (size_t)f * d.rcap;
    uint32_t *ymax = d.b_ymax + (size_t)f * d.rcap;
    for (uint32_t k = tid; k < K; k += K2_NT) { npx[k] = 0; nl[k] = 0; xmin[k] = 0xFFFFFFFFu; xmax[k] = 0; ymax[k] = 0; }
    __syncthreads();
    // 5. per-blob statistics
    for (uint32_t i = tid; i < R; i += K2_NT) {
        const tb_line r = runs[i];
        const uint32_t k = bidx[gpar[i]];
        atomicAdd(npx + k, (uint32_t)r.x1 - r.x0 + 1u);
        atomicAdd(nl + k, 1u);
        atomicMin(xmin + k, (uint32_t)r.x0);
        atomicMax(xmax + k, (uint32_t)r.x1);
        atomicMax(ymax + k, (uint32_t)r.y);
    }
    __syncthreads();
    // 6. size filter (BackgroundSubtraction.cpp:259: lo <= npx*cm^2 < hi in float; :306 line limit),
    //    offsets of the kept blobs inside the frame
    uint32_t *loff = d.b_loff + (size_t)f * d.rcap, *poff = d.b_poff + (size_t)f * d.rcap;
    uint32_t *kept = d.kept + (size_t)f * d.rcap;
    uint32_t kk = 0, kl = 0, kp = 0;
    for (uint32_t k0 = 0; k0 < K; k0 += K2_NT) {
        uint32_t k = k0 + tid, keep = 0, l = 0, px = 0;
        if (k < K) {
            l = nl[k]; px = npx[k];
            keep = d.n_ranges == 0;
            const double v = (double)((float)(d.keep_mask ? px : px * (uint32_t)d.opx) * d.sqcm);      // detect side: pixels->size() = payload BYTES (3 per rgb8 pixel), BackgroundSubtraction.cpp:247-259; tracker side: pixel counts
            for (int q = 0; q < d.n_ranges; ++q) keep |= (v >= d.lo[q] && v < d.hi[q]);
            keep = keep && l < 65535u && px * (uint32_t)d.opx > d.min_payload;
        }
        uint32_t t0, t1, t2;
        uint32_t e0 = block_excl_scan(keep, ws, t0);
        uint32_t e1 = block_excl_scan(keep ? l : 0u, ws, t1);
        uint32_t e2 = block_excl_scan(keep ? px * (uint32_t)d.opx : 0u, ws, t2);     // pixel arena offsets count bytes
        if (k < K && keep) { kept[kk + e0] = k; loff[k] = kl + e1; poff[k] = kp + e2; }
        kk += t0; kl += t1; kp += t2;
    }
    if (tid == 0) {
        uint32_t status = 0;
        if (kp > d.px_frame_cap) { status |= 2u; kk = 0; kl = 0; kp = 0; }
        tot[0] = kk; tot[1] = kl; tot[2] = kp; tot[3] = status;
    }
}


// three exclusive block scans behind one set of barriers (ws: 3 x 34 words)
__device__ __forceinline__ void block_excl_scan3(const uint32_t v[3], uint32_t *ws, uint32_t ex[3], uint32_t total[3])
{
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nw = (blockDim.x + 31u) >> 5;
    uint32_t inc[3] = {v[0], v[1], v[2]};
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc[q], d);
            if (lane >= (unsigned)d) inc[q] += t;
        }
    __syncthreads();
    if (lane == 31) { ws[warp] = inc[0]; ws[34 + warp] = inc[1]; ws[68 + warp] = inc[2]; }
    __syncthreads();
    if (warp < 3) {
        uint32_t *w = ws + 34 * warp;
        const uint32_t x = lane < nw ? w[lane] : 0u;
        uint32_t winc = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, winc, d);
            if (lane >= (unsigned)d) winc += t;
        }
        if (lane < nw) w[lane] = winc - x;
        if (lane == 31) w[32] = winc;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 3; ++q) { total[q] = ws[34 * q + 32]; ex[q] = ws[34 * q + warp] + inc[q] - v[q]; }
}

// K2, fast path (the common case: <= 8192 runs per frame, <= 4352 rows): the raster-ordered runs (packed x0 | x1 << 16
// and y), the union-find, the row table and the per-blob statistics all live in shared memory, every thread owns a
// contiguous chunk of runs when roots are numbered (one block scan for the whole frame) and the three prefix sums
// of the size filter share their barriers: ~16 barriers and no dependent global loads between them, instead of ~45
// phases with global-memory binary searches.  Frames beyond the limits take ccl_label_body.
constexpr int K2F_NT = 1024;
constexpr uint32_t K2F_RUNS = K2_SMEM_RUNS, K2F_ROWS = 4352, K2F_BLOBS = 512;
constexpr int K2F_SMEM = K2F_RUNS * 4 * 2 + K2F_RUNS * 2 + K2F_ROWS * 2 * 2 + K2F_BLOBS * 5 * 4;

__global__ void __launch_bounds__(K2F_NT, 1)
ccl_label_kernel(SegDev d)
{
    extern __shared__ __align__(16) uint8_t k2_dsm[];
    __shared__ uint32_t ws[3 * 34];
    uint32_t *s_par = reinterpret_cast<uint32_t *>(k2_dsm);
    uint32_t *s_x = s_par + K2F_RUNS;
    uint16_t *s_y = reinterpret_cast<uint16_t *>(s_x + K2F_RUNS);
    uint16_t *s_rs = s_y + K2F_RUNS, *s_re = s_rs + K2F_ROWS;
    uint32_t *s_st = reinterpret_cast<uint32_t *>(s_re + K2F_ROWS);          // [5][K2F_BLOBS]
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t R = d.run_count[f];
    if (R > K2F_RUNS || R > d.rcap || (uint32_t)d.H > K2F_ROWS || (uint32_t)d.n_bands > K2F_RUNS) { ccl_label_body(d, s_par, ws); return; }
    uint32_t *tot = d.frame_tot + (size_t)f * 4;
    const size_t o = (size_t)f * d.rcap;
    tb_line *runs = d.runs + o;
    const tb_line *raw = d.runs_raw + o;
    uint32_t *gpar = d.parent + o;
    uint32_t *rs = d.row_start + (size_t)f * d.H, *re = d.row_end + (size_t)f * d.H;
    const uint32_t *bbase = d.band_base + (size_t)f * d.n_bands, *bcnt = d.band_cnt + (size_t)f * d.n_bands;

    // 1. raster order: exclusive scan of the band counts (offsets parked in s_par), then copy band chunks (a warp per band)
    for (int y = tid; y < d.H; y += K2F_NT) { s_rs[y] = 0; s_re[y] = 0; }
    uint32_t carry = 0;
    for (int b0 = 0; b0 < d.n_bands; b0 += K2F_NT) {
        const int b = b0 + tid;
        uint32_t cnt = b < d.n_bands ? bcnt[b] : 0u, total;
        const uint32_t ex = block_excl_scan(cnt, ws, total);
        if (b < d.n_bands) s_par[b] = carry + ex;
        carry += total;
    }
    __syncthreads();
    for (int b = warp; b < d.n_bands; b += K2F_NT / 32) {
        const uint32_t n = bcnt[b];
        if (!n) continue;
        const uint32_t src = bbase[b], dst = s_par[b];
        for (uint32_t i = lane; i < n; i += 32) {
            const tb_line r = raw[src + i];
            runs[dst + i] = r;
            s_x[dst + i] = (uint32_t)r.x0 | ((uint32_t)r.x1 << 16);
            s_y[dst + i] = r.y;
        }
    }
    __syncthreads();
    // 2. union-find init, row table
    for (uint32_t i = tid; i < R; i += K2F_NT) {
        s_par[i] = i;
        const uint32_t y = s_y[i];
        if (i == 0 || s_y[i - 1] != y) s_rs[y] = (uint16_t)i;
        if (i + 1 == R || s_y[i + 1] != y) s_re[y] = (uint16_t)(i + 1);
    }
    __syncthreads();
    for (int y = tid; y < d.H; y += K2F_NT) { rs[y] = s_rs[y]; re[y] = s_re[y]; }      // K3 reads the global copy
    // 3. unions with the runs of row y-1 (8-connectivity, HLine.h:90-92)
    for (uint32_t i = tid; i < R; i += K2F_NT) {
        const uint32_t y = s_y[i];
        if (y == 0) continue;
        const uint32_t s = s_rs[y - 1], e = s_re[y - 1];
        if (s >= e) continue;
        const uint32_t x = s_x[i], x0 = x & 0xFFFFu, x1 = x >> 16;
        uint32_t lo = s, hi = e;                               // first j with x1[j] + 1 >= x0
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if ((s_x[mid] >> 16) + 1u < x0) lo = mid + 1; else hi = mid;
        }
        for (uint32_t j = lo; j < e && (s_x[j] & 0xFFFFu) <= x1 + 1u; ++j) uf_union(s_par, i, j);
    }
    __syncthreads();
    // 4. flatten (label = root run index), then number the roots in raster order: every thread owns a chunk of runs
    for (uint32_t i = tid; i < R; i += K2F_NT) {
        const uint32_t root = uf_find(s_par, i);
        gpar[i] = root;
        if (root != i) s_par[i] = root;
    }
    __syncthreads();
    const uint32_t per = (R + K2F_NT - 1) / K2F_NT, c0 = min(R, (uint32_t)tid * per), c1 = min(R, c0 + per);
    uint32_t nroot = 0;
    for (uint32_t i = c0; i < c1; ++i) nroot += s_par[i] == i;
    uint32_t K;
    uint32_t kx = block_excl_scan(nroot, ws, K);
    for (uint32_t i = c0; i < c1; ++i)
        if (s_par[i] == i) { s_par[i] = 0x80000000u | kx; d.b_root[o + kx] = i; ++kx; }     // a root's slot now holds its blob index
    // 5. per-blob statistics (shared-memory atomics when the frame has <= K2F_BLOBS components)
    const bool st_smem = K <= K2F_BLOBS;
    uint32_t *npx = st_smem ? s_st : d.b_npx + o, *nl = st_smem ? s_st + K2F_BLOBS : d.b_nl + o;
    uint32_t *xmin = st_smem ? s_st + 2 * K2F_BLOBS : d.b_xmin + o, *xmax = st_smem ? s_st + 3 * K2F_BLOBS : d.b_xmax + o;
    uint32_t *ymax = st_smem ? s_st + 4 * K2F_BLOBS : d.b_ymax + o;
    for (uint32_t k = tid; k < K; k += K2F_NT) { npx[k] = 0; nl[k] = 0; xmin[k] = 0xFFFFFFFFu; xmax[k] = 0; ymax[k] = 0; }
    __syncthreads();
    for (uint32_t i = tid; i < R; i += K2F_NT) {
        const uint32_t p = s_par[i];
        const uint32_t k = ((p & 0x80000000u) ? p : s_par[p]) & 0x7FFFFFFFu;
        const uint32_t x = s_x[i], x0 = x & 0xFFFFu, x1 = x >> 16;
        atomicAdd(npx + k, x1 - x0 + 1u);
        atomicAdd(nl + k, 1u);
        atomicMin(xmin + k, x0);
        atomicMax(xmax + k, x1);
        atomicMax(ymax + k, (uint32_t)s_y[i]);
    }
    __syncthreads();
    // 6. size filter (BackgroundSubtraction.cpp:259: lo <= npx*cm^2 < hi in float; :306 line limit), offsets of the kept blobs
    uint32_t *loff = d.b_loff + o, *poff = d.b_poff + o, *kept = d.kept + o;
    uint32_t kk = 0, kl = 0, kp = 0;
    for (uint32_t k0 = 0; k0 < K; k0 += K2F_NT) {
        const uint32_t k = k0 + tid;
        uint32_t keep = 0, l = 0, px = 0;
        if (k < K) {
            l = nl[k]; px = npx[k];
            keep = d.n_ranges == 0;
            const double v = (double)((float)(d.keep_mask ? px : px * (uint32_t)d.opx) * d.sqcm);      // detect side: pixels->size() = payload BYTES (3 per rgb8 pixel), BackgroundSubtraction.cpp:247-259; tracker side: pixel counts
            for (int q = 0; q < d.n_ranges; ++q) keep |= (v >= d.lo[q] && v < d.hi[q]);
            keep = keep && l < 65535u && px * (uint32_t)d.opx > d.min_payload;
            if (st_smem) { d.b_xmin[o + k] = xmin[k]; d.b_xmax[o + k] = xmax[k]; d.b_ymax[o + k] = ymax[k]; }
        }
        const uint32_t v3[3] = {keep, keep ? l : 0u, keep ? px * (uint32_t)d.opx : 0u};      // pixel arena offsets count bytes
        uint32_t e3[3], t3[3];
        block_excl_scan3(v3, ws, e3, t3);
        if (k < K && keep) { kept[kk + e3[0]] = k; loff[k] = kl + e3[1]; poff[k] = kp + e3[2]; }
        kk += t3[0]; kl += t3[1]; kp += t3[2];
    }
    if (tid == 0) {
        uint32_t status = 0;
        if (kp > d.px_frame_cap) { status |= 2u; kk = 0; kl = 0; kp = 0; }
        tot[0] = kk; tot[1] = kl; tot[2] = kp; tot[3] = status;
    }
}

// ------------------------------------------------------------------------------------------------
// K3: one CTA per frame, one warp per kept blob.  The warp walks the blob's bounding box, one lane
// per row, collecting the blob's runs in (y,x0) order (CPULabeling.cpp:256-323), then copies pixel
// bytes (:307) and renders the individual crop (FilterCache.cpp:158-235).  Arena offsets of the
// frame are the prefix sums over the batch of the totals K2 wrote.
// ------------------------------------------------------------------------------------------------
constexpr int K3_NT = 256, K3_SPLIT = 4;     // minimum CTAs per frame: blobs are dealt round-robin to split * 8 warps

__global__ void __launch_bounds__(K3_NT)
blob_emit_kernel(const uint8_t *__restrict__ frames, SegDev d)
{
    __shared__ uint32_t s_red[4][K3_NT / 32];
    __shared__ uint32_t s_base[4];
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int part = blockIdx.y, nparts = gridDim.y;
    // prefix over frames < f of (blobs, lines, pixels, crops)
    uint32_t acc[4] = {0, 0, 0, 0};
    for (int g = tid; g < f; g += K3_NT) {
        const uint32_t *t = d.frame_tot + (size_t)g * 4;
        acc[0] += t[0]; acc[1] += t[1]; acc[2] += t[2]; acc[3] += min(t[0], d.max_crops);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
        if (lane == 0) s_red[q][warp] = acc[q];
    }
    __syncthreads();
    if (tid < 4) { uint32_t s = 0; for (int w = 0; w < K3_NT / 32; ++w) s += s_red[tid][w]; s_base[tid] = s; }
    __syncthreads();
    const uint32_t *tot = d.frame_tot + (size_t)f * 4;
    uint32_t Kk = tot[0], Lk = tot[1], Pk = tot[2], status = tot[3];
    const uint32_t Bb = s_base[0], Lb = s_base[1], Pb = s_base[2], Cb = s_base[3];
    if (Bb + Kk > d.blobs_cap || Lb + Lk > d.lines_cap || Pb + Pk > d.px_cap) { status |= 8u; Kk = 0; Lk = 0; Pk = 0; }
    const uint32_t ncrop = min(Kk, d.max_crops);
    if (Kk > d.max_crops && d.max_crops) status |= 4u;
    if (tid == 0 && part == 0) {
        tb_frame_info fi;
        fi.blob_begin = Bb; fi.n_blobs = Kk; fi.line_begin = Lb; fi.n_lines = Lk;
        fi.px_begin = Pb; fi.n_pixels = Pk; fi.n_runs = d.run_count[f]; fi.status = status;
        d.infos[f] = fi;
        if (f == d.B - 1) {
            uint32_t tb = Bb + Kk, tl = Lb + Lk, tp = Pb + Pk, tc = Cb + ncrop;
            if (status & 8u) {          // a batch arena overflowed at or before this frame (every later frame is dropped too):
                tb = tl = tp = tc = 0;  // the totals cover the frames before the first overflow
                for (int g = 0; g < d.B; ++g) {
                    const uint32_t *t = d.frame_tot + (size_t)g * 4;
                    if (tb + t[0] > d.blobs_cap || tl + t[1] > d.lines_cap || tp + t[2] > d.px_cap) break;
                    tb += t[0]; tl += t[1]; tp += t[2]; tc += min(t[0], d.max_crops);
                }
            }
            d.totals[0] = tb; d.totals[1] = tl; d.totals[2] = tp; d.totals[3] = tc;
        }
    }
    if (Kk == 0) return;

    const tb_line *runs = d.runs + (size_t)f * d.rcap;
    const uint32_t *label = d.parent + (size_t)f * d.rcap;
    const uint32_t *rs = d.row_start + (size_t)f * d.H, *re = d.row_end + (size_t)f * d.H;
    const int CN = d.CN, opx = d.opx;
    const uint8_t *frame = frames + (size_t)f * d.W * d.H * CN;
    const size_t o = (size_t)f * d.rcap;
    const int cw = d.crop_w, ch = d.crop_h;

    for (uint32_t q = part * (K3_NT / 32) + warp; q < Kk; q += nparts * (K3_NT / 32)) {
        const uint32_t k = d.kept[o + q];
        const uint32_t root = d.b_root[o + k];
        const uint32_t bx0 = d.b_xmin[o + k], bx1 = d.b_xmax[o + k], by1 = d.b_ymax[o + k];
        const uint32_t by0 = runs[root].y;
        const uint32_t L0 = Lb + d.b_loff[o + k], P0 = Pb + d.b_poff[o + k];
        const uint32_t bh = by1 - by0 + 1, bw = bx1 - bx0 + 1;
        // phase A: lines in (y,x0) order
        uint32_t cl = 0, cp = 0;
        for (uint32_t rb = 0; rb < bh; rb += 32) {
            const uint32_t y = by0 + rb + lane;
            uint32_t j0 = 0, e = 0, nl = 0, np = 0;
            if (rb + lane < bh) {
                uint32_t s = rs[y]; e = re[y];
                uint32_t lo = s, hi = e;                   // first run with x1 >= bx0
                while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (runs[mid].x1 < bx0) lo = mid + 1; else hi = mid; }
                j0 = lo;
                for (uint32_t j = j0; j < e && runs[j].x0 <= bx1; ++j)
                    if (label[j] == root) { ++nl; np += (uint32_t)runs[j].x1 - runs[j].x0 + 1u; }
            }
            uint32_t tl, tp, el = warp_excl_scan(nl, tl), ep = warp_excl_scan(np * (uint32_t)opx, tp);
            if (nl) {
                uint32_t wl = L0 + cl + el, wp = P0 + cp + ep;
                for (uint32_t j = j0; j < e && runs[j].x0 <= bx1; ++j)
                    if (label[j] == root) {
                        d.lines[wl] = runs[j]; d.line_px[wl] = wp;
                        ++wl; wp += ((uint32_t)runs[j].x1 - runs[j].x0 + 1u) * (uint32_t)opx;
                    }
            }
            cl += tl; cp += tp;
        }
        const bool do_crop = q < ncrop;
        const bool render = do_crop && !d.crop_norm && d.crop_scale == 1.f;       // normalised / scaled crops are rendered by their own kernels
        const int cpx = d.cpx;
        uint8_t *crop = d.crops + (size_t)(Cb + q) * cw * ch * cpx;
        int offx = 0, offy = 0;
        if (render) {
            int dd;
            if ((int)bw < cw) { dd = cw - (int)bw; offx = dd - dd / 2; } else { dd = (int)bw - cw; offx = -(dd - dd / 2); }
            if ((int)bh < ch) { dd = ch - (int)bh; offy = dd - dd / 2; } else { dd = (int)bh - ch; offy = -(dd - dd / 2); }
            if (((cw * ch * cpx) & 15) == 0) {
                uint4 *c4 = reinterpret_cast<uint4 *>(crop);
                for (int i = lane; i < (cw * ch * cpx) >> 4; i += 32) c4[i] = make_uint4(0, 0, 0, 0);
            } else {
                for (int i = lane; i < cw * ch * cpx; i += 32) crop[i] = 0;
            }
        }
        if (lane == 0) {
            tb_blob_rec rec;
            rec.line_off = L0; rec.px_off = P0; rec.n_lines = cl; rec.n_pixels = cp / (uint32_t)opx;
            rec.x0 = (uint16_t)bx0; rec.y0 = (uint16_t)by0; rec.x1 = (uint16_t)bx1; rec.y1 = (uint16_t)by1;
            const tb_line fl = runs[root];
            rec.bid = ((((uint32_t)fl.x0 + fl.x1 + 1u) / 2u) << 19) | (((uint32_t)fl.y & 0x1FFFu) << 6) | ((cl & 0xFFu) % 64u);
            rec.frame = (uint32_t)f;
            d.recs[Bb + q] = rec;
            if (do_crop) d.crop_blob[Cb + q] = Bb + q;
        }
        __syncwarp();                                      // lines / line_px / zero fill visible to the warp
        // phase B: pixel bytes + crop.  32 lines are fetched at once (one per lane); the warp then walks their pixels as
        // one flat range, 64 pixels per step (two per lane): a pixel finds its line by a 5-step binary search over the
        // lines' first-pixel indices (shuffles), so that short lines do not leave lanes idle and the loads of a step are
        // all in flight before the first store.
        for (uint32_t lb = 0; lb < cl; lb += 32) {
            const uint32_t cnt = min(32u, cl - lb);
            tb_line mine = {0, 0, 0, 0}; uint32_t mypo = 0, myq = 0xFFFFFFFFu;
            if (lane < cnt) { mine = d.lines[L0 + lb + lane]; mypo = d.line_px[L0 + lb + lane]; myq = (mypo - P0) / (uint32_t)opx; }
            const uint32_t packed = (uint32_t)mine.x0 | ((uint32_t)mine.x1 << 16);
            const uint32_t qb = __shfl_sync(0xffffffffu, myq, 0);
            const uint32_t qe = __shfl_sync(0xffffffffu, myq + (uint32_t)mine.x1 - mine.x0 + 1u, (int)cnt - 1);
            for (uint32_t q0 = qb; q0 < qe; q0 += 64) {
                uint32_t qq[2] = {q0 + lane, q0 + 32 + lane};
                int li[2] = {0, 0};
#pragma unroll
                for (int step = 16; step > 0; step >>= 1) {
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const uint32_t v = __shfl_sync(0xffffffffu, myq, (li[u] + step) & 31);
                        if (li[u] + step < (int)cnt && v <= qq[u]) li[u] += step;
                    }
                }
                size_t pi[2]; uint32_t po[2]; int ci[2]; bool act[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const uint32_t xx = __shfl_sync(0xffffffffu, packed, li[u]);
                    const uint32_t ly = __shfl_sync(0xffffffffu, (uint32_t)mine.y, li[u]);
                    const uint32_t lq = __shfl_sync(0xffffffffu, myq, li[u]);
                    const uint32_t lp = __shfl_sync(0xffffffffu, mypo, li[u]);
                    act[u] = qq[u] < qe;
                    const uint32_t x = (xx & 0xFFFFu) + (qq[u] - lq);
                    pi[u] = (size_t)ly * d.W + x;
                    po[u] = lp + (qq[u] - lq) * (uint32_t)opx;
                    const int cx = (int)(x - bx0) + offx, cy = (int)(ly - by0) + offy;
                    ci[u] = (render && cx >= 0 && cx < cw && cy >= 0 && cy < ch) ? cy * cw + cx : -1;
                }
                if (opx == 1) {                                // gray encoding: the grey value (of a colour pixel: cvtColor / plane)
                    uint32_t v[2] = {0, 0}; int bgv[2] = {0, 0};
#pragma unroll
                    for (int u = 0; u < 2; ++u)
                        if (act[u]) {
                            v[u] = gray_px(frame + pi[u] * CN, CN, d.cc);
                            if (d.crop_method && ci[u] >= 0) bgv[u] = d.bg[pi[u]];
                        }
#pragma unroll
                    for (int u = 0; u < 2; ++u)
                        if (act[u]) {
                            d.pixels[po[u]] = (uint8_t)v[u];
                            if (ci[u] >= 0 && d.r3) {          // r3g3b2 blob: rendered as B,G,R = r3g3b2_to_vec(code), differences per channel
#pragma unroll                                               // against the background's expanded codes (Background.cpp:65-69, Background.h:113-116)
                                for (int k = 0; k < 3; ++k) {
                                    int val = r3g3b2_channel(v[u], k);
                                    if (d.crop_method) {
                                        const int bb = r3g3b2_channel((uint32_t)bgv[u], k);
                                        val = d.crop_method == 1 ? abs(bb - val) : max(0, bb - val);
                                    }
                                    crop[ci[u] * 3 + k] = (uint8_t)val;
                                }
                            } else if (ci[u] >= 0) {
                                int val = (int)v[u];
                                if (d.crop_method) val = d.crop_method == 1 ? abs(bgv[u] - val) : max(0, bgv[u] - val);
                                crop[ci[u]] = (uint8_t)val;
                            }
                        }
                } else {                                       // rgb8: B,G,R; crops difference per channel (Background.h:238-262)
#pragma unroll
                    for (int u = 0; u < 2; ++u)
                        if (act[u]) {
#pragma unroll
                            for (int k = 0; k < 3; ++k) {
                                const uint8_t v = frame[pi[u] * CN + k];
                                d.pixels[po[u] + k] = v;
                                if (ci[u] >= 0) {
                                    int val = v;
                                    if (d.crop_method) {
                                        const int bb = d.bg3[pi[u] * 3 + k];
                                        val = d.crop_method == 1 ? abs(bb - val) : max(0, bb - val);
                                    }
                                    crop[ci[u] * 3 + k] = (uint8_t)val;
                                }
                            }
                        }
                }
            }
        }
    }
}

}  // namespace tb

// ================================================================================================
// host side: the tb_seg handle
// ================================================================================================
using namespace tb;

struct tb_seg {
    tb_seg_config cfg{};
    tb_seg_params params{};
    SegDev d{};
    SegK k{};
    cudaStream_t stream = nullptr, own_stream = nullptr;
    cudaEvent_t ev_done = nullptr;
    bool has_bg = false;
    uint8_t *d_bg = nullptr, *d_frames = nullptr, *d_tmp = nullptr;
    uint8_t *d_bg3 = nullptr;                   // rgb8: the 3-channel background
    uint8_t *d_gray = nullptr, *d_nz = nullptr; // colour frames on the plane path: grey plane / non-zero plane of a batch (allocated on first use)
    bool gray_valid = false;                    // d_gray holds the grey plane of the last batch
    std::vector<void *> dev_allocs;
    // pinned host mirrors
    tb_frame_info *h_infos = nullptr; uint32_t *h_totals = nullptr;
    tb_blob_rec *h_recs = nullptr; tb_line *h_lines = nullptr; uint8_t *h_pixels = nullptr;
    uint8_t *h_crops = nullptr; uint32_t *h_crop_blob = nullptr;
    int last_n = 0, last_fetch = 0; bool pending = false, fetched_payload = false, fetched_crops = false;
    cudaStream_t last_stream = nullptr;
    uint64_t launches = 0;
    EventRing<3> prof;
    // optional morphology (use_closing / dilation_size): mask images of one batch, allocated on first use
    bool morph = false;
    uint8_t *m_a = nullptr, *m_b = nullptr, *m_diff = nullptr;
    uint32_t *box_hs = nullptr; int box_sub = 0;
    // posture chain (tb_seg_posture / tb_seg_outlines / tb_seg_midlines), allocated on first use
    uint8_t *o_visited = nullptr; uint32_t *o_rowfirst = nullptr; int4 *o_sel = nullptr; tb_outline_rec *o_recs = nullptr; uint32_t *o_totals = nullptr;
    float *o_raw = nullptr, *o_res = nullptr; uint32_t o_cap = 0, o_n = 0;
    tb_outline_rec *h_o_recs = nullptr; float *h_o_raw = nullptr, *h_o_res = nullptr; uint32_t *h_o_totals = nullptr;   // h_o_totals: raw, res, status
    float *m_pts = nullptr, *m_segs = nullptr; tb_midline_rec *m_recs = nullptr;
    tb_midline_norm *m_nrecs = nullptr; float *m_norm = nullptr; int m_res = 0;        // normalised midlines: m_res points per blob
    float *m_arena = nullptr; unsigned long long m_arena_floats = 0, *m_arena_used = nullptr; uint32_t *p_status = nullptr;
    uint8_t *crop_valid = nullptr, *h_crop_valid = nullptr;
    float *h_m_pts = nullptr, *h_m_segs = nullptr; tb_midline_rec *h_m_recs = nullptr; uint32_t m_n = 0;
    tb_midline_norm *h_m_nrecs = nullptr; float *h_m_norm = nullptr;
    bool p_pending = false, p_has_norm = false, p_has_crops = false; int p_fetch = 0, p_sms = 0; cudaStream_t p_stream = nullptr; uint32_t p_n = 0;
    EventRing<3> p_prof;
    // posture of thresholded sub-blobs (tb_seg_posture_thresholded): per-parent round state, allocated on first use
    unsigned long long *r_best = nullptr; uint32_t *r_index = nullptr, *r_sub_npx = nullptr, *r_remaining = nullptr, *h_r_remaining = nullptr;
    uint8_t *r_state = nullptr; tb_outline_rec *r_first = nullptr;
    tb_seg *p_parent = nullptr;
    float *d_recount = nullptr;                 // tb_seg_recount: one value per blob                 // the handle whose blobs the posture results are indexed by (nullptr: this one)
    const uint8_t *last_frames_dev = nullptr;   // frames of the last batch (device)
    uint8_t *keep_mask = nullptr;               // tracker-side handle: painted detection blobs of the batch
    double *d_coef = nullptr;                   // `moments` normalisation: inverted warp matrix per crop
    MorphEl el_close{}, el_dil{}, el_open{};
    uint8_t *d_meta = nullptr; tb_meta_layout meta{};   // headers | top-1 ids | top-1 probabilities | blob records (see tb_seg_metadata)
    int ws_ctas = 0, col_ctas = 0;              // grid of the persistent K1 (resident CTAs x SMs of THIS handle's device), set on first use
    unsigned long long *dbg = nullptr;          // TB_SEG_TIMELINE: per-CTA time stamps (debug)
};

static int seg_make_k(const tb_seg_params &p, SegK &k, std::string &why)
{
    if (p.use_closing && (p.closing_size < 1 || p.closing_size > 7)) { why = "closing_size must be 1..7 (element up to 15x15)"; return TB_ERR_INVALID; }
    if (p.dilation_size < -15 || p.dilation_size > 15) { why = "dilation_size must be -15..15"; return TB_ERR_INVALID; }
    if (p.n_size_ranges < 0 || p.n_size_ranges > 4) { why = "n_size_ranges must be 0..4"; return TB_ERR_INVALID; }
    auto rep = [](int v) { uint32_t b = (uint32_t)std::min(std::max(v, 0), 255); return b * 0x01010101u; };
    k.flags = (p.enable_difference ? F_DIFF : 0) | (p.detect_threshold_is_absolute ? F_ABS : 0) |
              (p.image_invert ? F_INV : 0) | (p.detect_threshold < 0 ? F_INVMASK : 0) |
              (p.threshold_maximum < 255 ? F_RANGE : 0);
    const int T = p.detect_threshold, aT = T < 0 ? -T : T;
    k.t4 = rep(aT);                       // d > |T|; |T| >= 255 can never be exceeded by a byte
    k.lo4 = rep(T); k.hi4 = rep(p.threshold_maximum);
    if (T > 255 || p.threshold_maximum < 0) { k.lo4 = rep(255); k.hi4 = rep(0); }   // empty range
    if (p.color_channel < -1) { why = "color_channel must be -1 (none) or a plane index"; return TB_ERR_INVALID; }
    return TB_OK;
}

// cv::getStructuringElement(MORPH_ELLIPSE, (2k+1, 2k+1)) as OpenCV computes it: row i spans
// |dx| <= round(k * sqrt(1 - (dy/k)^2))  (used at RawProcessing.cpp:442)
static MorphEl ellipse_element(int k)
{
    MorphEl e{};
    const int n = 2 * k + 1;
    e.kw = e.kh = n; e.ax = e.ay = k;
    const double inv_r2 = k ? 1.0 / ((double)k * k) : 0.0;
    for (int i = 0; i < n; ++i) {
        const int dy = i - k;
        const int dx = (int)std::lrint(k * std::sqrt((k * k - dy * dy) * inv_r2));
        const int j1 = std::max(k - dx, 0), j2 = std::min(k + dx + 1, n);
        for (int j = j1; j < j2; ++j) e.rows[i] |= 1u << j;
    }
    return e;
}
static MorphEl ones_element(int n)          // cv::Mat::ones(n, n), anchor at the centre (n / 2)
{
    MorphEl e{};
    e.kw = e.kh = n; e.ax = e.ay = n / 2;
    for (int i = 0; i < n; ++i) e.rows[i] = (1u << n) - 1u;
    return e;
}

extern "C" void tb_seg_default_params(tb_seg_params *p)
{
    std::memset(p, 0, sizeof(*p));
    p->detect_threshold = 15; p->threshold_maximum = 255; p->enable_difference = 1;
    p->detect_threshold_is_absolute = 1; p->closing_size = 3; p->cm_per_pixel = 1.f;
    p->n_size_ranges = 0; p->color_channel = -1; p->adaptive_threshold_scale = 2.f;
}

template <typename T>
static int seg_dev(tb_seg *h, T **p, size_t n)
{
    int r = dev_alloc(p, std::max<size_t>(n, 1));
    if (r == TB_OK) h->dev_allocs.push_back((void *)*p);
    return r;
}

extern "C" int tb_seg_create(const tb_seg_config *cfg, tb_seg **out)
{
    TB_REQUIRE(cfg && out, TB_ERR_INVALID, "tb_seg_create: null argument");
    TB_REQUIRE(cfg->width > 0 && cfg->height > 0 && cfg->width <= 16384 && cfg->height < 65535, TB_ERR_INVALID,
               "tb_seg_create: frame size must be 1..16384 x 1..65534");
    TB_REQUIRE(cfg->max_batch > 0 && cfg->max_batch <= 4096, TB_ERR_INVALID, "tb_seg_create: max_batch must be 1..4096");
    TB_REQUIRE(cfg->crop_method >= 0 && cfg->crop_method <= 2, TB_ERR_INVALID, "tb_seg_create: crop_method must be 0..2");
    TB_REQUIRE(cfg->channels == 0 || cfg->channels == 1 || cfg->channels == 3 || cfg->channels == 4, TB_ERR_INVALID,
               "tb_seg_create: channels must be 1 (gray), 3 (BGR) or 4 (BGRA)");
    TB_REQUIRE(cfg->encoding >= 0 && cfg->encoding <= 2, TB_ERR_INVALID, "tb_seg_create: encoding must be 0 (gray), 1 (rgb8) or 2 (r3g3b2)");
    TB_REQUIRE(cfg->crop_normalize == 0 || (cfg->crop_normalize >= 1 && cfg->crop_normalize <= 3 && cfg->encoding == 0), TB_ERR_INVALID,
               "tb_seg_create: crop_normalize must be 0 (none), 1 (moments), 2 (posture) or 3 (legacy); 1..3 are built for the gray encoding");
    TB_REQUIRE(cfg->encoding == 0 || cfg->channels >= 3, TB_ERR_INVALID,
               "tb_seg_create: rgb8 / r3g3b2 encoding needs colour frames (Invalid number of channels, BackgroundSubtraction.cpp:151-158,177-181)");
    TB_REQUIRE(cfg->encoding != 2 || cfg->crop_normalize == 0, TB_ERR_INVALID, "tb_seg_create: crop normalisation is built for the gray encoding");
    TB_REQUIRE(cfg->crop_scale >= 0.f && cfg->crop_scale <= 16.f, TB_ERR_INVALID, "tb_seg_create: crop_scale (individual_image_scale) must be in (0, 16]; 0 = 1");
    TB_REQUIRE(cfg->crop_scale == 0.f || cfg->crop_scale == 1.f || (cfg->encoding == 0 && cfg->crop_normalize == 0), TB_ERR_INVALID,
               "tb_seg_create: crop_scale != 1 is built for the gray encoding without crop normalisation");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_error("tb_seg_create: no CUDA device (there is no CPU fallback)");
        return TB_ERR_CUDA;
    }
    TB_REQUIRE(cfg->device >= 0 && cfg->device < ndev, TB_ERR_INVALID, "tb_seg_create: bad device ordinal");
    TB_CUDA(cudaSetDevice(cfg->device));
    tb_seg *h = new tb_seg();
    h->cfg = *cfg;
    tb_seg_default_params(&h->params);
    std::string why;
    seg_make_k(h->params, h->k, why);
    SegDev &d = h->d;
    d.W = cfg->width; d.H = cfg->height; d.B = cfg->max_batch;
    d.CN = cfg->channels > 1 ? cfg->channels : 1; d.r3 = cfg->encoding == 2; d.enc = cfg->encoding == 1; d.opx = d.enc ? 3 : 1;
    d.cc = d.r3 ? CC_R3G3B2 : -1; d.cpx = (d.enc || d.r3) ? 3 : 1;
    d.cpr = (d.W + 15) / 16;
    d.aligned = (d.W % 16) == 0;
    d.wpr = (d.cpr + 3) / 4 + 1;
    const int unit_chunks = d.CN == 1 ? K1_CHUNKS : 512;                         // colour units hold half the pixels (seg_rle_ws_kernel)
    d.rpt = std::max(1, std::min((d.cpr <= unit_chunks ? unit_chunks : K1_CHUNKS) / d.cpr, K1W_WORDS / d.wpr));   // rows per band: fits the chunk tile and the padded bit image
    d.n_bands = (d.H + d.rpt - 1) / d.rpt;
    d.rcap = cfg->max_runs_per_frame > 0 ? (uint32_t)cfg->max_runs_per_frame : 32768u;
    d.rcap = std::max<uint32_t>(d.rcap, (uint32_t)d.n_bands);
    d.px_frame_cap = cfg->max_pixels_per_frame > 0 ? (uint32_t)cfg->max_pixels_per_frame
                                                   : (uint32_t)std::max<int64_t>((int64_t)d.W * d.H / 8, std::min<int64_t>((int64_t)d.W * d.H, 1 << 18)) * (uint32_t)d.opx;
    d.max_crops = cfg->max_crops_per_frame > 0 ? (uint32_t)cfg->max_crops_per_frame : 0u;
    d.crop_w = cfg->crop_width > 0 ? cfg->crop_width : 80;
    d.crop_h = cfg->crop_height > 0 ? cfg->crop_height : 80;
    d.crop_method = cfg->crop_method;
    d.crop_norm = cfg->crop_normalize;
    d.crop_scale = cfg->crop_scale > 0.f ? cfg->crop_scale : 1.f;
    const size_t B = d.B;
    // batch arenas: an average budget per frame, but never less than one worst-case frame
    d.lines_cap = (uint32_t)std::min<uint64_t>(std::max<uint64_t>((uint64_t)B * std::min<uint32_t>(d.rcap, 8192u), d.rcap), 0x7FFFFFFFu);
    d.blobs_cap = (uint32_t)std::min<uint64_t>(std::max<uint64_t>((uint64_t)B * std::min<uint32_t>(d.rcap, 2048u), d.rcap), 0x7FFFFFFFu);
    d.px_cap = (uint32_t)std::min<uint64_t>((uint64_t)B * d.px_frame_cap, 0x7FFFFFFFu);
    d.crops_cap = (uint32_t)(B * d.max_crops);
    d.sqcm = 1.f; d.n_ranges = 0;

    int r = TB_OK;
#define A(p, n) if (r == TB_OK) r = seg_dev(h, &(p), (n))
    A(h->d_bg, (size_t)d.W * d.H + 16);
    A(h->d_frames, B * d.W * d.H * d.CN + 16);
    A(h->d_tmp, (size_t)d.W * d.H * d.opx);
    if (d.enc) A(h->d_bg3, (size_t)d.W * d.H * 3 + 16);
    A(d.run_count, B + 1); A(d.band_base, B * d.n_bands); A(d.band_cnt, B * d.n_bands);
    A(d.runs_raw, B * d.rcap); A(d.runs, B * d.rcap); A(d.parent, B * d.rcap); A(d.bidx, B * d.rcap);
    A(d.row_start, B * d.H); A(d.row_end, B * d.H);
    A(d.b_npx, B * d.rcap); A(d.b_nl, B * d.rcap); A(d.b_xmin, B * d.rcap); A(d.b_xmax, B * d.rcap);
    A(d.b_ymax, B * d.rcap); A(d.b_root, B * d.rcap); A(d.b_loff, B * d.rcap); A(d.b_poff, B * d.rcap);
    A(d.kept, B * d.rcap); A(d.frame_tot, B * 4);
    {   // per-frame headers, the identification outputs and the blob records share ONE block laid out as the multi-GPU metadata
        // unit (tb_meta_layout): K2 / K3 and the identification head write it in place, the all-gather reads its prefix -- no packing
        const size_t kb = (size_t)B * d.max_crops;
        h->meta.batch = (uint32_t)B; h->meta.kmax = d.max_crops;
        h->meta.off_infos = 0;
        h->meta.off_top_id = (B * sizeof(tb_frame_info) + 31) / 32 * 32;
        h->meta.off_top_p = (h->meta.off_top_id + kb * 4 + 31) / 32 * 32;
        h->meta.off_recs = (h->meta.off_top_p + kb * 4 + 31) / 32 * 32;
        h->meta.gather_bytes = h->meta.off_recs + kb * sizeof(tb_blob_rec);
        const size_t total = h->meta.off_recs + (size_t)d.blobs_cap * sizeof(tb_blob_rec);
        A(h->d_meta, total);
        if (r == TB_OK) {
            if (cudaMemset(h->d_meta, 0, total) != cudaSuccess) { set_error("cudaMemset failed"); r = TB_ERR_CUDA; }
            h->meta.base = h->d_meta;
            d.infos = (tb_frame_info *)(h->d_meta + h->meta.off_infos);
            d.recs = (tb_blob_rec *)(h->d_meta + h->meta.off_recs);
        }
    }
    A(d.lines, d.lines_cap); A(d.line_px, d.lines_cap);
    A(d.pixels, (size_t)d.px_cap + 16); A(d.crops, (size_t)d.crops_cap * d.crop_w * d.crop_h * d.cpx + 16);
    A(d.crop_blob, d.crops_cap); A(d.totals, 4);
    if (d.crop_norm) A(h->d_coef, (size_t)d.crops_cap * 6);
#undef A
    if (r == TB_OK) r = host_alloc(&h->h_infos, B);
    if (r == TB_OK) r = host_alloc(&h->h_totals, 4);
    if (r == TB_OK) r = host_alloc(&h->h_recs, d.blobs_cap);
    if (r == TB_OK) r = host_alloc(&h->h_lines, d.lines_cap);
    if (r == TB_OK) r = host_alloc(&h->h_pixels, (size_t)d.px_cap + 16);
    if (r == TB_OK) r = host_alloc(&h->h_crops, (size_t)d.crops_cap * d.crop_w * d.crop_h * d.cpx + 16);
    if (r == TB_OK) r = host_alloc(&h->h_crop_blob, std::max<size_t>(d.crops_cap, 1));
    if (r == TB_OK && cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("cudaStreamCreate failed"); r = TB_ERR_CUDA; }
    h->own_stream = h->stream;
    if (r == TB_OK && cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming) != cudaSuccess) { set_error("cudaEventCreate failed"); r = TB_ERR_CUDA; }
    if (r != TB_OK) { tb_seg_destroy(h); return r; }
    d.bg = h->d_bg; d.bg3 = h->d_bg3;
    std::memset(h->h_totals, 0, 16);
    *out = h;
    return TB_OK;
}

extern "C" void tb_seg_destroy(tb_seg *h)
{
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    cudaDeviceSynchronize();
    for (void *p : h->dev_allocs) cudaFree(p);
    void *hp[] = {h->h_infos, h->h_totals, h->h_recs, h->h_lines, h->h_pixels, h->h_crops, h->h_crop_blob,
                  h->h_o_recs, h->h_o_raw, h->h_o_res, h->h_o_totals, h->h_m_pts, h->h_m_segs, h->h_m_recs, h->h_m_nrecs, h->h_m_norm, h->h_crop_valid,
                  h->h_r_remaining};
    h->p_prof.destroy();
    for (void *p : hp) if (p) cudaFreeHost(p);
    h->prof.destroy();
    if (h->ev_done) cudaEventDestroy(h->ev_done);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
}

extern "C" int tb_seg_set_stream(tb_seg *h, void *stream)
{
    TB_REQUIRE(h, TB_ERR_INVALID, "tb_seg_set_stream: null handle");
    TB_REQUIRE(!h->pending, TB_ERR_STATE, "tb_seg_set_stream: a batch is pending (call tb_seg_wait first)");
    h->stream = stream ? (cudaStream_t)stream : h->own_stream;
    return TB_OK;
}

extern "C" int tb_seg_set_params(tb_seg *h, const tb_seg_params *p)
{
    TB_REQUIRE(h && p, TB_ERR_INVALID, "tb_seg_set_params: null argument");
    // 1. validate everything, 2. allocate what the new settings need, 3. only then touch the handle: a rejected call
    //    leaves the previous settings fully in place (the caller of update_settings may catch the error and go on)
    SegK k; std::string why;
    if (seg_make_k(*p, k, why) != TB_OK) { set_error("tb_seg_set_params: " + why); return TB_ERR_INVALID; }
    TB_REQUIRE(h->d.CN == 1 || p->color_channel < h->d.CN || p->color_channel >= 4, TB_ERR_INVALID, "tb_seg_set_params: color_channel beyond the frame's channels");
    TB_REQUIRE(!p->blur_difference || !h->d.enc, TB_ERR_INVALID, "tb_seg_set_params: blur_difference takes a 1-channel background (gray / r3g3b2 encoding)");
    TB_REQUIRE(!p->blur_difference || (h->d.W > 12 && h->d.H > 12), TB_ERR_INVALID, "tb_seg_set_params: blur_difference needs frames larger than its 25x25 window");
    TB_REQUIRE(!p->use_adaptive_threshold || (p->adaptive_threshold_scale >= 0.f && p->adaptive_threshold_scale <= 8.f), TB_ERR_INVALID,
               "tb_seg_set_params: adaptive_threshold_scale must be 0..8");
    TB_REQUIRE(p->open_size >= 0 && p->open_size <= 15, TB_ERR_INVALID, "tb_seg_set_params: open_size must be 0..15");
    const bool morph = p->use_closing || p->dilation_size != 0 || p->blur_difference || p->use_adaptive_threshold || p->open_size > 1;
    if (morph) {
        TB_CUDA(cudaSetDevice(h->cfg.device));
        const size_t bytes = (size_t)h->cfg.max_batch * h->d.W * h->d.H;
        if (!h->m_a || !h->m_b || !h->m_diff) {
            int r = TB_OK;
            if (!h->m_a) r = seg_dev(h, &h->m_a, bytes + 16);
            if (r == TB_OK && !h->m_b) r = seg_dev(h, &h->m_b, bytes + 16);
            if (r == TB_OK && !h->m_diff) r = seg_dev(h, &h->m_diff, bytes + 16);
            if (r != TB_OK) return r;
        }
        if ((p->blur_difference || p->use_adaptive_threshold) && !h->box_hs) {
            const int sub = std::min(h->cfg.max_batch, 8);
            int r = seg_dev(h, &h->box_hs, (size_t)sub * h->d.W * h->d.H);
            if (r != TB_OK) return r;
            h->box_sub = sub;
        }
    }
    // success path
    h->params = *p; h->k = k; h->morph = morph;
    h->d.cc = h->d.r3 ? CC_R3G3B2       // r3g3b2 ignores color_channel (BackgroundSubtraction.cpp:151-158)
              : ((h->d.CN > 1 && h->d.enc == 0 && p->color_channel >= 0 && p->color_channel < 4) ? p->color_channel : -1);   // >= 4: cvtColor (:161-163)
    if (p->use_closing) h->el_close = ellipse_element(p->closing_size);
    if (p->dilation_size) h->el_dil = ones_element(std::abs(p->dilation_size));
    if (p->open_size > 1) h->el_open = ones_element(p->open_size);
    h->d.sqcm = p->cm_per_pixel * p->cm_per_pixel;        // SQR(cm_per_pixel) in float, BackgroundSubtraction.cpp:139
    h->d.n_ranges = p->n_size_ranges;
    for (int i = 0; i < 4; ++i) { h->d.lo[i] = p->size_lo[i]; h->d.hi[i] = p->size_hi[i]; }
    return TB_OK;
}

extern "C" int tb_seg_set_background_c(tb_seg *h, const uint8_t *bg, int width, int height, int channels, int64_t stride)
{
    TB_REQUIRE(h && bg, TB_ERR_INVALID, "tb_seg_set_background: null argument");
    TB_REQUIRE(width == h->d.W && height == h->d.H, TB_ERR_INVALID, "tb_seg_set_background: size differs from the handle's frame size");
    TB_REQUIRE(channels == (h->d.enc ? 3 : 1), TB_ERR_INVALID,
               "tb_seg_set_background: the background must have 1 channel for gray and 3 for rgb8 encoding (RawProcessing.cpp:343)");
    TB_CUDA(cudaSetDevice(h->cfg.device));
    const size_t row = (size_t)width * channels;
    if (stride <= 0) stride = (int64_t)row;
    if (channels == 1) {
        TB_CUDA(cudaMemcpy2DAsync(h->d_bg, row, bg, (size_t)stride, row, (size_t)height, cudaMemcpyHostToDevice, h->stream));
    } else {                  // _grey_average = cvtColor(average, BGR2GRAY), RawProcessing.cpp:356-357
        TB_CUDA(cudaMemcpy2DAsync(h->d_bg3, row, bg, (size_t)stride, row, (size_t)height, cudaMemcpyHostToDevice, h->stream));
        to_gray_kernel<<<296, 256, 0, h->stream>>>(h->d_bg3, h->d_bg, nullptr, (size_t)width * height, 3, -1);
        h->launches += 1;
        TB_CUDA(cudaGetLastError());
    }
    TB_CUDA(cudaStreamSynchronize(h->stream));
    h->has_bg = true;
    return TB_OK;
}

extern "C" int tb_seg_set_background(tb_seg *h, const uint8_t *bg, int width, int height, int64_t stride)
{
    return tb_seg_set_background_c(h, bg, width, height, 1, stride);
}

// grey plane (and, for rgb8, non-zero plane) of n colour frames into the handle's plane buffers
static int seg_to_gray(tb_seg *h, const uint8_t *frames_dev, int n, cudaStream_t s, int tracker_formula = 0)
{
    const size_t px = (size_t)h->d.W * h->d.H;
    if (!h->d_gray) {
        int r = seg_dev(h, &h->d_gray, (size_t)h->cfg.max_batch * px + 16);
        if (r == TB_OK && h->d.enc) r = seg_dev(h, &h->d_nz, (size_t)h->cfg.max_batch * px + 16);
        if (r != TB_OK) return r;
    }
    to_gray_kernel<<<148 * 8, 256, 0, s>>>(frames_dev, h->d_gray, h->d.enc ? h->d_nz : nullptr, px * (size_t)n, h->d.CN, h->d.cc, tracker_formula);
    h->launches += 1;
    TB_CUDA(cudaGetLastError());
    h->gray_valid = true;
    return TB_OK;
}

// threshold mask -> closing -> dilation / erosion for n frames; returns the buffer holding the final mask
static int seg_morph(tb_seg *h, const uint8_t *frames_dev, int n, cudaStream_t s, const uint8_t **mask_out)
{
    const tb_seg_params &p = h->params;
    const size_t px = (size_t)h->d.W * h->d.H, total = px * (size_t)n;
    const int grid = 148 * 8, nt = 256;
    uint8_t *cur = h->m_a, *tmp = h->m_b;
    if (p.blur_difference) {               // RawProcessing.cpp:371-387: nothing else of the mask pipeline applies
        diff_tozero_kernel<<<grid, nt, 0, s>>>(frames_dev, h->d_bg, cur, px, total, h->k);
        int r = launch_box_mean(cur, tmp, h->box_hs, h->box_sub, h->d.W, h->d.H, n, 25, 1, s);
        if (r != TB_OK) return r;
        gt_mask_kernel<<<grid, nt, 0, s>>>(tmp, cur, total, h->k.t4 & 0xFFu);
        h->launches += 2 + 2 * (uint64_t)((n + h->box_sub - 1) / h->box_sub);
        TB_CUDA(cudaGetLastError());
        *mask_out = cur;
        return TB_OK;
    }
    thresh_mask_kernel<<<grid, nt, 0, s>>>(frames_dev, h->d_bg, cur, (p.dilation_size < 0 || p.use_adaptive_threshold) ? h->m_diff : nullptr, px, total, h->k);
    h->launches += 1;
    if (p.use_adaptive_threshold) {        // neighbourhood = int(cols * adaptive_threshold_scale), odd, >= 3 (:427-434)
        int nb = (int)((float)h->d.W * p.adaptive_threshold_scale);
        if (nb % 2 == 0) nb++;
        if (nb < 3) nb = 3;
        int r = launch_box_mean(h->m_diff, tmp, h->box_hs, h->box_sub, h->d.W, h->d.H, n, nb, 0, s);
        if (r != TB_OK) return r;
        adaptive_mask_kernel<<<grid, nt, 0, s>>>(h->m_diff, tmp, cur, total, p.detect_threshold);
        h->launches += 1 + 2 * (uint64_t)((n + h->box_sub - 1) / h->box_sub);
    }
    if (p.open_size > 1) {                 // optional n x n open of the threshold mask (north_star; not a reference stage): erode, dilate
        morph_kernel<<<grid, nt, 0, s>>>(cur, tmp, h->d.W, h->d.H, total, h->el_open, 0);
        morph_kernel<<<grid, nt, 0, s>>>(tmp, cur, h->d.W, h->d.H, total, h->el_open, 1);
        h->launches += 2;
    }
    auto closing = [&]() {
        morph_kernel<<<grid, nt, 0, s>>>(cur, tmp, h->d.W, h->d.H, total, h->el_close, 1);
        morph_kernel<<<grid, nt, 0, s>>>(tmp, cur, h->d.W, h->d.H, total, h->el_close, 0);
        h->launches += 2;
    };
    if (p.use_closing) closing();
    if (p.dilation_size > 0) {
        morph_kernel<<<grid, nt, 0, s>>>(cur, tmp, h->d.W, h->d.H, total, h->el_dil, 1);
        std::swap(cur, tmp);
        h->launches += 1;
    } else if (p.dilation_size < 0) {
        morph_kernel<<<grid, nt, 0, s>>>(cur, tmp, h->d.W, h->d.H, total, h->el_dil, 0);
        remask_kernel<<<grid, nt, 0, s>>>(tmp, h->m_diff, cur, total, h->k.t4 & 0xFFu);
        h->launches += 2;
        if (p.use_closing) closing();
    }
    TB_CUDA(cudaGetLastError());
    *mask_out = cur;
    return TB_OK;
}

static int seg_launch(tb_seg *h, const uint8_t *frames_dev, int n, cudaStream_t s, int fetch, const uint8_t *keep_mask = nullptr, uint32_t min_payload = 0)
{
    SegDev d = h->d;
    d.B = n;
    d.bg_stride = 0;
    d.keep_mask = nullptr;
    d.min_payload = 0;
    d.nz_plane = nullptr;
    h->last_frames_dev = frames_dev;
    h->gray_valid = false;
    TB_CUDA(cudaMemsetAsync(d.run_count, 0, sizeof(uint32_t) * ((size_t)n + 1), s));      // + the unit counter of the persistent K1
    static const int fpc_env = getenv("TB_SEG_FPC") ? atoi(getenv("TB_SEG_FPC")) : 0;
    const int fpc = fpc_env > 0 ? fpc_env : (n >= 64 ? 8 : (n >= 8 ? 2 : 1));
    dim3 g1((unsigned)d.n_bands, (unsigned)((n + fpc - 1) / fpc));
    const int slot = h->prof.begin(s);
    h->prof.mark(slot, 0);
    static const bool no_tma = getenv("TB_SEG_NO_TMA") != nullptr;      // bring-up switches
    static const bool no_ws = getenv("TB_SEG_NO_WS") != nullptr;
    const bool plain = !h->morph && h->k.flags == (F_DIFF | F_ABS) && (h->k.t4 & 0xFFu) <= 127u;
    SegK kk = h->k;
    if (plain) kk.lo4 = (127u - (kk.t4 & 0xFFu)) * 0x01010101u;      // SWAR addend of the fast path
    const bool ws_ok = d.aligned && d.wpr <= K1W_WORDS && n <= 65536 && d.n_bands < 65535 && !no_tma && !no_ws;
    // colour frames: fused cvtColor inside the persistent K1 for the default settings, else a grey plane first
    const bool fused_colour = d.CN > 1 && plain && !keep_mask && d.cc < 0 && ws_ok && d.rpt * d.cpr <= 512;
    const uint8_t *plane = frames_dev;          // what the 1-channel K1 variants read
    if (d.CN > 1 && !fused_colour) {
        // tracker-side re-threshold of rgb8 blobs compares cmn::bgr2gray of the pixel (not cvtColor) with the background's grey image
        int r = seg_to_gray(h, frames_dev, n, s, keep_mask && d.enc ? 1 : 0);
        if (r != TB_OK) return r;
        plane = h->d_gray;
        if (d.enc) d.nz_plane = h->d_nz;
    }
    if (fused_colour) {
        static DeviceOnce ctas_once[2];
        const int ci = d.CN == 3 ? 0 : 1;
        const int nt = (K1W_MW + 3 + 1) * 32, smem = k1w_smem(3, d.CN);
        if (ctas_once[ci].need()) {             // function attributes are per device
            if (d.CN == 3) {
                TB_CUDA(cudaFuncSetAttribute(seg_rle_ws_kernel<false, 3, 3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
                TB_CUDA(cudaFuncSetAttribute(seg_rle_ws_kernel<false, 3, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            } else {
                TB_CUDA(cudaFuncSetAttribute(seg_rle_ws_kernel<false, 3, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
                TB_CUDA(cudaFuncSetAttribute(seg_rle_ws_kernel<false, 3, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            }
            ctas_once[ci].done();
        }
        if (!h->col_ctas) {                     // per handle: the grid follows the handle's own device
            int per_sm = 0, sms = 0;
            if (d.CN == 3) TB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, seg_rle_ws_kernel<false, 3, 3, true>, nt, smem));
            else TB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, seg_rle_ws_kernel<false, 3, 4, true>, nt, smem));
            TB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->cfg.device));
            h->col_ctas = std::max(1, per_sm) * std::max(1, sms);
        }
        const unsigned units = (unsigned)d.n_bands * (unsigned)n;
        const unsigned grid = std::min<unsigned>(units, (unsigned)h->col_ctas);
        const uint32_t static_units = (uint32_t)((double)(units / grid) * 0.5);
        if (d.CN == 3 && !d.enc) seg_rle_ws_kernel<false, 3, 3, false><<<grid, nt, smem, s>>>(frames_dev, d, kk, static_units, nullptr);
        else if (d.CN == 3) seg_rle_ws_kernel<false, 3, 3, true><<<grid, nt, smem, s>>>(frames_dev, d, kk, static_units, nullptr);
        else if (!d.enc) seg_rle_ws_kernel<false, 3, 4, false><<<grid, nt, smem, s>>>(frames_dev, d, kk, static_units, nullptr);
        else seg_rle_ws_kernel<false, 3, 4, true><<<grid, nt, smem, s>>>(frames_dev, d, kk, static_units, nullptr);
    } else if (keep_mask) {                 // tracker-side re-threshold: comparison >=, restricted to the painted detection blobs
        TB_REQUIRE(!h->morph, TB_ERR_INVALID, "tb_seg_rethreshold: the tracker-side handle must not enable morphology");
        d.keep_mask = keep_mask;
        d.min_payload = min_payload;
        kk = h->k; kk.flags |= F_GE;
        seg_rle_kernel<true><<<g1, K1_NT, 0, s>>>(plane, d, kk, fpc);
    } else if (h->morph) {
        const uint8_t *mask = nullptr;
        int r = seg_morph(h, plane, n, s, &mask);
        if (r != TB_OK) return r;
        d.bg = mask; d.bg_stride = (size_t)d.W * d.H;
        kk.flags = F_PREMASK;
        seg_rle_kernel<true><<<g1, K1_NT, 0, s>>>(plane, d, kk, fpc);
    } else if (d.nz_plane) {          // rgb8 on the plane path: only the plain kernel reads the non-zero plane
        seg_rle_kernel<true><<<g1, K1_NT, 0, s>>>(plane, d, h->k, fpc);
    } else if (ws_ok && d.rpt * d.cpr <= K1_CHUNKS) {
        static const int ew = getenv("TB_SEG_EW") ? atoi(getenv("TB_SEG_EW")) : 3;                   // tuning knobs
        static const double static_frac = getenv("TB_SEG_STATIC") ? atof(getenv("TB_SEG_STATIC")) : 0.5;
        static DeviceOnce ws_once;
        const int nt = (K1W_MW + (ew == 4 ? 4 : 3) + 1) * 32, smem = k1w_smem(ew == 4 ? 4 : 3);
        if (ws_once.need()) {                   // function attributes are per device
            TB_CUDA(cudaFuncSetAttribute(seg_rle_ws_kernel<false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, k1w_smem(3)));
            TB_CUDA(cudaFuncSetAttribute(seg_rle_ws_kernel<true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, k1w_smem(3)));
            TB_CUDA(cudaFuncSetAttribute(seg_rle_ws_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, k1w_smem(4)));
            ws_once.done();
        }
        if (!h->ws_ctas) {                      // per handle: the grid follows the handle's own device
            int per_sm = 0, sms = 0;
            if (ew == 4) TB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, seg_rle_ws_kernel<false, 4>, nt, smem));
            else TB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, seg_rle_ws_kernel<false, 3>, nt, smem));
            TB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->cfg.device));
            h->ws_ctas = std::max(1, per_sm) * std::max(1, sms);
        }
        const int ws_ctas = h->ws_ctas;
        const unsigned units = (unsigned)d.n_bands * (unsigned)n;
        const unsigned grid = std::min<unsigned>(units, (unsigned)ws_ctas);
        const uint32_t static_units = (uint32_t)((double)(units / grid) * static_frac);
        static const bool timeline = getenv("TB_SEG_TIMELINE") != nullptr;                            // debug: per-CTA start / end times
        if (timeline && !h->dbg) { int r = seg_dev(h, &h->dbg, (size_t)4 * ws_ctas); if (r != TB_OK) return r; }
        unsigned long long *dbg = h->dbg;
        if (plain && ew == 4) seg_rle_ws_kernel<false, 4><<<grid, nt, smem, s>>>(plane, d, kk, static_units, dbg);
        else if (plain) seg_rle_ws_kernel<false, 3><<<grid, nt, smem, s>>>(plane, d, kk, static_units, dbg);
        else seg_rle_ws_kernel<true, 3><<<grid, (K1W_MW + 3 + 1) * 32, k1w_smem(3), s>>>(plane, d, h->k, static_units, dbg);
        if (timeline) {
            std::vector<unsigned long long> t(4 * grid);
            TB_CUDA(cudaStreamSynchronize(s));
            TB_CUDA(cudaMemcpy(t.data(), dbg, sizeof(unsigned long long) * 4 * grid, cudaMemcpyDeviceToHost));
            unsigned long long s0 = ~0ull, s1 = 0, p0 = ~0ull, p1 = 0, e0 = ~0ull, e1 = 0;
            for (unsigned i = 0; i < grid; ++i) {
                s0 = std::min(s0, t[4 * i]); s1 = std::max(s1, t[4 * i]);
                p0 = std::min(p0, t[4 * i + 1]); p1 = std::max(p1, t[4 * i + 1]);
                e0 = std::min(e0, t[4 * i + 2]); e1 = std::max(e1, t[4 * i + 2]);
            }
            fprintf(stderr, "[K1 timeline] grid %u static %u: CTA start spread %.2f us; producer done %.2f..%.2f us; extraction done %.2f..%.2f us after first start\n",
                    grid, static_units, (s1 - s0) * 1e-3, (p0 - s0) * 1e-3, (p1 - s0) * 1e-3, (e0 - s0) * 1e-3, (e1 - s0) * 1e-3);
        }
    } else if (d.aligned && !no_tma) {
        static DeviceOnce attr_done;
        if (attr_done.need()) {
            TB_CUDA(cudaFuncSetAttribute(seg_rle_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, K1T_SMEM));
            TB_CUDA(cudaFuncSetAttribute(seg_rle_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, K1T_SMEM));
            attr_done.done();
        }
        if (plain) seg_rle_tma_kernel<false><<<g1, K1T_NT, K1T_SMEM, s>>>(plane, d, kk, fpc);
        else seg_rle_tma_kernel<true><<<g1, K1T_NT, K1T_SMEM, s>>>(plane, d, h->k, fpc);
    } else {
        if (plain) seg_rle_kernel<false><<<g1, K1_NT, 0, s>>>(plane, d, h->k, fpc);
        else seg_rle_kernel<true><<<g1, K1_NT, 0, s>>>(plane, d, h->k, fpc);
    }
    h->prof.mark(slot, 1);
    static DeviceOnce k2_attr;
    if (k2_attr.need()) { TB_CUDA(cudaFuncSetAttribute(ccl_label_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, K2F_SMEM)); k2_attr.done(); }
    ccl_label_kernel<<<n, K2F_NT, K2F_SMEM, s>>>(d);
    h->prof.mark(slot, 2);
    d.bg = h->d_bg; d.bg_stride = 0; d.keep_mask = nullptr; d.nz_plane = nullptr;   // crops difference against the real background
    // CTAs per frame: enough to fill the GPU's warp slots (6 resident CTAs per SM) at small batches
    static const int k3_env = getenv("TB_SEG_K3_SPLIT") ? atoi(getenv("TB_SEG_K3_SPLIT")) : 0;
    const int k3_split = k3_env > 0 ? k3_env : std::min(16, std::max(K3_SPLIT, (148 * 8 + n - 1) / n));
    blob_emit_kernel<<<dim3((unsigned)n, (unsigned)k3_split), K3_NT, 0, s>>>(frames_dev, d);
    h->prof.mark(slot, 3);
    h->launches += 3;
    TB_CUDA(cudaGetLastError());
    if (d.crop_norm == 1 && d.max_crops) {     // individual_image_normalization = moments: orientation + cv::warpAffine per crop
        int r = launch_crop_moments(d.recs, d.totals, d.crop_blob, d.lines, d.line_px, d.pixels, h->d_bg, d.W, d.crop_method,
                                    d.crop_w, d.crop_h, d.crops, h->d_coef, n * (int)d.max_crops, s);
        if (r != TB_OK) return r;
        h->launches += 2;
    }
    if (d.crop_scale != 1.f && d.max_crops) {   // individual_image_scale: nearest-neighbour resize of the blob image before the pad / crop
        int r = launch_crop_scaled(d.recs, d.totals, d.crop_blob, d.lines, d.line_px, d.pixels, h->d_bg, d.W, d.crop_method,
                                   d.crop_w, d.crop_h, d.crop_scale, d.crops, n * (int)d.max_crops, s);
        if (r != TB_OK) return r;
        h->launches += 1;
    }
    TB_CUDA(cudaMemcpyAsync(h->h_totals, d.totals, 16, cudaMemcpyDeviceToHost, s));
    TB_CUDA(cudaMemcpyAsync(h->h_infos, d.infos, sizeof(tb_frame_info) * (size_t)n, cudaMemcpyDeviceToHost, s));
    h->last_n = n; h->last_fetch = fetch; h->pending = true; h->fetched_payload = false; h->fetched_crops = false; h->last_stream = s;
    h->o_n = 0; h->m_n = 0; h->p_n = 0; h->p_pending = false;      // posture results belong to the previous batch
    return TB_OK;
}

extern "C" int tb_seg_submit_device(tb_seg *h, const void *frames_dev, int n, void *stream, int fetch)
{
    TB_REQUIRE(h && frames_dev, TB_ERR_INVALID, "tb_seg_submit_device: null argument");
    TB_REQUIRE(n > 0 && n <= h->cfg.max_batch, TB_ERR_INVALID, "tb_seg_submit_device: n must be 1..max_batch");
    TB_REQUIRE(h->has_bg, TB_ERR_STATE, "tb_seg_submit_device: no background set (pipeline is paused until set_background)");
    TB_REQUIRE(((uintptr_t)frames_dev & 15) == 0 || !h->d.aligned, TB_ERR_INVALID, "tb_seg_submit_device: frames must be 16-byte aligned");
    TB_CUDA(cudaSetDevice(h->cfg.device));
    return seg_launch(h, (const uint8_t *)frames_dev, n, stream ? (cudaStream_t)stream : h->stream, fetch);
}

extern "C" int tb_seg_submit(tb_seg *h, const uint8_t *const *frames, int n, int64_t stride, int fetch)
{
    TB_REQUIRE(h && frames, TB_ERR_INVALID, "tb_seg_submit: null argument");
    TB_REQUIRE(n > 0 && n <= h->cfg.max_batch, TB_ERR_INVALID, "tb_seg_submit: n must be 1..max_batch");
    TB_REQUIRE(h->has_bg, TB_ERR_STATE, "tb_seg_submit: no background set (pipeline is paused until set_background)");
    TB_CUDA(cudaSetDevice(h->cfg.device));
    const size_t W = (size_t)h->d.W * h->d.CN, H = h->d.H;       // bytes per row
    if (stride <= 0) stride = (int64_t)W;
    bool contiguous = (size_t)stride == W;
    for (int i = 0; i < n; ++i) {
        TB_REQUIRE(frames[i], TB_ERR_INVALID, "tb_seg_submit: null frame pointer");
        if (i && frames[i] != frames[i - 1] + W * H) contiguous = false;
    }
    if (contiguous) {          // packed batch: one copy instead of n
        TB_CUDA(cudaMemcpyAsync(h->d_frames, frames[0], (size_t)n * W * H, cudaMemcpyHostToDevice, h->stream));
    } else {
        for (int i = 0; i < n; ++i)
            TB_CUDA(cudaMemcpy2DAsync(h->d_frames + (size_t)i * W * H, W, frames[i], (size_t)stride, W, H, cudaMemcpyHostToDevice, h->stream));
    }
    return seg_launch(h, h->d_frames, n, h->stream, fetch);
}

extern "C" int tb_seg_wait(tb_seg *h)
{
    TB_REQUIRE(h, TB_ERR_INVALID, "tb_seg_wait: null handle");
    TB_REQUIRE(h->pending || h->last_n > 0, TB_ERR_STATE, "tb_seg_wait: nothing submitted");
    if (!h->pending) return TB_OK;
    TB_CUDA(cudaSetDevice(h->cfg.device));
    cudaStream_t s = h->last_stream;
    TB_CUDA(cudaStreamSynchronize(s));
    if (h->last_fetch) {
        const uint32_t *t = h->h_totals;
        const SegDev &d = h->d;
        if (t[0]) TB_CUDA(cudaMemcpyAsync(h->h_recs, d.recs, sizeof(tb_blob_rec) * (size_t)t[0], cudaMemcpyDeviceToHost, s));
        if (t[1]) TB_CUDA(cudaMemcpyAsync(h->h_lines, d.lines, sizeof(tb_line) * (size_t)t[1], cudaMemcpyDeviceToHost, s));
        if (t[2]) TB_CUDA(cudaMemcpyAsync(h->h_pixels, d.pixels, (size_t)t[2], cudaMemcpyDeviceToHost, s));
        if (t[3] && h->last_fetch >= 2) {
            TB_CUDA(cudaMemcpyAsync(h->h_crops, d.crops, (size_t)t[3] * d.crop_w * d.crop_h * d.cpx, cudaMemcpyDeviceToHost, s));
            TB_CUDA(cudaMemcpyAsync(h->h_crop_blob, d.crop_blob, sizeof(uint32_t) * (size_t)t[3], cudaMemcpyDeviceToHost, s));
        }
        TB_CUDA(cudaStreamSynchronize(s));
        h->fetched_payload = true;
        h->fetched_crops = h->last_fetch >= 2;
    }
    h->pending = false;
    for (int i = 0; i < h->last_n; ++i)
        if (h->h_infos[i].status & ~4u) {
            set_error("tb_seg_wait: capacity exceeded for at least one frame (see tb_frame_info.status)");
            return TB_ERR_CAPACITY;
        }
    return TB_OK;
}

extern "C" int tb_seg_result(tb_seg *h, int i, tb_blob_view *out)
{
    TB_REQUIRE(h && out, TB_ERR_INVALID, "tb_seg_result: null argument");
    TB_REQUIRE(!h->pending && h->last_n > 0, TB_ERR_STATE, "tb_seg_result: call tb_seg_wait first");
    TB_REQUIRE(i >= 0 && i < h->last_n, TB_ERR_INVALID, "tb_seg_result: frame index out of range");
    TB_REQUIRE(h->fetched_payload, TB_ERR_STATE, "tb_seg_result: batch was submitted with fetch=0");
    const tb_frame_info &fi = h->h_infos[i];
    out->info = fi;
    out->recs = h->h_recs + fi.blob_begin;
    out->lines = h->h_lines + fi.line_begin;
    out->pixels = h->h_pixels + fi.px_begin;
    return TB_OK;
}

extern "C" int tb_seg_totals(tb_seg *h, uint32_t out[4])
{
    TB_REQUIRE(h && out, TB_ERR_INVALID, "tb_seg_totals: null argument");
    TB_REQUIRE(!h->pending && h->last_n > 0, TB_ERR_STATE, "tb_seg_totals: call tb_seg_wait first");
    std::memcpy(out, h->h_totals, 16);
    return TB_OK;
}

extern "C" int tb_seg_device_results(tb_seg *h, void **crops, void **n_crops_dev, void **crop_blob_index, void **recs, void **infos)
{
    TB_REQUIRE(h, TB_ERR_INVALID, "tb_seg_device_results: null handle");
    if (crops) *crops = h->d.crops;
    if (n_crops_dev) *n_crops_dev = h->d.totals + 3;
    if (crop_blob_index) *crop_blob_index = h->d.crop_blob;
    if (recs) *recs = h->d.recs;
    if (infos) *infos = h->d.infos;
    return TB_OK;
}

extern "C" int tb_seg_metadata(tb_seg *h, tb_meta_layout *out)
{
    TB_REQUIRE(h && out, TB_ERR_INVALID, "tb_seg_metadata: null argument");
    *out = h->meta;
    return TB_OK;
}

extern "C" int tb_seg_crops(tb_seg *h, const uint8_t **crops, const uint32_t **crop_blob_index, uint32_t *n)
{
    TB_REQUIRE(h && n, TB_ERR_INVALID, "tb_seg_crops: null argument");
    TB_REQUIRE(!h->pending && h->fetched_crops, TB_ERR_STATE, "tb_seg_crops: call tb_seg_wait after a fetch=2 submit first");
    if (crops) *crops = h->h_crops;
    if (crop_blob_index) *crop_blob_index = h->h_crop_blob;
    *n = h->h_totals[3];
    return TB_OK;
}

// ---- posture chain (N4): outlines -> midlines -> normalised midlines -> posture crops, all behind the batch's kernels
static int posture_alloc_crops(tb_seg *h)
{
    const SegDev &d = h->d;
    int r = TB_OK;
    if (d.crops_cap && !h->crop_valid) {
        r = seg_dev(h, &h->crop_valid, d.crops_cap);
        if (r == TB_OK) r = host_alloc(&h->h_crop_valid, d.crops_cap);
        if (r == TB_OK && !h->d_coef) r = seg_dev(h, &h->d_coef, (size_t)d.crops_cap * 6);
    }
    return r;
}

static int posture_alloc(tb_seg *h, int res)
{
    const SegDev &d = h->d;
    int r = TB_OK;
    if (!h->o_recs) {
        h->o_cap = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(1u << 20, (uint64_t)h->cfg.max_batch << 15), 1u << 26);
        static const size_t arena_env = getenv("TB_POSTURE_ARENA_MB") ? (size_t)atol(getenv("TB_POSTURE_ARENA_MB")) : 32;
        h->m_arena_floats = (unsigned long long)arena_env << 18;       // outlines too long for the shared-memory pool work here
        r = seg_dev(h, &h->o_visited, (size_t)d.px_cap / d.opx + 16);
        if (r == TB_OK) r = seg_dev(h, &h->o_rowfirst, d.lines_cap);
        if (r == TB_OK) r = seg_dev(h, &h->o_sel, d.blobs_cap);
        if (r == TB_OK) r = seg_dev(h, &h->o_recs, d.blobs_cap);
        if (r == TB_OK) r = seg_dev(h, &h->o_totals, 4);
        if (r == TB_OK) r = seg_dev(h, &h->o_raw, (size_t)h->o_cap * 2);
        if (r == TB_OK) r = seg_dev(h, &h->o_res, (size_t)h->o_cap * 2);
        if (r == TB_OK) r = seg_dev(h, &h->m_pts, (size_t)h->o_cap * 2);
        if (r == TB_OK) r = seg_dev(h, &h->m_segs, (size_t)h->o_cap * 4);
        if (r == TB_OK) r = seg_dev(h, &h->m_recs, d.blobs_cap);
        if (r == TB_OK) r = seg_dev(h, &h->m_nrecs, d.blobs_cap);
        if (r == TB_OK) r = seg_dev(h, &h->m_arena, (size_t)h->m_arena_floats);
        if (r == TB_OK) r = seg_dev(h, &h->m_arena_used, 1);
        if (r == TB_OK) r = host_alloc(&h->h_o_recs, d.blobs_cap);
        if (r == TB_OK) r = host_alloc(&h->h_o_raw, (size_t)h->o_cap * 2);
        if (r == TB_OK) r = host_alloc(&h->h_o_res, (size_t)h->o_cap * 2);
        if (r == TB_OK) r = host_alloc(&h->h_o_totals, 4);
        if (r == TB_OK) r = host_alloc(&h->h_m_pts, (size_t)h->o_cap * 2);
        if (r == TB_OK) r = host_alloc(&h->h_m_segs, (size_t)h->o_cap * 4);
        if (r == TB_OK) r = host_alloc(&h->h_m_recs, d.blobs_cap);
        if (r == TB_OK) r = host_alloc(&h->h_m_nrecs, d.blobs_cap);
        if (r == TB_OK) r = posture_alloc_crops(h);
        if (r == TB_OK) { h->p_status = h->o_totals + 2; TB_CUDA(cudaDeviceGetAttribute(&h->p_sms, cudaDevAttrMultiProcessorCount, h->cfg.device)); }
        if (r != TB_OK) return r;
    }
    if (res > h->m_res) {                      // normalised midlines: res float4 per blob (grown when midline_resolution grows)
        if (h->m_norm) { cudaFree(h->m_norm); h->dev_allocs.erase(std::find(h->dev_allocs.begin(), h->dev_allocs.end(), (void *)h->m_norm)); h->m_norm = nullptr; }
        if (h->h_m_norm) { cudaFreeHost(h->h_m_norm); h->h_m_norm = nullptr; }
        h->m_res = 0;
        r = seg_dev(h, &h->m_norm, (size_t)d.blobs_cap * res * 4);
        if (r == TB_OK) r = host_alloc(&h->h_m_norm, (size_t)d.blobs_cap * res * 4);
        if (r != TB_OK) return r;
        h->m_res = res;
    }
    return TB_OK;
}

static int posture_check_params(const tb_posture_params *p, const char *who)
{
    const std::string w(who);
    TB_REQUIRE(p->peak_mode == 0 || p->peak_mode == 1, TB_ERR_INVALID, w + ": peak_mode must be 0 (pointy) or 1 (broad)");
    TB_REQUIRE(p->outline_approximate >= 0 && p->outline_approximate <= 8, TB_ERR_INVALID, w + ": outline_approximate must be 0..8");
    TB_REQUIRE(p->outline_smooth_samples >= 0 && p->outline_smooth_samples <= 255 && p->outline_smooth_step >= 1 && p->outline_smooth_step <= 255,
               TB_ERR_INVALID, w + ": outline_smooth_samples 0..255, outline_smooth_step 1..255 (uint8 settings)");
    TB_REQUIRE(p->midline_resolution >= 2 && p->midline_resolution <= 256, TB_ERR_INVALID, w + ": midline_resolution must be 2..256");
    TB_REQUIRE(p->midline_stiff_percentage >= 0.f && p->midline_stiff_percentage <= 1.f, TB_ERR_INVALID, w + ": midline_stiff_percentage must be 0..1");
    return TB_OK;
}

// blob count of the batch: known on the host once tb_seg_wait has run, else read on the device (bounded by the record capacity)
static inline uint32_t posture_nb_max(const tb_seg *h) { return h->pending ? h->d.blobs_cap : h->h_totals[0]; }

static int posture_enqueue_outlines(tb_seg *h, float rd, cudaStream_t s)
{
    const SegDev &d = h->d;
    const uint32_t nb_max = posture_nb_max(h);
    TB_CUDA(cudaMemsetAsync(h->o_totals, 0, 16, s));
    int r = launch_outlines(d.recs, d.totals, nb_max, d.lines, d.line_px, d.opx, h->o_visited, (size_t)d.px_cap / d.opx + 16, rd,
                            h->o_rowfirst, h->o_sel, h->o_recs, h->o_totals, h->o_raw, h->o_res, h->o_cap, h->p_sms, s);
    if (r != TB_OK) return r;
    h->launches += nb_max ? 3 : 0;
    return TB_OK;
}

static int posture_enqueue_midlines(tb_seg *h, const tb_posture_params *p, int do_norm, const float *move_dir, const float *fix_len, cudaStream_t s)
{
    const uint32_t nb_max = posture_nb_max(h);
    int r = launch_midlines(h->o_recs, h->d.totals, nb_max, h->o_res, h->o_cap, p, do_norm, move_dir, fix_len, h->m_pts, h->m_segs, h->m_recs,
                            h->m_nrecs, h->m_norm, h->m_arena, h->m_arena_floats, h->m_arena_used, h->p_status, h->p_sms, s);
    if (r != TB_OK) return r;
    h->launches += nb_max ? 1 : 0;
    return TB_OK;
}

// after a stream synchronise: the arenas held everything?
static int posture_check_capacity(tb_seg *h, const char *who)
{
    const std::string w(who);
    TB_REQUIRE(h->h_o_totals[0] <= h->o_cap && h->h_o_totals[1] <= h->o_cap, TB_ERR_CAPACITY, w + ": the outline point arenas are too small for this batch");
    TB_REQUIRE(!(h->h_o_totals[2] & 1u), TB_ERR_CAPACITY, w + ": the work arena for long outlines is too small for this batch (TB_POSTURE_ARENA_MB)");
    return TB_OK;
}

extern "C" int tb_seg_outlines(tb_seg *h, float outline_resample)
{
    TB_REQUIRE(h, TB_ERR_INVALID, "tb_seg_outlines: null handle");
    TB_REQUIRE(!h->pending && h->last_n > 0, TB_ERR_STATE, "tb_seg_outlines: call tb_seg_wait on a submitted batch first");
    TB_REQUIRE(outline_resample < 255.f, TB_ERR_INVALID, "tb_seg_outlines: outline_resample must be < 255 (T/core/default_config.cpp:898)");
    TB_CUDA(cudaSetDevice(h->cfg.device));
    int r = posture_alloc(h, 0);
    if (r != TB_OK) return r;
    cudaStream_t s = h->last_stream ? h->last_stream : h->stream;
    const uint32_t nb = h->h_totals[0];
    h->o_n = 0; h->m_n = 0;
    if ((r = posture_enqueue_outlines(h, outline_resample, s)) != TB_OK) return r;
    TB_CUDA(cudaMemcpyAsync(h->h_o_totals, h->o_totals, 16, cudaMemcpyDeviceToHost, s));
    if (nb) TB_CUDA(cudaMemcpyAsync(h->h_o_recs, h->o_recs, sizeof(tb_outline_rec) * (size_t)nb, cudaMemcpyDeviceToHost, s));
    TB_CUDA(cudaStreamSynchronize(s));
    if ((r = posture_check_capacity(h, "tb_seg_outlines")) != TB_OK) return r;
    if (h->h_o_totals[0]) TB_CUDA(cudaMemcpyAsync(h->h_o_raw, h->o_raw, sizeof(float) * 2 * (size_t)h->h_o_totals[0], cudaMemcpyDeviceToHost, s));
    if (h->h_o_totals[1]) TB_CUDA(cudaMemcpyAsync(h->h_o_res, h->o_res, sizeof(float) * 2 * (size_t)h->h_o_totals[1], cudaMemcpyDeviceToHost, s));
    TB_CUDA(cudaStreamSynchronize(s));
    h->o_n = nb;
    return TB_OK;
}

extern "C" int tb_seg_outline_result(tb_seg *h, const tb_outline_rec **recs, const float **raw_points, const float **points, uint32_t *n_blobs)
{
    TB_REQUIRE(h && recs && raw_points && points && n_blobs, TB_ERR_INVALID, "tb_seg_outline_result: null argument");
    TB_REQUIRE(h->h_o_recs, TB_ERR_STATE, "tb_seg_outline_result: call tb_seg_outlines first");
    *recs = h->h_o_recs; *raw_points = h->h_o_raw; *points = h->h_o_res; *n_blobs = h->o_n;
    return TB_OK;
}

extern "C" void tb_posture_default_params(tb_posture_params *p)
{
    p->outline_smooth_samples = 4; p->outline_smooth_step = 1; p->outline_approximate = 3;
    p->outline_curvature_range_ratio = 0.03f; p->midline_walk_offset = 0.025f;
    p->peak_mode = 0; p->midline_start_with_head = 0; p->midline_invert = 0;
    p->midline_resolution = 25; p->midline_stiff_percentage = 0.15f;
}

extern "C" void tb_posture_default_request(tb_posture_request *r)
{
    std::memset(r, 0, sizeof(*r));
    tb_posture_default_params(&r->params);
    r->outline_resample = 1.f; r->normalize = 1; r->fetch = 1; r->individual_image_scale = 1.f;
}

extern "C" int tb_seg_midlines(tb_seg *h, const tb_posture_params *p)
{
    TB_REQUIRE(h && p, TB_ERR_INVALID, "tb_seg_midlines: null argument");
    TB_REQUIRE(h->h_o_recs && !h->pending && h->o_n == h->h_totals[0], TB_ERR_STATE, "tb_seg_midlines: call tb_seg_outlines on the batch first");
    int r = posture_check_params(p, "tb_seg_midlines");
    if (r != TB_OK) return r;
    TB_CUDA(cudaSetDevice(h->cfg.device));
    const uint32_t nb = h->o_n, total = h->h_o_totals[1];
    h->m_n = 0;
    cudaStream_t s = h->last_stream ? h->last_stream : h->stream;
    if ((r = posture_enqueue_midlines(h, p, 0, nullptr, nullptr, s)) != TB_OK) return r;
    TB_CUDA(cudaMemcpyAsync(h->h_o_totals + 2, h->p_status, 4, cudaMemcpyDeviceToHost, s));
    if (nb) {
        TB_CUDA(cudaMemcpyAsync(h->h_m_recs, h->m_recs, sizeof(tb_midline_rec) * (size_t)nb, cudaMemcpyDeviceToHost, s));
        if (total) {
            TB_CUDA(cudaMemcpyAsync(h->h_m_pts, h->m_pts, sizeof(float) * 2 * (size_t)total, cudaMemcpyDeviceToHost, s));
            TB_CUDA(cudaMemcpyAsync(h->h_m_segs, h->m_segs, sizeof(float) * 4 * (size_t)total, cudaMemcpyDeviceToHost, s));
        }
    }
    TB_CUDA(cudaStreamSynchronize(s));
    if ((r = posture_check_capacity(h, "tb_seg_midlines")) != TB_OK) return r;
    h->m_n = nb;
    return TB_OK;
}

extern "C" int tb_seg_midline_result(tb_seg *h, const tb_midline_rec **recs, const float **points, const float **segments, uint32_t *n_blobs)
{
    TB_REQUIRE(h && recs && points && segments && n_blobs, TB_ERR_INVALID, "tb_seg_midline_result: null argument");
    TB_REQUIRE(h->h_m_recs, TB_ERR_STATE, "tb_seg_midline_result: call tb_seg_midlines first");
    *recs = h->h_m_recs; *points = h->h_m_pts; *segments = h->h_m_segs; *n_blobs = h->m_n;
    return TB_OK;
}

extern "C" int tb_seg_posture(tb_seg *h, const tb_posture_request *q)
{
    TB_REQUIRE(h && q, TB_ERR_INVALID, "tb_seg_posture: null argument");
    TB_REQUIRE(h->last_n > 0, TB_ERR_STATE, "tb_seg_posture: submit a batch first");
    TB_REQUIRE(q->outline_resample < 255.f, TB_ERR_INVALID, "tb_seg_posture: outline_resample must be < 255 (T/core/default_config.cpp:898)");
    TB_REQUIRE(q->fetch >= 0 && q->fetch <= 2, TB_ERR_INVALID, "tb_seg_posture: fetch must be 0..2");
    int r = posture_check_params(&q->params, "tb_seg_posture");
    if (r != TB_OK) return r;
    TB_CUDA(cudaSetDevice(h->cfg.device));
    if ((r = posture_alloc(h, q->normalize ? q->params.midline_resolution : 0)) != TB_OK) return r;
    cudaStream_t s = h->last_stream ? h->last_stream : h->stream;
    const SegDev &d = h->d;
    h->o_n = 0; h->m_n = 0; h->p_n = 0;
    const int slot = h->p_prof.begin(s);
    h->p_prof.mark(slot, 0);
    if ((r = posture_enqueue_outlines(h, q->outline_resample, s)) != TB_OK) return r;
    h->p_prof.mark(slot, 1);
    if ((r = posture_enqueue_midlines(h, &q->params, q->normalize ? 1 : 0, q->move_direction_dev, q->fix_length_dev, s)) != TB_OK) return r;
    h->p_prof.mark(slot, 2);
    h->p_has_crops = false;
    if (q->normalize && d.crop_norm >= 2 && d.max_crops) {
        const float scale = q->individual_image_scale > 0.f ? q->individual_image_scale : 1.f;
        r = launch_posture_crops(d.recs, d.totals, d.crop_blob, d.lines, d.line_px, d.pixels, h->d_bg, d.W, d.crop_method, d.crop_w, d.crop_h,
                                 h->m_nrecs, q->median_midline_length_dev, q->median_midline_length_px, scale, d.crop_norm == 3,
                                 d.crops, h->d_coef, h->crop_valid, h->last_n * (int)d.max_crops, s);
        if (r != TB_OK) return r;
        h->launches += 2;
        h->p_has_crops = true;
    }
    h->p_prof.mark(slot, 3);
    TB_CUDA(cudaMemcpyAsync(h->h_o_totals, h->o_totals, 16, cudaMemcpyDeviceToHost, s));
    h->p_pending = true; h->p_fetch = q->fetch; h->p_has_norm = q->normalize != 0; h->p_stream = s; h->p_parent = nullptr;
    return TB_OK;
}

extern "C" int tb_seg_posture_wait(tb_seg *h)
{
    TB_REQUIRE(h, TB_ERR_INVALID, "tb_seg_posture_wait: null handle");
    TB_REQUIRE(h->p_pending || h->p_n > 0, TB_ERR_STATE, "tb_seg_posture_wait: call tb_seg_posture first");
    if (!h->p_pending) return TB_OK;
    int r = tb_seg_wait(h);                            // the blob count, and the batch's own payload as it was asked for
    if (r != TB_OK && r != TB_ERR_CAPACITY) return r;
    tb_seg *ch = h->p_parent ? h->p_parent : h;       // results are indexed by this handle's blobs; its crops were re-rendered
    if (ch != h && (r = tb_seg_wait(ch)) != TB_OK && r != TB_ERR_CAPACITY) return r;
    TB_CUDA(cudaSetDevice(h->cfg.device));
    cudaStream_t s = h->p_stream;
    TB_CUDA(cudaStreamSynchronize(s));
    h->p_pending = false;
    if ((r = posture_check_capacity(h, "tb_seg_posture_wait")) != TB_OK) return r;
    const uint32_t nb = ch->h_totals[0], nc = ch->h_totals[3];
    const SegDev &d = ch->d;
    if (nb && h->p_fetch >= 1) {
        TB_CUDA(cudaMemcpyAsync(h->h_m_recs, h->m_recs, sizeof(tb_midline_rec) * (size_t)nb, cudaMemcpyDeviceToHost, s));
        if (h->p_has_norm) {
            TB_CUDA(cudaMemcpyAsync(h->h_m_nrecs, h->m_nrecs, sizeof(tb_midline_norm) * (size_t)nb, cudaMemcpyDeviceToHost, s));
            TB_CUDA(cudaMemcpyAsync(h->h_m_norm, h->m_norm, sizeof(float) * 4 * (size_t)nb * h->m_res, cudaMemcpyDeviceToHost, s));
        }
        if (h->p_has_crops && nc) {
            TB_CUDA(cudaMemcpyAsync(ch->h_crop_valid, ch->crop_valid, nc, cudaMemcpyDeviceToHost, s));
            if (ch->last_fetch >= 2)                   // the crops tb_seg_wait fetched were the un-normalised ones
                TB_CUDA(cudaMemcpyAsync(ch->h_crops, d.crops, (size_t)nc * d.crop_w * d.crop_h * d.cpx, cudaMemcpyDeviceToHost, s));
        }
    }
    if (nb && h->p_fetch >= 2) {
        TB_CUDA(cudaMemcpyAsync(h->h_o_recs, h->o_recs, sizeof(tb_outline_rec) * (size_t)nb, cudaMemcpyDeviceToHost, s));
        if (h->h_o_totals[0]) TB_CUDA(cudaMemcpyAsync(h->h_o_raw, h->o_raw, sizeof(float) * 2 * (size_t)h->h_o_totals[0], cudaMemcpyDeviceToHost, s));
        if (h->h_o_totals[1]) {
            TB_CUDA(cudaMemcpyAsync(h->h_m_pts, h->m_pts, sizeof(float) * 2 * (size_t)h->h_o_totals[1], cudaMemcpyDeviceToHost, s));
            TB_CUDA(cudaMemcpyAsync(h->h_m_segs, h->m_segs, sizeof(float) * 4 * (size_t)h->h_o_totals[1], cudaMemcpyDeviceToHost, s));
        }
    }
    TB_CUDA(cudaStreamSynchronize(s));
    h->p_n = nb;
    return TB_OK;
}

extern "C" int tb_seg_posture_result(tb_seg *h, tb_posture_view *out)
{
    TB_REQUIRE(h && out, TB_ERR_INVALID, "tb_seg_posture_result: null argument");
    TB_REQUIRE(!h->p_pending && h->h_m_recs, TB_ERR_STATE, "tb_seg_posture_result: call tb_seg_posture and tb_seg_posture_wait first");
    std::memset(out, 0, sizeof(*out));
    out->n_blobs = h->p_n; out->midline_resolution = (uint32_t)h->m_res;
    if (h->p_fetch >= 1) {
        out->midlines = h->h_m_recs;
        if (h->p_has_norm) { out->normalized = h->h_m_nrecs; out->norm_points = h->h_m_norm; }
        if (h->p_has_crops) out->crop_valid = (h->p_parent ? h->p_parent : h)->h_crop_valid;
    }
    if (h->p_fetch >= 2) { out->outlines = h->h_o_recs; out->raw_points = h->h_o_raw; out->points = h->h_m_pts; out->segments = h->h_m_segs; }
    return TB_OK;
}

extern "C" int tb_seg_posture_device(tb_seg *h, void **outline_recs, void **midline_recs, void **normalized, void **norm_points,
                                     void **points, void **segments)
{
    TB_REQUIRE(h, TB_ERR_INVALID, "tb_seg_posture_device: null handle");
    TB_REQUIRE(h->o_recs, TB_ERR_STATE, "tb_seg_posture_device: call tb_seg_posture first");
    if (outline_recs) *outline_recs = h->o_recs;
    if (midline_recs) *midline_recs = h->m_recs;
    if (normalized) *normalized = h->m_nrecs;
    if (norm_points) *norm_points = h->m_norm;
    if (points) *points = h->m_pts;
    if (segments) *segments = h->m_segs;
    return TB_OK;
}

extern "C" int tb_seg_posture_ms(tb_seg *h, double out_ms[3], uint64_t *n_calls)
{
    TB_REQUIRE(h && out_ms && n_calls, TB_ERR_INVALID, "tb_seg_posture_ms: null argument");
    TB_CUDA(cudaSetDevice(h->cfg.device));
    if (h->p_prof.flush() != TB_OK) { set_error("tb_seg_posture_ms: event query failed"); return TB_ERR_CUDA; }
    for (int k = 0; k < 3; ++k) { out_ms[k] = h->p_prof.acc[k]; h->p_prof.acc[k] = 0; }
    *n_calls = h->p_prof.n; h->p_prof.n = 0;
    return TB_OK;
}

extern "C" int tb_seg_recount(tb_seg *h, int threshold, float *out, uint32_t n)
{
    TB_REQUIRE(h && out, TB_ERR_INVALID, "tb_seg_recount: null argument");
    TB_REQUIRE(!h->pending && h->last_n > 0, TB_ERR_STATE, "tb_seg_recount: call tb_seg_wait on a submitted batch first");
    TB_REQUIRE(threshold >= 0, TB_ERR_INVALID, "tb_seg_recount: threshold must be >= 0 (the cached forms recount(-1) are host state of pv::Blob)");
    TB_REQUIRE(!h->d.r3, TB_ERR_INVALID, "tb_seg_recount: built for gray and rgb8 blobs");
    TB_REQUIRE(h->d.sqcm != 0.f, TB_ERR_INVALID, "tb_seg_recount: cm_per_pixel is 0 (PVBlob.cpp:935-937)");
    const uint32_t nb = h->h_totals[0];
    TB_REQUIRE(n >= nb, TB_ERR_INVALID, "tb_seg_recount: the output array is smaller than the batch's blob count");
    if (nb == 0) return TB_OK;
    TB_CUDA(cudaSetDevice(h->cfg.device));
    if (!h->d_recount) { int r = seg_dev(h, &h->d_recount, h->d.blobs_cap); if (r != TB_OK) return r; }
    cudaStream_t s = h->last_stream ? h->last_stream : h->stream;
    const int method = !h->params.enable_difference ? 0 : (h->params.detect_threshold_is_absolute ? 1 : 2);
    recount_kernel<<<148 * 4, 256, 0, s>>>(h->d.recs, h->d.totals, h->d.lines, h->d.line_px, h->d.pixels, h->d.opx, h->d_bg, h->d.W, method, threshold,
                                           h->d.sqcm, h->d_recount);
    h->launches += 1;
    TB_CUDA(cudaGetLastError());
    TB_CUDA(cudaMemcpyAsync(out, h->d_recount, sizeof(float) * (size_t)nb, cudaMemcpyDeviceToHost, s));
    TB_CUDA(cudaStreamSynchronize(s));
    return TB_OK;
}

static int seg_rethreshold(tb_seg *det, tb_seg *trk, int fetch, uint32_t min_payload);
extern "C" int tb_seg_posture_thresholded(tb_seg *src, tb_seg *pst, const tb_posture_request *q, int track_posture_threshold)
{
    TB_REQUIRE(src && pst && q && src != pst, TB_ERR_INVALID, "tb_seg_posture_thresholded: need two distinct handles and a request");
    TB_REQUIRE(src->last_n > 0, TB_ERR_STATE, "tb_seg_posture_thresholded: the source handle has no batch");
    TB_REQUIRE(q->outline_resample < 255.f, TB_ERR_INVALID, "tb_seg_posture_thresholded: outline_resample must be < 255");
    TB_REQUIRE(q->fetch >= 0 && q->fetch <= 2, TB_ERR_INVALID, "tb_seg_posture_thresholded: fetch must be 0..2");
    TB_REQUIRE(pst->d.blobs_cap >= src->d.blobs_cap, TB_ERR_INVALID, "tb_seg_posture_thresholded: the posture handle needs the source handle's max_batch / max_runs_per_frame");
    int r = posture_check_params(&q->params, "tb_seg_posture_thresholded");
    if (r != TB_OK) return r;
    TB_CUDA(cudaSetDevice(pst->cfg.device));
    if ((r = posture_alloc(pst, q->normalize ? q->params.midline_resolution : 0)) != TB_OK) return r;
    if (!pst->r_best) {
        const size_t cap = pst->d.blobs_cap;
        r = seg_dev(pst, &pst->r_best, cap);
        if (r == TB_OK) r = seg_dev(pst, &pst->r_index, cap);
        if (r == TB_OK) r = seg_dev(pst, &pst->r_sub_npx, cap);
        if (r == TB_OK) r = seg_dev(pst, &pst->r_state, cap);
        if (r == TB_OK) r = seg_dev(pst, &pst->r_first, cap);
        if (r == TB_OK) r = seg_dev(pst, &pst->r_remaining, 1);
        if (r == TB_OK) r = host_alloc(&pst->h_r_remaining, 1);
        if (r != TB_OK) return r;
    }
    cudaStream_t s = src->last_stream ? src->last_stream : src->stream;
    const uint32_t np_max = posture_nb_max(src);
    const size_t cap = pst->d.blobs_cap;
    TB_CUDA(cudaMemsetAsync(pst->r_best, 0, sizeof(unsigned long long) * cap, s));
    TB_CUDA(cudaMemsetAsync(pst->r_state, 0, cap, s));
    TB_CUDA(cudaMemsetAsync(pst->r_first, 0, sizeof(tb_outline_rec) * cap, s));
    TB_CUDA(cudaMemsetAsync(pst->o_recs, 0, sizeof(tb_outline_rec) * cap, s));
    TB_CUDA(cudaMemsetAsync(pst->m_recs, 0, sizeof(tb_midline_rec) * cap, s));
    TB_CUDA(cudaMemsetAsync(pst->o_totals, 0, 16, s));
    pst->o_n = 0; pst->m_n = 0; pst->p_n = 0;
    const tb_seg_params saved = pst->params;
    PostureRound R{src->d.recs, src->d.infos, src->d.lines, src->d.totals, pst->d.recs, pst->d.lines, pst->d.totals,
                   pst->r_best, pst->r_index, pst->r_sub_npx, pst->r_state, pst->r_first, pst->r_remaining};
    const OutlineMap map{pst->r_index, src->d.recs, pst->r_state, 1};
    int rounds = 0;
    for (int thr = track_posture_threshold;; thr += 2) {       // Posture.cpp:326-379
        tb_seg_params p = saved;
        p.detect_threshold = thr;
        if ((r = tb_seg_set_params(pst, &p)) != TB_OK) break;
        if ((r = seg_rethreshold(src, pst, 0, 0u)) != TB_OK) break;      // every sub-blob of every source blob at this threshold (threshold_get_biggest_blob picks)
        const int last_round = thr + 2 >= track_posture_threshold + 100;
        if (cudaMemsetAsync(pst->r_remaining, 0, 4, s) != cudaSuccess) { set_error("cudaMemsetAsync failed"); r = TB_ERR_CUDA; break; }
        if ((r = launch_posture_parents(R, np_max, pst->d.blobs_cap, pst->p_sms, s)) != TB_OK) break;
        const SegDev &d = pst->d;
        if ((r = launch_outlines(d.recs, src->d.totals, np_max, d.lines, d.line_px, d.opx, pst->o_visited, (size_t)d.px_cap / d.opx + 16, q->outline_resample,
                                 pst->o_rowfirst, pst->o_sel, pst->o_recs, pst->o_totals, pst->o_raw, pst->o_res, pst->o_cap, pst->p_sms, s, &map)) != TB_OK) break;
        if ((r = launch_midlines(pst->o_recs, src->d.totals, np_max, pst->o_res, pst->o_cap, &q->params, q->normalize ? 1 : 0, q->move_direction_dev,
                                 q->fix_length_dev, pst->m_pts, pst->m_segs, pst->m_recs, pst->m_nrecs, pst->m_norm, pst->m_arena, pst->m_arena_floats,
                                 pst->m_arena_used, pst->p_status, pst->p_sms, s, pst->r_state)) != TB_OK) break;
        if ((r = launch_posture_round_end(R, np_max, pst->o_recs, pst->m_recs, pst->m_nrecs, q->normalize ? 1 : 0, last_round, pst->p_sms, s)) != TB_OK) break;
        pst->launches += 8;
        if (cudaMemcpyAsync(pst->h_r_remaining, pst->r_remaining, 4, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
            cudaStreamSynchronize(s) != cudaSuccess) { set_error("tb_seg_posture_thresholded: CUDA failure in the threshold loop"); r = TB_ERR_CUDA; break; }
        ++rounds;
        pst->pending = false;                          // the round's batch is complete (headers are on the host): the next round may re-submit
        if (*pst->h_r_remaining == 0 || last_round) break;
    }
    const int r2 = tb_seg_set_params(pst, &saved);
    if (r != TB_OK) return r;
    if (r2 != TB_OK) return r2;
    pst->p_has_crops = false;
    if (q->normalize && src->d.crop_norm >= 2 && src->d.max_crops) {    // the source blobs' crops through the midline transform
        if ((r = posture_alloc_crops(src)) != TB_OK) return r;
        const SegDev &sd = src->d;
        const float scale = q->individual_image_scale > 0.f ? q->individual_image_scale : 1.f;
        r = launch_posture_crops(sd.recs, sd.totals, sd.crop_blob, sd.lines, sd.line_px, sd.pixels, src->d_bg, sd.W, sd.crop_method, sd.crop_w, sd.crop_h,
                                 pst->m_nrecs, q->median_midline_length_dev, q->median_midline_length_px, scale, sd.crop_norm == 3,
                                 sd.crops, src->d_coef, src->crop_valid, src->last_n * (int)sd.max_crops, s);
        if (r != TB_OK) return r;
        src->launches += 2;
        pst->p_has_crops = true;
    }
    TB_CUDA(cudaMemcpyAsync(pst->h_o_totals, pst->o_totals, 16, cudaMemcpyDeviceToHost, s));
    pst->p_pending = true; pst->p_fetch = q->fetch; pst->p_has_norm = q->normalize != 0; pst->p_stream = s; pst->p_parent = src;
    return rounds;
}

extern "C" int tb_seg_debug_binary(tb_seg *h, const uint8_t *frame_host, uint8_t *out_host)
{
    TB_REQUIRE(h && frame_host && out_host, TB_ERR_INVALID, "tb_seg_debug_binary: null argument");
    TB_REQUIRE(h->has_bg, TB_ERR_STATE, "tb_seg_debug_binary: no background set");
    TB_CUDA(cudaSetDevice(h->cfg.device));
    const size_t n = (size_t)h->d.W * h->d.H;
    if (h->d.CN > 1) {
        TB_CUDA(cudaMemcpyAsync(h->d_frames, frame_host, n * h->d.CN, cudaMemcpyHostToDevice, h->stream));
        int r = seg_to_gray(h, h->d_frames, 1, h->stream);
        if (r != TB_OK) return r;
        h->gray_valid = false;
        const uint8_t *mask = nullptr;
        if (h->morph && (r = seg_morph(h, h->d_gray, 1, h->stream, &mask)) != TB_OK) return r;
        binary_image_color_kernel<<<592, 256, 0, h->stream>>>(h->d_frames, h->d_gray, h->d_bg, mask, h->d_tmp, n, h->k, h->d.CN, h->d.enc);
        h->launches += 1;
        TB_CUDA(cudaGetLastError());
        TB_CUDA(cudaMemcpyAsync(out_host, h->d_tmp, n * h->d.opx, cudaMemcpyDeviceToHost, h->stream));
        TB_CUDA(cudaStreamSynchronize(h->stream));
        return TB_OK;
    }
    TB_CUDA(cudaMemcpyAsync(h->d_frames, frame_host, n, cudaMemcpyHostToDevice, h->stream));
    if (h->morph) {
        const uint8_t *mask = nullptr;
        int r = seg_morph(h, h->d_frames, 1, h->stream, &mask);
        if (r != TB_OK) return r;
        mask_and_kernel<<<592, 256, 0, h->stream>>>(mask, h->d_frames, h->d_tmp, n);
    } else binary_image_kernel<<<592, 256, 0, h->stream>>>(h->d_frames, h->d_bg, h->d_tmp, n, h->k);
    h->launches += 1;
    TB_CUDA(cudaGetLastError());
    TB_CUDA(cudaMemcpyAsync(out_host, h->d_tmp, n, cudaMemcpyDeviceToHost, h->stream));
    TB_CUDA(cudaStreamSynchronize(h->stream));
    return TB_OK;
}

extern "C" uint64_t tb_seg_launch_count(tb_seg *h) { return h ? h->launches : 0; }

extern "C" int tb_seg_profile(tb_seg *h, int enable)
{
    TB_REQUIRE(h, TB_ERR_INVALID, "tb_seg_profile: null handle");
    TB_CUDA(cudaSetDevice(h->cfg.device));
    if (h->prof.enable(enable != 0) != TB_OK || h->p_prof.enable(enable != 0) != TB_OK) { set_error("tb_seg_profile: cudaEventCreate failed"); return TB_ERR_CUDA; }
    return TB_OK;
}

extern "C" int tb_seg_kernel_ms(tb_seg *h, double out_ms[3], uint64_t *n_batches)
{
    TB_REQUIRE(h && out_ms && n_batches, TB_ERR_INVALID, "tb_seg_kernel_ms: null argument");
    TB_CUDA(cudaSetDevice(h->cfg.device));
    if (h->prof.flush() != TB_OK) { set_error("tb_seg_kernel_ms: event query failed"); return TB_ERR_CUDA; }
    for (int k = 0; k < 3; ++k) { out_ms[k] = h->prof.acc[k]; h->prof.acc[k] = 0; }
    *n_batches = h->prof.n; h->prof.n = 0;
    return TB_OK;
}

// pixel::threshold_blob for every blob of the detection handle's last batch, on the device: `trk` is a second
// handle of the same geometry whose params carry the TRACKER settings (detect_threshold = track_threshold,
// enable_difference = track_background_subtraction, detect_threshold_is_absolute = track_threshold_is_absolute,
// size ranges = track_size_filter or none).  Results are read from `trk` like after a submit.
// min_payload: 1 = pixel::threshold_blob as the tracker calls it (sub-blobs of one payload byte are not handed on, PixelTree.cpp:350-352);
// 0 = every sub-blob (pixel::threshold_get_biggest_blob, :297-340: what the posture loop chooses from)
static int seg_rethreshold(tb_seg *det, tb_seg *trk, int fetch, uint32_t min_payload);
extern "C" int tb_seg_rethreshold(tb_seg *det, tb_seg *trk, int fetch) { return seg_rethreshold(det, trk, fetch, 1u); }
static int seg_rethreshold(tb_seg *det, tb_seg *trk, int fetch, uint32_t min_payload)
{
    TB_REQUIRE(det && trk && det != trk, TB_ERR_INVALID, "tb_seg_rethreshold: need two distinct handles");
    TB_REQUIRE(det->last_n > 0 && det->last_frames_dev, TB_ERR_STATE, "tb_seg_rethreshold: the detection handle has no batch");
    TB_REQUIRE(trk->d.W == det->d.W && trk->d.H == det->d.H && trk->cfg.device == det->cfg.device, TB_ERR_INVALID,
               "tb_seg_rethreshold: handles differ in frame size or device");
    TB_REQUIRE(det->last_n <= trk->cfg.max_batch, TB_ERR_INVALID, "tb_seg_rethreshold: tracker-side max_batch too small");
    TB_REQUIRE(trk->has_bg, TB_ERR_STATE, "tb_seg_rethreshold: the tracker-side handle has no background");
    TB_REQUIRE(!det->d.r3 && !trk->d.r3, TB_ERR_INVALID, "tb_seg_rethreshold: the tracker-side re-threshold is built for gray and rgb8 blobs");
    TB_REQUIRE((det->d.enc == 0 && trk->d.CN == 1) || (det->d.enc == 1 && trk->d.enc == 1 && trk->d.CN == det->d.CN), TB_ERR_INVALID,
               "tb_seg_rethreshold: the tracker-side handle takes the grey plane (channels = 1) for gray encoding, "
               "or the same colour frames (same channels, rgb8) for rgb8");
    TB_REQUIRE(!trk->pending, TB_ERR_STATE, "tb_seg_rethreshold: the tracker-side handle has a pending batch");
    TB_CUDA(cudaSetDevice(det->cfg.device));
    const size_t px = (size_t)det->d.W * det->d.H;
    if (!trk->keep_mask) {
        int r = seg_dev(trk, &trk->keep_mask, (size_t)trk->cfg.max_batch * px + 16);
        if (r != TB_OK) return r;
    }
    cudaStream_t s = det->last_stream ? det->last_stream : det->stream;
    const int n = det->last_n;
    TB_CUDA(cudaMemsetAsync(trk->keep_mask, 0, (size_t)n * px, s));
    paint_blobs_kernel<<<148 * 4, 256, 0, s>>>(det->d.recs, det->d.totals, det->d.lines, trk->keep_mask, det->d.W, det->d.H);
    trk->launches += 1;
    TB_CUDA(cudaGetLastError());
    const uint8_t *plane = det->last_frames_dev;
    if (det->d.enc == 1) return seg_launch(trk, plane, n, s, fetch, trk->keep_mask, min_payload);     // rgb8: trk converts with the tracker's grey formula
    if (det->d.CN > 1) {                       // colour frames, gray encoding: re-threshold the grey plane
        if (!det->gray_valid) { int r = seg_to_gray(det, det->last_frames_dev, n, s); if (r != TB_OK) return r; }
        plane = det->d_gray;
    }
    return seg_launch(trk, plane, n, s, fetch, trk->keep_mask, min_payload);
}
