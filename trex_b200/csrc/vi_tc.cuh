// Tensor-core (tcgen05 / TMEM) kernels of the V118_3 forward pass for sm_100a.
//
// The 5x5 "same" convolutions run as implicit GEMMs without an im2col buffer.  Activations live in
// channel-group planes  [image][hi|lo][group of 8 channels][flat padded position][8] bf16  (16 bytes per
// position and group, zero halo of 2 pixels, row pitch WP = W + 4).  In that layout the A operand of
// one filter tap (dy,dx) for 128 consecutive flat output positions q is simply the same plane read at
// position q + dy*WP + dx: a SWIZZLE_NONE K-major UMMA descriptor whose start address is shifted by the
// tap (8 consecutive positions x 16 B = one core matrix, SBO = 128 B, LBO = plane size).  Outputs at
// halo columns are garbage and are dropped in the epilogue.
//
// fp32 parity (1e-3 on logits, BASELINE.json north_star) is kept with a 2-term bf16 split of both
// operands and three MMAs per k-step: A_hi*B_hi + A_lo*B_hi + A_hi*B_lo (fp32 accumulation in TMEM).
//
// Warp roles per CTA: warp 0 = bulk-copy (TMA engine) producer, warp 1 = MMA issuer and TMEM owner,
// warps 2.. = epilogue (TMEM -> registers -> max-pool -> BN shift -> ReLU -> bf16 hi/lo planes for the
// next layer).  mbarrier pipelines connect them; CTAs are persistent.
#pragma once
#include "umma.cuh"

#include <cuda_fp8.h>

namespace tb { namespace tc {

// Arithmetic of conv2 / conv3 (template parameter MODE):
//   BF16X3  activations and weights split x = hi + lo (bf16 each), three MMAs per k-step (hi*hi + lo*hi + hi*lo)
//   FP16    one fp16 x fp16 MMA per k-step (the hi slots of the same buffers hold fp16)
//   FP16C   "fp16 + corrections": x = h + l with h = fp16(x); the main term h_a * h_w is one kind::f16 MMA (K = 16), the two
//           first-order correction terms l_a * h_w + h_a * l_w are ONE kind::f8f6f4 MMA (e5m2, K = 32 = [l_a * 2^8 | h_a * 2^-8] .
//           [h_w * 2^-8 | l_w * 2^8]; the power-of-two scales keep both factors inside e5m2's exponent range and cancel exactly)
//           into the same fp32 accumulator: two MMA slots per k-step instead of three, the same bytes as BF16X3 (the lo slot of a
//           channel-group pair holds 16 + 16 e5m2 bytes per position), error ~2^-3 of the fp16 rounding error.
enum : int { BF16X3 = 0, FP16 = 1, FP16C = 2 };
constexpr float FC_UP = 256.f, FC_DOWN = 1.f / 256.f;

// 8 fp32 values (one channel group of a position) -> 8 e5m2 bytes of the correction planes: L = (x - fp16(x)) * 2^8, H = fp16(x) * 2^-8
__device__ __forceinline__ void fp16c_pack(const float (&m)[8], uint32_t (&h16)[4], uint2 &l8, uint2 &h8)
{
    uint32_t lw[2] = {0u, 0u}, hw[2] = {0u, 0u};
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
        const __half a = __float2half_rn(m[j]), b = __float2half_rn(m[j + 1]);
        h16[j >> 1] = (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
        const float fa = __half2float(a), fb = __half2float(b);
        const uint32_t l2 = (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2((m[j] - fa) * FC_UP, (m[j + 1] - fb) * FC_UP), __NV_SATFINITE, __NV_E5M2);
        const uint32_t h2 = (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(fa * FC_DOWN, fb * FC_DOWN), __NV_SATFINITE, __NV_E5M2);
        lw[j >> 2] |= l2 << (16 * ((j >> 1) & 1));
        hw[j >> 2] |= h2 << (16 * ((j >> 1) & 1));
    }
    l8 = make_uint2(lw[0], lw[1]); h8 = make_uint2(hw[0], hw[1]);
}

constexpr int NT = 192;

__host__ __device__ constexpr int cmax(int a, int b) { return a > b ? a : b; }
__host__ __device__ constexpr int pow2_cols(int n) { return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512; }

// Geometry of the activation planes in global memory.
// conv2 input (conv1 output, 16 ch, 40x40): 2 groups, row pitch 44 (2-pixel zero halo on each side), 1960 positions.
struct Conv2Cfg {
    static constexpr int G = 2, WP = 44, PL = 1960;
    static constexpr size_t IMG_BYTES = (size_t)2 * G * PL * 16;
};
// conv3 input (conv2 output, 64 ch, 20x20): 8 groups, row pitch 22: the two zero columns between image rows
// serve as right halo of one row and left halo of the next, so a run of flat positions wastes only 2 of 22.
struct Conv3Cfg {
    static constexpr int G = 8, WP = 22, ROWS = 24, PL = 536;
    static constexpr size_t IMG_BYTES = (size_t)2 * G * PL * 16;
    static constexpr int WTAP_BYTES = 2 * G * 128 * 16;
};

// fc1 A operand: [hi|lo][kc block of 8][group of 8 images][kc in block][8 images][8] bf16, kc = c8 * 100 + pooled pixel.
// One (hi|lo, kc block) slab of 16 consecutive image groups is 16 KB contiguous: one bulk copy per pipeline stage.
constexpr int FC_K = 12800, FC_KC = FC_K / 8, FC_KB = FC_KC / 8, FC_N = 112, FC_NREAL = 100, FC_SPLIT = 4;
__host__ __device__ inline size_t fc_a_offset(int hl, int n_groups, int img, int kc)
{
    return (((((size_t)hl * FC_KB + (kc >> 3)) * n_groups + (img >> 3)) * 8 + (kc & 7)) * 8 + (img & 7)) * 16;
}

// ------------------------------------------------------------------------------------------------
// conv3, channel-major orientation:  D[128 cout][positions] = W[128][K] * X[positions][K]^T.
// The weights are the A operand (M = 128 output channels), a run of flat positions of the activation
// planes is the B operand (N = 224 covering 10 image rows x pitch 22), shifted per tap exactly as above.
// Each TMEM lane then holds ONE output channel and the columns are positions, so BN + ReLU + the 2x2
// max-pool are thread-local (columns q, q+1, q+22, q+23): no shuffles, no staging, no block barriers.
// 8 epilogue warps (2 per TMEM lane quarter, one per row half); the next image's planes are fetched
// while the epilogue drains TMEM.
// ------------------------------------------------------------------------------------------------
struct Conv3T {
    static constexpr int G = 8, NOUT = 128, H = 20, W = 20, WP = Conv3Cfg::WP;
    static constexpr int NT_ROWS = 10, TSTEP = NT_ROWS * WP, N = (TSTEP + 15) / 16 * 16, TILES = H / NT_ROWS;   // 2 tiles: 220 positions each, N = 224
    static constexpr int PIN = ((TILES - 1) * TSTEP + N + 4 * WP + 4 + 7) / 8 * 8;  // 536 positions staged
    static constexpr int PL = Conv3Cfg::PL;                                        // plane size in global memory (positions)
    // bf16x3 / fp16c (2 G planes per image): no room for two images, so the input is staged PER TILE -- tile t reads positions t * TSTEP ... + TPOS
    // of a plane, the two ranges (rows 0-13 and 10-23) are kept as separate blocks (the 4 shared rows twice) and block t of the next image loads
    // while the other tile of the current one is multiplied.  fp16 (G planes) double-buffers whole images.
    static constexpr int TPOS = (N + 4 * WP + 4 + 7) / 8 * 8;                      // 320 positions per tile block
    static constexpr int IN_BYTES = TILES * 2 * G * TPOS * 16 > 2 * G * PIN * 16 ? TILES * 2 * G * TPOS * 16 : 2 * G * PIN * 16;
    static constexpr int WTAP_BYTES = 2 * G * NOUT * 16;
    static constexpr int SMEM = (IN_BYTES + 127) / 128 * 128 + 2 * WTAP_BYTES + 128;
    static constexpr int TMEM_COLS = 512;
    static constexpr int THREADS = 64 + 512 + 32;          // weights, MMA issue, 16 epilogue warps: (TMEM lane quarter) x (tile) x (row-pair half), input loader
    static_assert((TILES - 1) * TSTEP + TPOS <= PL + 8 && TPOS * 16 % 16 == 0, "tile blocks stay inside the plane");
    static_assert(PIN <= PL, "staged positions exceed the plane");
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

// F16 ("fp16" precision): activations and weights are single fp16 planes (the hi slots of the same buffers) and every
// k-step is ONE MMA instead of three; max|dlogit| ~ 3e-4 against fp32 (tolerance 1e-3), see DESIGN.md.
// CL = 2: the kernel runs in clusters of two CTAs that share the weight stream.  The leader (rank 0) issues every
// tap's weights once as a MULTICAST bulk copy into both CTAs' rings when both have released the stage (the peer
// commits its MMAs onto the leader's "empty" barrier through the cluster address space); each CTA re-arms its own
// "full" barrier.  Half the L2 weight reads per SM make the tile-outer order (epilogue of a tile under the MMAs of
// the other tile) affordable for fp16 as well.  Both CTAs run the same number of image slots; a CTA without an
// image in the last slot only takes part in the weight hand-shake.
template <int MODE, int CL>
__global__ void __launch_bounds__(Conv3T::THREADS, 1)
conv3_t_kernel(const uint8_t *__restrict__ in, int n_max, const uint32_t *__restrict__ n_dev, int base,
               const uint8_t *__restrict__ wgt, const float *__restrict__ sc, const float *__restrict__ sh,
               uint8_t *__restrict__ out, int out_groups)
{
    using C = Conv3T;
    constexpr bool F16 = MODE == FP16;                    // one plane set per operand; FP16C / BF16X3 carry the lo slots too
    constexpr bool HALF_STAGES = MODE != BF16X3;          // weight ring in half-taps of 16 KB (hi slots | lo slots)
    constexpr int HALVES = MODE == FP16C ? 2 : 1;         // ring stages consumed per tap
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar_in_full[2], bar_in_empty[2], bar_acc_full[Conv3T::TILES], bar_acc_empty[Conv3T::TILES], bar_w_full[4], bar_w_empty[4];
    __shared__ uint32_t s_tmem;
    uint8_t *s_in = smem;
    uint8_t *s_w = smem + (C::IN_BYTES + 127) / 128 * 128;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_act = n_dev ? min((int)*n_dev - base, n_max) : n_max;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) { umma::mbar_init(&bar_in_full[i], 1); umma::mbar_init(&bar_in_empty[i], 1); }
        for (int i = 0; i < C::TILES; ++i) { umma::mbar_init(&bar_acc_full[i], 1); umma::mbar_init(&bar_acc_empty[i], 1); }
        for (int i = 0; i < 4; ++i) { umma::mbar_init(&bar_w_full[i], 1); umma::mbar_init(&bar_w_empty[i], CL); }
        umma::fence_mbar_init();
        if (CL == 2) for (int i = 0; i < (HALF_STAGES ? 4 : 2); ++i) umma::mbar_expect_tx(&bar_w_full[i], HALF_STAGES ? C::WTAP_BYTES / 2 : C::WTAP_BYTES);   // armed for the first use
    }
    if (warp == 1) umma::tmem_alloc(&s_tmem, C::TMEM_COLS);
    umma::fence_before_sync();
    __syncthreads();
    if (CL == 2) umma::cluster_sync();                    // the peer's barriers exist before anything is sent to them
    umma::fence_after_sync();
    const uint32_t tm = s_tmem;
    const uint32_t crank = CL == 2 ? umma::cluster_ctarank() : 0u;
    // image slots of this CTA: its own images, or (clusters) as many as the cluster's first CTA has
    const int first = (int)blockIdx.x - (int)crank;
    const int slots = CL == 2 ? (first < n_act ? (n_act - first + (int)gridDim.x - 1) / (int)gridDim.x : 0)
                              : ((int)blockIdx.x < n_act ? (n_act - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0);
    // fp16: the planes of one image are half the size, so the input is double buffered (stage = G planes) and the next
    // image loads under the MMAs of the current one; bf16x3 keeps one stage of 2 G planes
    constexpr int NBUF = F16 ? 2 : 1, STAGE_POS = C::G * C::PIN;
    constexpr bool SPLIT = !F16;                          // per-tile input blocks (see Conv3T::TPOS); bar_in_full / bar_in_empty are indexed by tile then
    constexpr int TBLK = 2 * C::G * C::TPOS;              // positions of one tile block (all planes)
    // weight ring: a tap is 32 KB (bf16 hi + lo: 2 stages, 1.4 us of MMAs per tap hide the L2 latency) or 16 KB (fp16:
    // 4 stages, because 0.46 us of MMAs per tap do not).  bf16x3 runs tile-outer (the epilogue of a tile under the
    // MMAs of the other one, taps streamed once per tile); fp16 runs tap-outer: twice the L2 weight stream measured slower.
    constexpr int WB = HALF_STAGES ? C::WTAP_BYTES / 2 : C::WTAP_BYTES, NWS = HALF_STAGES ? 4 : 2;
    constexpr bool TILE_OUTER = !F16 || CL == 2;
    static_assert(MODE != FP16C || TILE_OUTER, "FP16C runs tile-outer");

    if (warp == 0) {
        if (lane == 0 && (CL == 1 || crank == 0)) {                             // the leader streams the weights for both CTAs
            uint32_t tapc = 0;
            for (int k = 0; k < slots; ++k) {
                for (int tt = 0; tt < (TILE_OUTER ? 25 * C::TILES : 25) * HALVES; ++tt, ++tapc) {
                    const int tap = (tt / HALVES) % 25, half = tt % HALVES;     // FP16C: the hi slots (fp16) of a tap, then its lo slots (e5m2)
                    const uint32_t s = tapc % NWS;
                    const uint8_t *wsrc = wgt + (size_t)tap * C::WTAP_BYTES + (size_t)half * WB;
                    umma::mbar_wait(&bar_w_empty[s], ((tapc / NWS) & 1) ^ 1);   // CL = 2: both CTAs have released the stage
                    if (CL == 2) umma::bulk_g2s_multicast(s_w + s * WB, wsrc, WB, &bar_w_full[s], (uint16_t)3);
                    else {
                        umma::mbar_expect_tx(&bar_w_full[s], WB);
                        umma::bulk_g2s(s_w + s * WB, wsrc, WB, &bar_w_full[s]);
                    }
                }
            }
        }
    } else if (warp == 2 + 16) {
        // input loader (its own warp: waiting for a free block must not hold up the weight stream)
        if (lane == 0) {
            uint32_t it = 0;
            for (int img = blockIdx.x; img < n_act; img += gridDim.x, ++it) {
                const uint8_t *src = in + (size_t)img * Conv3Cfg::IMG_BYTES;
                if (SPLIT) {
                    for (int t = 0; t < C::TILES; ++t) {
                        umma::mbar_wait(&bar_in_empty[t], (it & 1) ^ 1);        // the MMAs of tile t of the previous image have retired
                        constexpr int NPL = 2 * C::G;
                        // the last block ends with the plane: it is short of TPOS by what the plane lacks (positions no stored output reads)
                        const int npos = min(C::TPOS, C::PL - t * C::TSTEP);
                        umma::mbar_expect_tx(&bar_in_full[t], (uint32_t)(NPL * npos * 16));
                        for (int p = 0; p < NPL; ++p)
                            umma::bulk_g2s(s_in + ((size_t)t * TBLK + (size_t)p * C::TPOS) * 16, src + ((size_t)p * C::PL + (size_t)t * C::TSTEP) * 16,
                                           (uint32_t)(npos * 16), &bar_in_full[t]);
                    }
                } else {
                    const uint32_t ib = it % NBUF, iph = (it / NBUF) & 1;
                    umma::mbar_wait(&bar_in_empty[ib], iph ^ 1);
                    constexpr int NPL = C::G;                                   // fp16: hi slots only
                    umma::mbar_expect_tx(&bar_in_full[ib], NPL * C::PIN * 16);
                    for (int p = 0; p < NPL; ++p)
                        umma::bulk_g2s(s_in + ((size_t)ib * STAGE_POS + (size_t)p * C::PIN) * 16, src + (size_t)p * C::PL * 16, C::PIN * 16, &bar_in_full[ib]);
                }
            }
        }
    } else if (warp == 1) {
        if (umma::elect_one()) {
            const uint32_t idesc = MODE != BF16X3 ? umma::idesc_f16_f32(128, C::N) : umma::idesc_bf16_f32(128, C::N);
            const uint32_t idesc8 = umma::idesc_e5m2_f32(128, C::N);
            const uint64_t x_base = umma::smem_desc(umma::smem_u32(s_in), (SPLIT ? C::TPOS : C::PIN) * 16, 128);
            const uint64_t w_base = umma::smem_desc(umma::smem_u32(s_w), C::NOUT * 16, 128);
            uint32_t it = 0, tapc = 0;
            // "empty" barrier of a weight stage: this CTA's own, or (clusters) the leader's, through the cluster window
            uint32_t w_empty_a[4];
            for (int i = 0; i < 4; ++i) w_empty_a[i] = CL == 2 ? umma::mapa(umma::smem_u32(&bar_w_empty[i]), 0u) : umma::smem_u32(&bar_w_empty[i]);
            int img = blockIdx.x;
            for (int k = 0; k < slots; ++k, img += gridDim.x) {
                const bool real = img < n_act;
                const uint32_t ib = it % NBUF, iph = (it / NBUF) & 1;
                const uint32_t xofs = ib * STAGE_POS;
                if (real && !SPLIT) {
                    umma::mbar_wait(&bar_in_full[ib], iph);
                    umma::fence_after_sync();
                }
                auto tap_mmas = [&](int t, int tap, uint32_t wofs) {
                    const uint32_t d = tm + (uint32_t)(t * C::N);
                    const uint32_t shift = (uint32_t)((tap / 5) * C::WP + (tap % 5));
#pragma unroll
                    for (int ks = 0; ks < C::G / 2; ++ks) {
                        const uint32_t x_hi = SPLIT ? t * TBLK + (0 * C::G + 2 * ks) * C::TPOS + shift : xofs + (0 * C::G + 2 * ks) * C::PIN + t * C::TSTEP + shift;
                        const uint32_t x_lo = SPLIT ? t * TBLK + (1 * C::G + 2 * ks) * C::TPOS + shift : (1 * C::G + 2 * ks) * C::PIN + t * C::TSTEP + shift;
                        const uint32_t w_hi = wofs + (0 * C::G + 2 * ks) * C::NOUT;
                        const uint32_t w_lo = wofs + (1 * C::G + 2 * ks) * C::NOUT;
                        umma::mma_bf16(d, w_base + w_hi, x_base + x_hi, idesc, (tap | ks) != 0);
                        if (MODE == BF16X3) {
                            umma::mma_bf16(d, w_base + w_hi, x_base + x_lo, idesc, 1);
                            umma::mma_bf16(d, w_base + w_lo, x_base + x_hi, idesc, 1);
                        }
                    }
                };
                // FP16C, second half-stage of a tap: the e5m2 correction MMAs (K = 32 over the lo-slot plane pair of a channel block);
                // the stage holds the tap's lo slots at its start
                auto tap_mmas8 = [&](int t, int tap, uint32_t wofs) {
                    const uint32_t d = tm + (uint32_t)(t * C::N);
                    const uint32_t shift = (uint32_t)((tap / 5) * C::WP + (tap % 5));
#pragma unroll
                    for (int ks = 0; ks < C::G / 2; ++ks) {
                        const uint32_t x_lo = t * TBLK + (1 * C::G + 2 * ks) * C::TPOS + shift;
                        const uint32_t w_lo = wofs + (2 * ks) * C::NOUT;
                        umma::mma_f8(d, w_base + w_lo, x_base + x_lo, idesc8, 1);
                    }
                };
                auto take_stage = [&](uint32_t s) {           // wait for the tap's weights; clusters: re-arm the barrier for its next use
                    umma::mbar_wait(&bar_w_full[s], (tapc / NWS) & 1);
                    if (CL == 2) umma::mbar_expect_tx(&bar_w_full[s], WB);
                    umma::fence_after_sync();
                };
                if (TILE_OUTER) {
#pragma unroll 1
                    for (int t = 0; t < C::TILES; ++t) {
                        if (real) {
                            if (SPLIT) umma::mbar_wait(&bar_in_full[t], it & 1);
                            umma::mbar_wait(&bar_acc_empty[t], (it & 1) ^ 1);
                            umma::fence_after_sync();
                        }
                        for (int tap = 0; tap < 25; ++tap, ++tapc) {
                            const uint32_t s = tapc % NWS;
                            take_stage(s);
                            if (real) tap_mmas(t, tap, (s * WB) >> 4);
                            umma::commit_a(w_empty_a[s]);
                            if (MODE == FP16C) {
                                ++tapc;
                                const uint32_t s8 = tapc % NWS;
                                take_stage(s8);
                                if (real) tap_mmas8(t, tap, (s8 * WB) >> 4);
                                umma::commit_a(w_empty_a[s8]);
                            }
                        }
                        if (real) {
                            umma::commit(&bar_acc_full[t]);
                            if (SPLIT) umma::commit(&bar_in_empty[t]);      // block t consumed: the loader may bring in the next image's
                        }
                    }
                } else {
                    for (int t = 0; t < C::TILES; ++t) umma::mbar_wait(&bar_acc_empty[t], (it & 1) ^ 1);
                    umma::fence_after_sync();
                    for (int tap = 0; tap < 25; ++tap, ++tapc) {
                        const uint32_t s = tapc % NWS;
                        take_stage(s);
#pragma unroll
                        for (int t = 0; t < C::TILES; ++t) tap_mmas(t, tap, (s * WB) >> 4);
                        umma::commit_a(w_empty_a[s]);
                    }
                    for (int t = 0; t < C::TILES; ++t) umma::commit(&bar_acc_full[t]);
                }
                if (real) {
                    if (!SPLIT) umma::commit(&bar_in_empty[ib]);          // planes consumed: the loader may refill this stage
                    ++it;
                }
            }
        }
    } else if (warp < 2 + 16) {
        // epilogue: 16 warps; a warp handles TMEM lane quarter (warp & 3) of tile ((ew >> 2) & 1), row pairs 0-2 or 3-4
        const int ew = warp - 2, quarter = warp & 3, t = (ew >> 2) & 1, sub = ew >> 3;
        const int pr0 = sub ? 3 : 0, pr1 = sub ? C::NT_ROWS / 2 : 3;
        const int c = quarter * 32 + lane;                    // output channel of this thread
        const float b = sh[c];
        const int c8 = c >> 3, e = c & 7;
        uint32_t it = 0;
        for (int img = blockIdx.x; img < n_act; img += gridDim.x, ++it) {
            umma::mbar_wait(&bar_acc_full[t], it & 1);
            umma::fence_after_sync();
            const size_t lo_ofs = fc_a_offset(1, out_groups, 0, 0);
#pragma unroll 1
            for (int pr = pr0; pr < pr1; ++pr) {                 // pairs of image rows: 48 consecutive columns
                uint32_t v[48];
                const uint32_t col = (uint32_t)(t * C::N + pr * 2 * C::WP);
                const uint32_t ta = tm + ((uint32_t)(quarter * 32) << 16) + col;
                umma::tmem_ld16(ta, *reinterpret_cast<uint32_t (*)[16]>(&v[0]));
                umma::tmem_ld16(ta + 16, *reinterpret_cast<uint32_t (*)[16]>(&v[16]));
                umma::tmem_ld16(ta + 32, *reinterpret_cast<uint32_t (*)[16]>(&v[32]));
                umma::tmem_ld_wait();
                const int py = t * (C::NT_ROWS / 2) + pr;
#pragma unroll
                for (int px = 0; px < C::W / 2; ++px) {
                    const float a0 = __uint_as_float(v[2 * px]), a1 = __uint_as_float(v[2 * px + 1]);
                    const float a2 = __uint_as_float(v[C::WP + 2 * px]), a3 = __uint_as_float(v[C::WP + 2 * px + 1]);
                    const float m = fmaxf(fmaxf(fmaxf(a0, a1), fmaxf(a2, a3)) + b, 0.f);      // BN scale folded into the weights
                    __nv_bfloat16 h, l;
                    umma::split_bf16(m, h, l);
                    uint8_t *o = out + fc_a_offset(0, out_groups, img, c8 * 100 + py * (C::W / 2) + px) + e * 2;
                    *reinterpret_cast<__nv_bfloat16 *>(o) = h;
                    *reinterpret_cast<__nv_bfloat16 *>(o + lo_ofs) = l;
                }
            }
            umma::fence_before_sync();
            asm volatile("bar.sync %0, 256;" :: "r"(1 + t) : "memory");      // the 8 warps of this tile
            if (quarter == 0 && sub == 0 && lane == 0) umma::mbar_arrive(&bar_acc_empty[t]);
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (CL == 2) umma::cluster_sync();                    // no CTA leaves while its peer may still write to it
    if (warp == 1) umma::tmem_dealloc(tm, C::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// conv2, position-major with 2-D tiles: an MMA tile is 8 pixels wide x 16 rows (SBO = row pitch), so the
// 2x2 pool partners of TMEM lane r are lanes r^1 and r^8 of the same warp: two shuffles, no staging.
// Work item = a band of 16 output rows of one image (5 tiles); input bands are double buffered, all
// 25 taps of weights stay resident, accumulators rotate through a ring of 8 TMEM buffers (64 columns
// each) so the 8 epilogue warps drain tile i while the tensor core works on tiles i+1..i+7.
// ------------------------------------------------------------------------------------------------
struct Conv2D {
    static constexpr int G = 2, NOUT = 64, H = 40, W = 40, WP = 44;
    static constexpr int BAND_ROWS = 16, IN_ROWS = BAND_ROWS + 4, BAND_POS = IN_ROWS * WP;     // 880 positions
    static constexpr int BANDS = 3, TILES = W / 8;                                              // y0 = 0, 16, 24
    static constexpr int IN_BYTES = 2 * G * BAND_POS * 16;
    static constexpr int WTAP_BYTES = 2 * G * NOUT * 16, W_BYTES = 25 * WTAP_BYTES;
    static constexpr int NACC = 4, ACC_COLS = 2 * NOUT;    // [A_hi*W_hi + A_lo*W_hi | A_hi*W_lo], summed in the epilogue (fp16: 8 x 64)
    static constexpr int SMEM = 2 * IN_BYTES + W_BYTES + NOUT * 8 + 128;
    static constexpr int EPI_SETS = 2, THREADS = 64 + EPI_SETS * 256;   // two sets of 8 epilogue warps take alternate tiles
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

template <int MODE>
__global__ void __launch_bounds__(Conv2D::THREADS, 1)
conv2_2d_kernel(const uint8_t *__restrict__ in, int n_max, const uint32_t *__restrict__ n_dev, int base,
                const uint8_t *__restrict__ wgt, const float *__restrict__ sc, const float *__restrict__ sh,
                uint8_t *__restrict__ out)
{
    using C = Conv2D;
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr bool F16 = MODE == FP16, ONE_ACC = MODE != BF16X3;      // FP16 / FP16C: every MMA of a tile lands in the same 64 columns
    constexpr int NACC = ONE_ACC ? 2 * C::NACC : C::NACC, ACC_COLS = ONE_ACC ? C::NOUT : C::ACC_COLS;
    __shared__ uint64_t bar_in_full[2], bar_in_empty[2], bar_acc_full[2 * C::NACC], bar_acc_empty[2 * C::NACC], bar_w_full;
    __shared__ uint32_t s_tmem;
    uint8_t *s_in = smem;                                            // [2][IN_BYTES]
    uint8_t *s_w = smem + 2 * C::IN_BYTES;
    float *s_sc = reinterpret_cast<float *>(s_w + C::W_BYTES), *s_sh = s_sc + C::NOUT;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_act = n_dev ? min((int)*n_dev - base, n_max) : n_max;
    const int n_items = max(n_act, 0) * C::BANDS;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) { umma::mbar_init(&bar_in_full[i], 1); umma::mbar_init(&bar_in_empty[i], 1); }
        for (int i = 0; i < NACC; ++i) { umma::mbar_init(&bar_acc_full[i], 1); umma::mbar_init(&bar_acc_empty[i], 8); }
        umma::mbar_init(&bar_w_full, 1);
        umma::fence_mbar_init();
    }
    if (warp == 1) umma::tmem_alloc(&s_tmem, 512);
    for (int i = tid; i < C::NOUT; i += C::THREADS) { s_sc[i] = sc[i]; s_sh[i] = sh[i]; }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tm = s_tmem;

    if (warp == 0) {
        if (lane == 0) {
            umma::mbar_expect_tx(&bar_w_full, C::W_BYTES);
            for (int o = 0; o < C::W_BYTES; o += C::WTAP_BYTES) umma::bulk_g2s(s_w + o, wgt + o, C::WTAP_BYTES, &bar_w_full);
            uint32_t it = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const int img = item / C::BANDS, band = item % C::BANDS;
                const int y0 = band == 2 ? 24 : band * 16;
                const uint32_t b = it & 1;
                umma::mbar_wait(&bar_in_empty[b], ((it >> 1) & 1) ^ 1);
                constexpr int NPL = F16 ? C::G : 2 * C::G;                 // fp16: hi planes only
                umma::mbar_expect_tx(&bar_in_full[b], NPL * C::BAND_POS * 16);
                const uint8_t *src = in + (size_t)img * Conv2Cfg::IMG_BYTES + (size_t)y0 * C::WP * 16;
                for (int p = 0; p < NPL; ++p)
                    umma::bulk_g2s(s_in + (size_t)b * C::IN_BYTES + (size_t)p * C::BAND_POS * 16, src + (size_t)p * Conv2Cfg::PL * 16,
                                   C::BAND_POS * 16, &bar_in_full[b]);
            }
        }
    } else if (warp == 1) {
        if (umma::elect_one()) {
            // weights per tap: [group][64 rows W_hi + 64 rows W_lo][8]; one N = 128 MMA gives A_hi*(W_hi | W_lo),
            // one N = 64 MMA adds A_lo*W_hi: 14 KB of operand reads per tap instead of 18 KB for three N = 64 MMAs
            const uint32_t idesc128 = umma::idesc_bf16_f32(128, 2 * C::NOUT);
            const uint32_t idesc64 = MODE != BF16X3 ? umma::idesc_f16_f32(128, C::NOUT) : umma::idesc_bf16_f32(128, C::NOUT);
            const uint32_t idesc8 = umma::idesc_e5m2_f32(128, C::NOUT);
            const uint64_t w_base = umma::smem_desc(umma::smem_u32(s_w), 2 * C::NOUT * 16, 128);
            umma::mbar_wait(&bar_w_full, 0);
            uint32_t it = 0, ai = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const uint32_t b = it & 1;
                const uint64_t a_base = umma::smem_desc(umma::smem_u32(s_in + (size_t)b * C::IN_BYTES), C::BAND_POS * 16, C::WP * 16);
                umma::mbar_wait(&bar_in_full[b], (it >> 1) & 1);
                umma::fence_after_sync();
#pragma unroll 1
                for (int tx = 0; tx < C::TILES; ++tx, ++ai) {
                    const uint32_t buf = ai % NACC;
                    umma::mbar_wait(&bar_acc_empty[buf], ((ai / NACC) & 1) ^ 1);
                    umma::fence_after_sync();
                    const uint32_t d = tm + buf * ACC_COLS;
                    // fully unrolled: per MMA the issuing thread executes two immediate adds and the MMA itself
                    const uint64_t a_tile = umma::desc_add(a_base, (uint32_t)(tx * 8));
#pragma unroll
                    for (int tap = 0; tap < 25; ++tap) {
                        const uint32_t pos = (uint32_t)((tap / 5) * C::WP + (tap % 5));
                        const uint32_t a_hi = 0 * C::G * C::BAND_POS + pos, a_lo = 1 * C::G * C::BAND_POS + pos;
                        const uint32_t w = (uint32_t)(tap * C::WTAP_BYTES >> 4);
                        if (MODE != BF16X3) {
                            umma::mma_bf16(d, umma::desc_add(a_tile, a_hi), umma::desc_add(w_base, w), idesc64, tap != 0);   // fp16 x fp16, one MMA per tap
                            // FP16C: + the e5m2 correction MMA: A = the lo-slot plane pair, B = the "lo rows" (64 ..127) of the two groups
                            if (MODE == FP16C) umma::mma_f8(d, umma::desc_add(a_tile, a_lo), umma::desc_add(w_base, w + (uint32_t)C::NOUT), idesc8, 1);
                        } else {
                            umma::mma_bf16(d, umma::desc_add(a_tile, a_hi), umma::desc_add(w_base, w), idesc128, tap != 0);
                            umma::mma_bf16(d, umma::desc_add(a_tile, a_lo), umma::desc_add(w_base, w), idesc64, 1);
                        }
                    }
                    umma::commit(&bar_acc_full[buf]);
                }
                umma::commit(&bar_in_empty[b]);
            }
        }
    } else {
        const int ew = (warp - 2) & 7, set = (warp - 2) >> 3, quarter = warp & 3, half = ew >> 2;     // half: which 32 of the 64 output channels
        const int r = quarter * 32 + lane, ty = r >> 3, tx8 = r & 7;
        const bool p1 = tx8 & 1, p2 = ty & 1;
        const int g = half * 4 + (p1 ? 2 : 0) + (p2 ? 1 : 0);           // channel group this lane ends up owning
        uint32_t ai = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int img = item / C::BANDS, band = item % C::BANDS;
            const int y0 = band == 2 ? 24 : band * 16, ymin = band == 2 ? 32 : y0;
            const int y = y0 + ty;
            const bool blk_ok = (y & ~1) >= ymin && (y & ~1) < C::H;      // the 2x2 block of this lane quad is stored
#pragma unroll 1
            for (int tx = 0; tx < C::TILES; ++tx, ++ai) {
                if ((int)(ai & (C::EPI_SETS - 1)) != set) continue;           // the other set's tile
                const uint32_t buf = ai % NACC;
                umma::mbar_wait(&bar_acc_full[buf], (ai / NACC) & 1);
                umma::fence_after_sync();
                uint32_t v[32], vb[32];
                const uint32_t ta = tm + ((uint32_t)(quarter * 32) << 16) + buf * ACC_COLS + half * 32;
                umma::tmem_ld16(ta, *reinterpret_cast<uint32_t (*)[16]>(&v[0]));
                umma::tmem_ld16(ta + 16, *reinterpret_cast<uint32_t (*)[16]>(&v[16]));
                if (!ONE_ACC) {
                    umma::tmem_ld16(ta + C::NOUT, *reinterpret_cast<uint32_t (*)[16]>(&vb[0]));
                    umma::tmem_ld16(ta + C::NOUT + 16, *reinterpret_cast<uint32_t (*)[16]>(&vb[16]));
                }
                umma::tmem_ld_wait();
                umma::fence_before_sync();
                __syncwarp();
                if (lane == 0) umma::mbar_arrive(&bar_acc_empty[buf]);     // values are in registers: buffer reusable
                if (!ONE_ACC) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(vb[j]));
                }
                // 2x2 max-pool as a butterfly over the lane quad (r, r^1, r^8, r^9): after the exchange with
                // lane^1 a lane keeps 16 of its 32 channels, after lane^8 it keeps 8 = one channel group.
                // BN scale is folded into the weights, so pooling runs on raw accumulators (+shift is monotone).
                float m1[16], m2[8];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float lo_v = __uint_as_float(v[j]), hi_v = __uint_as_float(v[j + 16]);
                    const float keep = p1 ? hi_v : lo_v, send = p1 ? lo_v : hi_v;
                    m1[j] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 1));
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float keep = p2 ? m1[j + 8] : m1[j], send = p2 ? m1[j] : m1[j + 8];
                    m2[j] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 8));
                }
                if (blk_ok) {
                    const float4 t0 = *reinterpret_cast<const float4 *>(s_sh + g * 8), t1 = *reinterpret_cast<const float4 *>(s_sh + g * 8 + 4);
                    const float tt[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int j = 0; j < 8; j += 2) {
                        __nv_bfloat16 h0, l0, h1, l1;
                        umma::split_bf16(fmaxf(m2[j] + tt[j], 0.f), h0, l0);
                        umma::split_bf16(fmaxf(m2[j + 1] + tt[j + 1], 0.f), h1, l1);
                        hi[j >> 1] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                        lo[j >> 1] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
                    }
                    constexpr int WPN = Conv3Cfg::WP, PLN = Conv3Cfg::PL, GN = Conv3Cfg::G;
                    const int pos = ((y >> 1) + 2) * WPN + ((tx * 8 + tx8) >> 1) + 2;
                    if (F16) {
#pragma unroll
                        for (int j = 0; j < 8; j += 2)
                            hi[j >> 1] = (uint32_t)__half_as_ushort(__float2half_rn(fmaxf(m2[j] + tt[j], 0.f))) |
                                         ((uint32_t)__half_as_ushort(__float2half_rn(fmaxf(m2[j + 1] + tt[j + 1], 0.f))) << 16);
                    }
                    if (MODE == FP16C) {        // fp16 in the hi slot; this group's 8 + 8 correction bytes in the lo-slot plane pair of its 16-channel block
                        float mm[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) mm[j] = fmaxf(m2[j] + tt[j], 0.f);
                        uint2 l8, h8;
                        fp16c_pack(mm, hi, l8, h8);
                        uint8_t *lo_base = out + ((((size_t)img * 2 + 1) * GN + (g & ~1)) * PLN + pos) * 16 + (g & 1) * 8;
                        *reinterpret_cast<uint2 *>(lo_base) = l8;
                        *reinterpret_cast<uint2 *>(lo_base + (size_t)PLN * 16) = h8;
                    }
                    *reinterpret_cast<uint4 *>(out + ((((size_t)img * 2 + 0) * GN + g) * PLN + pos) * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    if (MODE == BF16X3) *reinterpret_cast<uint4 *>(out + ((((size_t)img * 2 + 1) * GN + g) * PLN + pos) * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
            }
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 1) umma::tmem_dealloc(tm, 512);
}

// ------------------------------------------------------------------------------------------------
// conv2, "tap-pair" variant for the FP16 / FP16C arithmetic.  conv2_2d_kernel is bound by the tensor core's shared-memory operand
// feed (position-major: M = 128 positions, N = 64 channels, so every 32-cycle MMA reads 4 KB of A and 2 KB of B: ~97 B / cycle measured,
// 62 cycles per MMA).  Here two filter rows share one MMA: the B operand holds the weights of tap (dy, dx) in columns 0 .. 63 and of tap
// (dy + 1, dx) in columns 64 .. 127 (N = 128), the A operand is read once for both.  With the common A start q0 + dy * WP + dx the second
// block is the contribution of tap (dy + 1, dx) to the output position ONE ROW UP, so
//     out[ty][tx][c] = D[ty][tx][c] + D[ty + 1][tx][64 + c] :
// in the 8 x 16 tile TMEM lane r + 8 is the same pixel one row down -- a shuffle by 8 lanes inside a warp, and 8 lanes per warp boundary
// through a 1 KB shared-memory hand-over.  Row 15 of a tile has no partner: a band yields 14 output rows (bands at y0 = 0, 14, 26 cover the
// 40 rows like the 16-row bands at 0, 16, 24 did).  Per tile: 10 pair MMAs (filter rows 0+1, 2+3) + 5 single N = 64 MMAs (row 4):
// 110 KB of operand reads instead of 150 KB per arithmetic term.  Weights: 10 pair slots of 8 KB ([fp16: g0 128 rows | g1 128 rows]
// [e5m2: plane0 128 rows | plane1 128 rows]) then 5 single slots of 4 KB (64-row blocks), 100 KB like the other variant.
// ------------------------------------------------------------------------------------------------
#ifdef TB_CONV2_STATS
// bring-up counters (cycles, summed over CTAs): [0] MMA thread total, [1] wait own band, [2] wait peer band, [3] wait accumulator free,
// [4] epilogue warp total, [5] wait accumulator full, [6] hand-over barriers, [7] tiles seen by that warp, [8] producer wait ring slot
__device__ unsigned long long g_conv2_stats[16];
#define C2S_DECL long long c2s_t0 = clock64(), c2s_a = 0, c2s_b = 0, c2s_c = 0, c2s_d = 0, c2s_e = 0, c2s_f = 0, c2s_t
#define C2S_BEGIN c2s_t = clock64()
#define C2S_END(x) x += clock64() - c2s_t
#define C2S_FLUSH(i0, i1, i2, i3) do { atomicAdd(&g_conv2_stats[i0], (unsigned long long)(clock64() - c2s_t0)); atomicAdd(&g_conv2_stats[i1], (unsigned long long)c2s_a); \
    atomicAdd(&g_conv2_stats[i2], (unsigned long long)c2s_b); atomicAdd(&g_conv2_stats[i3], (unsigned long long)c2s_c); } while (0)
#else
#define C2S_DECL
#define C2S_BEGIN
#define C2S_END(x)
#define C2S_FLUSH(i0, i1, i2, i3)
#endif

struct Conv2P {
    using D = Conv2D;
    static constexpr int ROWS_OUT = 14, NACC = 4, ACC_COLS = 2 * D::NOUT;
    static constexpr int PAIR_BYTES = 8192, SINGLE_BYTES = 4096, W_BYTES = 10 * PAIR_BYTES + 5 * SINGLE_BYTES;
    static constexpr int XCH_BYTES = 2 * D::EPI_SETS * 2 * 3 * 32 * 8 * 4;    // [parity][set][half][boundary][32 threads][8 values] f32: alternate tiles of a set use alternate buffers
    // NCTA = 2 (CTA pair, cta_group::2): every CTA holds the half of each weight slot that its SM feeds -- rank r the 64 rows of tap (dy + r, dx) of a
    // pair slot, output channels 32 r .. 32 r + 31 of a single slot.
    // Input ring: only 18 of a band's 20 input rows feed stored outputs (row 13 = A[13] + B[14] reads rows <= 17), so a plane is loaded as 18 rows and
    // the planes sit 18 rows apart; the MMAs of the discarded tile rows 14, 15 read on into the next plane / the next region (mapped, never stored).
    static constexpr int LROWS = 18, PPOS = LROWS * D::WP;
    static constexpr int planes(bool f16) { return f16 ? D::G : 2 * D::G; }
    static constexpr int stage_bytes(bool f16) { return planes(f16) * PPOS * 16; }
    static constexpr int stages(bool f16, int ncta) { return ncta == 2 ? (f16 ? 4 : 3) : 2; }
    static constexpr int MAX_STAGES = 4;
    static constexpr int smem(bool f16, int ncta) { return stages(f16, ncta) * stage_bytes(f16) + W_BYTES / ncta + D::NOUT * 8 + XCH_BYTES + 128; }
    static_assert(W_BYTES == D::W_BYTES, "same weight buffer size as conv2_2d_kernel");
    __host__ __device__ static constexpr int y0(int band) { return band == 0 ? 0 : (band == 1 ? 14 : 26); }
    __host__ __device__ static constexpr int ymin(int band) { return band == 0 ? 0 : (band == 1 ? 14 : 28); }
};
static_assert(Conv2P::smem(false, 1) <= 227 * 1024 && Conv2P::smem(false, 2) <= 227 * 1024 && Conv2P::smem(true, 2) <= 227 * 1024, "shared memory budget");

template <int MODE, int NCTA>
__global__ void __launch_bounds__(Conv2D::THREADS, 1)
conv2_pair_kernel(const uint8_t *__restrict__ in, int n_max, const uint32_t *__restrict__ n_dev, int base,
                  const uint8_t *__restrict__ wgt, const float *__restrict__ sc, const float *__restrict__ sh,
                  uint8_t *__restrict__ out)
{
    static_assert(MODE == FP16 || MODE == FP16C, "the tap-pair kernel is built for the fp16 arithmetics");
    using C = Conv2D;
    using P = Conv2P;
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr bool F16 = MODE == FP16;
    constexpr int NACC = P::NACC, ACC_COLS = P::ACC_COLS;
    constexpr int WB = P::W_BYTES / NCTA, PAIR_B = P::PAIR_BYTES / NCTA, SINGLE_B = P::SINGLE_BYTES / NCTA;     // this CTA's share of the weights
    constexpr int STAGES = P::stages(F16, NCTA), STAGE_B = P::stage_bytes(F16), PPOS = P::PPOS;
    __shared__ uint64_t bar_in_full[P::MAX_STAGES], bar_in_empty[P::MAX_STAGES], bar_acc_full[P::NACC], bar_acc_empty[P::NACC], bar_w_full, bar_peer_full[P::MAX_STAGES];
    __shared__ uint32_t s_tmem;
    uint8_t *s_in = smem;                                            // [STAGES][STAGE_B]
    uint8_t *s_w = smem + STAGES * STAGE_B;
    float *s_sc = reinterpret_cast<float *>(s_w + WB), *s_sh = s_sc + C::NOUT;
    float *s_x = s_sh + C::NOUT;                                     // row hand-over between the warps of a tile
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_act = n_dev ? min((int)*n_dev - base, n_max) : n_max;
    const int n_items = max(n_act, 0) * C::BANDS;
    // NCTA = 2: the two CTAs of a pair walk the item list in lockstep (items 2 q + rank; the odd one out is a dummy: no loads, no stores), because
    // every MMA the leader issues runs on both SMs, each on its own band
    const int rank = NCTA == 2 ? (int)umma::cluster_ctarank() : 0;
    const int first = NCTA == 2 ? (int)(blockIdx.x >> 1) * 2 + rank : (int)blockIdx.x, stride = (int)gridDim.x;
    const int n_loop = NCTA == 2 ? (n_items + 1) & ~1 : n_items;       // item < n_loop: both CTAs of a pair iterate; item >= n_items: the dummy

    if (tid == 0) {
        for (int i = 0; i < STAGES; ++i) { umma::mbar_init(&bar_in_full[i], 1); umma::mbar_init(&bar_in_empty[i], 1); umma::mbar_init(&bar_peer_full[i], 1); }
        for (int i = 0; i < NACC; ++i) { umma::mbar_init(&bar_acc_full[i], 1); umma::mbar_init(&bar_acc_empty[i], 8 * NCTA); }
        umma::mbar_init(&bar_w_full, 1);
        umma::fence_mbar_init();
    }
    if (warp == 1) { if (NCTA == 2) umma::tmem_alloc2(&s_tmem, 512); else umma::tmem_alloc(&s_tmem, 512); }
    for (int i = tid; i < C::NOUT; i += C::THREADS) { s_sc[i] = sc[i]; s_sh[i] = sh[i]; }
    umma::fence_before_sync();
    __syncthreads();                                                   // s_tmem, s_sc / s_sh written (racecheck models this barrier, not the cluster one)
    if (NCTA == 2) umma::cluster_sync();                               // barriers of both CTAs initialised before any remote arrive
    umma::fence_after_sync();
    const uint32_t tm = s_tmem;

    if (warp == 0) {
        if (lane == 0) {
            umma::mbar_expect_tx(&bar_w_full, WB);
            for (int o = 0; o < WB; o += 2048) umma::bulk_g2s(s_w + o, wgt + (size_t)rank * WB + o, 2048, &bar_w_full);
            uint32_t it = 0;
            for (int item = first; item < n_loop; item += stride, ++it) {
                const int img = item / C::BANDS, band = item % C::BANDS;
                const int y0 = P::y0(band);
                const uint32_t b = it % STAGES;
                umma::mbar_wait(&bar_in_empty[b], ((it / STAGES) & 1) ^ 1);
                if (item >= n_items) { umma::mbar_arrive(&bar_in_full[b]); continue; }       // dummy: whatever the buffer holds is multiplied and dropped
                constexpr int NPL = P::planes(F16);                        // fp16: hi planes only
                umma::mbar_expect_tx(&bar_in_full[b], (uint32_t)STAGE_B);
                const uint8_t *src = in + (size_t)img * Conv2Cfg::IMG_BYTES + (size_t)y0 * C::WP * 16;
                for (int p = 0; p < NPL; ++p)
                    umma::bulk_g2s(s_in + (size_t)b * STAGE_B + (size_t)p * PPOS * 16, src + (size_t)p * Conv2Cfg::PL * 16, (uint32_t)(PPOS * 16), &bar_in_full[b]);
            }
        }
    } else if (warp == 1 && NCTA == 2 && rank == 1) {
        // the peer's MMA warp only relays: weights and every input band that landed in THIS CTA's shared memory are announced to the leader
        if (lane == 0) {
            umma::mbar_wait(&bar_w_full, 0);
            uint32_t it = 0;
            for (int item = first; item < n_loop; item += stride, ++it) {
                const uint32_t b = it % STAGES;
                umma::mbar_wait(&bar_in_full[b], (it / STAGES) & 1);
                umma::mbar_arrive_remote(umma::mapa(umma::smem_u32(&bar_peer_full[b]), 0));
            }
        }
    } else if (warp == 1) {
        if (umma::elect_one()) {
            const uint32_t id128 = umma::idesc_f16_f32(128 * NCTA, 2 * C::NOUT), id64 = umma::idesc_f16_f32(128 * NCTA, C::NOUT);
            const uint32_t id128_8 = umma::idesc_e5m2_f32(128 * NCTA, 2 * C::NOUT), id64_8 = umma::idesc_e5m2_f32(128 * NCTA, C::NOUT);
            const uint64_t wp_base = umma::smem_desc(umma::smem_u32(s_w), 2 * C::NOUT / NCTA * 16, 128);     // pair slots: 128-row blocks (64 per CTA of a pair)
            const uint64_t ws_base = umma::smem_desc(umma::smem_u32(s_w), C::NOUT / NCTA * 16, 128);         // single slots: 64-row blocks (32 per CTA of a pair)
            constexpr uint32_t F8P = (uint32_t)(PAIR_B / 2) >> 4, F8S = (uint32_t)(SINGLE_B / 2) >> 4;       // the e5m2 half of a slot, in 16-byte units
            umma::mbar_wait(&bar_w_full, 0);
            uint32_t it = 0, ai = 0;
            C2S_DECL;
            for (int item = first; item < n_loop; item += stride, ++it) {
                const uint32_t b = it % STAGES;
                const uint64_t a_base = umma::smem_desc(umma::smem_u32(s_in + (size_t)b * STAGE_B), PPOS * 16, C::WP * 16);
                C2S_BEGIN;
                umma::mbar_wait(&bar_in_full[b], (it / STAGES) & 1);
                C2S_END(c2s_a); C2S_BEGIN;
                if (NCTA == 2) umma::mbar_wait_cluster(&bar_peer_full[b], (it / STAGES) & 1);      // ... and the peer's band (its weights arrived before its first band)
                C2S_END(c2s_b);
                umma::fence_after_sync();
#pragma unroll 1
                for (int tx = 0; tx < C::TILES; ++tx, ++ai) {
                    const uint32_t buf = ai % NACC;
                    C2S_BEGIN;
                    umma::mbar_wait(&bar_acc_empty[buf], ((ai / NACC) & 1) ^ 1);
                    C2S_END(c2s_c);
                    umma::fence_after_sync();
                    const uint32_t d = tm + buf * ACC_COLS;
                    const uint64_t a_tile = umma::desc_add(a_base, (uint32_t)(tx * 8));
#pragma unroll
                    for (int sl = 0; sl < 10; ++sl) {                                   // filter rows (0, 1) and (2, 3), 5 columns each
                        const uint32_t pos = (uint32_t)((2 * (sl / 5)) * C::WP + (sl % 5));
                        const uint32_t w = (uint32_t)(sl * PAIR_B >> 4);
                        if (NCTA == 2) {
                            umma::mma2_f16(d, umma::desc_add(a_tile, pos), umma::desc_add(wp_base, w), id128, sl != 0);
                            if (MODE == FP16C) umma::mma2_f8(d, umma::desc_add(a_tile, (uint32_t)(C::G * PPOS) + pos), umma::desc_add(wp_base, w + F8P), id128_8, 1);
                        } else {
                            umma::mma_bf16(d, umma::desc_add(a_tile, pos), umma::desc_add(wp_base, w), id128, sl != 0);
                            if (MODE == FP16C) umma::mma_f8(d, umma::desc_add(a_tile, (uint32_t)(C::G * PPOS) + pos), umma::desc_add(wp_base, w + F8P), id128_8, 1);
                        }
                    }
#pragma unroll
                    for (int dx = 0; dx < 5; ++dx) {                                    // filter row 4: columns 0 .. 63 only
                        const uint32_t pos = (uint32_t)(4 * C::WP + dx);
                        const uint32_t w = (uint32_t)((10 * PAIR_B + dx * SINGLE_B) >> 4);
                        if (NCTA == 2) {
                            umma::mma2_f16(d, umma::desc_add(a_tile, pos), umma::desc_add(ws_base, w), id64, 1);
                            if (MODE == FP16C) umma::mma2_f8(d, umma::desc_add(a_tile, (uint32_t)(C::G * PPOS) + pos), umma::desc_add(ws_base, w + F8S), id64_8, 1);
                        } else {
                            umma::mma_bf16(d, umma::desc_add(a_tile, pos), umma::desc_add(ws_base, w), id64, 1);
                            if (MODE == FP16C) umma::mma_f8(d, umma::desc_add(a_tile, (uint32_t)(C::G * PPOS) + pos), umma::desc_add(ws_base, w + F8S), id64_8, 1);
                        }
                    }
                    if (NCTA == 2) umma::commit2(&bar_acc_full[buf], 3); else umma::commit(&bar_acc_full[buf]);
                }
                if (NCTA == 2) umma::commit2(&bar_in_empty[b], 3); else umma::commit(&bar_in_empty[b]);
            }
            C2S_FLUSH(0, 1, 2, 3);
        }
    } else {
        // Epilogue with the 16x256b TMEM load shape (the mma accumulator fragment): thread t of a warp holds, for pixel column tx = t / 4 of
        // the tile, the FOUR tile rows of the warp's lane quarter (lanes t/4, t/4 + 8 of each 16-lane half) and two adjacent channels per
        // 8-channel group -- so "lane r + 8" (the same pixel one row down) is in the same thread: the tap-pair sum and the vertical half of
        // the max-pool are thread-local, the horizontal half is one exchange with thread t ^ 4, and only a warp's last row takes its block-B
        // partner from the next warp (8 values per thread through shared memory).
        const int ew = (warp - 2) & 7, set = (warp - 2) >> 3, quarter = warp & 3, half = ew >> 2;     // half: which 32 of the 64 output channels
        const int txq = lane >> 2, cp = lane & 3;                      // pixel column inside the tile, channel pair inside a group
        constexpr int XCH_HALF = C::EPI_SETS * 2 * 3 * 32 * 8;        // floats per hand-over buffer; two buffers, so ONE barrier per tile orders
                                                                       // write -> read AND read -> next write into the same buffer (two tiles later)
        // [boundary][2 x 16-byte halves][32 lanes][4 floats]: every 128-bit access of a warp covers 512 contiguous bytes (no bank conflicts)
        float *xw = s_x + (size_t)((set * 2 + half) * 3 + (quarter - 1)) * 256 + lane * 4;       // written by quarters 1 .. 3
        const float *xr = s_x + (size_t)((set * 2 + half) * 3 + quarter) * 256 + lane * 4;       // read by quarters 0 .. 2
        const int prow = 2 * quarter + (txq & 1);                      // pooled row (inside the band) this thread ends up owning
        // this thread's 8 shifts (channels (half * 4 + j) * 8 + 2 cp + e), contiguous in shared memory: two 16-byte loads per tile instead of 8 live registers
        float *s_shp = s_sc;                                           // the scale slot is free (the BN scale is folded into the weights)
        if (set == 0 && quarter == 0 && lane < 4) {
#pragma unroll
            for (int j = 0; j < 4; ++j) { s_shp[(half * 4 + cp) * 8 + 2 * j] = sh[(half * 4 + j) * 8 + 2 * cp]; s_shp[(half * 4 + cp) * 8 + 2 * j + 1] = sh[(half * 4 + j) * 8 + 2 * cp + 1]; }
        }
        asm volatile("bar.sync 5, %0;" :: "r"(C::EPI_SETS * 256) : "memory");
        const float4 *shp = reinterpret_cast<const float4 *>(s_shp + (half * 4 + cp) * 8);
        constexpr int WPN = Conv3Cfg::WP, PLN = Conv3Cfg::PL, GN = Conv3Cfg::G;
        uint32_t ai = 0, par = 0;                                      // par: hand-over buffer of this set's next tile
        C2S_DECL;
        for (int item = first; item < n_loop; item += stride) {
            const int img = item / C::BANDS, band = item % C::BANDS;
            const int y0 = P::y0(band), ymin = P::ymin(band);
            const int y = y0 + 2 * prow;
            const bool blk_ok = 2 * prow + 1 < P::ROWS_OUT && y >= ymin && y < C::H && item < n_items;      // both rows of the 2x2 block are complete and stored by this band
            // stores of this item: fp16 plane of group half * 4 + j at o_hi + j * PLN * 16; e5m2 L plane of the group pair at o_lo + (j / 2) * 2 * PLN * 16
            // + (j & 1) * 8, H plane one plane further; + 64 bytes per tile
            const int pos0 = ((y >> 1) + 2) * WPN + (txq >> 1) + 2;
            uint8_t *o_hi = out + ((((size_t)img * 2 + 0) * GN + half * 4) * PLN + pos0) * 16 + cp * 4;
            uint8_t *o_lo = out + ((((size_t)img * 2 + 1) * GN + half * 4) * PLN + pos0) * 16 + cp * 2;
#pragma unroll 1
            for (int tx = 0; tx < C::TILES; ++tx, ++ai) {
                if ((int)(ai & (C::EPI_SETS - 1)) != set) continue;           // the other set's tile
                const uint32_t buf = ai % NACC;
                C2S_BEGIN;
                umma::mbar_wait(&bar_acc_full[buf], (ai / NACC) & 1);
                C2S_END(c2s_a);
                umma::fence_after_sync();
                uint32_t a0[16], a1[16], b0[16], b1[16];                   // rows (0, 1) and (2, 3) of the quarter: block A / block B
                C2S_BEGIN;
                const uint32_t ta = tm + ((uint32_t)(quarter * 32) << 16) + buf * ACC_COLS + half * 32;
                umma::tmem_ld_16x256b_x4(ta + C::NOUT, b0);
                umma::tmem_ld_16x256b_x4(ta + (16u << 16) + C::NOUT, b1);
                umma::tmem_ld_16x256b_x4(ta, a0);
                umma::tmem_ld_16x256b_x4(ta + (16u << 16), a1);
                umma::tmem_ld_wait();
                C2S_END(c2s_d);
                umma::fence_before_sync();
                __syncwarp();
                if (lane == 0) {                                           // values are in registers: buffer reusable (the leader's barrier counts both CTAs)
                    if (NCTA == 2 && rank == 1) umma::mbar_arrive_remote_relaxed(umma::mapa(umma::smem_u32(&bar_acc_empty[buf]), 0));
                    else umma::mbar_arrive(&bar_acc_empty[buf]);
                }
                if (quarter > 0) {                                         // block B of this warp's first row: the partner of the previous warp's last row
                    *reinterpret_cast<float4 *>(xw + par * XCH_HALF) = make_float4(__uint_as_float(b0[0]), __uint_as_float(b0[1]), __uint_as_float(b0[4]), __uint_as_float(b0[5]));
                    *reinterpret_cast<float4 *>(xw + par * XCH_HALF + 128) = make_float4(__uint_as_float(b0[8]), __uint_as_float(b0[9]), __uint_as_float(b0[12]), __uint_as_float(b0[13]));
                }
                C2S_BEGIN;
                asm volatile("bar.sync %0, 256;" :: "r"(1 + set) : "memory");
                C2S_END(c2s_b);
                float bn[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (quarter < 3) {
                    const float4 u0 = *reinterpret_cast<const float4 *>(xr + par * XCH_HALF), u1 = *reinterpret_cast<const float4 *>(xr + par * XCH_HALF + 128);
                    bn[0] = u0.x; bn[1] = u0.y; bn[2] = u0.z; bn[3] = u0.w; bn[4] = u1.x; bn[5] = u1.y; bn[6] = u1.z; bn[7] = u1.w;
                }
                par ^= 1u;
#ifdef TB_CONV2_STATS
                ++c2s_c;
#endif
                // out[k] = A[k] + B[k + 1]; pooled rows: max(out[0], out[1]) and max(out[2], out[3]); registers 4 j + 2 h + e
                C2S_BEGIN;
                float m[2][8];
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const float o0 = __uint_as_float(a0[4 * j + e]) + __uint_as_float(b0[4 * j + 2 + e]);
                        const float o1 = __uint_as_float(a0[4 * j + 2 + e]) + __uint_as_float(b1[4 * j + e]);
                        const float o2 = __uint_as_float(a1[4 * j + e]) + __uint_as_float(b1[4 * j + 2 + e]);
                        const float o3 = __uint_as_float(a1[4 * j + 2 + e]) + bn[2 * j + e];
                        m[0][2 * j + e] = fmaxf(o0, o1);
                        m[1][2 * j + e] = fmaxf(o2, o3);
                    }
                // horizontal half of the pool: the even pixel column keeps pooled row 0, the odd one pooled row 1
                const float4 s0 = shp[0], s1 = shp[1];
                const float shv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
                float mm[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float keep = (txq & 1) ? m[1][k] : m[0][k], send = (txq & 1) ? m[0][k] : m[1][k];
                    mm[k] = fmaxf(fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 4)) + shv[k], 0.f);      // BN scale is in the weights; + shift, ReLU
                }
                C2S_END(c2s_e); C2S_BEGIN;
                if (blk_ok) {
                    uint8_t *oh = o_hi + tx * 64, *ol = o_lo + tx * 64;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const __half h0 = __float2half_rn(mm[2 * j]), h1 = __float2half_rn(mm[2 * j + 1]);
                        *reinterpret_cast<uint32_t *>(oh + (size_t)j * PLN * 16) = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
                        if (MODE == FP16C) {
                            const float f0 = __half2float(h0), f1 = __half2float(h1);
                            const uint16_t l2 = (uint16_t)__nv_cvt_float2_to_fp8x2(make_float2((mm[2 * j] - f0) * FC_UP, (mm[2 * j + 1] - f1) * FC_UP), __NV_SATFINITE, __NV_E5M2);
                            const uint16_t h2 = (uint16_t)__nv_cvt_float2_to_fp8x2(make_float2(f0 * FC_DOWN, f1 * FC_DOWN), __NV_SATFINITE, __NV_E5M2);
                            uint8_t *lo_base = ol + (size_t)(j >> 1) * 2 * PLN * 16 + (j & 1) * 8;
                            *reinterpret_cast<uint16_t *>(lo_base) = l2;
                            *reinterpret_cast<uint16_t *>(lo_base + (size_t)PLN * 16) = h2;
                        }
                    }
                }
                C2S_END(c2s_f);
            }
        }
#ifdef TB_CONV2_STATS
        if (warp == 2 && lane == 0) { C2S_FLUSH(4, 5, 6, 7); atomicAdd(&g_conv2_stats[8], (unsigned long long)c2s_d); atomicAdd(&g_conv2_stats[9], (unsigned long long)c2s_e); atomicAdd(&g_conv2_stats[10], (unsigned long long)c2s_f); }
#endif
    }
    umma::fence_before_sync();
    if (NCTA == 2) umma::cluster_sync(); else __syncthreads();         // the leader's last MMAs have read the peer's shared memory
    if (warp == 1) { if (NCTA == 2) umma::tmem_dealloc2(tm, 512); else umma::tmem_dealloc(tm, 512); }
}

// ------------------------------------------------------------------------------------------------
// fc1 on tensor cores: h1[128 images][112] = A[128][12800] * B[112][12800]^T (+ bias), bf16x3.
// One CTA per 128 images, 3-stage bulk-copy pipeline over K (4 k-steps of 16 per stage).
// ------------------------------------------------------------------------------------------------
constexpr int FC_STAGES = 3, FC_KS = 4;                                   // k-steps (of 16) per stage
constexpr int FC_A_STAGE = 2 * 16 * (2 * FC_KS) * 128;                    // [hl][16 groups][8 kc][128 B]
constexpr int FC_B_STAGE = 2 * (2 * FC_KS) * FC_N * 16;                   // [hl][8 kc][112][16 B]
constexpr int FC_SMEM = FC_STAGES * (FC_A_STAGE + FC_B_STAGE) + 128;

__global__ void __launch_bounds__(NT, 1)
fc1_tc_kernel(const uint8_t *__restrict__ a, int n_groups, int n_max, const uint32_t *__restrict__ n_dev, int base,
              const uint8_t *__restrict__ b, float *__restrict__ h1p /*[FC_SPLIT][n_alloc][100] partial sums*/, int n_alloc)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar_full[FC_STAGES], bar_empty[FC_STAGES], bar_acc;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_act = n_dev ? min((int)*n_dev - base, n_max) : n_max;
    const int i0 = blockIdx.x * 128, split = blockIdx.y;
    if (i0 >= n_act) return;
    if (tid == 0) {
        for (int i = 0; i < FC_STAGES; ++i) { umma::mbar_init(&bar_full[i], 1); umma::mbar_init(&bar_empty[i], 1); }
        umma::mbar_init(&bar_acc, 1);
        umma::fence_mbar_init();
    }
    if (warp == 1) umma::tmem_alloc(&s_tmem, 128);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tm = s_tmem;
    constexpr int NSTEP = FC_KB / FC_SPLIT;                               // kc blocks handled by this CTA
    static_assert(FC_KB % FC_SPLIT == 0 && 2 * FC_KS == 8, "stage = one kc block");
    const int kb0 = split * NSTEP;
    if (warp == 0) {
        if (lane == 0) {
            for (int st = 0; st < NSTEP; ++st) {
                const int s = st % FC_STAGES;
                umma::mbar_wait(&bar_empty[s], ((st / FC_STAGES) & 1) ^ 1);
                umma::mbar_expect_tx(&bar_full[s], FC_A_STAGE + FC_B_STAGE);
                uint8_t *sa = smem + (size_t)s * (FC_A_STAGE + FC_B_STAGE), *sb = sa + FC_A_STAGE;
                const int kb = kb0 + st;
                for (int hl = 0; hl < 2; ++hl) {
                    umma::bulk_g2s(sa + (size_t)hl * (FC_A_STAGE / 2), a + (((size_t)hl * FC_KB + kb) * n_groups + blockIdx.x * 16) * 1024, FC_A_STAGE / 2, &bar_full[s]);
                    umma::bulk_g2s(sb + (size_t)hl * (FC_B_STAGE / 2), b + (((size_t)hl * FC_KC + kb * 8) * FC_N * 16), FC_B_STAGE / 2, &bar_full[s]);
                }
            }
        }
    } else if (warp == 1) {
        if (umma::elect_one()) {
            const uint32_t idesc = umma::idesc_bf16_f32(128, FC_N);
            for (int st = 0; st < NSTEP; ++st) {
                const int s = st % FC_STAGES;
                umma::mbar_wait(&bar_full[s], (st / FC_STAGES) & 1);
                umma::fence_after_sync();
                const uint8_t *sa = smem + (size_t)s * (FC_A_STAGE + FC_B_STAGE), *sb = sa + FC_A_STAGE;
                const uint64_t a_hi = umma::smem_desc(umma::smem_u32(sa), 128, 2 * FC_KS * 128);
                const uint64_t a_lo = umma::smem_desc(umma::smem_u32(sa + FC_A_STAGE / 2), 128, 2 * FC_KS * 128);
                const uint64_t b_hi = umma::smem_desc(umma::smem_u32(sb), FC_N * 16, 128);
                const uint64_t b_lo = umma::smem_desc(umma::smem_u32(sb + FC_B_STAGE / 2), FC_N * 16, 128);
#pragma unroll
                for (int j = 0; j < FC_KS; ++j) {
                    const uint32_t ao = (uint32_t)(2 * j * 128) >> 4, bo = (uint32_t)(2 * j * FC_N * 16) >> 4;
                    umma::mma_bf16(tm, a_hi + ao, b_hi + bo, idesc, (st | j) != 0);
                    umma::mma_bf16(tm, a_lo + ao, b_hi + bo, idesc, 1);
                    umma::mma_bf16(tm, a_hi + ao, b_lo + bo, idesc, 1);
                }
                umma::commit(&bar_empty[s]);
            }
            umma::commit(&bar_acc);
        }
    } else {
        const int quarter = warp & 3, row = quarter * 32 + lane;
        umma::mbar_wait(&bar_acc, 0);
        umma::fence_after_sync();
        const int img = i0 + row;
        float *dst = h1p + ((size_t)split * n_alloc + img) * FC_NREAL;
#pragma unroll 1
        for (int c0 = 0; c0 < FC_N; c0 += 16) {
            uint32_t v[16];
            umma::tmem_ld16(tm + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
            umma::tmem_ld_wait();
            if (img < n_act)
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (c0 + j < FC_NREAL) dst[c0 + j] = __uint_as_float(v[j]);
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 1) umma::tmem_dealloc(tm, 128);
}

// ------------------------------------------------------------------------------------------------
// conv1 on tensor cores, pooled-window formulation.  One GEMM row = one POOLED pixel (py,px); its K = 48
// operand is the 6x6 input window that feeds the 2x2 block of conv outputs (6 window rows x 8 columns, the
// last two columns meet zero weights); the N = 64 columns are (position inside the 2x2 block) x 16 output
// channels, each with the 5x5 filter shifted inside the window.  The max-pool is then a thread-local max
// over 4 column groups: no shuffles.  C_in = 1, so the "channel group" of a position is the pixel and its
// right neighbours, decimated by two:  P2[yy][px][e] = padded_crop[yy][2*px + e]  (bf16, exact for u8).
// A k-step (K = 16) covers window rows u, u+1 via LBO = one P2 row; consecutive pooled rows are two P2
// rows apart (SBO).  Tiles are 8 pooled px x 16 pooled rows; W_hi and W_lo accumulate into the same columns.
// ------------------------------------------------------------------------------------------------
struct Conv1T {
    static constexpr int H = 80, W = 80, PW = 40, PROWS = 84, IMG_PITCH = 96;
    static constexpr int P_BYTES = PROWS * PW * 16, IMG_BYTES = PROWS * IMG_PITCH;
    static constexpr int N = 64, W_BYTES = 2 * 3 * 2 * N * 16;       // [hi|lo][k-step][k-chunk][64 rows][8]
    static constexpr int NACC = 8, TILES_X = 5, TILES_Y = 3, TILES = TILES_X * TILES_Y;   // y0 = 0, 16, 24
    static constexpr int smem(int cin) { return cin * (P_BYTES + IMG_BYTES + W_BYTES) + 64 + 128; }   // 3 channels: 217 KB
    static constexpr int THREADS = 64 + 256;
};

// CIN = 3 (rgb8 crops, NHWC u8): one decimated plane, one padded image and one weight block per input channel;
// the three channels accumulate into the same TMEM columns (3 x 6 MMAs per tile).
template <int CIN>
__global__ void __launch_bounds__(Conv1T::THREADS, 1)
conv1_tc_kernel(const uint8_t *__restrict__ img, int n_max, const uint32_t *__restrict__ n_dev, int base,
                const uint8_t *__restrict__ wgt, const float *__restrict__ sh, uint8_t *__restrict__ out, int f16out)
{
    using C = Conv1T;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar_acc_full[C::NACC], bar_acc_empty[C::NACC];
    __shared__ uint32_t s_tmem;
    uint4 *s_p = reinterpret_cast<uint4 *>(smem);                      // [CIN] planes
    uint8_t *s_img = smem + CIN * C::P_BYTES;                          // [CIN] padded images
    uint8_t *s_w = s_img + CIN * C::IMG_BYTES;                         // [CIN] weight blocks
    float *s_sh = reinterpret_cast<float *>(s_w + CIN * C::W_BYTES);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_act = n_dev ? min((int)*n_dev - base, n_max) : n_max;

    if (tid == 0) {
        for (int i = 0; i < C::NACC; ++i) { umma::mbar_init(&bar_acc_full[i], 1); umma::mbar_init(&bar_acc_empty[i], 4); }
        umma::fence_mbar_init();
    }
    if (warp == 1) umma::tmem_alloc(&s_tmem, 512);
    for (int i = tid; i < CIN * C::W_BYTES / 4; i += C::THREADS) reinterpret_cast<uint32_t *>(s_w)[i] = reinterpret_cast<const uint32_t *>(wgt)[i];
    if (tid < 16) s_sh[tid] = sh[tid];
    for (int i = tid; i < CIN * C::IMG_BYTES / 4; i += C::THREADS) reinterpret_cast<uint32_t *>(s_img)[i] = 0u;   // zero halo, kept
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tm = s_tmem;
    uint32_t ai = 0;                                                   // running tile counter (ring phase)

    for (int n = blockIdx.x; n < n_act; n += gridDim.x) {
        // ---- stage the crop: interior of the zero-padded u8 image, then the decimated bf16 plane ----
        const uint32_t *src = reinterpret_cast<const uint32_t *>(img + (size_t)n * C::H * C::W * CIN);
        for (int i = tid; i < C::H * C::W / 4; i += C::THREADS) {
            const int y = i / (C::W / 4), x4 = i % (C::W / 4);
            if (CIN == 1) {
                const uint32_t v = src[i];
                uint16_t *d16 = reinterpret_cast<uint16_t *>(s_img + (y + 2) * C::IMG_PITCH + 2 + x4 * 4);   // 2-byte aligned
                d16[0] = (uint16_t)v; d16[1] = (uint16_t)(v >> 16);
            } else {                                                   // 4 interleaved B,G,R pixels = 12 bytes -> 4 bytes per channel
                const uint32_t w0 = src[3 * i], w1 = src[3 * i + 1], w2 = src[3 * i + 2];
                const uint32_t ch[3] = {__byte_perm(__byte_perm(w0, w1, 0x0630), w2, 0x5210),    // bytes 0,3,6,9
                                        __byte_perm(__byte_perm(w0, w1, 0x0741), w2, 0x6210),    // bytes 1,4,7,10
                                        __byte_perm(__byte_perm(w0, w1, 0x0052), w2, 0x7410)};   // bytes 2,5,8,11
#pragma unroll
                for (int ci = 0; ci < 3; ++ci) {
                    uint16_t *d16 = reinterpret_cast<uint16_t *>(s_img + ci * C::IMG_BYTES + (y + 2) * C::IMG_PITCH + 2 + x4 * 4);
                    d16[0] = (uint16_t)ch[ci]; d16[1] = (uint16_t)(ch[ci] >> 16);
                }
            }
        }
        __syncthreads();
        // four pooled positions per step: 14 source bytes -> bf16 once; position j uses bytes 2j .. 2j+7
        for (int i = tid; i < CIN * C::PROWS * (C::PW / 4); i += C::THREADS) {
            const int ci = i / (C::PROWS * (C::PW / 4)), ii = i % (C::PROWS * (C::PW / 4));
            const int yy = ii / (C::PW / 4), p4 = ii % (C::PW / 4);
            const uint2 wa = *reinterpret_cast<const uint2 *>(s_img + ci * C::IMG_BYTES + yy * C::IMG_PITCH + p4 * 8);
            const uint2 wb2 = *reinterpret_cast<const uint2 *>(s_img + ci * C::IMG_BYTES + yy * C::IMG_PITCH + p4 * 8 + 8);
            const uint32_t ws[4] = {wa.x, wa.y, wb2.x, wb2.y};
            uint32_t ev[7];
#pragma unroll
            for (int k = 0; k < 7; ++k) {                      // bf16 pair (byte 2k, byte 2k+1): high halves of the fp32 patterns
                const uint32_t b0 = (ws[(2 * k) >> 2] >> (8 * ((2 * k) & 3))) & 0xFFu, b1 = (ws[(2 * k + 1) >> 2] >> (8 * ((2 * k + 1) & 3))) & 0xFFu;
                ev[k] = __byte_perm(__float_as_uint((float)b0), __float_as_uint((float)b1), 0x7632);
            }
            uint4 *dst = s_p + ci * (C::P_BYTES / 16) + yy * C::PW + p4 * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) dst[j] = make_uint4(ev[j], ev[j + 1], ev[j + 2], ev[j + 3]);
        }
        umma::fence_async_smem();
        __syncthreads();
        umma::fence_after_sync();

        if (warp == 1) {
            if (umma::elect_one()) {
                const uint32_t idesc = umma::idesc_bf16_f32(128, C::N);
                const uint64_t a_base = umma::smem_desc(umma::smem_u32(s_p), C::PW * 16, 2 * C::PW * 16);
                const uint64_t w_base = umma::smem_desc(umma::smem_u32(s_w), C::N * 16, 128);
                uint32_t a2 = ai;
                for (int t = 0; t < C::TILES; ++t, ++a2) {
                    const int ty = t / C::TILES_X, tx = t % C::TILES_X;
                    const int y0 = ty == 2 ? 24 : ty * 16;
                    const uint32_t buf = a2 % C::NACC;
                    umma::mbar_wait(&bar_acc_empty[buf], ((a2 / C::NACC) & 1) ^ 1);
                    umma::fence_after_sync();
                    const uint32_t d = tm + buf * C::N;
#pragma unroll
                    for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
                        for (int j = 0; j < 3; ++j) {
                            const uint32_t pos = (uint32_t)(ci * (C::P_BYTES / 16) + (2 * y0 + 2 * j) * C::PW + tx * 8);
                            const uint32_t wo = (uint32_t)(ci * (C::W_BYTES / 16));
                            umma::mma_bf16(d, a_base + pos, w_base + wo + (uint32_t)((j * 2 * C::N * 16) >> 4), idesc, (ci | j) != 0);
                            umma::mma_bf16(d, a_base + pos, w_base + wo + (uint32_t)(((3 + j) * 2 * C::N * 16) >> 4), idesc, 1);
                        }
                    umma::commit(&bar_acc_full[buf]);
                }
            }
        } else if (warp >= 2) {
            const int es = (warp - 2) >> 2, quarter = warp & 3;
            const int r = quarter * 32 + lane, trow = r >> 3, tx8 = r & 7;
            for (int t = es; t < C::TILES; t += 2) {
                const uint32_t a2 = ai + t, buf = a2 % C::NACC;
                const int ty = t / C::TILES_X, tx = t % C::TILES_X;
                const int y0 = ty == 2 ? 24 : ty * 16, ymin = ty == 2 ? 32 : y0;
                umma::mbar_wait(&bar_acc_full[buf], (a2 / C::NACC) & 1);
                umma::fence_after_sync();
                uint32_t v[4][16];
                const uint32_t ta = tm + ((uint32_t)(quarter * 32) << 16) + buf * C::N;
#pragma unroll
                for (int q = 0; q < 4; ++q) umma::tmem_ld16(ta + 16 * q, v[q]);
                umma::tmem_ld_wait();
                umma::fence_before_sync();
                __syncwarp();
                if (lane == 0) umma::mbar_arrive(&bar_acc_empty[buf]);
                const int y = y0 + trow, px = tx * 8 + tx8;
                if (y >= ymin && y < C::PW) {
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int c = 0; c < 16; c += 2) {
                        float m[2];
#pragma unroll
                        for (int u = 0; u < 2; ++u)
                            m[u] = fmaxf(fmaxf(fmaxf(__uint_as_float(v[0][c + u]), __uint_as_float(v[1][c + u])),
                                               fmaxf(__uint_as_float(v[2][c + u]), __uint_as_float(v[3][c + u]))) + s_sh[c + u], 0.f);
                        if (f16out == 2) {                     // "fp16c": fp16 planes + the e5m2 correction planes (L: 16 channels, H: 16 channels)
                            const __half a = __float2half_rn(m[0]), b = __float2half_rn(m[1]);
                            hi[c >> 1] = (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
                            const float fa = __half2float(a), fb = __half2float(b);
                            const uint32_t l2 = (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2((m[0] - fa) * FC_UP, (m[1] - fb) * FC_UP), __NV_SATFINITE, __NV_E5M2);
                            const uint32_t h2 = (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(fa * FC_DOWN, fb * FC_DOWN), __NV_SATFINITE, __NV_E5M2);
                            // lo[0..3] = the 16 L bytes, lo[4..7] = the 16 H bytes (two channels per step)
                            if ((c & 2) == 0) { lo[c >> 2] = l2; lo[4 + (c >> 2)] = h2; }
                            else { lo[c >> 2] |= l2 << 16; lo[4 + (c >> 2)] |= h2 << 16; }
                        } else if (f16out) {                   // "fp16" precision: conv2 reads one fp16 plane
                            hi[c >> 1] = (uint32_t)__half_as_ushort(__float2half_rn(m[0])) | ((uint32_t)__half_as_ushort(__float2half_rn(m[1])) << 16);
                            lo[c >> 1] = 0u;
                        } else {
                            __nv_bfloat16 h0, l0, h1, l1;
                            umma::split_bf16(m[0], h0, l0); umma::split_bf16(m[1], h1, l1);
                            hi[c >> 1] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                            lo[c >> 1] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
                        }
                    }
                    const int pos = (y + 2) * Conv2Cfg::WP + px + 2;
                    uint8_t *o = out + (size_t)n * Conv2Cfg::IMG_BYTES + (size_t)pos * 16;
                    constexpr size_t PLB = (size_t)Conv2Cfg::PL * 16;
                    *reinterpret_cast<uint4 *>(o + 0 * PLB) = make_uint4(hi[0], hi[1], hi[2], hi[3]);      // hi, channels 0-7
                    *reinterpret_cast<uint4 *>(o + 1 * PLB) = make_uint4(hi[4], hi[5], hi[6], hi[7]);      // hi, channels 8-15
                    if (f16out != 1) {                         // bf16 lo planes, or (fp16c) the L and H correction planes
                        *reinterpret_cast<uint4 *>(o + 2 * PLB) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                        *reinterpret_cast<uint4 *>(o + 3 * PLB) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
                    }
                }
            }
        }
        ai += C::TILES;
        umma::fence_before_sync();
        __syncthreads();                                              // all MMAs of this crop retired (last acc_full seen)
        umma::fence_after_sync();
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 1) umma::tmem_dealloc(tm, 512);
}


// ------------------------------------------------------------------------------------------------
// conv1, C_in = 1, software-pipelined over crops (the default for grey crops).  Same GEMM formulation as above, but
// the three stages of a crop run in different warps and overlap across crops:
//   * 4 producer warps: crop n+1 from global memory -> padded u8 image -> decimated bf16 plane (double buffered),
//   * 1 MMA warp: 15 tiles x 6 MMAs of crop n into a ring of 8 TMEM accumulators; a tcgen05.commit after the last
//     tile of a crop frees its plane,
//   * 8 epilogue warps: tiles of crop n-1 / n out of TMEM: max-pool (thread-local), shift, ReLU, hi/lo split, store.
// No block-wide barrier inside the crop loop; everything is mbarrier hand-offs.
// ------------------------------------------------------------------------------------------------
struct Conv1P {
    using C = Conv1T;
    static constexpr int PRODUCERS = 128, EPI_SETS = 4, THREADS = 64 + EPI_SETS * 128 + PRODUCERS;     // 16 epilogue warps: four tiles drain concurrently
    static constexpr int SMEM = 2 * C::P_BYTES + C::IMG_BYTES + C::W_BYTES + 64 + 128;
};

__global__ void __launch_bounds__(Conv1P::THREADS, 1)
conv1_tc_pipe_kernel(const uint8_t *__restrict__ img, int n_max, const uint32_t *__restrict__ n_dev, int base,
                     const uint8_t *__restrict__ wgt, const float *__restrict__ sh, uint8_t *__restrict__ out, int f16out)
{
    using C = Conv1T;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar_acc_full[C::NACC], bar_acc_empty[C::NACC], bar_plane_full[2], bar_plane_empty[2];
    __shared__ uint32_t s_tmem;
    uint8_t *s_img = smem + 2 * C::P_BYTES;
    uint8_t *s_w = s_img + C::IMG_BYTES;
    float *s_sh = reinterpret_cast<float *>(s_w + C::W_BYTES);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_act = n_dev ? min((int)*n_dev - base, n_max) : n_max;

    if (tid == 0) {
        for (int i = 0; i < C::NACC; ++i) { umma::mbar_init(&bar_acc_full[i], 1); umma::mbar_init(&bar_acc_empty[i], 4); }
        for (int i = 0; i < 2; ++i) { umma::mbar_init(&bar_plane_full[i], 1); umma::mbar_init(&bar_plane_empty[i], 1); }
        umma::fence_mbar_init();
    }
    if (warp == 1) umma::tmem_alloc(&s_tmem, 512);
    for (int i = tid; i < C::W_BYTES / 4; i += Conv1P::THREADS) reinterpret_cast<uint32_t *>(s_w)[i] = reinterpret_cast<const uint32_t *>(wgt)[i];
    if (tid < 16) s_sh[tid] = sh[tid];
    for (int i = tid; i < C::IMG_BYTES / 4; i += Conv1P::THREADS) reinterpret_cast<uint32_t *>(s_img)[i] = 0u;   // zero halo, kept
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tm = s_tmem;

    if (warp >= 2 + 4 * Conv1P::EPI_SETS) {                            // ---- producers ----
        const int pt = tid - (64 + 128 * Conv1P::EPI_SETS);
        const uint32_t rot = ((uint32_t)pt >> 1) & 3u;                 // store rotation of the plane build (items of a thread are 128 apart: same rot)
        static_assert(Conv1P::PRODUCERS % 8 == 0, "the rotation is per thread");
        int it = 0;
        C2S_DECL;
        // the crop of the NEXT iteration is already in flight (13 words per thread) while this one is decimated: the global latency is off the
        // producer's critical path
        constexpr int NW = (C::H * C::W / 4 + Conv1P::PRODUCERS - 1) / Conv1P::PRODUCERS;
        uint32_t v[NW];
        auto fetch = [&](int n) {
            const uint32_t *src = reinterpret_cast<const uint32_t *>(img + (size_t)n * C::H * C::W);
#pragma unroll
            for (int k = 0; k < NW; ++k) {
                const int i = pt + k * Conv1P::PRODUCERS;
                v[k] = i < C::H * C::W / 4 ? src[i] : 0u;
            }
        };
        if ((int)blockIdx.x < n_act) fetch(blockIdx.x);
        for (int n = blockIdx.x; n < n_act; n += gridDim.x, ++it) {
            const int b = it & 1;
            uint4 *s_p = reinterpret_cast<uint4 *>(smem + b * C::P_BYTES);
#pragma unroll
            for (int k = 0; k < NW; ++k) {
                const int i = pt + k * Conv1P::PRODUCERS;
                if (i < C::H * C::W / 4) {
                    const int y = i / (C::W / 4), x4 = i % (C::W / 4);
                    uint16_t *d16 = reinterpret_cast<uint16_t *>(s_img + (y + 2) * C::IMG_PITCH + 2 + x4 * 4);   // 2-byte aligned
                    d16[0] = (uint16_t)v[k]; d16[1] = (uint16_t)(v[k] >> 16);
                }
            }
            if (n + (int)gridDim.x < n_act) fetch(n + gridDim.x);
            asm volatile("bar.sync 1, 128;" ::: "memory");
            C2S_BEGIN;
            umma::mbar_wait_suspend(&bar_plane_empty[b], (((uint32_t)it >> 1) & 1u) ^ 1u);       // the MMAs that read this plane have retired
            C2S_END(c2s_a);
            for (int i = pt; i < C::PROWS * (C::PW / 4); i += Conv1P::PRODUCERS) {
                const int yy = i / (C::PW / 4), p4 = i % (C::PW / 4);
                const uint2 wa = *reinterpret_cast<const uint2 *>(s_img + yy * C::IMG_PITCH + p4 * 8);
                const uint2 wb2 = *reinterpret_cast<const uint2 *>(s_img + yy * C::IMG_PITCH + p4 * 8 + 8);
                const uint32_t ws4[4] = {wa.x, wa.y, wb2.x, wb2.y};
                uint32_t ev[7];
#pragma unroll
                for (int k = 0; k < 7; ++k) {                  // bf16 pair (byte 2k, byte 2k+1): high halves of the fp32 patterns
                    const uint32_t b0 = (ws4[(2 * k) >> 2] >> (8 * ((2 * k) & 3))) & 0xFFu, b1 = (ws4[(2 * k + 1) >> 2] >> (8 * ((2 * k + 1) & 3))) & 0xFFu;
                    ev[k] = __byte_perm(__float_as_uint((float)b0), __float_as_uint((float)b1), 0x7632);
                }
                // item i owns bytes 64 i ... 64 i + 63 of the plane: with every thread storing chunk j in the same instruction, a quarter-warp touches only
                // two of the eight 16-byte bank groups (4-way conflict: ncu showed 77 % of the store wavefronts as replays).  Thread pairs therefore start
                // at different chunks: instruction jj stores chunk (jj + rot) & 3, rot = (i / 2) & 3 -- eight distinct groups per quarter-warp
                uint4 ch[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) ch[j] = make_uint4(ev[j], ev[j + 1], ev[j + 2], ev[j + 3]);
                if (rot & 1u) { const uint4 t0 = ch[0]; ch[0] = ch[1]; ch[1] = ch[2]; ch[2] = ch[3]; ch[3] = t0; }
                if (rot & 2u) { const uint4 t0 = ch[0], t1 = ch[1]; ch[0] = ch[2]; ch[1] = ch[3]; ch[2] = t0; ch[3] = t1; }
                uint4 *dst = s_p + yy * C::PW + p4 * 4;
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) dst[(jj + rot) & 3u] = ch[jj];
            }
            umma::fence_async_smem();
            asm volatile("bar.sync 1, 128;" ::: "memory");             // plane complete; s_img free for the next crop
            if (pt == 0) umma::mbar_arrive(&bar_plane_full[b]);
        }
#ifdef TB_CONV2_STATS
        if (pt == 0) { C2S_FLUSH(11, 12, 14, 14); }
#endif
    } else if (warp == 1) {                                            // ---- MMA issue ----
        if (umma::elect_one()) {
            const uint32_t idesc = umma::idesc_bf16_f32(128, C::N);
            const uint64_t w_base = umma::smem_desc(umma::smem_u32(s_w), C::N * 16, 128);
            uint32_t a2 = 0;
            int it = 0;
            C2S_DECL;
            for (int n = blockIdx.x; n < n_act; n += gridDim.x, ++it) {
                const int b = it & 1;
                const uint64_t a_base = umma::smem_desc(umma::smem_u32(smem + b * C::P_BYTES), C::PW * 16, 2 * C::PW * 16);
                C2S_BEGIN;
                umma::mbar_wait(&bar_plane_full[b], ((uint32_t)it >> 1) & 1u);
                C2S_END(c2s_a);
                umma::fence_after_sync();
                for (int t = 0; t < C::TILES; ++t, ++a2) {
                    const int ty = t / C::TILES_X, tx = t % C::TILES_X;
                    const int y0 = ty == 2 ? 24 : ty * 16;
                    const uint32_t buf = a2 % C::NACC;
                    C2S_BEGIN;
                    umma::mbar_wait(&bar_acc_empty[buf], ((a2 / C::NACC) & 1) ^ 1);
                    C2S_END(c2s_c);
                    umma::fence_after_sync();
                    const uint32_t d = tm + buf * C::N;
                    const uint64_t a_tile = umma::desc_add(a_base, (uint32_t)(2 * y0 * C::PW + tx * 8));
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const uint64_t ad = umma::desc_add(a_tile, (uint32_t)(2 * j * C::PW));
                        umma::mma_bf16(d, ad, umma::desc_add(w_base, (uint32_t)((j * 2 * C::N * 16) >> 4)), idesc, j != 0);
                        umma::mma_bf16(d, ad, umma::desc_add(w_base, (uint32_t)(((3 + j) * 2 * C::N * 16) >> 4)), idesc, 1);
                    }
                    umma::commit(&bar_acc_full[buf]);
                }
                umma::commit(&bar_plane_empty[b]);
            }
            C2S_FLUSH(0, 1, 2, 3);
        }
    } else if (warp >= 2) {                                            // ---- epilogue ----
        const int es = (warp - 2) >> 2, quarter = warp & 3;
        const int r = quarter * 32 + lane, trow = r >> 3, tx8 = r & 7;
        uint32_t ai = 0;
        C2S_DECL;
        for (int n = blockIdx.x; n < n_act; n += gridDim.x, ai += C::TILES) {
            for (int t = es; t < C::TILES; t += Conv1P::EPI_SETS) {
                const uint32_t a2 = ai + t, buf = a2 % C::NACC;
                const int ty = t / C::TILES_X, tx = t % C::TILES_X;
                const int y0 = ty == 2 ? 24 : ty * 16, ymin = ty == 2 ? 32 : y0;
                C2S_BEGIN;
                umma::mbar_wait_suspend(&bar_acc_full[buf], (a2 / C::NACC) & 1);
                C2S_END(c2s_a);
                umma::fence_after_sync();
                uint32_t v[4][16];
                const uint32_t ta = tm + ((uint32_t)(quarter * 32) << 16) + buf * C::N;
#pragma unroll
                for (int q = 0; q < 4; ++q) umma::tmem_ld16(ta + 16 * q, v[q]);
                umma::tmem_ld_wait();
                umma::fence_before_sync();
                __syncwarp();
                if (lane == 0) umma::mbar_arrive(&bar_acc_empty[buf]);
                const int y = y0 + trow, px = tx * 8 + tx8;
                if (y >= ymin && y < C::PW) {
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int c = 0; c < 16; c += 2) {
                        float m[2];
#pragma unroll
                        for (int u = 0; u < 2; ++u)
                            m[u] = fmaxf(fmaxf(fmaxf(__uint_as_float(v[0][c + u]), __uint_as_float(v[1][c + u])),
                                               fmaxf(__uint_as_float(v[2][c + u]), __uint_as_float(v[3][c + u]))) + s_sh[c + u], 0.f);
                        if (f16out == 2) {                     // "fp16c": fp16 planes + the e5m2 correction planes (L: 16 channels, H: 16 channels)
                            const __half a = __float2half_rn(m[0]), b = __float2half_rn(m[1]);
                            hi[c >> 1] = (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
                            const float fa = __half2float(a), fb = __half2float(b);
                            const uint32_t l2 = (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2((m[0] - fa) * FC_UP, (m[1] - fb) * FC_UP), __NV_SATFINITE, __NV_E5M2);
                            const uint32_t h2 = (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(fa * FC_DOWN, fb * FC_DOWN), __NV_SATFINITE, __NV_E5M2);
                            // lo[0..3] = the 16 L bytes, lo[4..7] = the 16 H bytes (two channels per step)
                            if ((c & 2) == 0) { lo[c >> 2] = l2; lo[4 + (c >> 2)] = h2; }
                            else { lo[c >> 2] |= l2 << 16; lo[4 + (c >> 2)] |= h2 << 16; }
                        } else if (f16out) {                   // "fp16" precision: conv2 reads one fp16 plane
                            hi[c >> 1] = (uint32_t)__half_as_ushort(__float2half_rn(m[0])) | ((uint32_t)__half_as_ushort(__float2half_rn(m[1])) << 16);
                            lo[c >> 1] = 0u;
                        } else {
                            __nv_bfloat16 h0, l0, h1, l1;
                            umma::split_bf16(m[0], h0, l0); umma::split_bf16(m[1], h1, l1);
                            hi[c >> 1] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                            lo[c >> 1] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
                        }
                    }
                    const int pos = (y + 2) * Conv2Cfg::WP + px + 2;
                    uint8_t *o = out + (size_t)n * Conv2Cfg::IMG_BYTES + (size_t)pos * 16;
                    constexpr size_t PLB = (size_t)Conv2Cfg::PL * 16;
                    *reinterpret_cast<uint4 *>(o + 0 * PLB) = make_uint4(hi[0], hi[1], hi[2], hi[3]);      // hi, channels 0-7
                    *reinterpret_cast<uint4 *>(o + 1 * PLB) = make_uint4(hi[4], hi[5], hi[6], hi[7]);      // hi, channels 8-15
                    if (f16out != 1) {                         // bf16 lo planes, or (fp16c) the L and H correction planes
                        *reinterpret_cast<uint4 *>(o + 2 * PLB) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                        *reinterpret_cast<uint4 *>(o + 3 * PLB) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
                    }
                }
            }
        }
#ifdef TB_CONV2_STATS
        if (warp == 2 && lane == 0) { C2S_FLUSH(4, 5, 6, 7); }
#endif
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 1) umma::tmem_dealloc(tm, 512);
}

}}  // namespace tb::tc
