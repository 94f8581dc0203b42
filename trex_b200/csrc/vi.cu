// Visual identification inference (V118_3) on sm_100a behind the tb_vi_* C ABI.
// Replaces Python::VINetwork::probabilities (T/ml/VisualIdentification.h:104-133, .cpp:440-494)
// -> predict_numpy (T/python/visual_recognition_torch.py:290-352) -> V118_3.forward
// (T/python/visual_identification_network_torch.py:216-258), eval mode:
//   conv5x5(C->16)+BN+ReLU+pool2 -> conv5x5(16->64)+BN+ReLU+pool2 -> conv5x5(64->128)+BN+ReLU+pool2
//   -> flatten (NCHW order) -> fc 12800->100 -> LayerNorm(100) -> ReLU -> fc 100->M -> softmax.
// Inputs are raw u8 crops cast to float (no /255, visual_recognition_torch.py:337).
// precision 0: fp32 CUDA-core kernels below.  precision 1: tcgen05 tensor-core kernels of vi_tc.cuh
// (bf16 hi/lo split, three MMAs per k-step, fp32 accumulation in TMEM) for conv2, conv3 and fc1.
#include "common.h"
#include "vi_tc.cuh"
#include "vi_nets.h"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace tb {

constexpr float BN_EPS = 1e-5f, LN_EPS = 1e-5f;

// ------------------------------------------------------------------------------------------------
// conv1: u8 crop [H][W][C] (C = 1 gray, 3 rgb8: NHWC as the reference hands it over) -> pooled [H/2][W/2][16] fp32.
// One CTA per image; each thread owns pooled pixels (a 2x2 window of conv outputs) x 16 output channels.
// ------------------------------------------------------------------------------------------------
constexpr int C1_NT = 256;

__global__ void __launch_bounds__(C1_NT)
conv1_kernel(const uint8_t *__restrict__ img, int H, int W, int C, int n_max, const uint32_t *__restrict__ n_dev, int base,
             const float *__restrict__ w /*[25][C][16]*/, const float *__restrict__ sc, const float *__restrict__ sh,
             float *__restrict__ out)
{
    extern __shared__ float sm[];
    const int n = blockIdx.x;
    const int n_act = n_dev ? min((int)*n_dev - base, n_max) : n_max;
    if (n >= n_act) return;
    const int PW = W + 4, PH = H + 4;
    float *patch = sm;                       // [C][PH][PW], zero halo
    float *sw = sm + C * PH * PW;            // [25][C][16]
    float *ssc = sw + 400 * C, *ssh = ssc + 16;
    const uint8_t *src = img + (size_t)n * H * W * C;
    for (int i = threadIdx.x; i < C * PH * PW; i += C1_NT) {
        int ci = i / (PH * PW), y = (i % (PH * PW)) / PW - 2, x = i % PW - 2;
        patch[i] = (y >= 0 && y < H && x >= 0 && x < W) ? (float)src[(y * W + x) * C + ci] : 0.f;
    }
    for (int i = threadIdx.x; i < 400 * C; i += C1_NT) sw[i] = w[i];
    if (threadIdx.x < 16) { ssc[threadIdx.x] = sc[threadIdx.x]; ssh[threadIdx.x] = sh[threadIdx.x]; }
    __syncthreads();
    const int OW = W / 2, OH = H / 2;
    float *dst = out + (size_t)n * OH * OW * 16;
    for (int q = threadIdx.x; q < OH * OW; q += C1_NT) {
        const int py = q / OW, px = q % OW;
        float acc[4][16];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 16; ++c) acc[a][c] = 0.f;
        for (int ci = 0; ci < C; ++ci)
        for (int dy = 0; dy < 5; ++dy) {
            float r0[6], r1[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                r0[i] = patch[(ci * PH + 2 * py + dy) * PW + 2 * px + i];
                r1[i] = patch[(ci * PH + 2 * py + dy + 1) * PW + 2 * px + i];
            }
#pragma unroll
            for (int dx = 0; dx < 5; ++dx) {
                const float4 *wp = reinterpret_cast<const float4 *>(sw + ((dy * 5 + dx) * C + ci) * 16);
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    const float4 wv = wp[c4];
                    const float ww[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        acc[0][c4 * 4 + j] = fmaf(r0[dx], ww[j], acc[0][c4 * 4 + j]);
                        acc[1][c4 * 4 + j] = fmaf(r0[dx + 1], ww[j], acc[1][c4 * 4 + j]);
                        acc[2][c4 * 4 + j] = fmaf(r1[dx], ww[j], acc[2][c4 * 4 + j]);
                        acc[3][c4 * 4 + j] = fmaf(r1[dx + 1], ww[j], acc[3][c4 * 4 + j]);
                    }
                }
            }
        }
        float o[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            float m = fmaxf(fmaxf(fmaf(acc[0][c], ssc[c], ssh[c]), fmaf(acc[1][c], ssc[c], ssh[c])),
                            fmaxf(fmaf(acc[2][c], ssc[c], ssh[c]), fmaf(acc[3][c], ssc[c], ssh[c])));
            o[c] = fmaxf(m, 0.f);
        }
        float4 *d4 = reinterpret_cast<float4 *>(dst + (size_t)q * 16);
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) d4[c4] = make_float4(o[c4 * 4], o[c4 * 4 + 1], o[c4 * 4 + 2], o[c4 * 4 + 3]);
    }
}

// ------------------------------------------------------------------------------------------------
// conv2 / conv3: NHWC fp32 [HW][HW][CIN] -> pooled [HW/2][HW/2][COUT].  A CTA computes a TWP x THP
// tile of pooled pixels (<= 64 windows of 2x2 conv outputs) x 64 output channels of one image;
// thread = (window, 16 couts).  Input patch (zero halo) and weights are staged through shared
// memory in chunks of 8 input channels.
// ------------------------------------------------------------------------------------------------
constexpr int CV_NT = 256, CV_CK = 8;

template <int TWP, int THP>
struct ConvTile {
    static constexpr int PR = 2 * THP + 4, PC = 2 * TWP + 4, PITCH = PC + 1;
    static constexpr int SMEM = (CV_CK * PR * PITCH + 25 * CV_CK * 64) * 4;
};

template <int CIN, int COUT, int HW, int TWP, int THP>
__global__ void __launch_bounds__(CV_NT)
conv_kernel(const float *__restrict__ in, int n_max, const uint32_t *__restrict__ n_dev, int base,
            const float *__restrict__ w /*[25][CIN][COUT]*/, const float *__restrict__ sc, const float *__restrict__ sh,
            float *__restrict__ out)
{
    using T = ConvTile<TWP, THP>;
    static_assert(TWP * THP <= 64, "tile has at most 64 windows");
    extern __shared__ float sm[];
    float *patch = sm;                                  // [CV_CK][PR][PITCH]
    float *sw = sm + CV_CK * T::PR * T::PITCH;          // [25][CV_CK][64]
    constexpr int OHW = HW / 2;
    constexpr int TX = (OHW + TWP - 1) / TWP, TY = (OHW + THP - 1) / THP;
    const int n = blockIdx.x / (TX * TY);
    const int n_act = n_dev ? min((int)*n_dev - base, n_max) : n_max;
    if (n >= n_act) return;
    const int tile = blockIdx.x % (TX * TY);
    const int ty = tile / TX, tx = tile % TX;
    const int co0 = blockIdx.y * 64;
    const int tid = threadIdx.x, win = tid & 63, cg = tid >> 6;
    const bool active = win < TWP * THP;
    const int wy = active ? win / TWP : 0, wx = active ? win % TWP : 0;
    const int y0 = ty * THP * 2 - 2, x0 = tx * TWP * 2 - 2;   // patch origin in the image
    const float *src = in + (size_t)n * HW * HW * CIN;

    float acc[4][16];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 16; ++c) acc[a][c] = 0.f;

    for (int c0 = 0; c0 < CIN; c0 += CV_CK) {
        __syncthreads();
        for (int i = tid; i < T::PR * T::PC * CV_CK; i += CV_NT) {
            const int c = i % CV_CK, p = i / CV_CK, px = p % T::PC, py = p / T::PC;
            const int y = y0 + py, x = x0 + px;
            float v = 0.f;
            if (y >= 0 && y < HW && x >= 0 && x < HW) v = src[((size_t)y * HW + x) * CIN + c0 + c];
            patch[(c * T::PR + py) * T::PITCH + px] = v;
        }
        for (int i = tid; i < 25 * CV_CK * 64; i += CV_NT) {
            const int co = i & 63, c = (i >> 6) % CV_CK, t = i / (64 * CV_CK);
            sw[i] = w[((size_t)t * CIN + c0 + c) * COUT + co0 + co];
        }
        __syncthreads();
#pragma unroll 1
        for (int c = 0; c < CV_CK; ++c) {
            const float *pc = patch + (c * T::PR + 2 * wy) * T::PITCH + 2 * wx;
#pragma unroll 1
            for (int dy = 0; dy < 5; ++dy) {
                float r0[6], r1[6];
#pragma unroll
                for (int i = 0; i < 6; ++i) { r0[i] = pc[dy * T::PITCH + i]; r1[i] = pc[(dy + 1) * T::PITCH + i]; }
#pragma unroll
                for (int dx = 0; dx < 5; ++dx) {
                    const float4 *wp = reinterpret_cast<const float4 *>(sw + ((dy * 5 + dx) * CV_CK + c) * 64 + cg * 16);
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        const float4 wv = wp[c4];
                        const float ww[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            acc[0][c4 * 4 + j] = fmaf(r0[dx], ww[j], acc[0][c4 * 4 + j]);
                            acc[1][c4 * 4 + j] = fmaf(r0[dx + 1], ww[j], acc[1][c4 * 4 + j]);
                            acc[2][c4 * 4 + j] = fmaf(r1[dx], ww[j], acc[2][c4 * 4 + j]);
                            acc[3][c4 * 4 + j] = fmaf(r1[dx + 1], ww[j], acc[3][c4 * 4 + j]);
                        }
                    }
                }
            }
        }
    }
    const int py = ty * THP + wy, px = tx * TWP + wx;
    if (active && py < OHW && px < OHW) {
        float o[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const float s = sc[co0 + cg * 16 + c], t = sh[co0 + cg * 16 + c];
            float m = fmaxf(fmaxf(fmaf(acc[0][c], s, t), fmaf(acc[1][c], s, t)), fmaxf(fmaf(acc[2][c], s, t), fmaf(acc[3][c], s, t)));
            o[c] = fmaxf(m, 0.f);
        }
        float4 *d4 = reinterpret_cast<float4 *>(out + (((size_t)n * OHW + py) * OHW + px) * COUT + co0 + cg * 16);
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) d4[c4] = make_float4(o[c4 * 4], o[c4 * 4 + 1], o[c4 * 4 + 2], o[c4 * 4 + 3]);
    }
}

// ------------------------------------------------------------------------------------------------
// fc1: [n][K] x Wt[K][100] + b -> [n][100].  CTA = 32 images x 104 outputs, K staged in chunks of 32.
// ------------------------------------------------------------------------------------------------
constexpr int FC_NT = 256, FC_IMG = 32, FC_KC = 32, FC_OUT = 100, FC_OPAD = 104;

__global__ void __launch_bounds__(FC_NT)
fc1_kernel(const float *__restrict__ x, int K, int n_max, const uint32_t *__restrict__ n_dev, int base,
           const float *__restrict__ wt /*[K][100]*/, const float *__restrict__ b, float *__restrict__ out)
{
    __shared__ float sx[FC_IMG][FC_KC + 1];
    __shared__ float swt[FC_KC][FC_OPAD];
    const int n_act = n_dev ? min((int)*n_dev - base, n_max) : n_max;
    const int i0 = blockIdx.x * FC_IMG;
    if (i0 >= n_act) return;
    const int tid = threadIdx.x, img = tid >> 3, og = tid & 7;
    float acc[13];
#pragma unroll
    for (int j = 0; j < 13; ++j) acc[j] = 0.f;
    for (int k0 = 0; k0 < K; k0 += FC_KC) {
        __syncthreads();
        for (int i = tid; i < FC_IMG * FC_KC; i += FC_NT) {
            const int r = i / FC_KC, c = i % FC_KC;
            sx[r][c] = (i0 + r < n_act) ? x[(size_t)(i0 + r) * K + k0 + c] : 0.f;
        }
        for (int i = tid; i < FC_KC * FC_OPAD; i += FC_NT) {
            const int r = i / FC_OPAD, c = i % FC_OPAD;
            swt[r][c] = c < FC_OUT ? wt[(size_t)(k0 + r) * FC_OUT + c] : 0.f;
        }
        __syncthreads();
#pragma unroll 4
        for (int k = 0; k < FC_KC; ++k) {
            const float xv = sx[img][k];
#pragma unroll
            for (int j = 0; j < 13; ++j) acc[j] = fmaf(xv, swt[k][og + 8 * j], acc[j]);
        }
    }
    if (i0 + img < n_act)
#pragma unroll
        for (int j = 0; j < 13; ++j) {
            const int o = og + 8 * j;
            if (o < FC_OUT) out[(size_t)(i0 + img) * FC_OUT + o] = acc[j] + b[o];
        }
}

// ------------------------------------------------------------------------------------------------
// head: (sum of fc1 partials + bias) -> LayerNorm(100) -> ReLU -> fc2 (100 -> M) -> softmax.
// One warp per image, 8 warps per CTA; fc2 weights staged in shared memory when they fit.
// ------------------------------------------------------------------------------------------------
constexpr int HD_WARPS = 8;

__global__ void __launch_bounds__(HD_WARPS * 32)
head_kernel(const float *__restrict__ h1, int nsplit, size_t split_stride, const float *__restrict__ b1,
            int M, int n_max, const uint32_t *__restrict__ n_dev, int base,
            const float *__restrict__ g, const float *__restrict__ be, const float *__restrict__ w2t /*[100][M]*/,
            const float *__restrict__ b2, float *__restrict__ probs, float *__restrict__ logits, int w_in_smem,
            uint32_t *__restrict__ top_id, float *__restrict__ top_p, int norm_mode)
{
    extern __shared__ float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_act = n_dev ? min((int)*n_dev - base, n_max) : n_max;
    if ((int)(blockIdx.x * HD_WARPS) >= n_act) return;
    float *sw = sm + HD_WARPS * (100 + M);
    if (w_in_smem) {
        for (int i = threadIdx.x; i < 100 * M; i += HD_WARPS * 32) sw[i] = w2t[i];
        __syncthreads();
    }
    const float *W = w_in_smem ? sw : w2t;
    const int n = blockIdx.x * HD_WARPS + warp;
    if (n >= n_act) return;
    float *sh = sm + warp * (100 + M), *sl = sh + 100;
    float v[4], s = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int i = lane + 32 * j;
        float a = 0.f;
        if (i < 100) {
            a = b1 ? b1[i] : 0.f;
            for (int q = 0; q < nsplit; ++q) a += h1[(size_t)q * split_stride + (size_t)n * 100 + i];
        }
        v[j] = a; s += a;
    }
    if (norm_mode) {           // V100 / V110 on the tensor path: no LayerNorm; g, be = BatchNorm1d folded to scale / shift (or 1, 0)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int i = lane + 32 * j;
            if (i < 100) sh[i] = fmaxf(fmaf(v[j], g[i], be[i]), 0.f);
        }
    } else {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / 100.f;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) { const int i = lane + 32 * j; const float dlt = i < 100 ? v[j] - mean : 0.f; q += dlt * dlt; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / 100.f + LN_EPS);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int i = lane + 32 * j;
        if (i < 100) sh[i] = fmaxf((v[j] - mean) * rstd * g[i] + be[i], 0.f);
    }
    }
    __syncwarp();
    float mx = -INFINITY; int arg = 0;
    for (int o = lane; o < M; o += 32) {
        float a = b2[o];
#pragma unroll 4
        for (int k = 0; k < 100; ++k) a = fmaf(sh[k], W[(size_t)k * M + o], a);
        sl[o] = a;
        if (logits) logits[(size_t)n * M + o] = a;
        if (a > mx) { mx = a; arg = o; }                    // first maximum per lane (o ascending)
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {                       // warp arg-max, ties -> smallest class index
        const float om = __shfl_xor_sync(0xffffffffu, mx, o);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
        if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
    }
    float sum = 0.f;
    for (int o = lane; o < M; o += 32) { const float e = expf(sl[o] - mx); sl[o] = e; sum += e; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.f / sum;
    for (int o = lane; o < M; o += 32) probs[(size_t)n * M + o] = sl[o] * inv;
    if (top_id && lane == 0) { top_id[n] = (uint32_t)arg; top_p[n] = inv; }     // exp(max - max) = 1
}

}  // namespace tb

// ================================================================================================
// host side: the tb_vi handle
// ================================================================================================
using namespace tb;

struct tb_vi {
    tb_vi_config cfg{};
    std::map<std::string, std::vector<float>> sd;
    bool committed = false;
    cudaStream_t stream = nullptr;
    int chunk = 0;
    std::vector<void *> dev_allocs;
    // fp32 parameters
    float *w1 = nullptr, *s1 = nullptr, *t1 = nullptr, *w2 = nullptr, *s2 = nullptr, *t2 = nullptr;
    float *w3 = nullptr, *s3 = nullptr, *t3 = nullptr, *wf1 = nullptr, *bf1 = nullptr, *lng = nullptr, *lnb = nullptr;
    float *wf2 = nullptr, *bf2 = nullptr;
    // activations (one chunk)
    float *a1 = nullptr, *a2 = nullptr, *a3 = nullptr, *h1 = nullptr;
    uint8_t *d_img = nullptr; float *d_probs = nullptr, *d_logits = nullptr;
    // tensor-core path (precision 1)
    uint8_t *in2 = nullptr, *in3 = nullptr, *fca = nullptr, *w1t = nullptr, *w2t = nullptr, *w2p = nullptr, *w3t = nullptr, *wfc = nullptr;
    int fc_groups = 0, n_sms = 148, head_w_smem = 0;
    uint32_t *top_id = nullptr; float *top_p = nullptr;     // optional device outputs: arg-max class and its probability per image
    uint64_t launches = 0;
    cudaStream_t last_stream = nullptr;
    EventRing<5> prof;
    ViNet *net = nullptr;                                    // arch != 0: V100 / V110 / V119 / V200 (vi_nets.cu)
};

static int head_smem(const tb_vi *h, int M) { return (HD_WARPS * (100 + M) + (h->head_w_smem ? 100 * M : 0)) * 4; }

template <typename T>
static int vi_dev(tb_vi *h, T **p, size_t n)
{
    int r = dev_alloc(p, std::max<size_t>(n, 1));
    if (r == TB_OK) h->dev_allocs.push_back((void *)*p);
    return r;
}

extern "C" int tb_vi_create(const tb_vi_config *cfg, tb_vi **out)
{
    TB_REQUIRE(cfg && out, TB_ERR_INVALID, "tb_vi_create: null argument");
    TB_REQUIRE(cfg->width == 80 && cfg->height == 80 && (cfg->channels == 1 || cfg->channels == 3), TB_ERR_INVALID,
               "tb_vi_create: this release builds V118_3 for 80x80 crops with 1 (meta_encoding gray) or 3 (rgb8) channels");
    TB_REQUIRE(cfg->num_classes > 0 && cfg->num_classes <= 1024, TB_ERR_INVALID, "tb_vi_create: num_classes must be 1..1024");
    TB_REQUIRE(cfg->max_images > 0, TB_ERR_INVALID, "tb_vi_create: max_images must be > 0");
    TB_REQUIRE(cfg->precision >= 0 && cfg->precision <= 3, TB_ERR_INVALID, "tb_vi_create: precision must be 0 (fp32), 1 (bf16x3), 2 (fp16) or 3 (fp16c: fp16 + e5m2 correction terms) on tensor cores");
    TB_REQUIRE(cfg->arch >= 0 && cfg->arch <= 4, TB_ERR_INVALID, "tb_vi_create: arch must be 0 (v118_3), 1 (v100), 2 (v110), 3 (v119) or 4 (v200)");
    TB_REQUIRE(cfg->arch <= 2 || cfg->precision == 0, TB_ERR_INVALID, "tb_vi_create: v119 / v200 run in fp32 (precision 0); the tensor-core precisions are built for v118_3, v100 and v110");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("tb_vi_create: no CUDA device (there is no CPU fallback)"); return TB_ERR_CUDA; }
    TB_REQUIRE(cfg->device >= 0 && cfg->device < ndev, TB_ERR_INVALID, "tb_vi_create: bad device ordinal");
    TB_CUDA(cudaSetDevice(cfg->device));
    tb_vi *h = new tb_vi();
    h->cfg = *cfg;
    // images per kernel sequence (measured: 4096 beats 16384 by ~4 % on conv3 although the round-robin tail is longer)
    static const int chunk_env = getenv("TB_VI_CHUNK") ? atoi(getenv("TB_VI_CHUNK")) : 0;
    h->chunk = std::min(cfg->max_images, chunk_env > 0 ? chunk_env : 4096);
    const size_t CH = h->chunk, M = cfg->num_classes, N = cfg->max_images, CI = cfg->channels;
    int r = TB_OK;
#define A(p, n) if (r == TB_OK) r = vi_dev(h, &(p), (n))
    if (cfg->arch != 0 && cfg->precision == 0) {        // fp32 layer-list executor; v100 / v110 with a tensor precision share v118_3's kernels below
        r = vinet_create(&h->net, cfg->arch, cfg->channels, cfg->num_classes, cfg->max_images, h->dev_allocs);
        A(h->d_img, N * 6400 * CI + 16); A(h->d_probs, N * M); A(h->d_logits, N * M);
        if (r == TB_OK && cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("cudaStreamCreate failed"); r = TB_ERR_CUDA; }
        if (r != TB_OK) { tb_vi_destroy(h); return r; }
        *out = h;
        return TB_OK;
    }
    A(h->w1, 400 * CI); A(h->s1, 16); A(h->t1, 16);
    A(h->w2, 25 * 16 * 64); A(h->s2, 64); A(h->t2, 64);
    A(h->w3, 25 * 64 * 128); A(h->s3, 128); A(h->t3, 128);
    A(h->wf1, 12800 * 100); A(h->bf1, 100); A(h->lng, 100); A(h->lnb, 100);
    A(h->wf2, 100 * M); A(h->bf2, M);
    if (cfg->precision == 0) { A(h->a1, CH * 40 * 40 * 16); A(h->a2, CH * 20 * 20 * 64); A(h->a3, CH * 10 * 10 * 128); }
    A(h->h1, CH * 100 * (cfg->precision >= 1 ? tc::FC_SPLIT : 1));
    if (cfg->precision >= 1) {
        h->fc_groups = (int)((CH + 127) / 128 * 16);
        A(h->in2, CH * tc::Conv2Cfg::IMG_BYTES + 256); A(h->in3, CH * tc::Conv3Cfg::IMG_BYTES + 256);
        A(h->fca, (size_t)2 * h->fc_groups * tc::FC_KC * 128);
        A(h->w1t, (size_t)tc::Conv1T::W_BYTES * CI);
        A(h->w2t, (size_t)tc::Conv2D::W_BYTES); A(h->w2p, (size_t)tc::Conv2P::W_BYTES); A(h->w3t, (size_t)25 * tc::Conv3Cfg::WTAP_BYTES);
        A(h->wfc, (size_t)2 * tc::FC_KC * tc::FC_N * 16);
        if (r == TB_OK) {   // halo positions are never written by the kernels: zero once
            if (cudaMemset(h->in2, 0, CH * tc::Conv2Cfg::IMG_BYTES + 256) != cudaSuccess || cudaMemset(h->in3, 0, CH * tc::Conv3Cfg::IMG_BYTES + 256) != cudaSuccess ||
                cudaMemset(h->fca, 0, (size_t)2 * h->fc_groups * tc::FC_KC * 128) != cudaSuccess) { set_error("cudaMemset failed"); r = TB_ERR_CUDA; }
        }
        int dev_sms = 0;
        if (cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, cfg->device) == cudaSuccess && dev_sms > 0) h->n_sms = dev_sms;
    }
    A(h->d_img, N * 6400 * CI + 16); A(h->d_probs, N * M); A(h->d_logits, N * M);
#undef A
    if (r == TB_OK && cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("cudaStreamCreate failed"); r = TB_ERR_CUDA; }
    if (r == TB_OK) {
        const int with_w = (HD_WARPS * (100 + (int)M) + 100 * (int)M) * 4;
        h->head_w_smem = with_w <= 200 * 1024;
        if (cudaFuncSetAttribute(head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, std::max(head_smem(h, (int)M), 48 * 1024)) != cudaSuccess) { set_error("head_kernel smem attribute failed"); r = TB_ERR_CUDA; }
    }
    if (r != TB_OK) { tb_vi_destroy(h); return r; }
    *out = h;
    return TB_OK;
}

extern "C" void tb_vi_destroy(tb_vi *h)
{
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (void *p : h->dev_allocs) cudaFree(p);
    h->prof.destroy();
    if (h->stream) cudaStreamDestroy(h->stream);
    vinet_destroy(h->net);
    delete h;
}

extern "C" int tb_vi_set_tensor(tb_vi *h, const char *name, const float *data, int64_t count)
{
    TB_REQUIRE(h && name && data && count > 0, TB_ERR_INVALID, "tb_vi_set_tensor: null/empty argument");
    h->sd[name].assign(data, data + count);
    h->committed = false;
    return TB_OK;
}

static int vi_need(tb_vi *h, const char *name, size_t count, const std::vector<float> **out)
{
    auto it = h->sd.find(name);
    if (it == h->sd.end()) { set_error(std::string("tb_vi_commit: missing tensor ") + name); return TB_ERR_STATE; }
    if (it->second.size() != count) {
        set_error(std::string("tb_vi_commit: tensor ") + name + " has " + std::to_string(it->second.size()) + " elements, expected " + std::to_string(count));
        return TB_ERR_INVALID;
    }
    *out = &it->second;
    return TB_OK;
}

static int vi_upload(float *dst, const std::vector<float> &v)
{
    TB_CUDA(cudaMemcpy(dst, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
    return TB_OK;
}

static inline uint16_t f2bf_host(float f)
{
    uint32_t u; std::memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fc0;
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
static inline void split_bf16_host(float x, uint16_t &hi, uint16_t &lo)
{
    hi = f2bf_host(x);
    uint32_t u = (uint32_t)hi << 16; float fh; std::memcpy(&fh, &u, 4);
    lo = f2bf_host(x - fh);
}
// operand encoding of the conv2 / conv3 weights: bf16 hi + lo ("bf16x3"), or one fp16 value in the hi slot ("fp16", "fp16c")
static inline void split_operand_host(float x, bool f16, uint16_t &hi, uint16_t &lo)
{
    if (f16) { hi = __half_as_ushort(__float2half_rn(x)); lo = 0; }
    else split_bf16_host(x, hi, lo);
}
// float -> e5m2 byte (round to nearest even through fp16's top byte; saturates to the largest finite value)
static inline uint8_t f2e5m2_host(float x)
{
    const uint32_t h = __half_as_ushort(__float2half_rn(x));
    uint32_t r = (h + 0x7Fu + ((h >> 8) & 1u)) >> 8;
    if ((r & 0x7Fu) >= 0x7Cu) r = (r & 0x80u) | 0x7Bu;
    return (uint8_t)r;
}
// "fp16c": the lo slots of a 16-channel block (two channel-group planes of one row: 16 bytes each) hold the e5m2 correction operands that
// pair with the activations' [l * 2^8 | h * 2^-8]: plane 0 = fp16(w) * 2^-8, plane 1 = (w - fp16(w)) * 2^8, 16 channels each
static inline void fp16c_weight_bytes(const float *w16 /* 16 channels */, uint8_t *plane0, uint8_t *plane1)
{
    for (int c = 0; c < 16; ++c) {
        const float h = __half2float(__float2half_rn(w16[c]));
        plane0[c] = f2e5m2_host(h * (1.f / 256.f));
        plane1[c] = f2e5m2_host((w16[c] - h) * 256.f);
    }
}
// tensor path: conv weight torch [Cout][Cin][25] -> B operand [tap][hi|lo][cin group][Cout][8] bf16
// scale: BatchNorm scale per output channel, folded into the weights (y = conv(x, w*s) + t), or nullptr
static int vi_upload_tc_conv(uint8_t *dst, const std::vector<float> &w, int G, int NOUT, const float *scale, bool f16, bool fp16c)
{
    const int cin = G * 8;
    std::vector<uint16_t> b((size_t)25 * 2 * G * NOUT * 8);
    for (int tap = 0; tap < 25; ++tap)
        for (int g = 0; g < G; ++g)
            for (int co = 0; co < NOUT; ++co)
                for (int e = 0; e < 8; ++e) {
                    uint16_t hi, lo;
                    split_operand_host(w[((size_t)co * cin + g * 8 + e) * 25 + tap] * (scale ? scale[co] : 1.f), f16, hi, lo);
                    b[((((size_t)tap * 2 + 0) * G + g) * NOUT + co) * 8 + e] = hi;
                    b[((((size_t)tap * 2 + 1) * G + g) * NOUT + co) * 8 + e] = lo;
                }
    if (fp16c) {
        uint8_t *bb = reinterpret_cast<uint8_t *>(b.data());
        for (int tap = 0; tap < 25; ++tap)
            for (int j = 0; j < G / 2; ++j)
                for (int co = 0; co < NOUT; ++co) {
                    float w16[16];
                    for (int c = 0; c < 16; ++c) w16[c] = w[((size_t)co * cin + j * 16 + c) * 25 + tap] * (scale ? scale[co] : 1.f);
                    fp16c_weight_bytes(w16, bb + ((((size_t)tap * 2 + 1) * G + 2 * j) * NOUT + co) * 16, bb + ((((size_t)tap * 2 + 1) * G + 2 * j + 1) * NOUT + co) * 16);
                }
    }
    TB_CUDA(cudaMemcpy(dst, b.data(), b.size() * 2, cudaMemcpyHostToDevice));
    return TB_OK;
}

// conv2 (2-D tile kernel): B operand [tap][cin group][Cout rows of W_hi, then Cout rows of W_lo][8] bf16
static int vi_upload_tc_conv_cat(uint8_t *dst, const std::vector<float> &w, int G, int NOUT, const float *scale, bool f16, bool fp16c)
{
    const int cin = G * 8;
    std::vector<uint16_t> b((size_t)25 * G * 2 * NOUT * 8);
    for (int tap = 0; tap < 25; ++tap)
        for (int g = 0; g < G; ++g)
            for (int co = 0; co < NOUT; ++co)
                for (int e = 0; e < 8; ++e) {
                    uint16_t hi, lo;
                    split_operand_host(w[((size_t)co * cin + g * 8 + e) * 25 + tap] * (scale ? scale[co] : 1.f), f16, hi, lo);
                    b[((((size_t)tap * G + g) * 2 + 0) * NOUT + co) * 8 + e] = hi;
                    b[((((size_t)tap * G + g) * 2 + 1) * NOUT + co) * 8 + e] = lo;
                }
    if (fp16c) {               // the "lo rows" of groups 2j / 2j + 1 = the two e5m2 planes of channel block j
        uint8_t *bb = reinterpret_cast<uint8_t *>(b.data());
        for (int tap = 0; tap < 25; ++tap)
            for (int j = 0; j < G / 2; ++j)
                for (int co = 0; co < NOUT; ++co) {
                    float w16[16];
                    for (int c = 0; c < 16; ++c) w16[c] = w[((size_t)co * cin + j * 16 + c) * 25 + tap] * (scale ? scale[co] : 1.f);
                    fp16c_weight_bytes(w16, bb + ((((size_t)tap * G + 2 * j) * 2 + 1) * NOUT + co) * 16, bb + ((((size_t)tap * G + 2 * j + 1) * 2 + 1) * NOUT + co) * 16);
                }
    }
    TB_CUDA(cudaMemcpy(dst, b.data(), b.size() * 2, cudaMemcpyHostToDevice));
    return TB_OK;
}

// conv2, tap-pair kernel (fp16 / fp16c): 10 pair slots of 8 KB -- [fp16: group 0: 64 rows of tap (dy, dx) then 64 rows of tap (dy + 1, dx) | group 1]
// [e5m2: plane 0 (fp16(w) * 2^-8): 128 rows | plane 1 ((w - fp16(w)) * 2^8): 128 rows] for dy = 0, 2 -- then 5 single slots of 4 KB (dy = 4, 64-row blocks)
static int vi_upload_tc_conv2_pair(uint8_t *dst, const std::vector<float> &w, const float *scale, bool fp16c, int ncta)
{
    constexpr int NOUT = 64, CIN = 16;
    std::vector<uint8_t> b(tc::Conv2P::W_BYTES, 0);
    auto wt = [&](int co, int ci, int dy, int dx) { return w[((size_t)co * CIN + ci) * 25 + dy * 5 + dx] * (scale ? scale[co] : 1.f); };
    // rows = rows per block of the slot; filter co0 .. co0 + nco - 1 of tap (dy, dx) go to rows row0 ...
    auto fill = [&](uint8_t *slot, int rows, int row0, int co0, int nco, int dy, int dx) {
        uint16_t *h16 = reinterpret_cast<uint16_t *>(slot);
        uint8_t *f8 = slot + (size_t)2 * rows * 16;
        for (int r = 0; r < nco; ++r) {
            float w16[16];
            for (int c = 0; c < 16; ++c) {
                w16[c] = wt(co0 + r, c, dy, dx);
                h16[(((size_t)(c >> 3) * rows) + row0 + r) * 8 + (c & 7)] = __half_as_ushort(__float2half_rn(w16[c]));
            }
            if (fp16c) fp16c_weight_bytes(w16, f8 + ((size_t)0 * rows + row0 + r) * 16, f8 + ((size_t)1 * rows + row0 + r) * 16);
        }
    };
    if (ncta == 1) {
        for (int pr = 0; pr < 2; ++pr)
            for (int dx = 0; dx < 5; ++dx) {
                uint8_t *slot = b.data() + (size_t)(pr * 5 + dx) * tc::Conv2P::PAIR_BYTES;
                fill(slot, 128, 0, 0, NOUT, 2 * pr, dx);
                fill(slot, 128, 64, 0, NOUT, 2 * pr + 1, dx);
            }
        for (int dx = 0; dx < 5; ++dx) fill(b.data() + (size_t)10 * tc::Conv2P::PAIR_BYTES + (size_t)dx * tc::Conv2P::SINGLE_BYTES, 64, 0, 0, NOUT, 4, dx);
    } else {
        // CTA pair: rank r of the pair feeds the second half of every B operand's rows: tap (dy + r, dx) of a pair slot, filters 32 r .. 32 r + 31 of a
        // single slot; the slots shrink to half and each rank's 50 KB are contiguous
        constexpr int PB = tc::Conv2P::PAIR_BYTES / 2, SB = tc::Conv2P::SINGLE_BYTES / 2, WB = tc::Conv2P::W_BYTES / 2;
        for (int r = 0; r < 2; ++r) {
            uint8_t *base = b.data() + (size_t)r * WB;
            for (int pr = 0; pr < 2; ++pr)
                for (int dx = 0; dx < 5; ++dx) fill(base + (size_t)(pr * 5 + dx) * PB, 64, 0, 0, NOUT, 2 * pr + r, dx);
            for (int dx = 0; dx < 5; ++dx) fill(base + (size_t)10 * PB + (size_t)dx * SB, 32, 0, 32 * r, 32, 4, dx);
        }
    }
    TB_CUDA(cudaMemcpy(dst, b.data(), b.size(), cudaMemcpyHostToDevice));
    return TB_OK;
}

#ifdef TB_CONV2_STATS
extern "C" __attribute__((visibility("default"))) int tbdbg_conv2_stats(unsigned long long *out16, int reset)
{
    cudaDeviceSynchronize();
    if (out16) cudaMemcpyFromSymbol(out16, tb::tc::g_conv2_stats, 16 * sizeof(unsigned long long));
    if (reset) { unsigned long long z[16] = {}; cudaMemcpyToSymbol(tb::tc::g_conv2_stats, z, sizeof(z)); }
    return 0;
}
#endif

// TB_VI_CONV2_PAIR: which conv2 kernel the fp16 / fp16c precisions use -- 2 (default): tap pairs on CTA pairs (cta_group::2), 1: tap pairs on single
// CTAs, 0: the one-tap-per-MMA kernel
static int vi_conv2_variant()
{
    static const int v = getenv("TB_VI_CONV2_PAIR") ? atoi(getenv("TB_VI_CONV2_PAIR")) : 2;
    return v;
}

// conv weight torch [Cout][Cin][5][5] -> [tap][Cin][Cout]; BN(eval) folded with the conv bias into
// y = conv * s + t,  s = gamma / sqrt(var + eps),  t = (bias - mean) * s + beta
// `cout` is the width the kernels are built for, `creal` the network's (V100 / V110: conv3 has 100 channels, padded with zero
// filters).  arch 0 (V118_3): conv -> BN -> ReLU -> pool.  arch 1 (V100): conv -> ReLU -> pool, s = 1, t = bias.
// arch 2 (V110): conv -> pool -> BN -> ReLU = relu(maxpool(conv * sB) + sB * bias + tB) for sB > 0 (checked).
static int vi_conv(tb_vi *h, int idx, int cin, int cout, int creal, float *dw, float *ds, float *dt, std::vector<float> *w_padded = nullptr)
{
    const std::string c = "model.conv" + std::to_string(idx), b = "model.bn" + std::to_string(idx);
    const std::vector<float> *w, *bias, *g = nullptr, *be = nullptr, *mu = nullptr, *var = nullptr;
    const int arch = h->cfg.arch;
    int r;
    if ((r = vi_need(h, (c + ".weight").c_str(), (size_t)creal * cin * 25, &w))) return r;
    if ((r = vi_need(h, (c + ".bias").c_str(), creal, &bias))) return r;
    if (arch != 1) {
        if ((r = vi_need(h, (b + ".weight").c_str(), creal, &g))) return r;
        if ((r = vi_need(h, (b + ".bias").c_str(), creal, &be))) return r;
        if ((r = vi_need(h, (b + ".running_mean").c_str(), creal, &mu))) return r;
        if ((r = vi_need(h, (b + ".running_var").c_str(), creal, &var))) return r;
    }
    std::vector<float> wt((size_t)25 * cin * cout, 0.f), s(cout, 1.f), t(cout, 0.f);
    for (int co = 0; co < creal; ++co) {
        for (int ci = 0; ci < cin; ++ci)
            for (int tap = 0; tap < 25; ++tap) wt[((size_t)tap * cin + ci) * cout + co] = (*w)[((size_t)co * cin + ci) * 25 + tap];
        if (arch == 1) { s[co] = 1.f; t[co] = (*bias)[co]; continue; }
        s[co] = (*g)[co] / std::sqrt((*var)[co] + BN_EPS);
        if (arch == 0) t[co] = ((*bias)[co] - (*mu)[co]) * s[co] + (*be)[co];
        else {
            if (!(s[co] > 0.f)) {
                set_error("tb_vi_commit: v110 on the tensor path folds BatchNorm (applied after the max-pool) into the filters, which needs "
                          "positive BatchNorm scales; " + b + ".weight has a non-positive entry -- use precision fp32");
                return TB_ERR_INVALID;
            }
            t[co] = s[co] * (*bias)[co] + ((*be)[co] - (*mu)[co] * s[co]);
        }
    }
    if (w_padded) {           // torch layout [cout][cin][25] at the padded width, for the tensor-path operand builders
        w_padded->assign((size_t)cout * cin * 25, 0.f);
        std::copy(w->begin(), w->end(), w_padded->begin());
    }
    if ((r = vi_upload(dw, wt))) return r;
    if ((r = vi_upload(ds, s))) return r;
    return vi_upload(dt, t);
}

extern "C" int tb_vi_commit(tb_vi *h)
{
    TB_REQUIRE(h, TB_ERR_INVALID, "tb_vi_commit: null handle");
    TB_CUDA(cudaSetDevice(h->cfg.device));
    int r;
    if (h->net) {
        if ((r = vinet_commit(h->net, h->sd))) return r;
        h->committed = true;
        return TB_OK;
    }
    const int CI = h->cfg.channels;
    // V100 / V110 (arch 1 / 2) share V118_3's kernels: their conv3 has 100 channels (padded to 128 with zero filters, zero
    // shift -> ReLU(0) = 0 feeds zero fc1 columns) and fc1 therefore 10000 inputs; the head has no LayerNorm
    const int arch = h->cfg.arch, C3 = arch == 0 ? 128 : 100;
    std::vector<float> c3_padded;
    if ((r = vi_conv(h, 1, CI, 16, 16, h->w1, h->s1, h->t1))) return r;
    if ((r = vi_conv(h, 2, 16, 64, 64, h->w2, h->s2, h->t2))) return r;
    if ((r = vi_conv(h, 3, 64, 128, C3, h->w3, h->s3, h->t3, &c3_padded))) return r;
    const int M = h->cfg.num_classes;
    const std::vector<float> *w, *b, *g, *be, *w2, *b2;
    std::vector<float> fc1_padded, ng, nb;
    if ((r = vi_need(h, "model.fc1.weight", (size_t)100 * C3 * 100, &w))) return r;
    if ((r = vi_need(h, "model.fc1.bias", 100, &b))) return r;
    if (arch == 0) {          // LayerNorm(100)
        if ((r = vi_need(h, "model.bn4.weight", 100, &g))) return r;
        if ((r = vi_need(h, "model.bn4.bias", 100, &be))) return r;
    } else {                  // V110: BatchNorm1d(100) in eval mode as scale / shift; V100: nothing (1, 0)
        ng.assign(100, 1.f); nb.assign(100, 0.f);
        if (arch == 2) {
            const std::vector<float> *bg, *bb, *bm, *bv;
            if ((r = vi_need(h, "model.bn4.weight", 100, &bg))) return r;
            if ((r = vi_need(h, "model.bn4.bias", 100, &bb))) return r;
            if ((r = vi_need(h, "model.bn4.running_mean", 100, &bm))) return r;
            if ((r = vi_need(h, "model.bn4.running_var", 100, &bv))) return r;
            for (int i = 0; i < 100; ++i) { ng[i] = (*bg)[i] / std::sqrt((*bv)[i] + BN_EPS); nb[i] = (*bb)[i] - (*bm)[i] * ng[i]; }
        }
        g = &ng; be = &nb;
        fc1_padded.assign((size_t)100 * 12800, 0.f);           // torch column k = c * 100 + p: the first C3 * 100 columns of each row
        for (int o = 0; o < 100; ++o) std::copy(w->begin() + (size_t)o * C3 * 100, w->begin() + (size_t)(o + 1) * C3 * 100, fc1_padded.begin() + (size_t)o * 12800);
        w = &fc1_padded;
    }
    if ((r = vi_need(h, "model.fc2.weight", (size_t)M * 100, &w2))) return r;
    if ((r = vi_need(h, "model.fc2.bias", M, &b2))) return r;
    // fc1: torch flattens NCHW (k = c*100 + y*10 + x); activations here are NHWC (k' = (y*10+x)*128 + c)
    std::vector<float> wt((size_t)12800 * 100);
    for (int o = 0; o < 100; ++o)
        for (int c = 0; c < 128; ++c)
            for (int p = 0; p < 100; ++p) wt[((size_t)p * 128 + c) * 100 + o] = (*w)[(size_t)o * 12800 + c * 100 + p];
    if ((r = vi_upload(h->wf1, wt))) return r;
    if ((r = vi_upload(h->bf1, *b))) return r;
    if ((r = vi_upload(h->lng, *g))) return r;
    if ((r = vi_upload(h->lnb, *be))) return r;
    std::vector<float> w2t((size_t)100 * M);
    for (int o = 0; o < M; ++o)
        for (int k = 0; k < 100; ++k) w2t[(size_t)k * M + o] = (*w2)[(size_t)o * 100 + k];
    if ((r = vi_upload(h->wf2, w2t))) return r;
    if ((r = vi_upload(h->bf2, *b2))) return r;
    if (h->cfg.precision >= 1) {
        const std::vector<float> *c2, *c3;
        if ((r = vi_need(h, "model.conv2.weight", (size_t)64 * 16 * 25, &c2))) return r;
        c3 = &c3_padded;
        {   // conv1 B operand [cin][hi|lo][k-step j][k-chunk c][row = wp*16 + cout][8]: window row u = 2j+c, window col e;
            // row (wp = 2i+jj, cout) holds the 5x5 filter shifted by (i, jj) inside the 6x8 window, BN scale folded
            const std::vector<float> *c1;
            if ((r = vi_need(h, "model.conv1.weight", (size_t)16 * CI * 25, &c1))) return r;
            std::vector<float> sc1(16);
            TB_CUDA(cudaMemcpy(sc1.data(), h->s1, 16 * 4, cudaMemcpyDeviceToHost));
            std::vector<uint16_t> wb(tc::Conv1T::W_BYTES / 2 * CI, 0);
            for (int ci = 0; ci < CI; ++ci)
            for (int j = 0; j < 3; ++j)
                for (int c = 0; c < 2; ++c)
                    for (int wp = 0; wp < 4; ++wp)
                        for (int co = 0; co < 16; ++co)
                            for (int e = 0; e < 8; ++e) {
                                const int u = 2 * j + c, dy = u - (wp >> 1), dx = e - (wp & 1);
                                if (dy < 0 || dy > 4 || dx < 0 || dx > 4) continue;
                                uint16_t hi, lo;
                                split_bf16_host((*c1)[((size_t)co * CI + ci) * 25 + dy * 5 + dx] * sc1[co], hi, lo);
                                const size_t idx = (size_t)ci * (tc::Conv1T::W_BYTES / 2) + ((((size_t)j * 2 + c) * tc::Conv1T::N) + wp * 16 + co) * 8 + e;
                                wb[idx] = hi;
                                wb[(size_t)3 * 2 * tc::Conv1T::N * 8 + idx] = lo;
                            }
            TB_CUDA(cudaMemcpy(h->w1t, wb.data(), wb.size() * 2, cudaMemcpyHostToDevice));
        }
        // the 2-D conv2 and channel-major conv3 kernels expect the BN scale inside the weights
        std::vector<float> sc2(64), sc3(128);
        TB_CUDA(cudaMemcpy(sc2.data(), h->s2, 64 * 4, cudaMemcpyDeviceToHost));
        TB_CUDA(cudaMemcpy(sc3.data(), h->s3, 128 * 4, cudaMemcpyDeviceToHost));
        const bool f16 = h->cfg.precision >= 2, fp16c = h->cfg.precision == 3;
        if ((r = vi_upload_tc_conv_cat(h->w2t, *c2, 2, 64, sc2.data(), f16, fp16c))) return r;
        if (f16 && vi_conv2_variant() && (r = vi_upload_tc_conv2_pair(h->w2p, *c2, sc2.data(), fp16c, vi_conv2_variant() == 2 ? 2 : 1))) return r;
        if ((r = vi_upload_tc_conv(h->w3t, *c3, 8, 128, sc3.data(), f16, fp16c))) return r;
        // fc1 B operand [hi|lo][kc = c8*100 + pp][112][8]; torch column = (c8*8+e)*100 + pp
        std::vector<uint16_t> wb((size_t)2 * tc::FC_KC * tc::FC_N * 8, 0);
        for (int kc = 0; kc < tc::FC_KC; ++kc)
            for (int n = 0; n < tc::FC_NREAL; ++n)
                for (int e = 0; e < 8; ++e) {
                    const int c8 = kc / 100, pp = kc % 100;
                    uint16_t hi, lo;
                    split_bf16_host((*w)[(size_t)n * 12800 + (c8 * 8 + e) * 100 + pp], hi, lo);
                    wb[(((size_t)0 * tc::FC_KC + kc) * tc::FC_N + n) * 8 + e] = hi;
                    wb[(((size_t)1 * tc::FC_KC + kc) * tc::FC_N + n) * 8 + e] = lo;
                }
        TB_CUDA(cudaMemcpy(h->wfc, wb.data(), wb.size() * 2, cudaMemcpyHostToDevice));
    }
    h->committed = true;
    return TB_OK;
}

static int vi_forward_tc(tb_vi *h, const uint8_t *img, int n_max, const uint32_t *n_dev, float *probs, float *logits, cudaStream_t s)
{
    using namespace tb::tc;
    const int M = h->cfg.num_classes;
    static DeviceOnce attr_done;
    if (attr_done.need()) {
        TB_CUDA(cudaFuncSetAttribute(conv1_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Conv1T::smem(1)));
        TB_CUDA(cudaFuncSetAttribute(conv1_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, Conv1T::smem(3)));
        TB_CUDA(cudaFuncSetAttribute(conv1_tc_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Conv1P::SMEM));
        TB_CUDA(cudaFuncSetAttribute(conv2_2d_kernel<BF16X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, Conv2D::SMEM));
        TB_CUDA(cudaFuncSetAttribute(conv2_2d_kernel<FP16>, cudaFuncAttributeMaxDynamicSharedMemorySize, Conv2D::SMEM));
        TB_CUDA(cudaFuncSetAttribute(conv2_2d_kernel<FP16C>, cudaFuncAttributeMaxDynamicSharedMemorySize, Conv2D::SMEM));
        TB_CUDA(cudaFuncSetAttribute(conv2_pair_kernel<FP16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Conv2P::smem(true, 1)));
        TB_CUDA(cudaFuncSetAttribute(conv2_pair_kernel<FP16C, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Conv2P::smem(false, 1)));
        TB_CUDA(cudaFuncSetAttribute(conv2_pair_kernel<FP16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Conv2P::smem(true, 2)));
        TB_CUDA(cudaFuncSetAttribute(conv2_pair_kernel<FP16C, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Conv2P::smem(false, 2)));
        TB_CUDA(cudaFuncSetAttribute(conv3_t_kernel<BF16X3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Conv3T::SMEM));
        TB_CUDA(cudaFuncSetAttribute(conv3_t_kernel<FP16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Conv3T::SMEM));
        TB_CUDA(cudaFuncSetAttribute(conv3_t_kernel<FP16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Conv3T::SMEM));
        TB_CUDA(cudaFuncSetAttribute(conv3_t_kernel<FP16C, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Conv3T::SMEM));
        TB_CUDA(cudaFuncSetAttribute(conv3_t_kernel<FP16C, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Conv3T::SMEM));
        TB_CUDA(cudaFuncSetAttribute(fc1_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FC_SMEM));
        attr_done.done();
    }
    const int prec = h->cfg.precision, f16 = prec == 2 ? 1 : (prec == 3 ? 2 : 0);        // conv1's output encoding: bf16 hi / lo, fp16, fp16 + e5m2 planes
    for (int base = 0; base < n_max; base += h->chunk) {
        const int n = std::min(h->chunk, n_max - base);
        const int slot = h->prof.begin(s);
        h->prof.mark(slot, 0);
        static const bool c1_nopipe = getenv("TB_VI_CONV1_NOPIPE") != nullptr;      // bring-up switch: the per-crop barrier variant
        if (h->cfg.channels == 1 && !c1_nopipe) conv1_tc_pipe_kernel<<<std::min(n, h->n_sms), Conv1P::THREADS, Conv1P::SMEM, s>>>(img + (size_t)base * 6400, n, n_dev, base, h->w1t, h->t1, h->in2, f16);
        else if (h->cfg.channels == 1) conv1_tc_kernel<1><<<std::min(n, h->n_sms), Conv1T::THREADS, Conv1T::smem(1), s>>>(img + (size_t)base * 6400, n, n_dev, base, h->w1t, h->t1, h->in2, f16);
        else conv1_tc_kernel<3><<<std::min(n, h->n_sms), Conv1T::THREADS, Conv1T::smem(3), s>>>(img + (size_t)base * 6400 * 3, n, n_dev, base, h->w1t, h->t1, h->in2, f16);
        h->prof.mark(slot, 1);
        {
            const int g2 = std::min(n * Conv2D::BANDS, h->n_sms);
            const int pair_env = vi_conv2_variant();
            if (prec >= 2 && pair_env == 2) {
                // CTA pairs: clusters of two, the leader's MMAs run on both SMs (each reads half of the weights)
                cudaLaunchConfig_t lc{};
                lc.gridDim = dim3((unsigned)std::min((n * Conv2D::BANDS + 1) & ~1, h->n_sms & ~1)); lc.blockDim = dim3(Conv2D::THREADS);
                lc.dynamicSmemBytes = Conv2P::smem(prec == 2, 2); lc.stream = s;
                cudaLaunchAttribute at{};
                at.id = cudaLaunchAttributeClusterDimension;
                at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
                lc.attrs = &at; lc.numAttrs = 1;
                const uint8_t *a_in = h->in2, *a_w = h->w2p; const float *a_s = h->s2, *a_t = h->t2; uint8_t *a_out = h->in3;
                if (prec == 2) TB_CUDA(cudaLaunchKernelEx(&lc, conv2_pair_kernel<FP16, 2>, a_in, n, n_dev, base, a_w, a_s, a_t, a_out));
                else TB_CUDA(cudaLaunchKernelEx(&lc, conv2_pair_kernel<FP16C, 2>, a_in, n, n_dev, base, a_w, a_s, a_t, a_out));
            }
            else if (prec == 2 && pair_env) conv2_pair_kernel<FP16, 1><<<g2, Conv2D::THREADS, Conv2P::smem(true, 1), s>>>(h->in2, n, n_dev, base, h->w2p, h->s2, h->t2, h->in3);
            else if (prec == 3 && pair_env) conv2_pair_kernel<FP16C, 1><<<g2, Conv2D::THREADS, Conv2P::smem(false, 1), s>>>(h->in2, n, n_dev, base, h->w2p, h->s2, h->t2, h->in3);
            else if (prec == 2) conv2_2d_kernel<FP16><<<g2, Conv2D::THREADS, Conv2D::SMEM, s>>>(h->in2, n, n_dev, base, h->w2t, h->s2, h->t2, h->in3);
            else if (prec == 3) conv2_2d_kernel<FP16C><<<g2, Conv2D::THREADS, Conv2D::SMEM, s>>>(h->in2, n, n_dev, base, h->w2t, h->s2, h->t2, h->in3);
            else conv2_2d_kernel<BF16X3><<<g2, Conv2D::THREADS, Conv2D::SMEM, s>>>(h->in2, n, n_dev, base, h->w2t, h->s2, h->t2, h->in3);
        }
        h->prof.mark(slot, 2);
        {
            // full grids run as clusters of two CTAs sharing one multicast weight stream (TB_VI_CONV3_CLUSTER=0: off)
            static const int cl_env = getenv("TB_VI_CONV3_CLUSTER") ? atoi(getenv("TB_VI_CONV3_CLUSTER")) : 1;
            const int grid3 = std::min(n, h->n_sms);
            const bool cluster = prec >= 2 && cl_env != 0 && grid3 == h->n_sms && (grid3 % 2) == 0;     // bf16x3: no gain measured (its taps hide the L2 latency)
            if (cluster) {
                cudaLaunchConfig_t lc{};
                lc.gridDim = dim3((unsigned)grid3); lc.blockDim = dim3(Conv3T::THREADS); lc.dynamicSmemBytes = Conv3T::SMEM; lc.stream = s;
                cudaLaunchAttribute at{};
                at.id = cudaLaunchAttributeClusterDimension;
                at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
                lc.attrs = &at; lc.numAttrs = 1;
                const uint8_t *a_in = h->in3, *a_w = h->w3t; const float *a_s = h->s3, *a_t = h->t3; uint8_t *a_out = h->fca; int a_g = h->fc_groups;
                if (prec == 2) TB_CUDA(cudaLaunchKernelEx(&lc, conv3_t_kernel<FP16, 2>, a_in, n, n_dev, base, a_w, a_s, a_t, a_out, a_g));
                else TB_CUDA(cudaLaunchKernelEx(&lc, conv3_t_kernel<FP16C, 2>, a_in, n, n_dev, base, a_w, a_s, a_t, a_out, a_g));
            } else if (prec == 2) conv3_t_kernel<FP16, 1><<<grid3, Conv3T::THREADS, Conv3T::SMEM, s>>>(h->in3, n, n_dev, base, h->w3t, h->s3, h->t3, h->fca, h->fc_groups);
            else if (prec == 3) conv3_t_kernel<FP16C, 1><<<grid3, Conv3T::THREADS, Conv3T::SMEM, s>>>(h->in3, n, n_dev, base, h->w3t, h->s3, h->t3, h->fca, h->fc_groups);
            else conv3_t_kernel<BF16X3, 1><<<grid3, Conv3T::THREADS, Conv3T::SMEM, s>>>(h->in3, n, n_dev, base, h->w3t, h->s3, h->t3, h->fca, h->fc_groups);
        }
        h->prof.mark(slot, 3);
        fc1_tc_kernel<<<dim3((n + 127) / 128, FC_SPLIT), NT, FC_SMEM, s>>>(h->fca, h->fc_groups, n, n_dev, base, h->wfc, h->h1, h->chunk);
        h->prof.mark(slot, 4);
        head_kernel<<<(n + HD_WARPS - 1) / HD_WARPS, HD_WARPS * 32, head_smem(h, M), s>>>(
            h->h1, FC_SPLIT, (size_t)h->chunk * 100, h->bf1, M, n, n_dev, base, h->lng, h->lnb, h->wf2, h->bf2,
            probs + (size_t)base * M, logits ? logits + (size_t)base * M : nullptr, h->head_w_smem,
            h->top_id ? h->top_id + base : nullptr, h->top_p ? h->top_p + base : nullptr, h->cfg.arch != 0);
        h->prof.mark(slot, 5);
        h->launches += 5;
    }
    TB_CUDA(cudaGetLastError());
    h->last_stream = s;
    return TB_OK;
}

static int vi_forward(tb_vi *h, const uint8_t *img, int n_max, const uint32_t *n_dev, float *probs, float *logits, cudaStream_t s)
{
    if (h->net) {
        int r = vinet_forward(h->net, img, n_max, n_dev, probs, logits, h->top_id, h->top_p, s, h->prof, h->launches);
        h->last_stream = s;
        return r;
    }
    if (h->cfg.precision >= 1) return vi_forward_tc(h, img, n_max, n_dev, probs, logits, s);
    const int M = h->cfg.num_classes;
    static DeviceOnce attr_done;
    auto k2 = conv_kernel<16, 64, 40, 20, 3>;           // pooled 20x20: 7 tiles of 20x3
    auto k3 = conv_kernel<64, 128, 20, 10, 6>;          // pooled 10x10: 2 tiles of 10x6
    constexpr int SM2 = ConvTile<20, 3>::SMEM, SM3 = ConvTile<10, 6>::SMEM;
    if (attr_done.need()) {
        TB_CUDA(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, SM2));
        TB_CUDA(cudaFuncSetAttribute(k3, cudaFuncAttributeMaxDynamicSharedMemorySize, SM3));
        attr_done.done();
    }
    for (int base = 0; base < n_max; base += h->chunk) {
        const int n = std::min(h->chunk, n_max - base);
        const int CI = h->cfg.channels, c1_smem = (CI * 84 * 84 + 400 * CI + 32) * 4;
        static DeviceOnce c1_attr;
        if (c1_attr.need()) { TB_CUDA(cudaFuncSetAttribute(conv1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (3 * 84 * 84 + 1200 + 32) * 4)); c1_attr.done(); }
        const int slot = h->prof.begin(s);
        h->prof.mark(slot, 0);
        conv1_kernel<<<n, C1_NT, c1_smem, s>>>(img + (size_t)base * 6400 * CI, 80, 80, CI, n, n_dev, base, h->w1, h->s1, h->t1, h->a1);
        h->prof.mark(slot, 1);
        k2<<<dim3(n * 7, 1), CV_NT, SM2, s>>>(h->a1, n, n_dev, base, h->w2, h->s2, h->t2, h->a2);
        h->prof.mark(slot, 2);
        k3<<<dim3(n * 2, 2), CV_NT, SM3, s>>>(h->a2, n, n_dev, base, h->w3, h->s3, h->t3, h->a3);
        h->prof.mark(slot, 3);
        fc1_kernel<<<(n + FC_IMG - 1) / FC_IMG, FC_NT, 0, s>>>(h->a3, 12800, n, n_dev, base, h->wf1, h->bf1, h->h1);
        h->prof.mark(slot, 4);
        head_kernel<<<(n + HD_WARPS - 1) / HD_WARPS, HD_WARPS * 32, head_smem(h, M), s>>>(
            h->h1, 1, 0, nullptr, M, n, n_dev, base, h->lng, h->lnb, h->wf2, h->bf2,
            probs + (size_t)base * M, logits ? logits + (size_t)base * M : nullptr, h->head_w_smem,
            h->top_id ? h->top_id + base : nullptr, h->top_p ? h->top_p + base : nullptr, h->cfg.arch != 0);
        h->prof.mark(slot, 5);
        h->launches += 5;
    }
    TB_CUDA(cudaGetLastError());
    h->last_stream = s;
    return TB_OK;
}

extern "C" int tb_vi_predict_device(tb_vi *h, const void *images_dev, int n_max, const void *n_dev, void *probs_dev, void *logits_dev, void *stream)
{
    TB_REQUIRE(h && images_dev && probs_dev, TB_ERR_INVALID, "tb_vi_predict_device: null argument");
    TB_REQUIRE(h->committed, TB_ERR_STATE, "tb_vi_predict_device: no weights loaded (tb_vi_set_tensor + tb_vi_commit)");
    TB_REQUIRE(n_max > 0, TB_ERR_INVALID, "tb_vi_predict_device: n_max must be > 0");
    TB_CUDA(cudaSetDevice(h->cfg.device));
    return vi_forward(h, (const uint8_t *)images_dev, n_max, (const uint32_t *)n_dev, (float *)probs_dev, (float *)logits_dev,
                      stream ? (cudaStream_t)stream : h->stream);
}

extern "C" int tb_vi_predict(tb_vi *h, const uint8_t *images, int n, float *probs, float *logits)
{
    TB_REQUIRE(h && images && probs, TB_ERR_INVALID, "tb_vi_predict: null argument");
    TB_REQUIRE(h->committed, TB_ERR_STATE, "tb_vi_predict: no weights loaded (the reference throws SoftException here)");
    TB_REQUIRE(n > 0, TB_ERR_INVALID, "tb_vi_predict: n must be > 0");
    TB_CUDA(cudaSetDevice(h->cfg.device));
    const int M = h->cfg.num_classes;
    for (int base = 0; base < n; base += h->cfg.max_images) {
        const int m = std::min(h->cfg.max_images, n - base);
        const size_t ib = (size_t)6400 * h->cfg.channels;
        TB_CUDA(cudaMemcpyAsync(h->d_img, images + (size_t)base * ib, (size_t)m * ib, cudaMemcpyHostToDevice, h->stream));
        int r = vi_forward(h, h->d_img, m, nullptr, h->d_probs, logits ? h->d_logits : nullptr, h->stream);
        if (r) return r;
        TB_CUDA(cudaMemcpyAsync(probs + (size_t)base * M, h->d_probs, (size_t)m * M * 4, cudaMemcpyDeviceToHost, h->stream));
        if (logits) TB_CUDA(cudaMemcpyAsync(logits + (size_t)base * M, h->d_logits, (size_t)m * M * 4, cudaMemcpyDeviceToHost, h->stream));
        TB_CUDA(cudaStreamSynchronize(h->stream));
    }
    return TB_OK;
}

extern "C" int tb_vi_wait(tb_vi *h)
{
    TB_REQUIRE(h, TB_ERR_INVALID, "tb_vi_wait: null handle");
    TB_CUDA(cudaSetDevice(h->cfg.device));
    TB_CUDA(cudaStreamSynchronize(h->last_stream ? h->last_stream : h->stream));
    return TB_OK;
}

extern "C" uint64_t tb_vi_launch_count(tb_vi *h) { return h ? h->launches : 0; }

extern "C" int tb_vi_profile(tb_vi *h, int enable)
{
    TB_REQUIRE(h, TB_ERR_INVALID, "tb_vi_profile: null handle");
    TB_CUDA(cudaSetDevice(h->cfg.device));
    if (h->prof.enable(enable != 0) != TB_OK) { set_error("tb_vi_profile: cudaEventCreate failed"); return TB_ERR_CUDA; }
    return TB_OK;
}

extern "C" int tb_vi_kernel_ms(tb_vi *h, double out_ms[5], uint64_t *n_chunks)
{
    TB_REQUIRE(h && out_ms && n_chunks, TB_ERR_INVALID, "tb_vi_kernel_ms: null argument");
    TB_CUDA(cudaSetDevice(h->cfg.device));
    if (h->prof.flush() != TB_OK) { set_error("tb_vi_kernel_ms: event query failed"); return TB_ERR_CUDA; }
    for (int k = 0; k < 5; ++k) { out_ms[k] = h->prof.acc[k]; h->prof.acc[k] = 0; }
    *n_chunks = h->prof.n; h->prof.n = 0;
    return TB_OK;
}

extern "C" int tb_vi_set_top1(tb_vi *h, void *ids_dev, void *probs_dev)
{
    TB_REQUIRE(h, TB_ERR_INVALID, "tb_vi_set_top1: null handle");
    TB_REQUIRE((ids_dev == nullptr) == (probs_dev == nullptr), TB_ERR_INVALID, "tb_vi_set_top1: give both outputs or none");
    h->top_id = (uint32_t *)ids_dev; h->top_p = (float *)probs_dev;
    return TB_OK;
}
