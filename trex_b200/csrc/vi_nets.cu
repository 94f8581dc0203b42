// V100 / V110 / V119 / V200 (T/python/visual_identification_network_torch.py:328-386, :262-325, :106-181, :30-103), eval mode,
// behind the same tb_vi_* entry points as V118_3 (tb_vi_config.arch).  One layer-list executor on fp32 CUDA cores:
//   conv block   y = pool_k( conv(x) * sA + tA ),  then optionally  y * sB + tB,  then ReLU
//                "pre"  (V200, V119): conv -> BN -> ReLU -> pool   sA,tA = folded BatchNorm + conv bias
//                "post" (V110):       conv -> pool -> BN -> ReLU   sA,tA = 1, bias;  sB,tB = folded BatchNorm
//                none   (V100):       conv -> ReLU -> pool         sA,tA = 1, bias   (ReLU and max-pool commute)
//   V200 only    global average pool over the last 2x2 map
//   head         fc1 (+ BatchNorm1d) + ReLU -> fc2 -> softmax (predict_numpy, T/python/visual_recognition_torch.py:337-345)
// Activations are NHWC fp32; a thread owns a WIN x WIN window of conv outputs (WIN = pool size, 2 when there is no pool)
// x CPT output channels, so the max-pool is thread-local.  Dropout layers are identities in eval mode.
#include "vi_nets.h"

#include <algorithm>
#include <cmath>

namespace tb {

constexpr float NET_BN_EPS = 1e-5f;
constexpr int GC_NT = 256, GC_CK = 8;

struct GConvArgs {
    const float *in; float *out;
    const float *w, *sa, *ta, *sb, *tb;     // w [KS*KS][CIN][COUT]; sb == nullptr: no second affine
    int H, W, CIN, COUT;                    // conv input = conv output size ('same' padding)
    int OH, OW;                             // stored map: H/pool x W/pool (floor, as nn.MaxPool2d)
    int TWP, THP, IPB, TX, TY;              // windows per tile, images per CTA (small maps), tiles per image
    int n_max, base; const uint32_t *n_dev;
};

template <int KS, int WIN, int CPT, bool POOL>
__global__ void __launch_bounds__(GC_NT)
gconv_kernel(GConvArgs a)
{
    extern __shared__ float sm[];
    constexpr int HALO = KS / 2, NCO = 4 * CPT, RW = WIN + KS - 1;
    const int PR = WIN * a.THP + KS - 1, PC = WIN * a.TWP + KS - 1, PITCH = PC | 1;
    float *patch = sm;                                       // [IPB][GC_CK][PR][PITCH], zero halo
    float *sw = sm + a.IPB * GC_CK * PR * PITCH;             // [KS*KS][GC_CK][NCO]
    const int n_act = a.n_dev ? min((int)*a.n_dev - a.base, a.n_max) : a.n_max;
    int img0, ty = 0, tx = 0;
    if (a.IPB > 1) img0 = blockIdx.x * a.IPB;
    else { const int tiles = a.TX * a.TY, t = blockIdx.x % tiles; img0 = blockIdx.x / tiles; ty = t / a.TX; tx = t % a.TX; }
    if (img0 >= n_act) return;
    const int co0 = blockIdx.y * NCO;
    const int tid = threadIdx.x, win = tid & 63, cg = tid >> 6;
    const int wpi = a.TWP * a.THP;
    const int il_raw = win / wpi, wl = win % wpi;
    const bool active = il_raw < a.IPB && img0 + il_raw < n_act;
    const int il = active ? il_raw : 0;
    const int wy = wl / a.TWP, wx = wl % a.TWP;
    const int y0 = ty * a.THP * WIN - HALO, x0 = tx * a.TWP * WIN - HALO;

    float acc[WIN * WIN][CPT];
#pragma unroll
    for (int k = 0; k < WIN * WIN; ++k)
#pragma unroll
        for (int j = 0; j < CPT; ++j) acc[k][j] = 0.f;

    const int per_img = PR * PC * GC_CK;
    for (int c0 = 0; c0 < a.CIN; c0 += GC_CK) {
        const int cn = min(GC_CK, a.CIN - c0);
        __syncthreads();
        for (int i = tid; i < a.IPB * per_img; i += GC_NT) {
            const int im = i / per_img, r = i % per_img, c = r % GC_CK, p = r / GC_CK, px = p % PC, py = p / PC;
            const int y = y0 + py, x = x0 + px;
            float v = 0.f;
            if (c < cn && img0 + im < n_act && y >= 0 && y < a.H && x >= 0 && x < a.W)
                v = a.in[(((size_t)(img0 + im) * a.H + y) * a.W + x) * a.CIN + c0 + c];
            patch[((im * GC_CK + c) * PR + py) * PITCH + px] = v;
        }
        for (int i = tid; i < KS * KS * GC_CK * NCO; i += GC_NT) {
            const int co = i % NCO, c = (i / NCO) % GC_CK, t = i / (NCO * GC_CK);
            sw[i] = (c < cn && co0 + co < a.COUT) ? a.w[((size_t)t * a.CIN + c0 + c) * a.COUT + co0 + co] : 0.f;
        }
        __syncthreads();
#pragma unroll 1
        for (int c = 0; c < cn; ++c) {
            const float *pc = patch + ((il * GC_CK + c) * PR + WIN * wy) * PITCH + WIN * wx;
#pragma unroll 1
            for (int dy = 0; dy < KS; ++dy) {
                float r[WIN][RW];
#pragma unroll
                for (int u = 0; u < WIN; ++u)
#pragma unroll
                    for (int i = 0; i < RW; ++i) r[u][i] = pc[(dy + u) * PITCH + i];
#pragma unroll
                for (int dx = 0; dx < KS; ++dx) {
                    const float4 *wp = reinterpret_cast<const float4 *>(sw + ((dy * KS + dx) * GC_CK + c) * NCO + cg * CPT);
#pragma unroll
                    for (int c4 = 0; c4 < CPT / 4; ++c4) {
                        const float4 wv = wp[c4];
                        const float ww[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j)
#pragma unroll
                            for (int u = 0; u < WIN; ++u)
#pragma unroll
                                for (int v = 0; v < WIN; ++v)
                                    acc[u * WIN + v][c4 * 4 + j] = fmaf(r[u][dx + v], ww[j], acc[u * WIN + v][c4 * 4 + j]);
                    }
                }
            }
        }
    }
    if (!active) return;
    const int gy = ty * a.THP + wy, gx = tx * a.TWP + wx;           // window coordinates inside the image
    float *dst = a.out + (size_t)(img0 + il) * a.OH * a.OW * a.COUT;
#pragma unroll
    for (int c4 = 0; c4 < CPT / 4; ++c4) {
        const int co = co0 + cg * CPT + c4 * 4;
        if (co >= a.COUT) continue;                                   // COUT is a multiple of 4
        const float4 sa = *reinterpret_cast<const float4 *>(a.sa + co), ta = *reinterpret_cast<const float4 *>(a.ta + co);
        const float s[4] = {sa.x, sa.y, sa.z, sa.w}, t[4] = {ta.x, ta.y, ta.z, ta.w};
        if (POOL) {
            if (gy >= a.OH || gx >= a.OW) continue;
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float m = fmaf(acc[0][c4 * 4 + j], s[j], t[j]);
#pragma unroll
                for (int k = 1; k < WIN * WIN; ++k) m = fmaxf(m, fmaf(acc[k][c4 * 4 + j], s[j], t[j]));
                if (a.sb) m = fmaf(m, a.sb[co + j], a.tb[co + j]);
                o[j] = fmaxf(m, 0.f);
            }
            *reinterpret_cast<float4 *>(dst + ((size_t)gy * a.OW + gx) * a.COUT + co) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
            for (int u = 0; u < WIN; ++u)
#pragma unroll
                for (int v = 0; v < WIN; ++v) {
                    const int y = gy * WIN + u, x = gx * WIN + v;
                    if (y >= a.OH || x >= a.OW) continue;
                    float o[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float m = fmaf(acc[u * WIN + v][c4 * 4 + j], s[j], t[j]);
                        if (a.sb) m = fmaf(m, a.sb[co + j], a.tb[co + j]);
                        o[j] = fmaxf(m, 0.f);
                    }
                    *reinterpret_cast<float4 *>(dst + ((size_t)y * a.OW + x) * a.COUT + co) = make_float4(o[0], o[1], o[2], o[3]);
                }
        }
    }
}

// u8 NHWC crops -> fp32 (no scaling, visual_recognition_torch.py:337)
__global__ void u8_to_f32_kernel(const uint8_t *__restrict__ img, float *__restrict__ out, size_t per_img, int n_max,
                                 const uint32_t *__restrict__ n_dev, int base)
{
    const int n_act = n_dev ? min((int)*n_dev - base, n_max) : n_max;
    const size_t total = (size_t)max(n_act, 0) * per_img;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) out[i] = (float)img[i];
}

// nn.AdaptiveAvgPool2d((1,1)): [n][P][C] -> [n][C]
__global__ void gap_kernel(const float *__restrict__ in, float *__restrict__ out, int P, int C, int n_max, const uint32_t *__restrict__ n_dev, int base)
{
    const int n_act = n_dev ? min((int)*n_dev - base, n_max) : n_max;
    const size_t total = (size_t)max(n_act, 0) * C;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t n = i / C, c = i % C;
        float s = 0.f;
        for (int p = 0; p < P; ++p) s += in[(n * P + p) * C + c];
        out[i] = s / (float)P;
    }
}

// Linear (+ folded BatchNorm1d) (+ ReLU): out[n][o] = act( (x[n] . wt[:, o]) * s[o] + t[o] ).  CTA = 32 images x 128 outputs.
constexpr int GF_NT = 256, GF_IMG = 32, GF_KC = 32, GF_OT = 128;

__global__ void __launch_bounds__(GF_NT)
gfc_kernel(const float *__restrict__ x, int K, int OUT, int n_max, const uint32_t *__restrict__ n_dev, int base,
           const float *__restrict__ wt /*[K][OUT]*/, const float *__restrict__ s, const float *__restrict__ t, int relu,
           float *__restrict__ out)
{
    __shared__ float sx[GF_IMG][GF_KC + 1];
    __shared__ float swt[GF_KC][GF_OT];
    const int n_act = n_dev ? min((int)*n_dev - base, n_max) : n_max;
    const int i0 = blockIdx.x * GF_IMG, o0 = blockIdx.y * GF_OT;
    if (i0 >= n_act) return;
    const int tid = threadIdx.x, img = tid >> 3, og = tid & 7;
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
    for (int k0 = 0; k0 < K; k0 += GF_KC) {
        __syncthreads();
        for (int i = tid; i < GF_IMG * GF_KC; i += GF_NT) {
            const int r = i / GF_KC, c = i % GF_KC;
            sx[r][c] = (i0 + r < n_act && k0 + c < K) ? x[(size_t)(i0 + r) * K + k0 + c] : 0.f;
        }
        for (int i = tid; i < GF_KC * GF_OT; i += GF_NT) {
            const int r = i / GF_OT, c = i % GF_OT;
            swt[r][c] = (k0 + r < K && o0 + c < OUT) ? wt[(size_t)(k0 + r) * OUT + o0 + c] : 0.f;
        }
        __syncthreads();
#pragma unroll 4
        for (int k = 0; k < GF_KC; ++k) {
            const float xv = sx[img][k];
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] = fmaf(xv, swt[k][og + 8 * j], acc[j]);
        }
    }
    if (i0 + img < n_act)
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int o = o0 + og + 8 * j;
            if (o < OUT) {
                const float v = fmaf(acc[j], s[o], t[o]);
                out[(size_t)(i0 + img) * OUT + o] = relu ? fmaxf(v, 0.f) : v;
            }
        }
}

// softmax(dim=1) + arg-max per image; one warp per image
__global__ void softmax_kernel(const float *__restrict__ logits, float *__restrict__ probs, int M, int n_max, const uint32_t *__restrict__ n_dev, int base,
                               uint32_t *__restrict__ top_id, float *__restrict__ top_p)
{
    const int n_act = n_dev ? min((int)*n_dev - base, n_max) : n_max;
    const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (n >= n_act) return;
    const float *l = logits + (size_t)n * M;
    float mx = -INFINITY; int arg = 0;
    for (int o = lane; o < M; o += 32) { const float v = l[o]; if (v > mx) { mx = v; arg = o; } }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, mx, o);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
        if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
    }
    float sum = 0.f;
    for (int o = lane; o < M; o += 32) sum += expf(l[o] - mx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.f / sum;
    for (int o = lane; o < M; o += 32) probs[(size_t)n * M + o] = expf(l[o] - mx) * inv;
    if (top_id && lane == 0) { top_id[n] = (uint32_t)arg; top_p[n] = inv; }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
enum BnMode { BN_NONE = 0, BN_PRE = 1, BN_POST = 2 };

struct NetConv {
    int cin, cout, ks, pool, bn, hw_in, hw_out;
    float *w = nullptr, *sa = nullptr, *ta = nullptr, *sb = nullptr, *tb = nullptr;
};

struct ViNet {
    int arch = 0, CI = 1, M = 0, chunk = 0;
    std::vector<NetConv> conv;
    bool gap = false;
    int fc1_in = 0, fc1_out = 0; const char *fc_bn = nullptr;
    float *wf1 = nullptr, *sf1 = nullptr, *tf1 = nullptr, *wf2 = nullptr, *sf2 = nullptr, *tf2 = nullptr;
    float *act[2] = {nullptr, nullptr}, *h1 = nullptr, *lg = nullptr;
};

template <typename T>
static int net_dev(std::vector<void *> &allocs, T **p, size_t n)
{
    int r = dev_alloc(p, std::max<size_t>(n, 1));
    if (r == TB_OK) allocs.push_back((void *)*p);
    return r;
}

int vinet_create(ViNet **out, int arch, int channels, int num_classes, int max_images, std::vector<void *> &allocs)
{
    ViNet *n = new ViNet();
    n->arch = arch; n->CI = channels; n->M = num_classes;
    struct Blk { int cout, ks, pool, bn; };
    std::vector<Blk> blks;
    switch (arch) {
    case 1: blks = {{16, 5, 2, BN_NONE}, {64, 5, 2, BN_NONE}, {100, 5, 2, BN_NONE}}; n->fc1_out = 100; n->fc_bn = nullptr; break;
    case 2: blks = {{16, 5, 2, BN_POST}, {64, 5, 2, BN_POST}, {100, 5, 2, BN_POST}}; n->fc1_out = 100; n->fc_bn = "model.bn4"; break;
    case 3: blks = {{256, 5, 2, BN_PRE}, {128, 5, 2, BN_PRE}, {32, 5, 2, BN_PRE}, {128, 5, 2, BN_PRE}}; n->fc1_out = 1024; n->fc_bn = "model.bn5"; break;
    case 4: blks = {{64, 3, 1, BN_PRE}, {128, 3, 3, BN_PRE}, {256, 3, 1, BN_PRE}, {512, 3, 3, BN_PRE}, {512, 3, 3, BN_PRE}};
            n->fc1_out = 1024; n->fc_bn = "model.bn6"; n->gap = true; break;
    default: delete n; set_error("tb_vi_create: arch must be 0 (v118_3), 1 (v100), 2 (v110), 3 (v119) or 4 (v200)"); return TB_ERR_INVALID;
    }
    int cin = channels, hw = 80;
    size_t act_max = (size_t)6400 * channels;
    for (const Blk &b : blks) {
        NetConv c{};
        c.cin = cin; c.cout = b.cout; c.ks = b.ks; c.pool = b.pool; c.bn = b.bn; c.hw_in = hw; c.hw_out = hw / b.pool;
        n->conv.push_back(c);
        cin = b.cout; hw = c.hw_out;
        act_max = std::max(act_max, (size_t)hw * hw * cin);
    }
    n->fc1_in = n->gap ? cin : cin * hw * hw;
    n->chunk = std::min(max_images, 256);
    const size_t CH = n->chunk, M = num_classes;
    int r = TB_OK;
#define A(p, cnt) if (r == TB_OK) r = net_dev(allocs, &(p), (cnt))
    for (NetConv &c : n->conv) {
        A(c.w, (size_t)c.ks * c.ks * c.cin * c.cout); A(c.sa, c.cout); A(c.ta, c.cout);
        if (c.bn == BN_POST) { A(c.sb, c.cout); A(c.tb, c.cout); }
    }
    A(n->wf1, (size_t)n->fc1_in * n->fc1_out); A(n->sf1, n->fc1_out); A(n->tf1, n->fc1_out);
    A(n->wf2, (size_t)n->fc1_out * M); A(n->sf2, M); A(n->tf2, M);
    A(n->act[0], CH * act_max); A(n->act[1], CH * act_max);
    A(n->h1, CH * (size_t)std::max(n->fc1_out, cin)); A(n->lg, CH * M);
#undef A
    if (r != TB_OK) { delete n; return r; }
    *out = n;
    return TB_OK;
}

void vinet_destroy(ViNet *n) { delete n; }          // device memory belongs to the owning tb_vi handle

static int net_need(const std::map<std::string, std::vector<float>> &sd, const std::string &name, size_t count, const std::vector<float> **out)
{
    auto it = sd.find(name);
    if (it == sd.end()) { set_error("tb_vi_commit: missing tensor " + name); return TB_ERR_STATE; }
    if (it->second.size() != count) {
        set_error("tb_vi_commit: tensor " + name + " has " + std::to_string(it->second.size()) + " elements, expected " + std::to_string(count));
        return TB_ERR_INVALID;
    }
    *out = &it->second;
    return TB_OK;
}

static int net_upload(float *dst, const std::vector<float> &v)
{
    TB_CUDA(cudaMemcpy(dst, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
    return TB_OK;
}

// BatchNorm (eval) as y = x * s + t
static int net_bn(const std::map<std::string, std::vector<float>> &sd, const std::string &name, int c, std::vector<float> &s, std::vector<float> &t)
{
    const std::vector<float> *g, *be, *mu, *var;
    int r;
    if ((r = net_need(sd, name + ".weight", c, &g))) return r;
    if ((r = net_need(sd, name + ".bias", c, &be))) return r;
    if ((r = net_need(sd, name + ".running_mean", c, &mu))) return r;
    if ((r = net_need(sd, name + ".running_var", c, &var))) return r;
    s.resize(c); t.resize(c);
    for (int i = 0; i < c; ++i) { s[i] = (*g)[i] / std::sqrt((*var)[i] + NET_BN_EPS); t[i] = (*be)[i] - (*mu)[i] * s[i]; }
    return TB_OK;
}

int vinet_commit(ViNet *n, const std::map<std::string, std::vector<float>> &sd)
{
    int r, idx = 0;
    for (NetConv &c : n->conv) {
        ++idx;
        const std::string cn = "model.conv" + std::to_string(idx), bn = "model.bn" + std::to_string(idx);
        const int taps = c.ks * c.ks;
        const std::vector<float> *w, *bias;
        if ((r = net_need(sd, cn + ".weight", (size_t)c.cout * c.cin * taps, &w))) return r;
        if ((r = net_need(sd, cn + ".bias", c.cout, &bias))) return r;
        std::vector<float> wt((size_t)taps * c.cin * c.cout), sa(c.cout, 1.f), ta(*bias);
        for (int co = 0; co < c.cout; ++co)
            for (int ci = 0; ci < c.cin; ++ci)
                for (int t = 0; t < taps; ++t) wt[((size_t)t * c.cin + ci) * c.cout + co] = (*w)[((size_t)co * c.cin + ci) * taps + t];
        if ((r = net_upload(c.w, wt))) return r;
        if (c.bn != BN_NONE) {
            std::vector<float> s, t;
            if ((r = net_bn(sd, bn, c.cout, s, t))) return r;
            if (c.bn == BN_PRE) { for (int i = 0; i < c.cout; ++i) { sa[i] = s[i]; ta[i] = (*bias)[i] * s[i] + t[i]; } }
            else { if ((r = net_upload(c.sb, s))) return r; if ((r = net_upload(c.tb, t))) return r; }
        }
        if ((r = net_upload(c.sa, sa))) return r;
        if ((r = net_upload(c.ta, ta))) return r;
    }
    const NetConv &last = n->conv.back();
    const int C = last.cout, P = n->gap ? 1 : last.hw_out * last.hw_out, K = n->fc1_in, O = n->fc1_out, M = n->M;
    const std::vector<float> *w1, *b1, *w2, *b2;
    if ((r = net_need(sd, "model.fc1.weight", (size_t)O * K, &w1))) return r;
    if ((r = net_need(sd, "model.fc1.bias", O, &b1))) return r;
    if ((r = net_need(sd, "model.fc2.weight", (size_t)M * O, &w2))) return r;
    if ((r = net_need(sd, "model.fc2.bias", M, &b2))) return r;
    // nn.Flatten on NCHW: k = c*P + p; the activations here are NHWC: k' = p*C + c
    std::vector<float> wt((size_t)K * O), s1(O, 1.f), t1(*b1);
    for (int o = 0; o < O; ++o)
        for (int c = 0; c < C; ++c)
            for (int p = 0; p < P; ++p) wt[((size_t)p * C + c) * O + o] = (*w1)[(size_t)o * K + (size_t)c * P + p];
    if (n->fc_bn) {
        std::vector<float> s, t;
        if ((r = net_bn(sd, n->fc_bn, O, s, t))) return r;
        for (int i = 0; i < O; ++i) { s1[i] = s[i]; t1[i] = (*b1)[i] * s[i] + t[i]; }
    }
    if ((r = net_upload(n->wf1, wt))) return r;
    if ((r = net_upload(n->sf1, s1))) return r;
    if ((r = net_upload(n->tf1, t1))) return r;
    std::vector<float> w2t((size_t)O * M), ones(M, 1.f);
    for (int o = 0; o < M; ++o)
        for (int k = 0; k < O; ++k) w2t[(size_t)k * M + o] = (*w2)[(size_t)o * O + k];
    if ((r = net_upload(n->wf2, w2t))) return r;
    if ((r = net_upload(n->sf2, ones))) return r;
    return net_upload(n->tf2, *b2);
}

// windows per tile along one axis: the value in [lo, 8] that wastes the fewest window slots
static int pick_tile(int nw, int cap)
{
    if (nw <= cap) return nw;
    int best = cap, waste = (nw + cap - 1) / cap * cap - nw;
    for (int t = cap - 1; t >= std::max(1, cap / 2); --t) {
        const int w = (nw + t - 1) / t * t - nw;
        if (w < waste) { waste = w; best = t; }
    }
    return best;
}

template <int KS, int WIN, int CPT, bool POOL>
static int launch_gconv(ViNet *n, GConvArgs &a, int n_img, cudaStream_t s)
{
    const int NWX = POOL ? a.OW : (a.W + WIN - 1) / WIN, NWY = POOL ? a.OH : (a.H + WIN - 1) / WIN;
    if (NWX * NWY <= 32) { a.TWP = NWX; a.THP = NWY; a.IPB = 64 / (NWX * NWY); a.TX = a.TY = 1; }
    else {
        a.TWP = pick_tile(NWX, 8); a.THP = pick_tile(NWY, 64 / a.TWP); a.IPB = 1;
        a.TX = (NWX + a.TWP - 1) / a.TWP; a.TY = (NWY + a.THP - 1) / a.THP;
    }
    const int PR = WIN * a.THP + KS - 1, PC = WIN * a.TWP + KS - 1, PITCH = PC | 1;
    const int smem = (a.IPB * GC_CK * PR * PITCH + KS * KS * GC_CK * 4 * CPT) * 4;
    // the largest configuration of this instantiation: 8x8 windows, one image
    constexpr int PRM = WIN * 8 + KS - 1, SMEM_MAX = (GC_CK * PRM * (PRM | 1) * 2 + KS * KS * GC_CK * 4 * CPT) * 4;
    static DeviceOnce attr;
    if (attr.need()) { TB_CUDA(cudaFuncSetAttribute(gconv_kernel<KS, WIN, CPT, POOL>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX)); attr.done(); }
    TB_REQUIRE(smem <= SMEM_MAX, TB_ERR_INVALID, "vinet: conv tile does not fit the shared-memory budget");
    const unsigned gx = a.IPB > 1 ? (unsigned)((n_img + a.IPB - 1) / a.IPB) : (unsigned)(n_img * a.TX * a.TY);
    gconv_kernel<KS, WIN, CPT, POOL><<<dim3(gx, (unsigned)((a.COUT + 4 * CPT - 1) / (4 * CPT))), GC_NT, smem, s>>>(a);
    (void)n;
    return TB_OK;
}

int vinet_forward(ViNet *n, const uint8_t *img, int n_max, const uint32_t *n_dev, float *probs, float *logits,
                  uint32_t *top_id, float *top_p, cudaStream_t s, EventRing<5> &prof, uint64_t &launches)
{
    const int M = n->M;
    for (int base = 0; base < n_max; base += n->chunk) {
        const int cnt = std::min(n->chunk, n_max - base);
        const int slot = prof.begin(s);
        prof.mark(slot, 0);
        const size_t per_img = (size_t)6400 * n->CI;
        u8_to_f32_kernel<<<148 * 4, 256, 0, s>>>(img + (size_t)base * per_img, n->act[0], per_img, cnt, n_dev, base);
        ++launches;
        int cur = 0, li = 0;
        for (NetConv &c : n->conv) {
            GConvArgs a{};
            a.in = n->act[cur]; a.out = n->act[cur ^ 1];
            a.w = c.w; a.sa = c.sa; a.ta = c.ta; a.sb = c.sb; a.tb = c.tb;
            a.H = a.W = c.hw_in; a.CIN = c.cin; a.COUT = c.cout; a.OH = a.OW = c.hw_out;
            a.n_max = cnt; a.base = base; a.n_dev = n_dev;
            int r;
            if (c.ks == 5 && c.pool == 2) r = launch_gconv<5, 2, 16, true>(n, a, cnt, s);
            else if (c.ks == 3 && c.pool == 3) r = launch_gconv<3, 3, 8, true>(n, a, cnt, s);
            else if (c.ks == 3 && c.pool == 1) r = launch_gconv<3, 2, 16, false>(n, a, cnt, s);
            else { set_error("vinet: unsupported conv block"); r = TB_ERR_INVALID; }
            if (r != TB_OK) return r;
            ++launches;
            cur ^= 1;
            if (li < 2) prof.mark(slot, li + 1);          // slots: conv1, conv2, remaining convs, fc1, head
            ++li;
        }
        prof.mark(slot, 3);
        const float *feat = n->act[cur];
        if (n->gap) {
            const NetConv &last = n->conv.back();
            gap_kernel<<<(cnt * last.cout + 255) / 256, 256, 0, s>>>(feat, n->h1, last.hw_out * last.hw_out, last.cout, cnt, n_dev, base);
            ++launches;
            feat = n->h1;
        }
        float *hid = n->act[cur ^ 1];
        gfc_kernel<<<dim3((cnt + GF_IMG - 1) / GF_IMG, (n->fc1_out + GF_OT - 1) / GF_OT), GF_NT, 0, s>>>(
            feat, n->fc1_in, n->fc1_out, cnt, n_dev, base, n->wf1, n->sf1, n->tf1, 1, hid);
        prof.mark(slot, 4);
        float *lg = logits ? logits + (size_t)base * M : n->lg;
        gfc_kernel<<<dim3((cnt + GF_IMG - 1) / GF_IMG, (M + GF_OT - 1) / GF_OT), GF_NT, 0, s>>>(
            hid, n->fc1_out, M, cnt, n_dev, base, n->wf2, n->sf2, n->tf2, 0, lg);
        softmax_kernel<<<(cnt + 7) / 8, 256, 0, s>>>(lg, probs + (size_t)base * M, M, cnt, n_dev, base,
                                                      top_id ? top_id + base : nullptr, top_p ? top_p + base : nullptr);
        prof.mark(slot, 5);
        launches += 3;
    }
    TB_CUDA(cudaGetLastError());
    return TB_OK;
}

}  // namespace tb
