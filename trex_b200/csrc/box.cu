// Normalised box filter of 8-bit planes as cv::boxFilter / cv::blur compute it -- the arithmetic behind two optional stages
// of RawProcessing::generate_binary (C/processing/RawProcessing.cpp): blur_difference (:371-387, cv::blur 25x25,
// BORDER_REFLECT_101) and use_adaptive_threshold (:427-434,487,526, cv::adaptiveThreshold(MEAN_C): box mean of the
// neighbourhood with BORDER_REPLICATE).  mean = (window sum) / k^2 rounded to nearest (k odd: never a tie), exact in integers.
// Two passes over a sub-batch of frames: row window sums from a shared-memory prefix of the row (closed form for the
// border part, so the neighbourhood may exceed the image: the reference's default adaptive_threshold_scale = 2 asks for
// 2 * cols + 1), then a sliding window down the columns.
#include "common.h"

namespace tb {

constexpr int BX_NT = 256, BX_SEG = 64;

// sum of a[lo..hi] under the border rule, from the inclusive-exclusive prefix P (P[i] = a[0] + .. + a[i-1]); lo <= hi, the
// range contains at least one in-range index.  border 0: replicate, 1: reflect-101 (needs hi - (n-1) <= n-1 and -lo <= n-1).
__device__ __forceinline__ uint32_t border_range_sum(const uint32_t *P, int n, int lo, int hi, int border)
{
    const int l = max(lo, 0), r = min(hi, n - 1);
    uint32_t s = P[r + 1] - P[l];
    if (lo < 0) s += border == 0 ? (uint32_t)(-lo) * (P[1] - P[0]) : P[-lo + 1] - P[1];
    if (hi > n - 1) s += border == 0 ? (uint32_t)(hi - (n - 1)) * (P[n] - P[n - 1]) : P[n - 1] - P[2 * (n - 1) - hi];
    return s;
}

// pass 1: one CTA per image row.  hs[y][x] = sum of the row over [x - p, x + p]
__global__ void __launch_bounds__(BX_NT)
box_rows_kernel(const uint8_t *__restrict__ src, uint32_t *__restrict__ hs, int W, int p, int border)
{
    extern __shared__ uint32_t P[];          // W + 1 prefix sums, then 33 words of scan scratch
    uint32_t *ws = P + W + 1;
    const uint8_t *row = src + (size_t)blockIdx.x * W;
    const int chunk = (W + BX_NT - 1) / BX_NT, x0 = threadIdx.x * chunk, x1 = min(x0 + chunk, W);
    uint32_t local = 0;
    for (int x = x0; x < x1; ++x) local += row[x];
    uint32_t total;
    uint32_t run = block_excl_scan(local, ws, total);
    for (int x = x0; x < x1; ++x) { P[x] = run; run += row[x]; }
    if (threadIdx.x == 0) P[W] = total;
    __syncthreads();
    uint32_t *out = hs + (size_t)blockIdx.x * W;
    for (int x = threadIdx.x; x < W; x += BX_NT) out[x] = border_range_sum(P, W, x - p, x + p, border);
}

__device__ __forceinline__ int border_index(int i, int n, int border)
{
    if (border == 0) return min(max(i, 0), n - 1);
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}

// pass 2: thread = (column, segment of BX_SEG rows) of one frame: direct window sum for the first row, then sliding
__global__ void __launch_bounds__(BX_NT)
box_cols_kernel(const uint32_t *__restrict__ hs, uint8_t *__restrict__ dst, int W, int H, int p, int border)
{
    const int x = blockIdx.x * BX_NT + threadIdx.x;
    if (x >= W) return;
    const int y0 = blockIdx.y * BX_SEG, y1 = min(y0 + BX_SEG, H);
    const uint32_t *col = hs + (size_t)blockIdx.z * W * H + x;
    uint8_t *out = dst + (size_t)blockIdx.z * W * H + x;
    const unsigned long long kk = (unsigned long long)(2 * p + 1) * (unsigned long long)(2 * p + 1);
    unsigned long long s = 0;
    {
        const int lo = y0 - p, hi = y0 + p, l = max(lo, 0), r = min(hi, H - 1);
        for (int i = l; i <= r; ++i) s += col[(size_t)i * W];
        if (border == 0) {
            if (lo < 0) s += (unsigned long long)(-lo) * col[0];
            if (hi > H - 1) s += (unsigned long long)(hi - (H - 1)) * col[(size_t)(H - 1) * W];
        } else {
            for (int i = lo; i < 0; ++i) s += col[(size_t)(-i) * W];
            for (int i = H; i <= hi; ++i) s += col[(size_t)(2 * (H - 1) - i) * W];
        }
    }
    const bool small = kk * 511ull < 0xFFFFFFFFull;          // 2 * s + kk fits 32 bits: cheap division
    for (int y = y0; y < y1; ++y) {
        if (y > y0) s += (unsigned long long)col[(size_t)border_index(y + p, H, border) * W] - (unsigned long long)col[(size_t)border_index(y - 1 - p, H, border) * W];
        out[(size_t)y * W] = small ? (uint8_t)(((uint32_t)s * 2u + (uint32_t)kk) / (2u * (uint32_t)kk)) : (uint8_t)((2ull * s + kk) / (2ull * kk));
    }
}

// mean of the k x k neighbourhood of every pixel of n planes (W x H each); hs: scratch of sub * W * H words
int launch_box_mean(const uint8_t *src, uint8_t *dst, uint32_t *hs, int sub, int W, int H, int n, int k, int border, cudaStream_t s)
{
    const int p = k / 2;
    TB_REQUIRE(k >= 1 && (k & 1), TB_ERR_INVALID, "box filter: the neighbourhood must be odd");
    TB_REQUIRE(border == 0 || (p <= W - 1 && p <= H - 1), TB_ERR_INVALID, "box filter: the blur window exceeds the frame");
    const int smem = (W + 1 + 33) * 4;
    static DeviceOnce attr;
    if (attr.need()) { TB_CUDA(cudaFuncSetAttribute(box_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (16384 + 1 + 33) * 4)); attr.done(); }
    for (int f0 = 0; f0 < n; f0 += sub) {
        const int m = min(sub, n - f0);
        box_rows_kernel<<<(unsigned)(m * H), BX_NT, smem, s>>>(src + (size_t)f0 * W * H, hs, W, p, border);
        box_cols_kernel<<<dim3((unsigned)((W + BX_NT - 1) / BX_NT), (unsigned)((H + BX_SEG - 1) / BX_SEG), (unsigned)m), BX_NT, 0, s>>>(
            hs, dst + (size_t)f0 * W * H, W, H, p, border);
    }
    TB_CUDA(cudaGetLastError());
    return TB_OK;
}

}  // namespace tb
