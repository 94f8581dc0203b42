// Debug entry point: one tcgen05 "shifted GEMM", the building block of the implicit-GEMM convolutions.
//   D[128][N] = sum_k A[shift + m][k] * B[n][k]
// A is stored as channel-group planes [n_cg][n_pos][8] bf16 (16 bytes per position and group), i.e. the
// SWIZZLE_NONE K-major canonical layout with SBO = 128 B and LBO = plane size; a tap of the convolution
// is just a different start address.  B is [n_cg][N][8].  Used by tests/test_gpu_umma.py.
#include "common.h"
#include "umma.cuh"

namespace tb {

__global__ void __launch_bounds__(128)
umma_shifted_gemm_kernel(const uint4 *__restrict__ a, int n_pos, int n_cg, int shift,
                         const uint4 *__restrict__ b, int N, float *__restrict__ dout, float *__restrict__ raw16)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base;
    uint4 *sa = reinterpret_cast<uint4 *>(smem);                     // [n_cg][n_pos]
    uint4 *sb = sa + (size_t)n_cg * n_pos;                           // [n_cg][N]
    const int tid = threadIdx.x, warp = tid >> 5;
    uint32_t ncols = 32; while ((int)ncols < N) ncols <<= 1;
    if (warp == 0) umma::tmem_alloc(&tmem_base, ncols);
    if (tid == 0) { umma::mbar_init(&mbar, 1); umma::fence_mbar_init(); }
    for (int i = tid; i < n_cg * n_pos; i += 128) sa[i] = a[i];
    for (int i = tid; i < n_cg * N; i += 128) sb[i] = b[i];
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tm = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = umma::idesc_bf16_f32(128, N);
        for (int ks = 0; ks < n_cg / 2; ++ks) {                      // one MMA = K 16 = two channel groups
            const uint64_t ad = umma::smem_desc(umma::smem_u32(sa + (size_t)(2 * ks) * n_pos + shift), (uint32_t)n_pos * 16u, 128u);
            const uint64_t bd = umma::smem_desc(umma::smem_u32(sb + (size_t)(2 * ks) * N), (uint32_t)N * 16u, 128u);
            umma::mma_bf16(tm, ad, bd, idesc, ks > 0);
        }
        umma::commit(&mbar);
    }
    umma::mbar_wait(&mbar, 0);
    umma::fence_after_sync();
    const int row = tid;                                             // TMEM lane = D row
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8];
        umma::tmem_ld8(tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        umma::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) dout[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
    }
    if (raw16) {            // the same accumulator through the 16x256b.x4 shape: raw16[warp][c0 / 32][lane half][thread][16 registers]
        for (int c0 = 0; c0 + 32 <= N; c0 += 32)
            for (int lh = 0; lh < 2; ++lh) {
                uint32_t v[16];
                umma::tmem_ld_16x256b_x4(tm + ((uint32_t)(warp * 32 + lh * 16) << 16) + (uint32_t)c0, v);
                umma::tmem_ld_wait();
                float *o = raw16 + ((((size_t)warp * (N / 32) + c0 / 32) * 2 + lh) * 32 + (tid & 31)) * 16;
#pragma unroll
                for (int j = 0; j < 16; ++j) o[j] = __uint_as_float(v[j]);
            }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tm, ncols);
}

// Mixed-kind accumulation (bring-up of the "fp16c" precision): D[128][N] = A16 . B16^T (kind::f16, fp16 operands in planes of 8
// channels) + A8 . B8^T (kind::f8f6f4, e5m2 operands in planes of 16 channels; one K = 32 MMA spans planes 2b, 2b+1 through LBO),
// both into the SAME fp32 TMEM columns, the start address shifted like a convolution tap.
__global__ void __launch_bounds__(128)
umma_mixed_gemm_kernel(const uint4 *__restrict__ a16, int n_pos, int n_cg, int shift, const uint4 *__restrict__ b16,
                       const uint4 *__restrict__ a8, const uint4 *__restrict__ b8, int n_pl8, int N, float *__restrict__ dout)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base;
    uint4 *sa = reinterpret_cast<uint4 *>(smem);                     // [n_cg][n_pos] fp16 x 8
    uint4 *sb = sa + (size_t)n_cg * n_pos;                           // [n_cg][N]
    uint4 *sa8 = sb + (size_t)n_cg * N;                              // [n_pl8][n_pos] e5m2 x 16
    uint4 *sb8 = sa8 + (size_t)n_pl8 * n_pos;                        // [n_pl8][N]
    const int tid = threadIdx.x, warp = tid >> 5;
    uint32_t ncols = 32; while ((int)ncols < N) ncols <<= 1;
    if (warp == 0) umma::tmem_alloc(&tmem_base, ncols);
    if (tid == 0) { umma::mbar_init(&mbar, 1); umma::fence_mbar_init(); }
    for (int i = tid; i < n_cg * n_pos; i += 128) sa[i] = a16[i];
    for (int i = tid; i < n_cg * N; i += 128) sb[i] = b16[i];
    for (int i = tid; i < n_pl8 * n_pos; i += 128) sa8[i] = a8[i];
    for (int i = tid; i < n_pl8 * N; i += 128) sb8[i] = b8[i];
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tm = tmem_base;
    if (tid == 0) {
        const uint32_t i16 = umma::idesc_f16_f32(128, N), i8 = umma::idesc_e5m2_f32(128, N);
        for (int ks = 0; ks < n_cg / 2; ++ks) {
            const uint64_t ad = umma::smem_desc(umma::smem_u32(sa + (size_t)(2 * ks) * n_pos + shift), (uint32_t)n_pos * 16u, 128u);
            const uint64_t bd = umma::smem_desc(umma::smem_u32(sb + (size_t)(2 * ks) * N), (uint32_t)N * 16u, 128u);
            umma::mma_bf16(tm, ad, bd, i16, ks > 0);
        }
        for (int b = 0; b < n_pl8 / 2; ++b) {                        // K = 32: core matrix 0 from plane 2b, core matrix 1 from plane 2b + 1
            const uint64_t ad = umma::smem_desc(umma::smem_u32(sa8 + (size_t)(2 * b) * n_pos + shift), (uint32_t)n_pos * 16u, 128u);
            const uint64_t bd = umma::smem_desc(umma::smem_u32(sb8 + (size_t)(2 * b) * N), (uint32_t)N * 16u, 128u);
            umma::mma_f8(tm, ad, bd, i8, 1);
        }
        umma::commit(&mbar);
    }
    umma::mbar_wait(&mbar, 0);
    umma::fence_after_sync();
    const int row = tid;
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8];
        umma::tmem_ld8(tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        umma::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) dout[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tm, ncols);
}

}  // namespace tb

#define TB_DEBUG_EXPORTS 1
// bring-up entry point: exported for tests/test_gpu_umma.py, declared in the header only under TB_DEBUG_EXPORTS (not product ABI)
static int umma_shifted_gemm_run(const void *a_host, int n_pos, int n_cg, int shift, const void *b_host, int N, float *d_host, float *raw16_host);
extern "C" __attribute__((visibility("default"))) int tbdbg_umma_shifted_gemm(const void *a_host, int n_pos, int n_cg, int shift, const void *b_host, int N, float *d_host)
{
    return umma_shifted_gemm_run(a_host, n_pos, n_cg, shift, b_host, N, d_host, nullptr);
}
// the same GEMM; additionally returns the accumulator as the 16x256b.x4 TMEM load shape delivers it: raw16[4 warps][N / 32][2 lane halves][32 threads][16]
extern "C" __attribute__((visibility("default"))) int tbdbg_tmem_ld_16x256b(const void *a_host, int n_pos, int n_cg, int shift, const void *b_host, int N, float *d_host, float *raw16_host)
{
    TB_REQUIRE(raw16_host && N % 32 == 0, TB_ERR_INVALID, "tbdbg_tmem_ld_16x256b: need the raw output and N % 32 == 0");
    return umma_shifted_gemm_run(a_host, n_pos, n_cg, shift, b_host, N, d_host, raw16_host);
}
static int umma_shifted_gemm_run(const void *a_host, int n_pos, int n_cg, int shift, const void *b_host, int N, float *d_host, float *raw16_host)
{
    using namespace tb;
    TB_REQUIRE(a_host && b_host && d_host, TB_ERR_INVALID, "tbdbg_umma_shifted_gemm: null argument");
    TB_REQUIRE(n_cg >= 2 && n_cg % 2 == 0 && N >= 16 && N <= 256 && N % 16 == 0 && shift >= 0 && shift + 128 <= n_pos,
               TB_ERR_INVALID, "tbdbg_umma_shifted_gemm: bad shape");
    const size_t abytes = (size_t)n_cg * n_pos * 16, bbytes = (size_t)n_cg * N * 16;
    TB_REQUIRE(abytes + bbytes <= 200 * 1024, TB_ERR_INVALID, "tbdbg_umma_shifted_gemm: operands exceed shared memory");
    void *da = nullptr, *db = nullptr; float *dd = nullptr, *dr = nullptr;
    TB_CUDA(cudaMalloc(&da, abytes)); TB_CUDA(cudaMalloc(&db, bbytes)); TB_CUDA(cudaMalloc((void **)&dd, (size_t)128 * N * 4));
    if (raw16_host) TB_CUDA(cudaMalloc((void **)&dr, (size_t)128 * N * 4));
    TB_CUDA(cudaMemcpy(da, a_host, abytes, cudaMemcpyHostToDevice));
    TB_CUDA(cudaMemcpy(db, b_host, bbytes, cudaMemcpyHostToDevice));
    TB_CUDA(cudaFuncSetAttribute(umma_shifted_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(abytes + bbytes)));
    umma_shifted_gemm_kernel<<<1, 128, abytes + bbytes>>>((const uint4 *)da, n_pos, n_cg, shift, (const uint4 *)db, N, dd, dr);
    TB_CUDA(cudaGetLastError());
    TB_CUDA(cudaDeviceSynchronize());
    TB_CUDA(cudaMemcpy(d_host, dd, (size_t)128 * N * 4, cudaMemcpyDeviceToHost));
    if (raw16_host) TB_CUDA(cudaMemcpy(raw16_host, dr, (size_t)128 * N * 4, cudaMemcpyDeviceToHost));
    cudaFree(da); cudaFree(db); cudaFree(dd); if (dr) cudaFree(dr);
    return TB_OK;
}

// a16 / b16: fp16 bits [n_cg][n_pos | N][8]; a8 / b8: e5m2 bytes [n_pl8][n_pos | N][16] (n_pl8 even: pairs of planes form one K = 32 step)
extern "C" __attribute__((visibility("default"))) int tbdbg_umma_mixed_gemm(const void *a16, int n_pos, int n_cg, int shift, const void *b16,
                                                                            const void *a8, const void *b8, int n_pl8, int N, float *d_host)
{
    using namespace tb;
    TB_REQUIRE(a16 && b16 && a8 && b8 && d_host, TB_ERR_INVALID, "tbdbg_umma_mixed_gemm: null argument");
    TB_REQUIRE(n_cg >= 2 && n_cg % 2 == 0 && n_pl8 >= 2 && n_pl8 % 2 == 0 && N >= 16 && N <= 256 && N % 16 == 0 && shift >= 0 && shift + 128 <= n_pos,
               TB_ERR_INVALID, "tbdbg_umma_mixed_gemm: bad shape");
    const size_t ab = (size_t)n_cg * n_pos * 16, bb = (size_t)n_cg * N * 16, a8b = (size_t)n_pl8 * n_pos * 16, b8b = (size_t)n_pl8 * N * 16;
    TB_REQUIRE(ab + bb + a8b + b8b <= 200 * 1024, TB_ERR_INVALID, "tbdbg_umma_mixed_gemm: operands exceed shared memory");
    void *p[4] = {nullptr, nullptr, nullptr, nullptr}; float *dd = nullptr;
    const void *src[4] = {a16, b16, a8, b8}; const size_t sz[4] = {ab, bb, a8b, b8b};
    for (int i = 0; i < 4; ++i) { TB_CUDA(cudaMalloc(&p[i], sz[i])); TB_CUDA(cudaMemcpy(p[i], src[i], sz[i], cudaMemcpyHostToDevice)); }
    TB_CUDA(cudaMalloc((void **)&dd, (size_t)128 * N * 4));
    TB_CUDA(cudaFuncSetAttribute(umma_mixed_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(ab + bb + a8b + b8b)));
    umma_mixed_gemm_kernel<<<1, 128, ab + bb + a8b + b8b>>>((const uint4 *)p[0], n_pos, n_cg, shift, (const uint4 *)p[1], (const uint4 *)p[2], (const uint4 *)p[3], n_pl8, N, dd);
    TB_CUDA(cudaGetLastError());
    TB_CUDA(cudaDeviceSynchronize());
    TB_CUDA(cudaMemcpy(d_host, dd, (size_t)128 * N * 4, cudaMemcpyDeviceToHost));
    for (int i = 0; i < 4; ++i) cudaFree(p[i]);
    cudaFree(dd);
    return TB_OK;
}
