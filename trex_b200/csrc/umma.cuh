// Thin inline-PTX layer for Blackwell tcgen05 / TMEM / mbarrier / bulk copies (sm_100a).
// Descriptor encodings follow the PTX ISA "tcgen05 matrix descriptor" / "instruction descriptor".
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace tb { namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Shared-memory matrix descriptor, SWIZZLE_NONE, K-major canonical layout:
//   core matrix = 8 rows x 16 bytes stored as 128 contiguous bytes (row pitch 16 B);
//   LBO = byte distance between the core matrices adjacent in K,
//   SBO = byte distance between the 8-row groups adjacent in M/N.
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;                     // descriptor version (sm_100)
    return d;                                   // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}

// descriptor + offset (in 16-byte units) with ONE 32-bit add: the start-address field sits in bits 0-13 of the low
// word and never carries out of it (shared memory < 256 KB), so the single MMA-issuing thread spends 1 instead of 2
// dependent integer instructions per operand -- that thread's instruction latency bounds kernels of small MMAs
__device__ __forceinline__ uint64_t desc_add(uint64_t base, uint32_t ofs16)
{
    return (base & 0xFFFFFFFF00000000ull) | (uint64_t)((uint32_t)base + ofs16);
}

// Instruction descriptor for kind::f16 with BF16 A/B (K-major both), FP32 accumulate.
__host__ __device__ constexpr uint32_t idesc_bf16_f32(int M, int N)
{
    return (1u << 4)                    // D format F32
         | (1u << 7) | (1u << 10)       // A, B format BF16
         | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// The same with FP16 A/B (format code 0): 11-bit significands, used by the single-MMA "fp16" precision.
__host__ __device__ constexpr uint32_t idesc_f16_f32(int M, int N)
{
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// kind::f8f6f4 with E5M2 A/B (format code 1; K-major both; one MMA covers K = 32 = two 16-byte core matrices), FP32 accumulate.
// Used for the correction terms of the "fp16c" precision (vi_tc.cuh): twice the K per instruction at the same issue cost.
__host__ __device__ constexpr uint32_t idesc_e5m2_f32(int M, int N)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void tmem_alloc(uint32_t *smem_dst, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma operand fetch, bulk copies)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// One elected lane of a converged warp (elect.sync).  Unlike `lane == 0`, the compiler knows that exactly one lane
// passes, so the tcgen05.mma / commit instructions (uniform datapath) inside the branch are emitted straight instead
// of inside a per-lane election loop (~6 extra dependent instructions per MMA for the single issuing thread).
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// the same for 8-bit float operands (kind::f8f6f4, K = 32 per instruction); accumulates into the same fp32 TMEM columns
__device__ __forceinline__ void mma_f8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void commit(uint64_t *mbar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(mbar)) : "memory");
}

// ---- CTA pairs (cta_group::2): the MMA of the leader CTA (cluster rank 0) runs on the tensor cores of BOTH SMs of the pair.  M = 256 = 128 rows of A
// from each CTA's own shared memory (the descriptor offsets are applied in both), N/2 rows of B from each CTA -- so every SM reads only half of the
// B operand -- and each CTA's TMEM receives its 128 rows x N columns.  Both CTAs' warps of the same index allocate / free together. ----
__device__ __forceinline__ void tmem_alloc2(uint32_t *smem_dst, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma2_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma2_f8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on the mbarrier at this shared-memory offset in every CTA of cta_mask when all MMAs issued so far by this thread have completed
__device__ __forceinline__ void commit2(uint64_t *mbar, uint16_t cta_mask)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(smem_u32(mbar)), "h"(cta_mask) : "memory");
}
// arrive on an mbarrier of another CTA of the cluster (address from mapa()); wait with cluster-scope acquire
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr)
{
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(cluster_addr) : "memory");
}
// the same without memory ordering: for hand-shakes that publish no generic-proxy data (e.g. "my tcgen05.ld of this accumulator has completed").
// The .release.cluster form costs a MEMBAR.ALL.GPU, which waits for every earlier global store of the thread.
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint32_t cluster_addr)
{
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" :: "r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *mbar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(mbar)), "r"(parity) : "memory");
    } while (!ok);
}

// the same on a shared::cluster address (e.g. a barrier of the peer CTA of a cluster, see mapa())
__device__ __forceinline__ void commit_a(uint32_t cluster_addr)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(cluster_addr) : "memory");
}
// thread-block clusters: rank of this CTA, address of a shared-memory object in CTA `rank`, cluster-wide barrier
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa(uint32_t cta_addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 1-D bulk copy global -> the same shared-memory offset of every CTA in cta_mask; each destination CTA's mbarrier (same
// offset) receives the complete_tx
__device__ __forceinline__ void bulk_g2s_multicast(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *mbar, uint16_t cta_mask)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(mbar)), "h"(cta_mask) : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t *mbar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(mbar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t *mbar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(mbar)), "r"(parity) : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_wait(uint64_t *mbar, uint32_t parity) { while (!mbar_try_wait(mbar, parity)) { } }
// same, but the thread is suspended by the hardware until the phase completes (or the time hint expires) instead of
// polling: waiting warps do not take issue slots from the warps they wait for
__device__ __forceinline__ void mbar_wait_suspend(uint64_t *mbar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
                 "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
                 :: "r"(smem_u32(mbar)), "r"(parity), "r"(0x989680u) : "memory");
}
// variants taking 32-bit shared-window addresses (no generic -> shared conversion in inner loops)
__device__ __forceinline__ void mbar_wait_suspend_a(uint32_t a, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
                 "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
                 :: "r"(a), "r"(parity), "r"(0x989680u) : "memory");
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t a) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(a) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx_a(uint32_t a, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s_a(uint32_t smem_dst, const void *gsrc, uint32_t bytes, uint32_t mbar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *mbar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(mbar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *mbar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(mbar)), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> shared (TMA engine, no tensor map), completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *mbar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(mbar)) : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 8 / 16 consecutive columns (one row per thread)
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr));
}
// TMEM -> registers, shape 16x256b.x4: 16 lanes starting at the address' lane x 32 consecutive columns.  Thread t receives, for column group
// j = 0 .. 3 (8 columns each), v[4 j + 0 .. 1] = lane (t / 4), columns 8 j + 2 (t % 4) + {0, 1} and v[4 j + 2 .. 3] = lane (t / 4) + 8, same columns
// (the mma accumulator fragment; checked on hardware by tests/test_gpu_umma.py::test_tmem_ld_16x256b_layout)
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// fp32 -> bf16 hi + bf16 lo (x ~= hi + lo, |err| ~ 2^-17 |x|)
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16 &hi, __nv_bfloat16 &lo)
{
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

}}  // namespace tb::umma
