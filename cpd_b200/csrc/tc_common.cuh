// PTX wrappers shared by the tcgen05 kernels (sm_100a): mbarrier, proxy fences, tcgen05.mma /
// commit / ld, TMEM allocation, and the 3xTF32 operand split.
#pragma once
#include "common.cuh"

namespace cpd {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(a_desc), "l"(b_desc),
        "r"(idesc), "r"(accumulate) : "memory");
}
// kind::f16 with bf16 operands (fp32 accumulate): K = 16 per instruction
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(a_desc), "l"(b_desc),
        "r"(idesc), "r"(accumulate) : "memory");
}
// One lane of a converged warp (the form the compiler turns into a single uniform branch around the UMMA stream;
// `lane == 0` makes it wrap every tcgen05.mma in an elect / broadcast loop).
__device__ __forceinline__ bool elect_one()
{
    uint32_t p;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(p));
    return p != 0;
}
// accumulate variants without the predicate register dance
__device__ __forceinline__ void umma_bf16_acc(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
                 "l"(a_desc), "l"(b_desc), "r"(idesc) : "memory");
}
__device__ __forceinline__ void umma_bf16_set(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
                 "l"(a_desc), "l"(b_desc), "r"(idesc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}


// hi = x with the low 13 mantissa bits cleared (exactly representable in tf32); lo = x - hi (exact).
__device__ __forceinline__ void split4(const float4 v, float4 &h, float4 &l)
{
    h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); l.x = v.x - h.x;
    h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); l.y = v.y - h.y;
    h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); l.z = v.z - h.z;
    h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); l.w = v.w - h.w;
}

// ---- bf16x3 operand split: x = hi + lo + O(2^-18 |x|), hi = RN_bf16(x), lo = RN_bf16(x - hi) (x - hi is exact in fp32).
// hi.hi + hi.lo + lo.hi on the bf16 tensor cores then carries ~2^-17 relative error per product, at half the
// shared-memory bytes and half the tensor time of the 3xTF32 split.
__device__ __forceinline__ uint32_t bf16x2_rn(float e0, float e1)      // e0 -> low half (lower address), e1 -> high half
{
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(e1), "f"(e0));
    return r;
}
__device__ __forceinline__ void split2(float e0, float e1, uint32_t &h, uint32_t &l)
{
    h = bf16x2_rn(e0, e1);
    l = bf16x2_rn(e0 - __uint_as_float(h << 16), e1 - __uint_as_float(h & 0xffff0000u));
}
// 8 consecutive fp32 -> 8 bf16 hi (16 bytes) + 8 bf16 lo (16 bytes)
__device__ __forceinline__ void split8(const float4 v0, const float4 v1, uint4 &h, uint4 &l)
{
    split2(v0.x, v0.y, h.x, l.x);
    split2(v0.z, v0.w, h.y, l.y);
    split2(v1.x, v1.y, h.z, l.z);
    split2(v1.z, v1.w, h.w, l.w);
}
// 4 consecutive fp32 -> 4 bf16 hi (8 bytes) + 4 bf16 lo (8 bytes)
__device__ __forceinline__ void split4b(const float4 v, uint2 &h, uint2 &l)
{
    split2(v.x, v.y, h.x, l.x);
    split2(v.z, v.w, h.y, l.y);
}

// ---- 1-D bulk copy global -> shared with mbarrier transaction accounting (TMA engine, no tensor map) ----
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t smem_dst, const void *gmem_src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst),
                 "l"(gmem_src), "r"(bytes), "r"(bar) : "memory");
}

// ---- tiled TMA load of a 4-D box (tensor map built by cuTensorMapEncodeTiled; coordinates innermost first; out-of-range
//      elements -- negative coordinates included -- are written as zeros and still count towards complete_tx) ----
__device__ __forceinline__ void tma_load_4d(uint32_t smem_dst, const void *tensor_map, uint32_t bar, int c0, int c1, int c2, int c3)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_dst), "l"(tensor_map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void *tensor_map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(tensor_map) : "memory");
}

// ---- 16-byte asynchronous copies global -> shared (LDGSTS); src_size = 0 zero-fills the destination ----
__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void *gmem_src, uint32_t src_size)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gmem_src), "r"(src_size) : "memory");
}
// explicit shared-space load (volatile: stays ordered with the barrier waits / copies around it, but carries no "memory"
// clobber, so several of them can be issued back to back before the first dependent instruction)
__device__ __forceinline__ int32_t lds_i32(uint32_t smem_addr)
{
    int32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(smem_addr));
    return v;
}
// L2 eviction-priority policies (createpolicy): the gathered operand is re-read ~K/2 times per row and should stay, the
// neighbour table streams through once
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void cp_async16_hint(uint32_t smem_dst, const void *gmem_src, uint32_t src_size, uint64_t policy)
{
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2, %3;" ::"r"(smem_dst), "l"(gmem_src), "r"(src_size), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s_hint(uint32_t smem_dst, const void *gmem_src, uint32_t bytes, uint32_t bar, uint64_t policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_dst),
                 "l"(gmem_src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// mbarrier arrival triggered when all prior cp.async of this thread have completed (counted in the barrier's init count)
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t cols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols));
}
// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread (lane = row)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&u)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
          "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

}  // namespace tc
}  // namespace cpd
