// tcgen05 gather-GEMM for sm_100a: y[o,:] = epi( sum_k W[:,k,:] x[nbr[o,k],:] ) with fp32-class
// accuracy on the 5th-generation tensor cores (3xTF32 operand splitting).
//
// Serves spconv.SubMConv3d / SparseConv3d forward + input-gradient (call sites
// cpd/models/backbones_3d/spconv_backbone.py:17,20-21,108-115) and, through cpd_conv2d_table,
// the dense BEV convolutions (cpd/models/backbones_2d/base_bev_backbone.py:31-59,
// cpd/models/dense_heads/center_head.py:11-45,73-80).
//
// One CTA = 128 output rows x all BN = C_out columns, accumulator in TMEM (BN columns).
//   warps 0-7  producers: gather A rows (x[nbr[o,k]], 128 B per row per k-block) and the
//              W[:,k,c0:c0+32] slab straight from global/L2 with 16-byte loads, split every
//              fp32 into tf32 hi + lo parts in registers and store both into shared memory in
//              the canonical K-major SWIZZLE_128B layout that UMMA descriptors address
//              (gathered rows are not TMA-tileable; the split has to pass through registers
//              anyway).  fence.proxy.async + mbarrier arrive hand the stage to the MMA warp.
//              Taps for which no row of the tile has a neighbour are skipped altogether.
//   warp 8     allocates TMEM, then one elected lane issues per k-block 4 x 3
//              tcgen05.mma.cta_group::1.kind::tf32 (A_hi.B_hi + A_lo.B_hi + A_hi.B_lo, M=128,
//              N=BN, K=8) and tcgen05.commit's the stage back to the producers.
//   warps 0-7  epilogue: tcgen05.ld the accumulator (lane = row), + bias, stage through shared
//              memory (padded rows, conflict-free), then coalesced float4 stores with the folded
//              BatchNorm affine / residual / ReLU applied on the way out and per-channel
//              sum / sum-of-squares taken from the staged tile.
//
// TF32 keeps 10 mantissa bits: a single-pass product would miss the 1e-4 parity bar
// (SURVEY.md H4); hi = x & 0xffffe000, lo = x - hi (exact) restores ~2^-21 relative error.
#include "tc_common.cuh"

namespace cpd {
namespace {
using namespace tc;

constexpr int BM = 128;       // UMMA M
constexpr int BK = 32;        // fp32 per k-block = 128 bytes = one swizzle row
constexpr int NPW = 8;             // producer / epilogue warps (two per SM sub-partition, so their issue stalls overlap)
constexpr int NPROD = NPW * 32;
constexpr int NTHREADS = NPROD + 32;   // + MMA warp (warp NPW)
constexpr int RSTEP = NPROD / 8;   // row stride between the chunks one producer thread owns (8 x 16 B chunks per row)
constexpr int MAX_TAPS = 32;

__host__ __device__ constexpr int stages_for(int bn) { return bn >= 256 ? 2 : bn >= 128 ? 3 : 2; }
__host__ __device__ constexpr int ctas_per_sm(int bn) { return bn <= 64 ? 2 : 1; }
__host__ __device__ constexpr int stage_bytes(int bn) { return 2 * BM * BK * 4 + 2 * bn * BK * 4; }
// TMEM accumulators per tile: n_main(bn) "main" ones (A_hi.B_hi, k-blocks dealt round-robin) + 1
// "correction" one (A_lo.B_hi + A_hi.B_lo, ~2^-11 of the main magnitude).  The tensor core
// truncates when it adds into the fp32 accumulator; with thousands of adds into one accumulator
// that bias reaches ~1e-4 (measured).  Spreading the adds over separate accumulators and summing
// them in fp32 registers in the epilogue cuts it by 3 * n_main for free (TMEM columns are idle).
__host__ __device__ constexpr int n_main(int bn) { return bn <= 128 ? 2 : 1; }
__host__ __device__ constexpr int tmem_cols(int bn)
{
    int need = (n_main(bn) + 1) * bn, c = 32;
    while (c < need) c <<= 1;
    return c;
}

// K-major, SWIZZLE_128B, 8-row groups 1024 B apart (SBO), LBO unused (=1), descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr)
{
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// byte offset of 16-byte chunk c (0..7) of row r inside a [rows x 128 B] swizzled tile
__device__ __forceinline__ uint32_t swz(int r, int c) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)); }

__device__ __forceinline__ void split_store(uint8_t *hi_tile, uint8_t *lo_tile, uint32_t off, float4 v)
{
    float4 h, l;
    h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); l.x = v.x - h.x;
    h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); l.y = v.y - h.y;
    h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); l.z = v.z - h.z;
    h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); l.w = v.w - h.w;
    *reinterpret_cast<float4 *>(hi_tile + off) = h;
    *reinterpret_cast<float4 *>(lo_tile + off) = l;
}

struct TcArgs {
    const float *x, *w, *bias, *scale, *shift, *residual;
    const int32_t *nbr;
    float *stats, *y;
    long long m_out;
    int cin, K, cout, relu;
};

template <int BN>
__global__ void __launch_bounds__(NTHREADS, ctas_per_sm(BN)) gather_gemm_tc_kernel(TcArgs a)
{
    constexpr int STAGES = stages_for(BN);
    constexpr int NMAIN = n_main(BN);
    constexpr int A_BYTES = BM * BK * 4, B_BYTES = BN * BK * 4, STAGE = stage_bytes(BN);
    constexpr int OUT_LD = BN + 4;   // padded staging row (floats): conflict-free 16-byte stores
    static_assert(BM * OUT_LD * 4 <= STAGES * STAGE, "epilogue staging must fit in the pipeline buffers");
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

    extern __shared__ uint8_t smem_raw[];
    uint8_t *tiles = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    int32_t *nbr_s = reinterpret_cast<int32_t *>(tiles + STAGES * STAGE);          // [K][BM]
    uint64_t *bars = reinterpret_cast<uint64_t *>(nbr_s + a.K * BM);              // full[S], empty[S], accum
    uint32_t *misc = reinterpret_cast<uint32_t *>(bars + 2 * STAGES + 1);         // [0] tmem base, [1] tap mask, [2..] active tap list
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES), accum_bar = smem_u32(bars + 2 * STAGES);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long row0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;                    // output-channel tile (cout > 256 is split over grid.y)

    if (tid == 0) misc[1] = 0u;
    if (warp == NPW) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, NPW); mbar_init(empty0 + 8 * s, 1); }
            mbar_init(accum_bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(misc)), "r"(tmem_cols(BN)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    __syncthreads();
    if (tid < BM) {   // neighbour tile -> smem, and the set of taps that have any work in this tile
        const long long row = row0 + tid;
        uint32_t mine = 0u;
        for (int k = 0; k < a.K; ++k) {
            int32_t idx = row < a.m_out ? __ldg(a.nbr + row * a.K + k) : -1;
            nbr_s[k * BM + tid] = idx;
            if (idx >= 0) mine |= 1u << k;
        }
        mine = __reduce_or_sync(0xffffffffu, mine);
        if (lane == 0 && mine) atomicOr(&misc[1], mine);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = misc[0];
    const uint32_t tap_mask = misc[1];
    const int kblocks = (a.cin + BK - 1) / BK;
    const int n_iters = __popc(tap_mask) * kblocks;
    if (tid < a.K && ((tap_mask >> tid) & 1u)) misc[2 + __popc(tap_mask & ((1u << tid) - 1u))] = (uint32_t)tid;   // compact tap list
    asm volatile("bar.sync 2, %0;" ::"n"(NTHREADS) : "memory");

    if (warp < NPW) {
        // ================= producers =================
        // Software-pipelined: the gather loads of k-block it+1 are in flight while k-block it is
        // split and stored (register double buffering).
        constexpr int A_V = BM / RSTEP, B_V = (BN + RSTEP - 1) / RSTEP;   // 16-byte chunks per thread per k-block
        const int c = tid & 7, r_base = tid >> 3;   // 16-byte chunk, first row (rows r_base + RSTEP j)
        const bool b_own = BN >= RSTEP || r_base < BN;
        uint32_t soff[A_V];                         // swizzled byte offsets of this thread's chunks (loop invariant;
#pragma unroll                                      //  the B tile reuses them, 128 rows = 16 KB apart)
        for (int j = 0; j < A_V; ++j) soff[j] = swz(r_base + RSTEP * j, c);
        auto load = [&](int it, float4(&av)[A_V], float4(&bv)[B_V]) {
            const int k = (int)misc[2 + it / kblocks];               // it/kblocks-th active tap
            const int col = (it % kblocks) * BK + c * 4;
            const bool col_ok = col < a.cin;
#pragma unroll
            for (int j = 0; j < A_V; ++j) {
                const int32_t idx = nbr_s[k * BM + r_base + RSTEP * j];
                av[j] = (idx >= 0 && col_ok) ? __ldg(reinterpret_cast<const float4 *>(a.x + (size_t)idx * a.cin + col))
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int j = 0; j < B_V; ++j) {
                const int n = n0 + r_base + RSTEP * j;
                bv[j] = (col_ok && b_own) ? __ldg(reinterpret_cast<const float4 *>(a.w + ((size_t)n * a.K + k) * a.cin + col))
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        auto store = [&](int it, const float4(&av)[A_V], const float4(&bv)[B_V]) {
            const int s = it % STAGES;
            mbar_wait(empty0 + 8 * s, ((it / STAGES) & 1) ^ 1);
            uint8_t *st = tiles + s * STAGE;
#pragma unroll
            for (int j = 0; j < A_V; ++j) split_store(st, st + A_BYTES, soff[j], av[j]);
            if (b_own) {
#pragma unroll
                for (int j = 0; j < B_V; ++j)      // rows r_base + RSTEP j; 128 rows = 16 KB of swizzled tile
                    split_store(st + 2 * A_BYTES, st + 2 * A_BYTES + B_BYTES, soff[j % A_V] + (uint32_t)(j / A_V) * 16384u, bv[j]);
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(full0 + 8 * s);      // one arrival per producer warp
        };
        if (n_iters > 0) {
            float4 a0[A_V], b0[B_V], a1[A_V], b1[B_V];
            load(0, a0, b0);
            for (int it = 0; it < n_iters; it += 2) {
                if (it + 1 < n_iters) load(it + 1, a1, b1);
                store(it, a0, b0);
                if (it + 1 < n_iters) {
                    if (it + 2 < n_iters) load(it + 2, a0, b0);
                    store(it + 1, a1, b1);
                }
            }
        }
        // ================= epilogue =================
        float *stage_out = reinterpret_cast<float *>(tiles);
        if (n_iters > 0) {
            mbar_wait(accum_bar, 0);
            tc_fence_after();
        }
        // warp w may read TMEM lanes 32*(w%4)..+31; the two warps sharing a lane quarter split the columns
        const int my_row = (warp & 3) * 32 + lane;
        const int n_acc = n_iters == 0 ? 0 : (n_iters < NMAIN ? n_iters : NMAIN) + 1;   // mains in use + correction
        constexpr int COLS_PER_GROUP = BN / (NPW / 4) >= 16 ? BN / (NPW / 4) : 16;
        const int c_begin = (warp >> 2) * COLS_PER_GROUP;
#pragma unroll
        for (int cc = 0; cc < COLS_PER_GROUP; cc += 16) {
            const int c0 = c_begin + cc;
            if (c0 >= BN) break;
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = 0.f;
            for (int acc = 0; acc < n_acc; ++acc) {
                const int slot = acc == n_acc - 1 ? NMAIN : acc;                              // last one read = correction
                const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(slot * BN + c0);
                uint32_t u[16];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                    : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                      "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] += __uint_as_float(u[j]);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float4 o;
                o.x = v[4 * q + 0]; o.y = v[4 * q + 1];
                o.z = v[4 * q + 2]; o.w = v[4 * q + 3];
                if (a.bias) {
                    const float4 b = __ldg(reinterpret_cast<const float4 *>(a.bias + n0 + c0 + 4 * q));
                    o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
                }
                *reinterpret_cast<float4 *>(stage_out + my_row * OUT_LD + c0 + 4 * q) = o;
            }
        }
        tc_fence_before();
        asm volatile("bar.sync 1, %0;" ::"n"(NPROD) : "memory");
        if (a.stats && tid < BN) {   // training-mode BatchNorm statistics of the pre-affine output
            float s = 0.f, q = 0.f;
            const int rows = (int)min((long long)BM, a.m_out - row0);
            for (int r = 0; r < rows; ++r) { float t = stage_out[r * OUT_LD + tid]; s += t; q += t * t; }
            atomicAdd(a.stats + n0 + tid, s);
            atomicAdd(a.stats + a.cout + n0 + tid, q);
        }
        if (BN > NPROD && a.stats && tid + NPROD < BN) {
            float s = 0.f, q = 0.f;
            const int rows = (int)min((long long)BM, a.m_out - row0);
            for (int r = 0; r < rows; ++r) { float t = stage_out[r * OUT_LD + tid + NPROD]; s += t; q += t * t; }
            atomicAdd(a.stats + n0 + tid + NPROD, s);
            atomicAdd(a.stats + a.cout + n0 + tid + NPROD, q);
        }
        constexpr int V_PER_ROW = BN / 4;
        for (int t = tid; t < BM * V_PER_ROW; t += NPROD) {
            const int r = t / V_PER_ROW, cl = (t % V_PER_ROW) * 4, cv = n0 + cl;
            const long long row = row0 + r;
            if (row >= a.m_out) break;
            float4 o = *reinterpret_cast<const float4 *>(stage_out + r * OUT_LD + cl);
            if (a.scale) {
                const float4 sc = __ldg(reinterpret_cast<const float4 *>(a.scale + cv));
                const float4 sh = __ldg(reinterpret_cast<const float4 *>(a.shift + cv));
                o.x = fmaf(o.x, sc.x, sh.x); o.y = fmaf(o.y, sc.y, sh.y); o.z = fmaf(o.z, sc.z, sh.z); o.w = fmaf(o.w, sc.w, sh.w);
            }
            if (a.residual) {
                const float4 rs = __ldg(reinterpret_cast<const float4 *>(a.residual + row * a.cout + cv));
                o.x += rs.x; o.y += rs.y; o.z += rs.z; o.w += rs.w;
            }
            if (a.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            *reinterpret_cast<float4 *>(a.y + row * a.cout + cv) = o;
        }
    } else {
        // ================= MMA issuer (warp NPW) =================
        for (int it = 0; it < n_iters; ++it) {
            const int s = it % STAGES;
            mbar_wait(full0 + 8 * s, (it / STAGES) & 1);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t st = smem_u32(tiles + s * STAGE);
                const uint64_t a_hi = make_desc(st), a_lo = make_desc(st + A_BYTES);
                const uint64_t b_hi = make_desc(st + 2 * A_BYTES), b_lo = make_desc(st + 2 * A_BYTES + B_BYTES);
                const int kb = it % kblocks;
                const int k8n = min(BK / 8, (a.cin - kb * BK + 7) / 8);
                const uint32_t d_main = tmem_base + (uint32_t)((it % NMAIN) * BN), d_corr = tmem_base + (uint32_t)(NMAIN * BN);
                for (int k8 = 0; k8 < k8n; ++k8) {
                    const uint64_t adv = (uint64_t)((k8 * 32) >> 4);   // +32 bytes along K inside the swizzle row
                    umma_tf32(d_main, a_hi + adv, b_hi + adv, IDESC, (it >= NMAIN || k8) ? 1u : 0u);
                    umma_tf32(d_corr, a_lo + adv, b_hi + adv, IDESC, (it | k8) ? 1u : 0u);
                    umma_tf32(d_corr, a_hi + adv, b_lo + adv, IDESC, 1u);
                }
                umma_commit(empty0 + 8 * s);            // frees the stage once these MMAs have read it
                if (it == n_iters - 1) umma_commit(accum_bar);
            }
            __syncwarp();
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == NPW) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols(BN)));
    }
}

template <int BN>
size_t smem_bytes(int K) { return 1024 + (size_t)stages_for(BN) * stage_bytes(BN) + (size_t)K * BM * 4 + (2 * stages_for(BN) + 1) * 8 + (2 + MAX_TAPS) * 4; }

template <int BN>
int32_t launch_tc(const TcArgs &a, cudaStream_t stream)
{
    static bool configured = false;
    const size_t smem = smem_bytes<BN>(MAX_TAPS);
    if (!configured) {
        CPD_CUDA(cudaFuncSetAttribute(gather_gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    dim3 grid((unsigned)div_up(a.m_out, BM), (unsigned)(a.cout / BN));
    gather_gemm_tc_kernel<BN><<<grid, NTHREADS, smem_bytes<BN>(a.K), stream>>>(a);
    count_launch();
    return launch_status("cpd_gather_gemm[tcgen05]");
}

}  // namespace

bool gather_gemm_tc_supported(int32_t cin, int32_t K, int32_t cout)
{
    return cin % 8 == 0 && cin >= 8 && K <= MAX_TAPS &&
           (cout == 16 || cout == 32 || cout == 64 || cout == 128 || (cout >= 256 && cout % 256 == 0 && cout <= 2048));
}

size_t gather_gemm_tc_workspace(int64_t, int32_t, int32_t, int32_t) { return 16; }

int32_t gather_gemm_tc(const float *x, int64_t, int32_t cin, const float *w, int32_t K, int32_t cout, const int32_t *nbr,
                       int64_t m_out, const float *bias, const float *scale, const float *shift, const float *residual,
                       int32_t relu, float *stats, float *y, void *, size_t, cudaStream_t stream)
{
    CPD_REQUIRE(gather_gemm_tc_supported(cin, K, cout), CPD_ERR_UNSUPPORTED, "tcgen05 gather-GEMM: unsupported shape");
    CPD_REQUIRE((((uintptr_t)x | (uintptr_t)w | (uintptr_t)y | (uintptr_t)bias | (uintptr_t)scale | (uintptr_t)shift |
                  (uintptr_t)residual) & 15) == 0, CPD_ERR_MISALIGNED, "tcgen05 gather-GEMM: pointers must be 16-byte aligned");
    TcArgs a{x, w, bias, scale, shift, residual, nbr, stats, y, m_out, cin, K, cout, relu};
    switch (cout) {
        case 16: return launch_tc<16>(a, stream);
        case 32: return launch_tc<32>(a, stream);
        case 64: return launch_tc<64>(a, stream);
        case 128: return launch_tc<128>(a, stream);
        default: return launch_tc<256>(a, stream);
    }
}

}  // namespace cpd
