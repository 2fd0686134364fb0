// tcgen05 weight-gradient for sm_100a, PER-TAP PAIR-LIST variant (used for C_in >= 64; the row-stationary
// variant in wgrad_tc.cu wins for narrower layers):  dW[co, k, ci] = sum_o dY[o, co] * X[nbr[o, k], ci]
// (spconv Appendix A.5 `dW_k = sum_pairs dY[o] x[i]^T`; also the wgrad of the dense BEV convs).
//
// Per tap k this is a GEMM  D_k[cout x cin] = dY^T [cout x P_k] . Xg_k [P_k x cin]  whose
// contraction runs over the PAIRS of the tap.  Both operands are "MN-major" for the tensor
// core (rows of dY / X are contiguous along cout / cin, i.e. along M / N), which tcgen05
// supports for tf32 through the MN-major SWIZZLE_128B_BASE32B canonical layout -- no transposes.
//
// One CTA = (tap k, slice of the output rows, <=128-wide cout tile (AM), <=256-wide cin tile (BN)):
//   warps 0-3  scan their slice of the (tap-major) neighbour table 1024 rows at a time, COMPACT
//              the rows that have a neighbour at this tap into a pair list (ballot + prefix),
//              then for every 32 pairs gather the dY rows (A) and the X rows (B), split into
//              tf32 hi/lo and store them in the MN-major swizzled layout.  NB k-blocks are
//              loaded per batch and batches are register double-buffered, so up to 2*NB
//              k-blocks of gathers are in flight per CTA (the narrow layers are latency bound);
//   warp 4     issues 4 x 3 tcgen05.mma kind::tf32 (M=128, N=BN, K=8) per 32-pair block
//              into separate main / correction TMEM accumulators (see spconv_tc.cu);
//   warps 0-3  finally tcgen05.ld the tile and add it to dW with fp32 atomics (split-K over
//              row slices; dW is zeroed by the host wrapper first).
// The MMA M is fixed at 128.  For cout tiles narrower than 128 only AM = 32/64 columns of the
// A tile exist in shared memory: the descriptor's 32-column blocks beyond AM alias whatever
// follows in shared memory, which only pollutes accumulator lanes >= AM that are never read.
#include "tc_common.cuh"

namespace cpd {
namespace {
using namespace tc;

constexpr int KB = 32;          // pairs per k-block (4 MMA K-steps of 8)
constexpr int WIN = 1024;       // table rows scanned per compaction window
constexpr int NPROD = 128;
constexpr int RPT = WIN / NPROD;
constexpr int NTHREADS = 160;
constexpr uint32_t END_MARK = 0xffffffffu;
constexpr int SLACK = 4 * 1024;    // max over-read past the stage ring by the aliased A blocks (AM=32, BN=32, A_lo tile)

__host__ __device__ constexpr int w_stage_bytes(int bn, int am) { return 2 * KB * am * 4 + 2 * KB * bn * 4; }
__host__ __device__ constexpr int w_nb(int bn, int am)          // k-blocks per producer batch (<= 16 float4 per thread)
{
    return (am + bn) / 16 <= 4 ? 4 : (am + bn) / 16 <= 8 ? 2 : 1;
}
__host__ __device__ constexpr int w_stages(int bn, int am)
{
    const int sb = w_stage_bytes(bn, am), nb = w_nb(bn, am);
    if (sb <= 32 * 1024) {                       // small stages: ring <= 96 KB so that two CTAs share an SM
        int s = (96 * 1024) / sb, lo = nb + (nb >= 4 ? 2 : 1);
        return s > 6 ? 6 : (s < lo ? lo : s);
    }
    int s = (190 * 1024) / sb;                   // one CTA per SM
    return s > 3 ? 3 : (s < 2 ? 2 : s);
}
__host__ __device__ constexpr int w_ctas_per_sm(int bn, int am) { return w_stages(bn, am) * w_stage_bytes(bn, am) <= 100 * 1024 ? 2 : 1; }
__host__ __device__ constexpr int w_nmain(int bn) { return bn <= 128 ? 2 : 1; }
__host__ __device__ constexpr int w_tmem_cols(int bn)
{
    int need = (w_nmain(bn) + 1) * bn, c = 32;
    while (c < need) c <<= 1;
    return c;
}

// MN-major tf32 operands have exactly one legal shared-memory layout: SWIZZLE_128B_BASE32B
// (descriptor layout type 1; CUTLASS: "for mn-major tf32 operands, SW128_32B is the only available
// smem layout").  Atom = 4 K-rows x 128 B (32 consecutive M/N elements per row); the 32-byte chunk
// index (address bits 5-6) is XOR-ed with the row index inside the atom (address bits 7-8).
// Tile = [32 pair rows x cols]: 32-column blocks LBO = 4096 B apart, 4-row atoms SBO = 512 B apart.
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr)
{
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(4096 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) |
           (1ull << 46) | (1ull << 61);
}
// byte offset of the 16-byte chunk holding columns [4*c4, 4*c4+4) of row r in a [32 rows x cols] MN-major tile
__device__ __forceinline__ uint32_t swz_mn(int r, int c4)
{
    const int c32 = (c4 & 7) >> 1;                       // 32-byte chunk inside the 128-byte row
    return (uint32_t)((c4 >> 3) * 4096 + r * 128 + ((c32 ^ (r & 3)) << 5) + ((c4 & 1) << 4));
}

struct WgArgs {
    const float *x, *dy;
    const int32_t *nbr;
    float *dw;
    long long m_out;
    int cin, cout, K, rows_per_cta, ci_tiles, tap_major;
};

template <int BN, int AM>
__global__ void __launch_bounds__(NTHREADS, w_ctas_per_sm(BN, AM)) gather_wgrad_pairs_kernel(WgArgs a)
{
    constexpr int STAGES = w_stages(BN, AM), STAGE = w_stage_bytes(BN, AM), NMAIN = w_nmain(BN), NB = w_nb(BN, AM);
    constexpr int A_BYTES = KB * AM * 4, B_BYTES = KB * BN * 4;
    constexpr int A_V = (KB * AM / 4) / NPROD, B_V = (KB * BN / 4) / NPROD;   // float4 per thread per k-block
    // a_major = b_major = MN (bits 15, 16), fp32 accumulate, tf32 operands, M = 128
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(BN >> 3) << 17) |
                               ((uint32_t)(128 >> 4) << 24);

    extern __shared__ uint8_t smem_raw[];
    uint8_t *tiles = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    int32_t *pair_o = reinterpret_cast<int32_t *>(tiles + STAGES * STAGE);      // [WIN]   (also the over-read slack)
    int32_t *pair_i = pair_o + WIN;                                             // [WIN]
    uint64_t *bars = reinterpret_cast<uint64_t *>(pair_i + WIN);                // full[S], empty[S], accum
    uint32_t *info = reinterpret_cast<uint32_t *>(bars + 2 * STAGES + 1);       // [S] k8 steps valid in the stage / END
    uint32_t *misc = info + STAGES;                                             // [0] tmem base, [1..64] per-warp pair counts
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES), accum_bar = smem_u32(bars + 2 * STAGES);
    static_assert(2 * WIN * 4 >= SLACK, "pair lists double as the over-read slack");

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tap = blockIdx.x;
    const int co0 = (blockIdx.z / a.ci_tiles) * 128, ci0 = (blockIdx.z % a.ci_tiles) * BN;
    const long long r_begin = (long long)blockIdx.y * a.rows_per_cta;
    const long long r_end = min(r_begin + (long long)a.rows_per_cta, a.m_out);

    if (warp == 4) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, NPROD); mbar_init(empty0 + 8 * s, 1); }
            mbar_init(accum_bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        tmem_alloc(smem_u32(misc), w_tmem_cols(BN));
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = misc[0];

    if (warp < 4) {
        // Zero all operand stages once: columns beyond cout / cin inside the AM / BN tiles are never
        // written again (their loads and stores are skipped below).
        for (int e = tid; e < STAGES * STAGE / 16; e += NPROD) reinterpret_cast<float4 *>(tiles)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        asm volatile("bar.sync 1, 128;" ::: "memory");
        // loop-invariant element -> (row, 16-byte chunk) mapping: e = tid + 128 j over the 32 x (cols/4) chunks of a tile
        uint32_t a_off[A_V], b_off[B_V];
        int a_row[A_V], b_row[B_V], a_col[A_V], b_col[B_V];
        bool a_live[A_V], b_live[B_V];
#pragma unroll
        for (int j = 0; j < A_V; ++j) {
            const int e = tid + NPROD * j;
            a_row[j] = e / (AM / 4); a_col[j] = co0 + (e % (AM / 4)) * 4;
            a_off[j] = swz_mn(a_row[j], e % (AM / 4)); a_live[j] = a_col[j] < a.cout;
        }
#pragma unroll
        for (int j = 0; j < B_V; ++j) {
            const int e = tid + NPROD * j;
            b_row[j] = e / (BN / 4); b_col[j] = ci0 + (e % (BN / 4)) * 4;
            b_off[j] = swz_mn(b_row[j], e % (BN / 4)); b_live[j] = b_col[j] < a.cin;
        }
        // ================= producers =================
        int it = 0;   // k-blocks produced so far
        auto load = [&](int b0, int total, float4(&av)[A_V], float4(&bv)[B_V]) {
            const int nvalid = min(KB, total - b0);
#pragma unroll
            for (int j = 0; j < A_V; ++j)          // A: dY rows of the pairs
                av[j] = (a_live[j] && a_row[j] < nvalid)
                            ? __ldg(reinterpret_cast<const float4 *>(a.dy + (r_begin + pair_o[b0 + a_row[j]]) * a.cout + a_col[j]))
                            : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < B_V; ++j)          // B: gathered X rows of the pairs
                bv[j] = (b_live[j] && b_row[j] < nvalid)
                            ? __ldg(reinterpret_cast<const float4 *>(a.x + (size_t)pair_i[b0 + b_row[j]] * a.cin + b_col[j]))
                            : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        auto store = [&](int b0, int total, const float4(&av)[A_V], const float4(&bv)[B_V]) {
            const int s = it % STAGES;
            mbar_wait(empty0 + 8 * s, ((it / STAGES) & 1) ^ 1);
            uint8_t *st = tiles + s * STAGE;
#pragma unroll
            for (int j = 0; j < A_V; ++j) {
                if (!a_live[j]) continue;
                float4 h, l;
                split4(av[j], h, l);
                *reinterpret_cast<float4 *>(st + a_off[j]) = h;
                *reinterpret_cast<float4 *>(st + A_BYTES + a_off[j]) = l;
            }
#pragma unroll
            for (int j = 0; j < B_V; ++j) {
                if (!b_live[j]) continue;
                float4 h, l;
                split4(bv[j], h, l);
                *reinterpret_cast<float4 *>(st + 2 * A_BYTES + b_off[j]) = h;
                *reinterpret_cast<float4 *>(st + 2 * A_BYTES + B_BYTES + b_off[j]) = l;
            }
            if (tid == 0) info[s] = (uint32_t)((min(KB, total - b0) + 7) / 8);
            fence_async_smem();
            mbar_arrive(full0 + 8 * s);
            ++it;
        };
        // batch = NB consecutive k-blocks starting at pair b0
        auto load_batch = [&](int b0, int total, float4(&av)[NB][A_V], float4(&bv)[NB][B_V]) {
#pragma unroll
            for (int q = 0; q < NB; ++q)
                if (b0 + q * KB < total) load(b0 + q * KB, total, av[q], bv[q]);
        };
        auto store_batch = [&](int b0, int total, const float4(&av)[NB][A_V], const float4(&bv)[NB][B_V]) {
#pragma unroll
            for (int q = 0; q < NB; ++q)
                if (b0 + q * KB < total) store(b0 + q * KB, total, av[q], bv[q]);
        };
        for (long long w0 = r_begin; w0 < r_end; w0 += WIN) {
            // ---- compact the active (row, neighbour) pairs of this window ----
            int32_t idx[RPT];
            unsigned bal[RPT];
#pragma unroll
            for (int h = 0; h < RPT; ++h) {
                const long long o = w0 + h * NPROD + tid;
                idx[h] = o < r_end ? __ldg(a.nbr + (a.tap_major ? (long long)tap * a.m_out + o : o * a.K + tap)) : -1;
            }
#pragma unroll
            for (int h = 0; h < RPT; ++h) bal[h] = __ballot_sync(0xffffffffu, idx[h] >= 0);
            asm volatile("bar.sync 1, 128;" ::: "memory");                 // previous window's list fully consumed
            if (lane == 0) {
#pragma unroll
                for (int h = 0; h < RPT; ++h) misc[1 + h * 4 + warp] = __popc(bal[h]);
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            int total = 0;
#pragma unroll
            for (int h = 0; h < RPT; ++h) {
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const int cnt = misc[1 + h * 4 + w];
                    if (w == warp && idx[h] >= 0) {
                        const int p = total + __popc(bal[h] & ((1u << lane) - 1u));
                        pair_o[p] = (int32_t)(w0 + h * NPROD + tid - r_begin); pair_i[p] = idx[h];
                    }
                    total += cnt;
                }
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            // ---- NB k-blocks per batch, batches register double-buffered ----
            constexpr int STEP = NB * KB;
            if constexpr ((A_V + B_V) * NB <= 16) {
                if (total > 0) {
                    float4 a0[NB][A_V], b0v[NB][B_V], a1[NB][A_V], b1v[NB][B_V];
                    load_batch(0, total, a0, b0v);
                    for (int b0 = 0; b0 < total; b0 += 2 * STEP) {
                        if (b0 + STEP < total) load_batch(b0 + STEP, total, a1, b1v);
                        store_batch(b0, total, a0, b0v);
                        if (b0 + STEP < total) {
                            if (b0 + 2 * STEP < total) load_batch(b0 + 2 * STEP, total, a0, b0v);
                            store_batch(b0 + STEP, total, a1, b1v);
                        }
                    }
                }
            } else {
                for (int b0 = 0; b0 < total; b0 += STEP) {
                    float4 av[NB][A_V], bv[NB][B_V];
                    load_batch(b0, total, av, bv);
                    store_batch(b0, total, av, bv);
                }
            }
        }
        // ---- end marker, then epilogue ----
        {
            const int s = it % STAGES;
            mbar_wait(empty0 + 8 * s, ((it / STAGES) & 1) ^ 1);
            if (tid == 0) info[s] = END_MARK;
            mbar_arrive(full0 + 8 * s);
        }
        if (it > 0) {
            mbar_wait(accum_bar, 0);
            tc_fence_after();
            const int n_acc = (it < NMAIN ? it : NMAIN) + 1;
            const int co = co0 + warp * 32 + lane;
            if (warp * 32 < AM) {                       // accumulator lanes >= AM hold aliased garbage
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 16) {
                    float v[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = 0.f;
                    for (int acc = 0; acc < n_acc; ++acc) {
                        const int slot = acc == n_acc - 1 ? NMAIN : acc;
                        uint32_t u[16];
                        tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(slot * BN + c0), u);
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] += __uint_as_float(u[j]);
                    }
                    if (co < a.cout) {
                        float *dst = a.dw + ((size_t)co * a.K + tap) * a.cin + ci0 + c0;
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (ci0 + c0 + j < a.cin && v[j] != 0.f) atomicAdd(dst + j, v[j]);
                    }
                }
            }
        }
        tc_fence_before();
    } else {
        // ================= MMA issuer =================
        int it = 0;
        for (;; ++it) {
            const int s = it % STAGES;
            mbar_wait(full0 + 8 * s, (it / STAGES) & 1);
            const uint32_t k8n = *reinterpret_cast<volatile uint32_t *>(&info[s]);
            if (k8n == END_MARK) break;
            tc_fence_after();
            if (lane == 0) {
                const uint32_t st = smem_u32(tiles + s * STAGE);
                const uint64_t a_hi = make_desc_mn(st), a_lo = make_desc_mn(st + A_BYTES);
                const uint64_t b_hi = make_desc_mn(st + 2 * A_BYTES), b_lo = make_desc_mn(st + 2 * A_BYTES + B_BYTES);
                const uint32_t d_main = tmem_base + (uint32_t)((it % NMAIN) * BN), d_corr = tmem_base + (uint32_t)(NMAIN * BN);
                for (uint32_t k8 = 0; k8 < k8n; ++k8) {
                    const uint64_t adv = (uint64_t)((k8 * 1024) >> 4);     // next 8-row K atom
                    umma_tf32(d_main, a_hi + adv, b_hi + adv, IDESC, (it >= NMAIN || k8) ? 1u : 0u);
                    umma_tf32(d_corr, a_lo + adv, b_hi + adv, IDESC, (it | (int)k8) ? 1u : 0u);
                    umma_tf32(d_corr, a_hi + adv, b_lo + adv, IDESC, 1u);
                }
                umma_commit(empty0 + 8 * s);
            }
            __syncwarp();
        }
        if (it > 0 && lane == 0) umma_commit(accum_bar);
        __syncwarp();
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, w_tmem_cols(BN));
    }
}

template <int BN, int AM>
constexpr size_t wg_smem()
{
    return 1024 + (size_t)w_stages(BN, AM) * w_stage_bytes(BN, AM) + 2 * WIN * 4 + (2 * w_stages(BN, AM) + 1) * 8 +
           (w_stages(BN, AM) + 1 + 4 * RPT + 8) * 4;
}

template <int BN, int AM>
int32_t launch_wg(const WgArgs &a, dim3 grid, cudaStream_t stream)
{
    static bool configured = false;
    if (!configured) {
        CPD_CUDA(cudaFuncSetAttribute(gather_wgrad_pairs_kernel<BN, AM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wg_smem<BN, AM>()));
        configured = true;
    }
    gather_wgrad_pairs_kernel<BN, AM><<<grid, NTHREADS, wg_smem<BN, AM>(), stream>>>(a);
    count_launch();
    return launch_status("cpd_gather_wgrad[tcgen05]");
}

template <int BN>
int32_t launch_wg_am(int am, const WgArgs &a, dim3 grid, cudaStream_t stream)
{
    if (am <= 32) return launch_wg<BN, 32>(a, grid, stream);
    if (am <= 64) return launch_wg<BN, 64>(a, grid, stream);
    return launch_wg<BN, 128>(a, grid, stream);
}

}  // namespace

bool gather_wgrad_pairs_supported(int32_t cin, int32_t K, int32_t cout)
{
    return cin % 4 == 0 && cout % 4 == 0 && cin >= 8 && cout >= 8 && K <= 64;
}

// dw must be zeroed by the caller (split-K atomics).
int32_t gather_wgrad_pairs_tc(const float *x, int32_t cin, const float *dy, int64_t m_out, int32_t cout, const int32_t *nbr,
                        int32_t K, int32_t tap_major, float *dw, cudaStream_t stream)
{
    CPD_REQUIRE(gather_wgrad_pairs_supported(cin, K, cout), CPD_ERR_UNSUPPORTED, "tcgen05 wgrad: unsupported shape");
    CPD_REQUIRE((((uintptr_t)x | (uintptr_t)dy) & 15) == 0, CPD_ERR_MISALIGNED, "tcgen05 wgrad: pointers must be 16-byte aligned");
    const int bn = cin <= 32 ? 32 : cin <= 64 ? 64 : cin <= 128 ? 128 : 256;
    const int am = cout <= 32 ? 32 : cout <= 64 ? 64 : 128;
    const int ci_tiles = (int)div_up(cin, bn), co_tiles = (int)div_up(cout, 128);
    int S = (int)div_up(148 * 2, (long long)K * ci_tiles * co_tiles);
    const int max_s = (int)div_up(m_out, 2 * WIN);
    if (S > max_s) S = max_s;
    if (S < 1) S = 1;
    int rows = (int)(div_up(div_up(m_out, S), WIN) * WIN);
    S = (int)div_up(m_out, rows);
    WgArgs a{x, dy, nbr, dw, m_out, cin, cout, K, rows, ci_tiles, tap_major};
    dim3 grid(K, S, ci_tiles * co_tiles);
    switch (bn) {
        case 32: return launch_wg_am<32>(am, a, grid, stream);
        case 64: return launch_wg_am<64>(am, a, grid, stream);
        case 128: return launch_wg_am<128>(am, a, grid, stream);
        default: return launch_wg_am<256>(am, a, grid, stream);
    }
}

}  // namespace cpd
