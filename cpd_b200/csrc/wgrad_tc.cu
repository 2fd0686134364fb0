// tcgen05 weight-gradient for sm_100a:  dW[co, k, ci] = sum_o dY[o, co] * X[nbr[o, k], ci]
// (spconv Appendix A.5 `dW_k = sum_pairs dY[o] x[i]^T`; also the wgrad of the dense BEV convs).
//
// Row-stationary formulation with the taps concatenated along N:
//     D[co, (t, ci)] = sum over output rows o of   dY[o, co] * Xg[o, (t, ci)],   Xg[o,(t,ci)] = X[nbr[o,t], ci] or 0
// i.e. ONE GEMM per tap group with  M = cout tile (MMA M is fixed at 128),  N = T*cin <= 256,  K = rows.
// Compared with a per-tap pair-list GEMM this reads every dY row once per tap group instead of once
// per pair, needs no compaction pass, keeps the contraction loop long and regular, and uses the
// N = 256 MMA shape even for 16/32-channel layers; the price is multiplying the zero rows of missing
// neighbours, which costs tensor time that these gather-bound layers have to spare.
//
// Both operands are "MN-major" for the tensor core (rows of dY / X are contiguous along cout / cin).
// fp32-class accuracy comes from the bf16x3 split (hi = RN(x), lo = RN(x - hi); hi.hi + lo.hi + hi.lo,
// see spconv_tc.cu): half the shared-memory bytes and half the tensor time of a 3xTF32 split.
//
// One CTA = (tap group, slice of the output rows, <=128-wide cout tile AM, BN-wide (tap,cin) tile):
//   warps 0-7  per 32-row k-block: copy the dY rows (A, contiguous) and gather the X rows of each tap
//              of the group (B; neighbour indices come from a shared-memory copy of the tap-major
//              table, prefetched one 128-row window ahead) from the split-row images (split.cu) with
//              16-byte cp.async straight into the swizzled tiles -- no register staging, missing
//              neighbours zero-filled by the copy, STAGES-1 k-blocks in flight per thread;
//   warp 8     issues 2 x 3 tcgen05.mma kind::f16 (K = 16 rows) per k-block into a main and a correction TMEM
//              accumulator (A_hi.B_hi | A_lo.B_hi + A_hi.B_lo, see spconv_tc.cu);
//   warps 0-7  finally tcgen05.ld both accumulators and add them into dW with vector fp32 reductions
//              (split-K over row slices; dW is zeroed by the host wrapper first).
// For cout tiles narrower than 128 only one 64-column block of the A tile exists in shared memory: the
// descriptor's second block aliases whatever follows, which only pollutes accumulator
// lanes >= AM that are never read.
#include <algorithm>
#include <stdlib.h>
#include "tc_common.cuh"

namespace cpd {
namespace {
using namespace tc;

constexpr int KB = 32;           // rows per k-block (2 MMA K-steps of 16)
constexpr int WINR = 128;        // rows per neighbour-table window (4 k-blocks)
constexpr int WSTR = WINR + 1;   // pitch of a tap's row in the shared-memory window: the lanes of a warp read the SAME row of up to
                                 // 16 different taps, a pitch of 128 words put all of them in one bank (ncu: 17.6 M conflicts per launch)
constexpr int NPW = 8;            // producer / epilogue warps
constexpr int NPROD = NPW * 32;
constexpr int NTHREADS = NPROD + 32;
constexpr int MAX_T = 32;        // taps per group (cin = 8 -> 32 taps in N = 256)
constexpr int MAX_ROWS_PER_CTA = 8192;   // = 512 K-steps accumulated in one TMEM accumulator (bounds the truncation bias)

__host__ __device__ constexpr int w_a_bytes(int am) { return KB * (am < 64 ? 64 : am) * 2; }   // whole 64-column blocks
__host__ __device__ constexpr int w_stage_bytes(int bn, int am) { return 2 * w_a_bytes(am) + 2 * KB * bn * 2; }
__host__ __device__ constexpr int w_stages(int bn, int am)
{
    int s = (160 * 1024) / w_stage_bytes(bn, am);
    return s > 4 ? 4 : (s < 2 ? 2 : s);
}
__host__ __device__ constexpr int w_tmem_cols(int bn) { return 2 * bn < 32 ? 32 : 2 * bn; }   // main + correction

// SWIZZLE_128B, MN-major bf16: atom = 8 K-rows x 128 B (64 consecutive M/N elements per row); the 16-byte
// chunk index (address bits 4-6) is XOR-ed with the row index inside the atom (bits 7-9).
// Tile = [32 rows x cols]: 64-column blocks LBO = 4096 B apart, 8-row atoms SBO = 1024 B apart.
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr)
{
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(4096 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
           (1ull << 46) | (2ull << 61);
}
// byte offset of the 16 bytes holding columns [8*c8, 8*c8+8) of row r in a [32 rows x cols] MN-major bf16 tile
__device__ __forceinline__ uint32_t swz_mn(int r, int c8)
{
    return (uint32_t)((c8 >> 3) * 4096 + r * 128 + (((c8 & 7) ^ (r & 7)) << 4));
}

__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

struct WgArgs {
    const uint8_t *xs, *dys;     // split-row images: row i = [hi(c) | lo(c)] bf16 (split.cu)
    const int32_t *nbr_t;        // tap-major table (K, m_out)
    float *dw;
    long long m_out;
    int cin, cout, K, rows_per_cta, ci_tiles, taps_per_group;
    int skip_zero;               // skip the zero-fill copy of slots that are already zero (CPD_WGRAD_SKIP_ZERO, default on)
};

template <int BN, int AM>
__global__ void __launch_bounds__(NTHREADS, 1) gather_wgrad_rows_kernel(WgArgs a)
{
    constexpr int STAGES = w_stages(BN, AM), STAGE = w_stage_bytes(BN, AM);
    constexpr int A_BYTES = w_a_bytes(AM), B_BYTES = KB * BN * 2;
    constexpr bool SWAP = BN == 256 && AM <= 64;     // narrow cout: the gathered tile is the M operand (see the MMA issuer)
    static_assert(!SWAP || A_BYTES == 4096, "swapped roles expect one 64-column dY block per hi / lo tile");
    constexpr int A_C8 = AM / 8, A_V = (KB * A_C8 + NPROD - 1) / NPROD;        // 16-byte chunks per row / per thread per k-block
    constexpr int B_C8 = BN / 8, B_RSTEP = NPROD / B_C8, B_V = KB / B_RSTEP;   // chunks per row; row stride between a thread's chunks
    static_assert(NPROD % B_C8 == 0 && KB % B_RSTEP == 0, "B mapping");
    // a_major = b_major = MN (bits 15, 16), fp32 accumulate, bf16 operands, M = 128
    constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(BN >> 3) << 17) |
                               ((uint32_t)(128 >> 4) << 24);

    extern __shared__ uint8_t smem_raw[];
    uint8_t *tiles = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    int32_t *nbr_w = reinterpret_cast<int32_t *>(tiles + STAGES * STAGE);       // [2][MAX_T][WSTR]  (also the over-read slack)
    uint64_t *bars = reinterpret_cast<uint64_t *>(nbr_w + 2 * MAX_T * WSTR + 2);  // (+2: keeps the barriers 8-byte aligned)    // full[S], empty[S], accum
    uint32_t *info = reinterpret_cast<uint32_t *>(bars + 2 * STAGES + 1);       // [S] k8 steps valid in the stage / END
    uint32_t *misc = info + STAGES;                                             // [0] tmem base
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES), accum_bar = smem_u32(bars + 2 * STAGES);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tap0 = blockIdx.x * a.taps_per_group;
    const int T = min(a.taps_per_group, a.K - tap0);
    const int co0 = (blockIdx.z / a.ci_tiles) * 128, ci0 = (blockIdx.z % a.ci_tiles) * BN;   // ci0 != 0 only when cin > BN (T == 1)
    const long long r_begin = (long long)blockIdx.y * a.rows_per_cta;
    const long long r_end = min(r_begin + (long long)a.rows_per_cta, a.m_out);
    const int n_blocks = (int)((r_end - r_begin + KB - 1) / KB);

    if (warp == NPW) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, NPROD); mbar_init(empty0 + 8 * s, 1); }
            mbar_init(accum_bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        tmem_alloc(smem_u32(misc), w_tmem_cols(BN));
    }
    if (warp < NPW) {
        // Zero all operand stages once: columns beyond cout / (T*cin) are never written again.
        for (int e = tid; e < STAGES * STAGE / 16; e += NPROD) reinterpret_cast<float4 *>(tiles)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        fence_async_smem();                          // generic-proxy zeros -> visible to the tensor core's async-proxy reads
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = misc[0];

    if (warp < NPW) {
        // loop-invariant mapping.  A: e = tid + 256 j over 32 x (AM/8) 16-byte chunks.
        uint32_t a_off[A_V];
        int a_row[A_V], a_col[A_V];
        bool a_live[A_V];
#pragma unroll
        for (int j = 0; j < A_V; ++j) {
            const int e = tid + NPROD * j;
            a_row[j] = e / A_C8; a_col[j] = co0 + (e % A_C8) * 8;
            a_off[j] = swz_mn(a_row[j], e % A_C8); a_live[j] = a_row[j] < KB && a_col[j] < a.cout;
        }
        // B: this thread owns ONE column chunk (fixed tap and channel offset) and rows b_row0 + B_RSTEP*j.
        const int b_c8 = tid % B_C8, b_row0 = tid / B_C8;
        const int b_colv = b_c8 * 8;
        const int b_t = a.cin >= BN ? 0 : b_colv / a.cin;                  // tap inside the group
        const int b_ci = a.cin >= BN ? ci0 + b_colv : b_colv % a.cin;      // input channel of the chunk
        const bool b_live = b_t < T && b_ci < a.cin;
        uint32_t b_off[B_V];
#pragma unroll
        for (int j = 0; j < B_V; ++j) b_off[j] = swz_mn(b_row0 + B_RSTEP * j, b_c8);
        const int my_nbr_off = b_t * WSTR + b_row0;
        const uint32_t tiles_u32 = smem_u32(tiles);
        const size_t dy_row = (size_t)a.cout * 4;                                  // image rows: hi(c) | lo(c)
        const uint32_t x_row32 = (uint32_t)a.cin * 4, nbr_w_u32 = smem_u32(nbr_w);
        uint32_t bstate = 0u;                       // bit (stage * B_V + j): this thread's slot j of the stage holds gathered data
        static_assert(w_stages(BN, AM) * B_V <= 32, "slot-state bits");
        const uint32_t x_lo = (uint32_t)a.cin * 2, dy_lo = (uint32_t)a.cout * 2;

        // ---- neighbour-table windows (T x 128 entries): fetched into registers two windows ahead, published into
        //      a double-buffered shared-memory copy one window ahead of the gathers that read it ----
        constexpr int NW = (MAX_T * WINR) / NPROD;      // table entries per thread per window (upper bound)
        static_assert(NPROD == 2 * WINR, "window mapping: thread -> (tap parity, row), taps advance by 2 per entry");
        const int wt = tid / WINR, wr = tid % WINR;     // entry q of this thread = (tap wt + 2 q, row wr) of the window
        int32_t nreg[NW];
        auto fetch_window = [&](long long w0) {
            const long long o = w0 + wr;
            const bool in = o < r_end;
            const int32_t *p = a.nbr_t + (long long)(tap0 + wt) * a.m_out + o;
            const long long step = 2 * a.m_out;
            // T (taps of the group) is uniform over the CTA: leave the unrolled loop at the group's last tap instead of
            // issuing NW predicated-off loads (T = 8 for 32 channels uses 4 of the 16 entries; this loop was 24 % of the
            // kernel's instructions)
#pragma unroll
            for (int q = 0; q < NW; ++q) {
                if (2 * q >= T) break;
                nreg[q] = (in && wt + 2 * q < T) ? __ldg(p + q * step) : -1;
            }
        };
        auto publish_window = [&](int buf) {
            int32_t *dstw = nbr_w + buf * MAX_T * WSTR + wt * WSTR + wr;
#pragma unroll
            for (int q = 0; q < NW; ++q) {
                if (2 * q >= T) break;
                if (wt + 2 * q < T) dstw[2 * WSTR * q] = nreg[q];
            }
        };
        auto issue = [&](int blk) {
            const int s = blk % STAGES;
            mbar_wait(empty0 + 8 * s, ((blk / STAGES) & 1) ^ 1);
            const long long r0 = r_begin + (long long)blk * KB;
            const int nvalid = (int)min((long long)KB, r_end - r0);
            const uint32_t tab = nbr_w_u32 + (uint32_t)((((blk / (WINR / KB)) & 1) * MAX_T * WSTR + my_nbr_off + (blk % (WINR / KB)) * KB) * 4);
            const uint32_t dst = tiles_u32 + (uint32_t)(s * STAGE);
            // neighbour indices FIRST, back to back (each copy below is a compiler barrier: interleaved with the copies the
            // loads serialise into B_V dependent load -> address -> copy chains per k-block and the warp runs latency-bound)
            int32_t idx[B_V];
            if (b_live) {
#pragma unroll
                for (int j = 0; j < B_V; ++j) idx[j] = lds_i32(tab + (uint32_t)(B_RSTEP * j * 4));   // rows past r_end carry -1 in the table copy
            }
#pragma unroll
            for (int j = 0; j < A_V; ++j) {
                if (!a_live[j]) continue;
                const bool ok = a_row[j] < nvalid;
                const uint8_t *src = a.dys + (size_t)(ok ? r0 + a_row[j] : 0) * dy_row + (size_t)a_col[j] * 2;
                cp_async16(dst + a_off[j], src, ok ? 16u : 0u);
                cp_async16(dst + A_BYTES + a_off[j], src + dy_lo, ok ? 16u : 0u);
            }
            if (b_live) {
                // 40-85 % of the (row, tap) slots have no neighbour.  All stages start zeroed, so a missing neighbour needs a
                // zero-fill copy only when the slot still holds data from the stage's previous use (bstate: one bit per
                // (stage, row of this thread)): the copies -- the LSU / shared-memory wavefronts that bound this kernel --
                // drop from one per slot to one per present slot + one per data -> empty transition.
                const uint8_t *col_hi = a.xs + (size_t)b_ci * 2, *col_lo = col_hi + x_lo;
                const uint32_t held = (bstate >> (s * B_V)) & ((1u << B_V) - 1u);
                uint32_t now = 0u;
#pragma unroll
                for (int j = 0; j < B_V; ++j) {
                    const bool present = idx[j] >= 0;
                    if (present || ((held >> j) & 1u) || !a.skip_zero) {
                        const uint32_t sz = present ? 16u : 0u;
                        const uint64_t off = (uint64_t)(uint32_t)max(idx[j], 0) * x_row32;
                        cp_async16(dst + 2 * A_BYTES + b_off[j], col_hi + off, sz);
                        cp_async16(dst + 2 * A_BYTES + B_BYTES + b_off[j], col_lo + off, sz);
                    }
                    now |= (present ? 1u : 0u) << j;
                }
                bstate = (bstate & ~(((1u << B_V) - 1u) << (s * B_V))) | (now << (s * B_V));
            }
        };
        // before block `blk` is ISSUED its window's table must be published; windows are 4 blocks long
        auto prepare = [&](int blk) {
            if (blk % (WINR / KB) != 0 || blk >= n_blocks) return;
            const int w = blk / (WINR / KB);
            asm volatile("bar.sync 1, %0;" ::"n"(NPROD) : "memory");   // buffer (w & 1) no longer read (window w-2 is long done)
            publish_window(w & 1);
            asm volatile("bar.sync 1, %0;" ::"n"(NPROD) : "memory");
            const long long next = r_begin + (long long)(w + 1) * WINR;
            if (next < r_end) fetch_window(next);
        };
        fetch_window(r_begin);
        // Each thread's copies signal the stage's mbarrier themselves when they land (cp.async.mbarrier.arrive.noinc):
        // the producers never wait for data, only for a free stage.  The generic -> async proxy fence is executed by
        // the MMA warp after it has observed the barrier (a producer-side wait_group + fence.proxy.async hand-over
        // serialises the pipeline: the proxy fence waits for all of the thread's copies in flight).
        for (int blk = 0; blk < n_blocks; ++blk) {
            prepare(blk);
            issue(blk);
            cp_async_arrive_noinc(full0 + 8 * (blk % STAGES));
        }
        cp_async_wait_all();
        // ---- epilogue ----
        if (n_blocks > 0 && SWAP) {
            // swapped roles: TMEM lane = (tap, ci) column of half h = warp / 4, accumulator columns = cout
            mbar_wait(accum_bar, 0);
            tc_fence_after();
            const int h = warp >> 2, col = h * 128 + (warp & 3) * 32 + lane;
            const int t = a.cin >= BN ? 0 : col / a.cin;
            const int ci = a.cin >= BN ? ci0 + col : col % a.cin;
            const bool live = t < T && ci < a.cin;
            const uint32_t tacc = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(h * 128);
#pragma unroll 1
            for (int c0 = 0; c0 < AM; c0 += 16) {
                uint32_t u[16], v[16];
                tmem_ld16(tacc + (uint32_t)c0, u);            // X_hi . dY_hi
                tmem_ld16(tacc + (uint32_t)(64 + c0), v);     // X_hi . dY_lo + X_lo . dY_hi
                if (live) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int co = co0 + c0 + j;          // consecutive lanes -> consecutive ci: coalesced reductions
                        if (co < a.cout)
                            atomicAdd(a.dw + ((size_t)co * a.K + tap0 + t) * a.cin + ci, __uint_as_float(u[j]) + __uint_as_float(v[j]));
                    }
                }
            }
        } else if (n_blocks > 0) {
            mbar_wait(accum_bar, 0);
            tc_fence_after();
            const int co = co0 + (warp & 3) * 32 + lane;      // warps w and w+4 share TMEM lanes and split the columns
            if ((warp & 3) * 32 < AM) {                       // accumulator lanes >= AM hold aliased garbage
#pragma unroll 1
                for (int c0 = (warp >> 2) * (BN / 2); c0 < ((warp >> 2) + 1) * (BN / 2); c0 += 16) {
                    uint32_t u[16], v[16];
                    tmem_ld16(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c0, u);
                    tmem_ld16(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(BN + c0), v);
                    if (co < a.cout) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int col = c0 + 4 * q;
                            const int t = a.cin >= BN ? 0 : col / a.cin;
                            const int ci = a.cin >= BN ? ci0 + col : col % a.cin;
                            if (t < T && ci < a.cin)
                                red_add_v4(a.dw + ((size_t)co * a.K + tap0 + t) * a.cin + ci,
                                           __uint_as_float(u[4 * q]) + __uint_as_float(v[4 * q]),
                                           __uint_as_float(u[4 * q + 1]) + __uint_as_float(v[4 * q + 1]),
                                           __uint_as_float(u[4 * q + 2]) + __uint_as_float(v[4 * q + 2]),
                                           __uint_as_float(u[4 * q + 3]) + __uint_as_float(v[4 * q + 3]));
                        }
                    }
                }
            }
        }
        tc_fence_before();
    } else {
        // ================= MMA issuer =================
        constexpr uint32_t IDESC2 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                                    ((uint32_t)((BN <= 128 ? 2 * BN : BN) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        constexpr uint32_t IDESC_SW2 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(128 >> 3) << 17) |
                                       ((uint32_t)(128 >> 4) << 24);      // swapped roles: N = [dY_hi | dY_lo] = 128
        constexpr uint32_t IDESC_SW1 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) |
                                       ((uint32_t)(128 >> 4) << 24);      // N = dY_hi = 64
        const uint32_t tiles_u32 = smem_u32(tiles);
        int it = 0;
        for (; it < n_blocks; ++it) {
            const int s = it % STAGES;
            mbar_wait(full0 + 8 * s, (it / STAGES) & 1);
            const int nvalid = (int)min((long long)KB, r_end - (r_begin + (long long)it * KB));
            const uint32_t k16n = (uint32_t)((nvalid + 15) / 16);
            fence_async_smem();        // producers' cp.async / zero-fill writes (generic proxy), observed through the barrier -> async proxy
            tc_fence_after();
            // one elected lane, descriptors = constant + stage offset, K loop unrolled: the issue rate of the UMMA
            // stream matters (tools/micro/mma_bench2.cu).  For BN = 128 the B tile [B_hi | B_lo] (adjacent 64-column
            // blocks) is ONE N = 256 operand: A_hi.[B_hi|B_lo] lands in [main | corr], A_lo.B_hi is added into corr.
            if (elect_one()) {
                const uint64_t a_hi = make_desc_mn(tiles_u32 + (uint32_t)(s * STAGE)), a_lo = a_hi + (A_BYTES >> 4);
                const uint64_t b_hi = a_hi + (2 * A_BYTES >> 4), b_lo = b_hi + (B_BYTES >> 4);
                const uint32_t d_main = tmem_base, d_corr = tmem_base + (uint32_t)BN;
#pragma unroll
                for (int k16 = 0; k16 < KB / 16; ++k16) {
                    if (k16 < (int)k16n) {
                        const uint64_t adv = (uint64_t)((k16 * 2048) >> 4);     // next 16-row K step (two 8-row atoms)
                        if (SWAP) {
                            // narrow cout: M = 128 lanes would be 3/4 (1/2) empty with dY as the A operand.  Swap the roles:
                            // A = 128 gathered (tap, ci) columns (two halves of the X tile), B = [dY_hi | dY_lo] as ONE
                            // N = 128 operand (its two 64-column blocks are adjacent) -> [main | corr], then X_lo . dY_hi
                            // into corr: 4 narrow instructions (2 x (64 + 48) clk) instead of 3 N = 256 ones (3 x 128 clk).
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const uint64_t xh = b_hi + adv + (uint64_t)(h * (2 * 4096 >> 4)), xl = b_lo + adv + (uint64_t)(h * (2 * 4096 >> 4));
                                const uint32_t d = tmem_base + (uint32_t)(h * 128);
                                if (it == 0 && k16 == 0) umma_bf16_set(d, xh, a_hi, IDESC_SW2);
                                else umma_bf16_acc(d, xh, a_hi + adv, IDESC_SW2);
                                umma_bf16_acc(d + 64u, xl, a_hi + adv, IDESC_SW1);
                            }
                        } else if (BN <= 128) {
                            if (it == 0 && k16 == 0) umma_bf16_set(d_main, a_hi, b_hi, IDESC2);
                            else umma_bf16_acc(d_main, a_hi + adv, b_hi + adv, IDESC2);
                            umma_bf16_acc(d_corr, a_lo + adv, b_hi + adv, IDESC);
                        } else {
                            if (it == 0 && k16 == 0) {
                                umma_bf16_set(d_main, a_hi, b_hi, IDESC);
                                umma_bf16_set(d_corr, a_lo, b_hi, IDESC);
                            } else {
                                umma_bf16_acc(d_main, a_hi + adv, b_hi + adv, IDESC);
                                umma_bf16_acc(d_corr, a_lo + adv, b_hi + adv, IDESC);
                            }
                            umma_bf16_acc(d_corr, a_hi + adv, b_lo + adv, IDESC);
                        }
                    }
                }
                umma_commit(empty0 + 8 * s);
                if (it == n_blocks - 1) umma_commit(accum_bar);
            }
            __syncwarp();
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == NPW) {
        tc_fence_after();
        tmem_dealloc(tmem_base, w_tmem_cols(BN));
    }
}

template <int BN, int AM>
constexpr size_t wg_smem()
{
    return 1024 + (size_t)w_stages(BN, AM) * w_stage_bytes(BN, AM) + (2 * MAX_T * WSTR + 2) * 4 + (2 * w_stages(BN, AM) + 1) * 8 +
           (w_stages(BN, AM) + 8) * 4;
}

template <int BN, int AM>
int32_t launch_wg(const WgArgs &a, dim3 grid, cudaStream_t stream)
{
    static_assert(wg_smem<BN, AM>() <= 227 * 1024, "shared memory budget");
    static PerDevice pd;
    CPD_CUDA(opt_in_smem(pd, gather_wgrad_rows_kernel<BN, AM>, wg_smem<BN, AM>()));
    gather_wgrad_rows_kernel<BN, AM><<<grid, NTHREADS, wg_smem<BN, AM>(), stream>>>(a);
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

bool gather_wgrad_rows_supported(int32_t cin, int32_t K, int32_t cout)
{
    // cin must tile N: either a divisor-friendly small width (cin | 256 or cin | 128) or a multiple of 256
    const bool cin_ok = cin >= 8 && cin % 8 == 0 && ((cin <= 256 && 256 % cin == 0) || cin % 256 == 0);
    return cin_ok && cout % 8 == 0 && cout >= 8 && K <= 64;
}

// dw must be zeroed by the caller (split-K reductions).  nbr_t: tap-major (K, m_out) table.
int32_t gather_wgrad_rows_tc(const void *xs, int32_t cin, const void *dys, int64_t m_out, int32_t cout, const int32_t *nbr_t,
                        int32_t K, float *dw, cudaStream_t stream)
{
    CPD_REQUIRE(gather_wgrad_rows_supported(cin, K, cout), CPD_ERR_UNSUPPORTED, "tcgen05 wgrad: unsupported shape");
    CPD_REQUIRE((((uintptr_t)xs | (uintptr_t)dys | (uintptr_t)dw) & 15) == 0, CPD_ERR_MISALIGNED, "tcgen05 wgrad: pointers must be 16-byte aligned");
    // N tile: 256 when the taps of a group (or cin itself) fill it, else 128
    const int bn = (cin >= 256 || (long long)K * cin > 128) ? 256 : 128;
    const int am = cout <= 32 ? 32 : cout <= 64 ? 64 : 128;
    const int ci_tiles = cin > bn ? cin / bn : 1;
    int tpg = cin >= bn ? 1 : bn / cin;                       // taps per group
    if (tpg > MAX_T) tpg = MAX_T;
    if (tpg > K) tpg = K;
    const int groups = (int)div_up(K, tpg), co_tiles = (int)div_up(cout, 128);
    // Row slices (split-K): one CTA per SM at a time, so the launch costs waves x (rows per CTA + a fixed prologue /
    // epilogue); pick the slice count that minimises it -- a grid of 2.5 waves wastes half a wave.
    const long long per_slice = (long long)groups * ci_tiles * co_tiles;
    const long long min_s = std::max<long long>(1, div_up(m_out, (long long)MAX_ROWS_PER_CTA));
    const long long max_s = std::max<long long>(min_s, std::min<long long>(div_up(m_out, WINR), div_up(148 * 6, per_slice)));
    long long S = min_s, best = -1;
    for (long long s = min_s; s <= max_s; ++s) {
        const long long r = div_up(div_up(m_out, s), WINR) * WINR, ctas = div_up(m_out, r) * per_slice;
        const long long cost = div_up(ctas, 148) * (r + 384);          // 384 rows ~ prologue (zero fill) + epilogue (TMEM -> atomics)
        if (best < 0 || cost < best) { best = cost; S = s; }
    }
    long long rows = div_up(div_up(m_out, S), WINR) * WINR;
    S = div_up(m_out, rows);
    CPD_REQUIRE(S <= 65535, CPD_ERR_UNSUPPORTED, "tcgen05 wgrad: too many row slices");
    static const int skip_zero = getenv("CPD_WGRAD_SKIP_ZERO") ? atoi(getenv("CPD_WGRAD_SKIP_ZERO")) : 1;
    WgArgs a{reinterpret_cast<const uint8_t *>(xs), reinterpret_cast<const uint8_t *>(dys), nbr_t, dw, m_out, cin, cout, K, (int)rows, ci_tiles, tpg, skip_zero};
    dim3 grid(groups, (unsigned)S, ci_tiles * co_tiles);
    if (bn == 256) return launch_wg_am<256>(am, a, grid, stream);
    return launch_wg_am<128>(am, a, grid, stream);
}

}  // namespace cpd
