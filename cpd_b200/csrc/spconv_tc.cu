// tcgen05 gather-GEMM for sm_100a: y[o,:] = epi( sum_k W[:,k,:] x[nbr[o,k],:] ) with fp32-class
// accuracy on the 5th-generation tensor cores (bf16x3 operand splitting, fp32 accumulation in TMEM).
//
// Serves spconv.SubMConv3d / SparseConv3d forward + input-gradient (call sites
// cpd/models/backbones_3d/spconv_backbone.py:17,20-21,108-115) and, through cpd_conv2d_table,
// the dense BEV convolutions (cpd/models/backbones_2d/base_bev_backbone.py:31-59,
// cpd/models/dense_heads/center_head.py:11-45,73-80).
//
// The contraction index is the flattened (tap, channel) pair f = k * cin + c, cut into k-blocks of
// 64 (= one 128-byte bf16 swizzle row): a 16- or 32-channel layer packs 4 or 2 taps into one
// k-block, a 128-channel layer spends two k-blocks per tap.  W (cout, K, cin) is already contiguous
// along f, so the B operand of k-block kb is simply W[:, 64 kb : 64 kb + 64].
//
// PERSISTENT, warp-specialised: one CTA per SM pulls (128-row x BN-channel) output tiles from a global
// atomic counter (dynamic scheduling: a CTA that starts late or shares its SM with another stream's
// kernels -- NCCL under DDP, the prefetched input stage -- simply takes fewer tiles); the operand
// pipeline never drains between tiles and the epilogue of tile i overlaps the main loop of tile i+1
// (double-buffered TMEM accumulators).
//   weight_split_kernel (one tiny launch before the GEMM): W -> bf16 hi / lo images stored in
//              global memory ALREADY in the swizzled shared-memory tile layout, one contiguous
//              [hi | lo] block of 2 * BN * 128 bytes per (k-block, cout tile).
//   warp 13    loader: per k-block ONE cp.async.bulk (TMA engine; no tensor map needed because the
//              image is pre-swizzled) brings the B tile in; per tile one more bulk copy prefetches
//              the NEXT tile's [128 x K] slice of the neighbour table into a double-buffered
//              shared-memory copy.
//   warps 4-11 producers: gather A rows from the split-row image of x (split.cu: every row is split
//              into bf16 hi | lo ONCE per layer, not once per rulebook pair) with 16-byte cp.async
//              straight into the canonical K-major SWIZZLE_128B layout that UMMA descriptors address:
//              no register staging, no conversion work in the loop, missing neighbours zero-filled by
//              the copy itself (src-size 0), STAGES-1 k-blocks (up to 4 x 32 KB) in flight per SM --
//              the gather is L2-latency bound, so bytes in flight are what buys bandwidth.
//              Each thread's copies signal the stage's mbarrier themselves when they land
//              (cp.async.mbarrier.arrive.noinc): the producers never wait for data.  (A producer-side
//              cp.async.wait_group + fence.proxy.async hand-over serialises the pipeline -- the proxy
//              fence waits for ALL of the thread's copies in flight -- and ran 2x slower.)  The
//              generic -> async proxy fence is executed by the MMA warp after it has observed the barrier.
//   warp 12    allocates TMEM, then one elected lane issues per k-block 4 x 3
//              tcgen05.mma.cta_group::1.kind::f16 (A_hi.B_hi + A_lo.B_hi + A_hi.B_lo, M=128,
//              N=BN, K=16), tcgen05.commit's the stage back and, per tile, the accumulator over.
//   warps 0-3  epilogue: tcgen05.ld the accumulators (lane = row), + bias, per-channel sum /
//              sum-of-squares (training BatchNorm statistics: butterfly transpose-reduce across the
//              warp, kept in registers across tiles, ONE atomic per column per warp at the end),
//              folded-BatchNorm affine / residual / ReLU, 64-byte row segments stored directly.
//
// Why bf16x3: one bf16 product keeps 8 mantissa bits, far from the 1e-4 parity bar;
// hi = RN(x), lo = RN(x - hi) restores ~2^-17 relative error per product (measured ~2e-5 on the
// layer outputs, tools/tc_check.py) and moves half the shared-memory bytes of a 3xTF32 split --
// and shared-memory bandwidth (operand stores + UMMA operand reads), not the tensor pipe, is what
// bounds a 3-product emulation at M = N = 128.
#include <cuda.h>
#include <stdlib.h>
#include "tc_common.cuh"

namespace cpd {
namespace {
using namespace tc;

constexpr int BM = 128;       // UMMA M
constexpr int BKE = 64;       // bf16 elements per k-block = 128 bytes = one swizzle row
constexpr int NEPI = 128;     // warps 0-3: epilogue (TMEM lane quarter = warp index)
constexpr int NPW = 8;        // warps 4-11: producers
constexpr int NPROD_SP = NPW * 32;                 // sparse kernel: warp 12 = MMA issuer, warp 13 = loader
constexpr int NTHREADS_SP = NEPI + NPROD_SP + 64;
constexpr int NTHREADS_DENSE = NEPI + 64;          // dense kernel: no producers; warp 4 = MMA issuer, warp 5 = loader
constexpr int RSTEP = NPROD_SP / 8;   // row stride between the chunks one producer thread owns (8 x 16 B chunks per row)
constexpr int A_V = BM / RSTEP;    // rows per producer thread per k-block
constexpr int MAX_TAPS = 27;
constexpr int SCHED_R = 4;                     // depth of the tile-id ring
constexpr int KBW = 4;                         // 32-bit words of the per-tile active-k-block bitmap (n_kb <= 128)

__host__ __device__ constexpr int stage_bytes(int bn) { return 2 * BM * 128 + 2 * bn * 128; }
__host__ __device__ constexpr int stages_for(int bn) { return bn >= 256 ? 2 : bn >= 128 ? 3 : bn >= 32 ? 4 : 5; }
// TMEM: per accumulator buffer one "main" (A_hi.B_hi) and one "correction" (A_lo.B_hi + A_hi.B_lo, ~2^-9 of
// the main magnitude) accumulator of BN columns each; two buffers (tile i / tile i+1) when they fit in 512 columns.
__host__ __device__ constexpr int acc_bufs(int bn) { return bn <= 128 ? 2 : 1; }
__host__ __device__ constexpr int tmem_cols(int bn)
{
    int need = acc_bufs(bn) * 2 * bn, c = 32;
    while (c < need) c <<= 1;
    return c;
}

// K-major, SWIZZLE_128B, 8-row groups 1024 B apart (SBO), LBO unused (=1), descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr)
{
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// byte offset of 16-byte chunk c (0..7) of row r inside a [rows x 128 B] swizzled tile
__host__ __device__ __forceinline__ uint32_t swz(int r, int c) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)); }

struct TcArgs {
    const float *bias, *scale, *shift, *residual;
    const uint8_t *xs;          // split-row image of x: row i = [hi(cin) | lo(cin)] bf16 (split.cu)
    const uint8_t *wsplit;      // [k-block][cout tile][hi | lo][BN rows x 128 B, swizzled]
    const int32_t *nbr;         // (m_out, K)
    unsigned int *tile_counter; // dynamic tile scheduler (zeroed by weight_split_kernel)
    const uint32_t *tile_masks; // NULL ok: per 128-row tile, bit k = tap k has a neighbour in the tile (cpd_tile_tap_masks)
    const int32_t *out_rows;    // NULL ok: row r of the table is written to y[out_rows[r]] (tables visited in a permuted order)
    float *stats, *y;
    long long m_out;
    int cin, K, cout, relu;
    int cin_shift;              // log2(cin) when cin is a power of two (every CPD layer), else -1
    int l2_hints;               // gathers with L2 evict_last, table slices with evict_first (CPD_L2_HINTS; measured neutral, default off)
    int debug;                  // CPD_TC_DEBUG (timing experiments only, results are WRONG): 1 = no zero-fill copies, 2 = no MMAs, 4 = no epilogue stores, 8 = no B-tile reloads
};

// DENSE variant (cpd_conv2d_fwd / cpd_conv2d_dgrad: the dense BEV convolutions, stride 1): the rows of an output tile are a
// DT_W x DT_H pixel patch of one image and the A operand of tap (ky, kx), channel block cb is the SAME patch shifted by
// (ky - pad, kx - pad) -- a 4-D box {64 channels, DT_W, DT_H, 1} of the NHWC split-row image that ONE tiled TMA load
// (cp.async.bulk.tensor.4d, tensor map over [2 C, W, H, N]) lands in the K-major SWIZZLE_128B layout, zero-filling the
// conv padding and the ragged image border by itself.  No neighbour table, no producer warps, no LSU work.
constexpr int DT_W = 16, DT_H = 8;          // 128 pixels per tile
static_assert(DT_W * DT_H == BM, "dense tile = one UMMA M");
struct DenseGeo {
    int n, h, w;                // input image batch (NHWC rows)
    int ho, wo;                 // conv output size: h + 2 pad - kh + 1, w + 2 pad - kw + 1
    int kh, kw, pad;
    int tiles_x, tiles_y;
    int out_h, out_w, out_sy, out_sx, out_oy, out_ox;   // output pixel (y, x) is written to (y * out_sy + out_oy, x * out_sx + out_ox) of an
                                                        // (n, out_h, out_w) map: (ho, wo, 1, 1, 0, 0) for a conv, the pixel shuffle of ConvTranspose k == s
};

// W (cout, Kf) fp32 -> pre-swizzled bf16 hi / lo tile images.  One thread per 16-byte output chunk.
// Also clears the (2, cout) BatchNorm statistics accumulators of the GEMM that follows on the stream.
__global__ void weight_split_kernel(const float *__restrict__ w, int cout, int Kf, int n_kb, int bn, uint8_t *__restrict__ out,
                                    float *__restrict__ stats, unsigned int *__restrict__ tile_counter)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (stats && t < 2 * cout) stats[t] = 0.f;
    if (t == 0) *tile_counter = 0u;
    if (t >= (long long)n_kb * cout * 8) return;
    const int c = (int)(t & 7);
    const int n = (int)((t >> 3) % cout), kb = (int)((t >> 3) / cout);
    const int f = kb * BKE + c * 8;
    float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
    if (f < Kf) {                                   // Kf % 8 == 0: a chunk is entirely inside or outside
        const float4 *p = reinterpret_cast<const float4 *>(w + (size_t)n * Kf + f);
        v0 = __ldg(p); v1 = __ldg(p + 1);
    }
    uint4 h, l;
    split8(v0, v1, h, l);
    const int ntiles = cout / bn, nt = n / bn, nl = n % bn;
    uint8_t *base = out + ((size_t)kb * ntiles + nt) * (size_t)(2 * bn * 128);
    const uint32_t off = swz(nl, c);
    *reinterpret_cast<uint4 *>(base + off) = h;
    *reinterpret_cast<uint4 *>(base + (size_t)bn * 128 + off) = l;
}

// Sum x[0..31] of every lane across the 32 lanes of a warp: lane L returns the total of x[L] (31 shuffles).
__device__ __forceinline__ float warp_transpose_reduce(float (&x)[32], int lane)
{
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float send = upper ? x[i] : x[i + off];
            const float keep = upper ? x[i + off] : x[i];
            x[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return x[0];
}

template <int BN, bool DENSE>
__global__ void __launch_bounds__(DENSE ? NTHREADS_DENSE : NTHREADS_SP, 1) gather_gemm_tc_kernel(TcArgs a, const __grid_constant__ CUtensorMap xmap, DenseGeo dg)
{
    // warp roles: epilogue 0-3 | producers (sparse only) | MMA | loader
    constexpr int NPROD = DENSE ? 0 : NPROD_SP, MMA_WARP = (NEPI + NPROD) / 32, NCONS = (NEPI + NPROD + 64) / 32;   // every warp consumes the tile sequence
    constexpr int STAGES = stages_for(BN), ACC = acc_bufs(BN);
    constexpr int A_BYTES = BM * 128, B_BYTES = BN * 128, STAGE = stage_bytes(BN);
    // fp32 accumulate (bit 4), bf16 A and B (bits 7, 10), K-major both, N >> 3, M >> 4
    constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

    extern __shared__ uint8_t smem_raw[];
    uint8_t *tiles = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int tbl_ints = DENSE ? 0 : BM * a.K;
    int32_t *nbr_s = reinterpret_cast<int32_t *>(tiles + STAGES * STAGE);          // [2][BM][K]
    uint64_t *bars = reinterpret_cast<uint64_t *>(nbr_s + 2 * tbl_ints);          // full[S] empty[S] tbl_full[2] tbl_empty[2] acc_full[2] acc_empty[2]
    uint32_t *misc = reinterpret_cast<uint32_t *>(bars + 2 * STAGES + 8 + 2 * SCHED_R);   // [0] tmem base, [4..) tile ring
    volatile int32_t *sched_tile = reinterpret_cast<volatile int32_t *>(misc + 4);       // [R] tile id
    volatile uint32_t *sched_kb = reinterpret_cast<volatile uint32_t *>(misc + 4 + SCHED_R);   // [R][KBW] active k-blocks of the tile
    const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * STAGES, tblf0 = empty0 + 8 * STAGES, tble0 = tblf0 + 16,
                   accf0 = tble0 + 16, acce0 = accf0 + 16, schf0 = acce0 + 16, sche0 = schf0 + 8 * SCHED_R;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Kf = a.K * a.cin;
    const int n_kb = (Kf + BKE - 1) / BKE;
    const int ntn = a.cout / BN;                                   // cout tiles
    const long long total_tiles = (DENSE ? (long long)dg.n * dg.tiles_y * dg.tiles_x : (a.m_out + BM - 1) / BM) * ntn;

    if (warp == MMA_WARP) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, NPROD + 1); mbar_init(empty0 + 8 * s, 1); }
            for (int b = 0; b < 2; ++b) {
                mbar_init(tblf0 + 8 * b, 1); mbar_init(tble0 + 8 * b, NPW);
                mbar_init(accf0 + 8 * b, 1); mbar_init(acce0 + 8 * b, NEPI / 32);
            }
            for (int r = 0; r < SCHED_R; ++r) { mbar_init(schf0 + 8 * r, 1); mbar_init(sche0 + 8 * r, NCONS); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(misc)), "r"(tmem_cols(BN)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = misc[0];
    // Every warp walks the same tile sequence: entry ti of the ring is published by the loader warp (atomicAdd on the
    // global counter) and released once all NCONS warps have read it.  A value >= total_tiles ends the sequence.
    // The entry also carries the bitmap of the tile's ACTIVE k-blocks (those with at least one tap that has a neighbour
    // somewhere in the tile, from tile_masks): every role skips the others -- z-boundary tiles of a SubM conv lose a third
    // of their taps, a parity-sorted strided input-gradient keeps <= 8 of 27.
    uint32_t km0 = 0, km1 = 0, km2 = 0, km3 = 0;   // scalars, not an array: must stay in registers
    static_assert(KBW == 4, "bitmap words are spelled out");
    auto fetch_tile = [&](int ti) -> long long {
        const int slot = ti % SCHED_R;
        mbar_wait(schf0 + 8 * slot, (ti / SCHED_R) & 1);
        const long long t = sched_tile[slot];
        km0 = sched_kb[slot * KBW]; km1 = sched_kb[slot * KBW + 1]; km2 = sched_kb[slot * KBW + 2]; km3 = sched_kb[slot * KBW + 3];
        __syncwarp();
        if (lane == 0) mbar_arrive(sche0 + 8 * slot);
        return t;
    };
    auto kb_bit = [](int kb, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) -> bool {
        const int w = kb >> 5;
        const uint32_t word = w == 0 ? w0 : w == 1 ? w1 : w == 2 ? w2 : w3;
        return (word >> (kb & 31)) & 1u;
    };
    auto kb_active = [&](int kb) -> bool { return kb_bit(kb, km0, km1, km2, km3); };

    if (warp < NEPI / 32) {
        // ================= epilogue (warps 0-3; TMEM lanes 32 * warp .. + 31 = tile rows) =================
        float st_acc[BN / 16];                      // running BatchNorm statistics: lane L < 16: sum of column 16 i + L; L >= 16: sum of squares
#pragma unroll
        for (int i = 0; i < BN / 16; ++i) st_acc[i] = 0.f;
        int ti = 0;
        for (long long t = fetch_tile(0); t < total_tiles; t = fetch_tile(++ti)) {
            const int buf = ACC == 2 ? (ti & 1) : 0;
            mbar_wait(accf0 + 8 * buf, ACC == 2 ? ((ti >> 1) & 1) : (ti & 1));
            tc_fence_after();
            const int n0 = (int)(t % ntn) * BN;
            bool valid;
            long long row;                                                 // row of y (and of the residual)
            if (DENSE) {
                const int pt = (int)(t / ntn), r = warp * 32 + lane;       // pixel tile, pixel inside it (x fastest)
                const int img = pt / (dg.tiles_x * dg.tiles_y), rem = pt - img * (dg.tiles_x * dg.tiles_y);
                const int y = (rem / dg.tiles_x) * DT_H + r / DT_W, x = (rem % dg.tiles_x) * DT_W + r % DT_W;
                valid = y < dg.ho && x < dg.wo;
                row = ((long long)img * dg.out_h + (y * dg.out_sy + dg.out_oy)) * dg.out_w + (x * dg.out_sx + dg.out_ox);
            } else {
                const long long trow = (t / ntn) * BM + warp * 32 + lane;  // row of the table / tile
                valid = trow < a.m_out;
                row = (valid && a.out_rows) ? (long long)__ldg(a.out_rows + trow) : trow;
            }
            const uint32_t tacc = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * 2 * BN);
#pragma unroll
            for (int i = 0; i < BN / 16; ++i) {
                uint32_t u[16], v[16];
                tmem_ld16(tacc + (uint32_t)(16 * i), u);                 // main
                tmem_ld16(tacc + (uint32_t)(BN + 16 * i), v);            // correction
                if (i == BN / 16 - 1) {                                  // accumulator fully read: hand the buffer back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acce0 + 8 * buf);
                }
                float o[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) o[j] = __uint_as_float(u[j]) + __uint_as_float(v[j]);
                const int cv = n0 + 16 * i;
                if (a.bias) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 b = __ldg(reinterpret_cast<const float4 *>(a.bias + cv + 4 * q));
                        o[4 * q] += b.x; o[4 * q + 1] += b.y; o[4 * q + 2] += b.z; o[4 * q + 3] += b.w;
                    }
                }
                if (a.stats) {                                           // statistics of the pre-affine output
                    float x[32];
#pragma unroll
                    for (int j = 0; j < 16; ++j) { x[j] = valid ? o[j] : 0.f; x[16 + j] = x[j] * x[j]; }
                    st_acc[i] += warp_transpose_reduce(x, lane);
                }
                if (valid && !((a.debug & 4) && o[0] != 12345.678f)) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float4 r = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
                        if (a.scale) {
                            const float4 sc = __ldg(reinterpret_cast<const float4 *>(a.scale + cv + 4 * q));
                            const float4 sh = __ldg(reinterpret_cast<const float4 *>(a.shift + cv + 4 * q));
                            r.x = fmaf(r.x, sc.x, sh.x); r.y = fmaf(r.y, sc.y, sh.y); r.z = fmaf(r.z, sc.z, sh.z); r.w = fmaf(r.w, sc.w, sh.w);
                        }
                        if (a.residual) {
                            const float4 rs = __ldg(reinterpret_cast<const float4 *>(a.residual + row * a.cout + cv + 4 * q));
                            r.x += rs.x; r.y += rs.y; r.z += rs.z; r.w += rs.w;
                        }
                        if (a.relu) { r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f); }
                        *reinterpret_cast<float4 *>(a.y + row * a.cout + cv + 4 * q) = r;
                    }
                }
            }
            if (a.stats && ntn > 1) {               // the cout tile changes from one tile to the next: flush per tile
#pragma unroll
                for (int i = 0; i < BN / 16; ++i) {
                    atomicAdd(a.stats + (lane < 16 ? 0 : a.cout) + n0 + 16 * i + (lane & 15), st_acc[i]);
                    st_acc[i] = 0.f;
                }
            }
        }
        if (a.stats && ntn == 1) {
#pragma unroll
            for (int i = 0; i < BN / 16; ++i) atomicAdd(a.stats + (lane < 16 ? 0 : a.cout) + 16 * i + (lane & 15), st_acc[i]);
        }
    } else if (!DENSE && warp < MMA_WARP) {
        // ================= producers (warps 4-11; sparse kernel only) =================
        const int ptid = tid - NEPI;
        const int c = ptid & 7, r_base = ptid >> 3;   // 16-byte smem chunk (8 channels), first row (rows r_base + RSTEP j)
        uint32_t soff[A_V];                           // swizzled byte offsets of this thread's chunks (loop invariant)
#pragma unroll
        for (int j = 0; j < A_V; ++j) soff[j] = swz(r_base + RSTEP * j, c);
        const uint32_t tiles_u32 = smem_u32(tiles);
        const uint32_t row_bytes = (uint32_t)a.cin * 4;            // image row = hi(cin) | lo(cin) bf16
        const uint32_t lo_off = (uint32_t)a.cin * 2;
        const uint32_t tab_stride = (uint32_t)(RSTEP * a.K * 4);   // bytes between the table rows of this thread's chunks
        const uint32_t tab0 = smem_u32(nbr_s) + (uint32_t)(r_base * a.K * 4);
        const uint64_t pol_keep = l2_policy_evict_last();          // a feature row is gathered by ~K/2 output rows: keep it in L2
        int g = 0, ti = 0;                            // k-blocks issued so far (all tiles), tiles started
        for (long long t = fetch_tile(0); t < total_tiles; t = fetch_tile(++ti)) {
            const int tb = ti & 1;
            mbar_wait(tblf0 + 8 * tb, (ti >> 1) & 1);                              // this tile's slice of the neighbour table has landed
            const uint32_t tab = tab0 + (uint32_t)(tb * tbl_ints * 4);            // (rows past m_out carry -1: see load_table)
            for (int kb = 0; kb < n_kb; ++kb) {
                if (!kb_active(kb)) continue;
                const int s = g % STAGES;
                mbar_wait(empty0 + 8 * s, ((g / STAGES) & 1) ^ 1);
                const int f = kb * BKE + c * 8;        // flattened (tap, channel) index of this thread's chunk
                int k, ch;                              // (a runtime integer division here cost 12 % of the kernel's instructions)
                if (a.cin_shift >= 0) { k = f >> a.cin_shift; ch = f & (a.cin - 1); }
                else { k = f / a.cin; ch = f - k * a.cin; }
                // chunks past the last tap (f >= K cin, last k-block only) meet zero weights (weight_split_kernel) or lie beyond
                // the K steps the MMA warp issues: any FINITE data will do there, so they re-read tap K-1 instead of branching
                k = min(k, a.K - 1);
                // The whole producer loop is latency-bound per warp (ncu: one generic table load -> address -> copy chain per row,
                // serialised by the copies' "memory" clobbers): fetch the A_V indices FIRST, back to back, then issue the copies.
                int32_t idx[A_V];
                const uint32_t tk = tab + (uint32_t)(k * 4);
#pragma unroll
                for (int j = 0; j < A_V; ++j) idx[j] = lds_i32(tk + (uint32_t)j * tab_stride);
                const uint32_t dst = tiles_u32 + (uint32_t)(s * STAGE);
                const uint8_t *col_hi = a.xs + (size_t)ch * 2, *col_lo = col_hi + lo_off;
                if (a.l2_hints) {
#pragma unroll
                    for (int j = 0; j < A_V; ++j) {
                        const uint32_t sz = idx[j] >= 0 ? 16u : 0u;      // 0 -> the copy writes 16 zero bytes
                        const uint64_t off = (uint64_t)(uint32_t)max(idx[j], 0) * row_bytes;
                        cp_async16_hint(dst + soff[j], col_hi + off, sz, pol_keep);
                        cp_async16_hint(dst + A_BYTES + soff[j], col_lo + off, sz, pol_keep);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < A_V; ++j) {
                        const uint32_t sz = (idx[j] >= 0 || (a.debug & 1)) ? 16u : 0u;
                        const uint64_t off = (uint64_t)(uint32_t)max(idx[j], 0) * row_bytes;
                        cp_async16(dst + soff[j], col_hi + off, sz);
                        cp_async16(dst + A_BYTES + soff[j], col_lo + off, sz);
                    }
                }
                cp_async_arrive_noinc(full0 + 8 * s);    // this thread's arrival fires when its copies above have landed
                ++g;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(tble0 + 8 * tb);                            // table copy no longer needed by this warp
        }
        cp_async_wait_all();
    } else if (warp == MMA_WARP) {
        // ================= MMA issuer =================
        // Issue rate matters as much as tensor time here (measured, tools/micro/mma_bench2.cu): one thread
        // issues a tcgen05.mma every ~52 clk at best, an M=128 instruction needs max(N/2, (4096 + 32 N)/128)
        // clk of tensor / operand-read time.  So: (1) for BN <= 128 the B tile [B_hi | B_lo] (contiguous in
        // the stage) is ONE N = 2 BN operand: A_hi.[B_hi|B_lo] lands in [main | corr] with one instruction
        // and A_lo.B_hi is added into corr -- 2 instructions per K step instead of 3; (2) descriptors are a
        // constant plus the stage offset, the K loop is unrolled, and one elected lane runs the stream.
        constexpr uint32_t IDESC2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((BN <= 128 ? 2 * BN : BN) >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        constexpr uint64_t DESC_HI = (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
        const uint32_t tiles_u32 = smem_u32(tiles);
        const int last_k16 = (Kf - (n_kb - 1) * BKE + 15) / 16;       // K steps of the last (possibly partial) k-block
        int g = 0, ti = 0;
        for (long long t = fetch_tile(0); t < total_tiles; t = fetch_tile(++ti)) {
            const int buf = ACC == 2 ? (ti & 1) : 0;
            mbar_wait(acce0 + 8 * buf, (ACC == 2 ? ((ti >> 1) & 1) : (ti & 1)) ^ 1);   // epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t d_main = tmem_base + (uint32_t)(buf * 2 * BN), d_corr = d_main + (uint32_t)BN;
            bool first = true;                        // the tile's first instruction overwrites the accumulators
            for (int kb = 0; kb < n_kb; ++kb) {
                if (!kb_active(kb)) continue;
                const int s = g % STAGES;
                mbar_wait(full0 + 8 * s, (g / STAGES) & 1);
                fence_async_smem();        // the producers' cp.async writes (generic proxy), observed through the barrier -> async proxy
                tc_fence_after();
                if (elect_one()) {
                    const bool no_mma = (a.debug & 2) && !(first);          // (timing experiment: only the tile's first k-block runs)
                    const uint64_t a_hi = DESC_HI | (uint64_t)(((tiles_u32 + (uint32_t)(s * STAGE)) & 0x3FFFFu) >> 4);
                    const uint64_t a_lo = a_hi + (A_BYTES >> 4), b_hi = a_hi + (2 * A_BYTES >> 4), b_lo = b_hi + (B_BYTES >> 4);
                    const int k16n = kb == n_kb - 1 ? last_k16 : BKE / 16;
#pragma unroll
                    for (int k16 = 0; k16 < BKE / 16; ++k16) {
                        if (k16 < k16n && !no_mma) {
                            const uint64_t adv = (uint64_t)(k16 * 2);            // +32 bytes along K inside the swizzle row
                            if (BN <= 128) {
                                if (first && k16 == 0) umma_bf16_set(d_main, a_hi, b_hi, IDESC2);
                                else umma_bf16_acc(d_main, a_hi + adv, b_hi + adv, IDESC2);           // [main | corr] += A_hi.[B_hi | B_lo]
                                umma_bf16_acc(d_corr, a_lo + adv, b_hi + adv, IDESC);                // corr += A_lo.B_hi
                            } else {
                                if (first && k16 == 0) {
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
                    umma_commit(empty0 + 8 * s);            // frees the stage once these MMAs have read it
                }
                __syncwarp();
                first = false;
                ++g;
            }
            if (elect_one()) umma_commit(accf0 + 8 * buf);  // all of the tile's MMAs done -> epilogue
            __syncwarp();
        }
        tc_fence_before();
    } else {
        // ================= loader: B tiles + neighbour-table slices, all by bulk copy =================
        const uint32_t tile_bytes = (uint32_t)(2 * B_BYTES);
        const uint64_t pol_stream = l2_policy_evict_first();       // a table slice is read exactly once
        auto load_table = [&](long long t, int tb) {          // rows of tile t -> nbr_s[tb]
            const long long row0 = (t / ntn) * BM;
            const int rows = (int)min((long long)BM, a.m_out - row0);
            const int32_t *src = a.nbr + row0 * a.K;
            int32_t *dst = nbr_s + tb * tbl_ints;
            if (rows == BM) {                                  // BM * K * 4 bytes: a multiple of 16, 16-byte aligned source
                if (lane == 0) {
                    mbar_arrive_expect_tx(tblf0 + 8 * tb, (uint32_t)(tbl_ints * 4));
                    if (a.l2_hints) bulk_copy_g2s_hint(smem_u32(dst), src, (uint32_t)(tbl_ints * 4), tblf0 + 8 * tb, pol_stream);
                    else bulk_copy_g2s(smem_u32(dst), src, (uint32_t)(tbl_ints * 4), tblf0 + 8 * tb);
                }
            } else {                                           // ragged last tile: plain copy by the warp
                for (int e = lane; e < BM * a.K; e += 32) dst[e] = e < rows * a.K ? __ldg(src + e) : -1;   // producers do not test the row count
                __threadfence_block();
                __syncwarp();
                if (lane == 0) mbar_arrive(tblf0 + 8 * tb);
            }
        };
        auto publish = [&](int j) {                          // tile-id ring entry j <- next tile from the global counter
            const int slot = j % SCHED_R;
            mbar_wait(sche0 + 8 * slot, ((j / SCHED_R) & 1) ^ 1);
            unsigned int t = 0;
            if (lane == 0) t = atomicAdd(a.tile_counter, 1u);
            t = __shfl_sync(0xffffffffu, t, 0);
            uint32_t mask = 0xffffffffu;
            if (a.tile_masks && (long long)t < total_tiles) mask = __ldg(a.tile_masks + t / ntn);
            bool any = false;
            for (int w = 0; w < KBW; ++w) {               // lane l decides k-block 32 w + l
                const int kb = 32 * w + lane;
                bool act = false;
                if (kb < n_kb) {
                    const int t0 = (kb * BKE) / a.cin;
                    int t1 = (kb * BKE + BKE - 1) / a.cin;
                    if (t1 > a.K - 1) t1 = a.K - 1;
                    const uint32_t upto = t1 >= 31 ? 0xffffffffu : ((1u << (t1 + 1)) - 1u);
                    act = (mask & upto & ~((1u << t0) - 1u)) != 0u;
                }
                uint32_t bits = __ballot_sync(0xffffffffu, act);
                any |= bits != 0u;
                if (lane == 0) sched_kb[slot * KBW + w] = bits;
            }
            if (!any && lane == 0) sched_kb[slot * KBW] = 1u;   // an empty tile still runs k-block 0 (all zero rows -> zero output)
            if (lane == 0) {
                sched_tile[slot] = (int32_t)(t < 0x7fffffffu ? t : 0x7fffffffu);
                mbar_arrive(schf0 + 8 * slot);
            }
            __syncwarp();
        };
        if (DENSE && lane == 0) tma_prefetch_desc(&xmap);
        publish(0);
        long long t = fetch_tile(0);
        if (!DENSE && t < total_tiles) load_table(t, 0);
        const int cpb = a.cin / BKE;                                   // DENSE: 64-channel blocks per tap (cin % 64 == 0)
        int g = 0, ti = 0;
        for (; t < total_tiles; ++ti) {
            const uint32_t c0 = km0, c1 = km1, c2 = km2, c3 = km3;     // this tile's k-block map (fetch_tile overwrites km*)
            auto kb_active_cur = [&](int kb) -> bool { return kb_bit(kb, c0, c1, c2, c3); };
            publish(ti + 1);
            const long long t_next = fetch_tile(ti + 1);
            if (!DENSE && t_next < total_tiles) {
                const int tbn = (ti + 1) & 1;
                mbar_wait(tble0 + 8 * tbn, (((ti + 1) >> 1) & 1) ^ 1);             // producers are done with the tile that used this buffer
                load_table(t_next, tbn);
            }
            const int nt = (int)(t % ntn);
            int img = 0, x0 = 0, y0 = 0;                               // DENSE: origin of the tile's pixel patch
            if (DENSE) {
                const int pt = (int)(t / ntn);
                img = pt / (dg.tiles_x * dg.tiles_y);
                const int rem = pt - img * (dg.tiles_x * dg.tiles_y);
                y0 = (rem / dg.tiles_x) * DT_H - dg.pad; x0 = (rem % dg.tiles_x) * DT_W - dg.pad;
            }
            for (int kb = 0; kb < n_kb; ++kb) {
                if (!kb_active_cur(kb)) continue;
                const int s = g % STAGES;
                if (lane == 0) {
                    mbar_wait(empty0 + 8 * s, ((g / STAGES) & 1) ^ 1);
                    const uint8_t *src = a.wsplit + ((size_t)kb * ntn + nt) * (size_t)tile_bytes;
                    const uint32_t stage_u32 = smem_u32(tiles + s * STAGE);
                    if (DENSE) {
                        // A_hi / A_lo: the tile's pixel patch shifted by the tap, 64 channels of the hi / lo half of the image rows
                        const int tap = kb / cpb, ch = (kb - tap * cpb) * BKE;
                        const int ky = tap / dg.kw, kx = tap - ky * dg.kw;
                        mbar_arrive_expect_tx(full0 + 8 * s, tile_bytes + 2 * A_BYTES);
                        tma_load_4d(stage_u32, &xmap, full0 + 8 * s, ch, x0 + kx, y0 + ky, img);
                        tma_load_4d(stage_u32 + A_BYTES, &xmap, full0 + 8 * s, a.cin + ch, x0 + kx, y0 + ky, img);
                    } else if ((a.debug & 8) && g >= STAGES) {      // (timing experiment: B tiles are loaded once per stage only)
                        mbar_arrive(full0 + 8 * s);
                    } else {
                        mbar_arrive_expect_tx(full0 + 8 * s, tile_bytes);
                    }
                    if (DENSE || !((a.debug & 8) && g >= STAGES)) bulk_copy_g2s(stage_u32 + 2 * A_BYTES, src, tile_bytes, full0 + 8 * s);
                }
                __syncwarp();
                ++g;
            }
            t = t_next;
        }
    }
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols(BN)));
    }
}

template <int BN>
size_t smem_bytes(int K)
{
    return 1024 + (size_t)stages_for(BN) * stage_bytes(BN) + (size_t)2 * K * BM * 4 + (2 * stages_for(BN) + 8 + 2 * SCHED_R) * 8 + 16 +
           SCHED_R * 4 * (1 + KBW);
}

int num_sms()
{
    static PerDevice pd;
    return device_sms(pd);
}

template <int BN, bool DENSE>
int32_t launch_tc(const TcArgs &a, const CUtensorMap &xmap, const DenseGeo &dg, cudaStream_t stream)
{
    static_assert(stages_for(BN) * stage_bytes(BN) + 2 * MAX_TAPS * BM * 4 + 1024 + 256 <= 227 * 1024, "shared memory budget");
    static PerDevice pd;
    CPD_CUDA(opt_in_smem(pd, gather_gemm_tc_kernel<BN, DENSE>, smem_bytes<BN>(DENSE ? 0 : MAX_TAPS)));
    const long long tiles = (DENSE ? (long long)dg.n * dg.tiles_y * dg.tiles_x : div_up(a.m_out, BM)) * (a.cout / BN);
    const unsigned grid = (unsigned)(tiles < num_sms() ? tiles : num_sms());
    gather_gemm_tc_kernel<BN, DENSE><<<grid, DENSE ? NTHREADS_DENSE : NTHREADS_SP, smem_bytes<BN>(DENSE ? 0 : a.K), stream>>>(a, xmap, dg);
    count_launch();
    return launch_status(DENSE ? "cpd_conv2d[tcgen05+TMA]" : "cpd_gather_gemm[tcgen05]");
}

inline int bn_for(int cout) { return cout >= 256 ? 256 : cout; }

}  // namespace

bool gather_gemm_tc_supported(int32_t cin, int32_t K, int32_t cout)
{
    return cin % 8 == 0 && cin >= 8 && K <= MAX_TAPS && (long long)K * cin <= 32ll * KBW * BKE &&
           (cout == 16 || cout == 32 || cout == 64 || cout == 128 || (cout >= 256 && cout % 256 == 0 && cout <= 2048));
}

// workspace = the pre-swizzled bf16 hi / lo weight image
size_t gather_gemm_tc_workspace(int32_t cin, int32_t K, int32_t cout)
{
    const size_t n_kb = (size_t)div_up((long long)K * cin, BKE);
    return 512 + n_kb * (size_t)cout * 256;              // tile counter + weight image
}

// One bit per tap and 128-row tile: does any row of the tile have a neighbour at that tap?
__global__ void tile_masks_kernel(const int32_t *__restrict__ nbr, long long m, int K, uint32_t *__restrict__ masks)
{
    __shared__ uint32_t acc;
    if (threadIdx.x == 0) acc = 0u;
    __syncthreads();
    const long long row = (long long)blockIdx.x * 128 + threadIdx.x;
    uint32_t mine = 0u;
    if (row < m)
        for (int k = 0; k < K; ++k)
            if (__ldg(nbr + row * K + k) >= 0) mine |= 1u << k;
    mine = __reduce_or_sync(0xffffffffu, mine);
    if ((threadIdx.x & 31) == 0 && mine) atomicOr(&acc, mine);
    __syncthreads();
    if (threadIdx.x == 0) masks[blockIdx.x] = acc;
}

// Sort key of a table row = bitmap of the k-blocks it needs: bit j set iff the row has a neighbour at any of the taps
// [j * tpb, (j + 1) * tpb).  Rows with equal keys keep each other's k-blocks busy: visiting the table in key order
// (Rulebook.sorted_table) makes every 128-row tile skip the k-blocks none of its rows needs.
__global__ void tap_block_keys_kernel(const int32_t *__restrict__ nbr, long long m, int K, int tpb, int32_t *__restrict__ keys)
{
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= m) return;
    uint32_t key = 0u;
    for (int k = 0; k < K; ++k)
        if (__ldg(nbr + row * K + k) >= 0) key |= 1u << (k / tpb);
    keys[row] = (int32_t)key;
}

int32_t tap_block_keys(const int32_t *nbr, int64_t m, int32_t K, int32_t tpb, int32_t *keys, cudaStream_t stream)
{
    CPD_REQUIRE(K >= 1 && tpb >= 1 && (K + tpb - 1) / tpb <= 31, CPD_ERR_UNSUPPORTED, "cpd_tap_block_keys: at most 31 tap blocks");
    if (m == 0) return CPD_OK;
    tap_block_keys_kernel<<<(unsigned)div_up(m, 256), 256, 0, stream>>>(nbr, m, K, tpb, keys);
    count_launch();
    return launch_status("cpd_tap_block_keys");
}

int32_t tile_tap_masks(const int32_t *nbr, int64_t m, int32_t K, uint32_t *masks, cudaStream_t stream)
{
    CPD_REQUIRE(K >= 1 && K <= 32, CPD_ERR_UNSUPPORTED, "cpd_tile_tap_masks: at most 32 taps");
    if (m == 0) return CPD_OK;
    tile_masks_kernel<<<(unsigned)div_up(m, 128), 128, 0, stream>>>(nbr, m, K, masks);
    count_launch();
    return launch_status("cpd_tile_tap_masks");
}

int32_t gather_gemm_tc(const void *xs, int32_t cin, const float *w, int32_t K, int32_t cout, const int32_t *nbr,
                       const uint32_t *tile_masks, const int32_t *out_rows, int64_t m_out, const float *bias, const float *scale, const float *shift, const float *residual,
                       int32_t relu, float *stats, float *y, void *ws, size_t ws_bytes, cudaStream_t stream)
{
    CPD_REQUIRE(gather_gemm_tc_supported(cin, K, cout), CPD_ERR_UNSUPPORTED, "tcgen05 gather-GEMM: unsupported shape");
    CPD_REQUIRE((((uintptr_t)xs | (uintptr_t)w | (uintptr_t)y | (uintptr_t)bias | (uintptr_t)scale | (uintptr_t)shift |
                  (uintptr_t)residual | (uintptr_t)nbr) & 15) == 0, CPD_ERR_MISALIGNED, "tcgen05 gather-GEMM: pointers must be 16-byte aligned");
    CPD_REQUIRE(ws && ws_bytes >= gather_gemm_tc_workspace(cin, K, cout), CPD_ERR_WORKSPACE, "tcgen05 gather-GEMM: workspace too small");
    unsigned int *tile_counter = reinterpret_cast<unsigned int *>(((uintptr_t)ws + 255) & ~(uintptr_t)255);
    uint8_t *wsplit = reinterpret_cast<uint8_t *>(tile_counter) + 256;
    const int Kf = K * cin, n_kb = (int)div_up(Kf, BKE), bn = bn_for(cout);
    const long long chunks = (long long)n_kb * cout * 8;
    weight_split_kernel<<<(unsigned)div_up(chunks, 256), 256, 0, stream>>>(w, cout, Kf, n_kb, bn, wsplit, stats, tile_counter);
    count_launch();
    int cin_shift = -1;
    if ((cin & (cin - 1)) == 0)
        for (cin_shift = 0; (1 << cin_shift) < cin; ++cin_shift) {}
    static const int l2_hints = getenv("CPD_L2_HINTS") ? atoi(getenv("CPD_L2_HINTS")) : 0;
    static const int debug = getenv("CPD_TC_DEBUG") ? atoi(getenv("CPD_TC_DEBUG")) : 0;
    TcArgs a{bias, scale, shift, residual, reinterpret_cast<const uint8_t *>(xs), wsplit, nbr, tile_counter, tile_masks, out_rows, stats, y, m_out, cin, K, cout, relu, cin_shift, l2_hints, debug};
    static const CUtensorMap no_map{};
    const DenseGeo no_geo{};
    switch (cout) {
        case 16: return launch_tc<16, false>(a, no_map, no_geo, stream);
        case 32: return launch_tc<32, false>(a, no_map, no_geo, stream);
        case 64: return launch_tc<64, false>(a, no_map, no_geo, stream);
        case 128: return launch_tc<128, false>(a, no_map, no_geo, stream);
        default: return launch_tc<256, false>(a, no_map, no_geo, stream);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Dense stride-1 convolution over an NHWC image batch (the BEV backbone / CenterHead convs and their input-gradients).
// ---------------------------------------------------------------------------------------------------------------
bool conv2d_tc_supported(int32_t cin, int32_t kh, int32_t kw, int32_t cout)
{
    const long long K = (long long)kh * kw;
    return cin >= BKE && cin % BKE == 0 && K >= 1 && K <= MAX_TAPS && K * cin <= 32ll * KBW * BKE &&
           (cout == 16 || cout == 32 || cout == 64 || cout == 128 || (cout >= 256 && cout % 256 == 0 && cout <= 2048));
}

size_t conv2d_tc_workspace(int32_t cin, int32_t kh, int32_t kw, int32_t cout) { return gather_gemm_tc_workspace(cin, kh * kw, cout); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {                       // the driver entry point, resolved at run time: the library does not link libcuda
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int32_t conv2d_tc(const void *xs, int32_t n, int32_t h, int32_t w_, int32_t cin, const float *w, int32_t kh, int32_t kw, int32_t pad,
                  int32_t cout, const float *bias, const float *scale, const float *shift, const float *residual, int32_t relu,
                  float *stats, int32_t clear_stats, float *y, int32_t out_h, int32_t out_w, int32_t out_sy, int32_t out_sx, int32_t out_oy,
                  int32_t out_ox, void *ws, size_t ws_bytes, cudaStream_t stream)
{
    CPD_REQUIRE(conv2d_tc_supported(cin, kh, kw, cout), CPD_ERR_UNSUPPORTED, "cpd_conv2d: needs cin %% 64 == 0 and cout in {16,32,64,128,256k}");
    CPD_REQUIRE((((uintptr_t)xs | (uintptr_t)w | (uintptr_t)y | (uintptr_t)bias | (uintptr_t)scale | (uintptr_t)shift | (uintptr_t)residual) & 15) == 0,
                CPD_ERR_MISALIGNED, "cpd_conv2d: pointers must be 16-byte aligned");
    CPD_REQUIRE(ws && ws_bytes >= conv2d_tc_workspace(cin, kh, kw, cout), CPD_ERR_WORKSPACE, "cpd_conv2d: workspace too small");
    const int ho = h + 2 * pad - kh + 1, wo = w_ + 2 * pad - kw + 1;
    CPD_REQUIRE(ho >= 1 && wo >= 1 && pad >= 0, CPD_ERR_BAD_ARG, "cpd_conv2d: empty output");
    CPD_REQUIRE(out_sy >= 1 && out_sx >= 1 && out_oy >= 0 && out_ox >= 0 && (ho - 1) * out_sy + out_oy < out_h && (wo - 1) * out_sx + out_ox < out_w,
                CPD_ERR_BAD_ARG, "cpd_conv2d: output placement outside the (out_h, out_w) map");
    CPD_REQUIRE((long long)n * out_h * out_w < (1ll << 31) && (long long)n * h * w_ < (1ll << 31), CPD_ERR_UNSUPPORTED, "cpd_conv2d: image batch too large");
    EncodeTiledFn enc = encode_tiled();
    CPD_REQUIRE(enc, CPD_ERR_CUDA, "cpd_conv2d: cuTensorMapEncodeTiled is not available from this driver");
    // split-row image as a 4-D bf16 tensor [n][h][w][2 cin] (row = hi(cin) | lo(cin)); box = 64 channels x DT_W x DT_H pixels
    CUtensorMap xmap;
    const cuuint64_t dims[4] = {(cuuint64_t)(2 * cin), (cuuint64_t)w_, (cuuint64_t)h, (cuuint64_t)n};
    const cuuint64_t strides[3] = {(cuuint64_t)cin * 4, (cuuint64_t)cin * 4 * w_, (cuuint64_t)cin * 4 * w_ * h};
    const cuuint32_t box[4] = {(cuuint32_t)BKE, (cuuint32_t)DT_W, (cuuint32_t)DT_H, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = enc(&xmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(xs), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CPD_REQUIRE(r == CUDA_SUCCESS, CPD_ERR_CUDA, "cpd_conv2d: cuTensorMapEncodeTiled failed (%d)", (int)r);
    unsigned int *tile_counter = reinterpret_cast<unsigned int *>(((uintptr_t)ws + 255) & ~(uintptr_t)255);
    uint8_t *wsplit = reinterpret_cast<uint8_t *>(tile_counter) + 256;
    const int K = kh * kw, Kf = K * cin, n_kb = Kf / BKE;
    static const int bn_wide = getenv("CPD_DENSE_BN") ? atoi(getenv("CPD_DENSE_BN")) : 256;      // tuning knob: N tile for cout >= 256
    const int bn = cout >= 256 ? (bn_wide == 128 ? 128 : 256) : cout;
    const long long chunks = (long long)n_kb * cout * 8;
    weight_split_kernel<<<(unsigned)div_up(chunks, 256), 256, 0, stream>>>(w, cout, Kf, n_kb, bn, wsplit, clear_stats ? stats : nullptr, tile_counter);
    count_launch();
    int cin_shift = -1;
    TcArgs a{bias, scale, shift, residual, reinterpret_cast<const uint8_t *>(xs), wsplit, nullptr, tile_counter, nullptr, nullptr, stats, y,
             (long long)n * ho * wo, cin, K, cout, relu, cin_shift, 0, 0};
    DenseGeo dg{n, h, w_, ho, wo, kh, kw, pad, (int)div_up(wo, DT_W), (int)div_up(ho, DT_H), out_h, out_w, out_sy, out_sx, out_oy, out_ox};
    switch (bn) {
        case 16: return launch_tc<16, true>(a, xmap, dg, stream);
        case 32: return launch_tc<32, true>(a, xmap, dg, stream);
        case 64: return launch_tc<64, true>(a, xmap, dg, stream);
        case 128: return launch_tc<128, true>(a, xmap, dg, stream);
        default: return launch_tc<256, true>(a, xmap, dg, stream);
    }
}

}  // namespace cpd
