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
// One CTA = 128 output rows x BN output channels, accumulators in TMEM.
//   weight_split_kernel (one tiny launch before the GEMM): W -> bf16 hi / lo images stored in
//              global memory ALREADY in the swizzled shared-memory tile layout, one contiguous
//              [hi | lo] block of 2 * BN * 128 bytes per (k-block, cout tile).
//   warp 9     per k-block ONE cp.async.bulk (TMA engine, no tensor map needed because the image
//              is pre-swizzled) brings the B tile in, completing on the stage's full barrier.
//   warps 0-7  producers: gather A rows from the split-row image of x (split.cu: every row is split
//              into bf16 hi | lo ONCE per layer, not once per rulebook pair) with 16-byte cp.async
//              straight into the canonical K-major SWIZZLE_128B layout that UMMA descriptors address:
//              no register staging, no conversion work in the loop, missing neighbours zero-filled by
//              the copy itself (src-size 0), STAGES-1 k-blocks in flight per thread.  Completed groups
//              are handed over with cp.async.wait_group + fence.proxy.async + mbarrier arrive.
//              k-blocks none of whose taps has a neighbour in the tile are skipped altogether.
//   warp 8     allocates TMEM, then one elected lane issues per k-block 4 x 3
//              tcgen05.mma.cta_group::1.kind::f16 (A_hi.B_hi + A_lo.B_hi + A_hi.B_lo, M=128,
//              N=BN, K=16) and tcgen05.commit's the stage back.
//   warps 0-7  epilogue: tcgen05.ld the accumulators (lane = row), + bias, stage through shared
//              memory (padded rows, conflict-free), then coalesced float4 stores with the folded
//              BatchNorm affine / residual / ReLU applied on the way out and per-channel
//              sum / sum-of-squares taken from the staged tile.
//
// Why bf16x3: one bf16 product keeps 8 mantissa bits, far from the 1e-4 parity bar;
// hi = RN(x), lo = RN(x - hi) restores ~2^-17 relative error per product (measured ~1e-5 on the
// layer outputs, tools/tc_check.py) and moves half the shared-memory bytes of a 3xTF32 split --
// and shared-memory bandwidth (operand stores + UMMA operand reads), not the tensor pipe, is what
// bounds a 3-product emulation at M = N = 128.
#include "tc_common.cuh"

namespace cpd {
namespace {
using namespace tc;

constexpr int BM = 128;       // UMMA M
constexpr int BKE = 64;       // bf16 elements per k-block = 128 bytes = one swizzle row
constexpr int NPW = 8;             // producer / epilogue warps (two per SM sub-partition, so their issue stalls overlap)
constexpr int NPROD = NPW * 32;
constexpr int NTHREADS = NPROD + 64;   // + MMA warp (warp NPW) + B-loader warp (warp NPW + 1)
constexpr int RSTEP = NPROD / 8;   // row stride between the chunks one producer thread owns (8 x 16 B chunks per row)
constexpr int A_V = BM / RSTEP;    // rows per producer thread per k-block
constexpr int MAX_TAPS = 32;
constexpr int MAX_KB = 512;        // k-blocks per tile (K * cin / 64)

__host__ __device__ constexpr int stages_for(int bn) { return bn >= 256 ? 2 : bn >= 128 ? 3 : 2; }
__host__ __device__ constexpr int ctas_per_sm(int bn) { return bn <= 64 ? 2 : 1; }
__host__ __device__ constexpr int stage_bytes(int bn) { return 2 * BM * 128 + 2 * bn * 128; }
// TMEM accumulators per tile: n_main(bn) "main" ones (A_hi.B_hi, k-blocks dealt round-robin) + 1
// "correction" one (A_lo.B_hi + A_hi.B_lo, ~2^-9 of the main magnitude).  The tensor core
// truncates when it adds into the fp32 accumulator; with hundreds of adds into one accumulator
// that bias becomes visible at the 1e-4 level.  Spreading the adds over separate accumulators and
// summing them in fp32 registers in the epilogue cuts it for free (TMEM columns are idle).
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
__host__ __device__ __forceinline__ uint32_t swz(int r, int c) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)); }

struct TcArgs {
    const float *bias, *scale, *shift, *residual;
    const uint8_t *xs;          // split-row image of x: row i = [hi(cin) | lo(cin)] bf16
    const uint8_t *wsplit;      // [k-block][cout tile][hi | lo][BN rows x 128 B, swizzled]
    const int32_t *nbr;
    float *stats, *y;
    long long m_out;
    int cin, K, cout, relu;
};

// W (cout, Kf) fp32 -> pre-swizzled bf16 hi / lo tile images.  One thread per 16-byte output chunk.
__global__ void weight_split_kernel(const float *__restrict__ w, int cout, int Kf, int n_kb, int bn, uint8_t *__restrict__ out)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
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

template <int BN>
__global__ void __launch_bounds__(NTHREADS, ctas_per_sm(BN)) gather_gemm_tc_kernel(TcArgs a)
{
    constexpr int STAGES = stages_for(BN);
    constexpr int NMAIN = n_main(BN);
    constexpr int A_BYTES = BM * 128, B_BYTES = BN * 128, STAGE = stage_bytes(BN);
    constexpr int OUT_LD = BN + 4;   // padded staging row (floats): conflict-free 16-byte stores
    static_assert(BM * OUT_LD * 4 <= STAGES * STAGE, "epilogue staging must fit in the pipeline buffers");
    // fp32 accumulate (bit 4), bf16 A and B (bits 7, 10), K-major both, N >> 3, M >> 4
    constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

    extern __shared__ uint8_t smem_raw[];
    uint8_t *tiles = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    int32_t *nbr_s = reinterpret_cast<int32_t *>(tiles + STAGES * STAGE);          // [K][BM]
    uint64_t *bars = reinterpret_cast<uint64_t *>(nbr_s + a.K * BM);              // full[S], empty[S], accum
    uint32_t *misc = reinterpret_cast<uint32_t *>(bars + 2 * STAGES + 1);         // [0] tmem base, [1] tap mask, [2] active k-blocks
    uint16_t *kb_list = reinterpret_cast<uint16_t *>(misc + 4);                   // [n_kb] active k-block indices
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES), accum_bar = smem_u32(bars + 2 * STAGES);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long row0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;                    // output-channel tile (cout > 256 is split over grid.y)
    const int Kf = a.K * a.cin;
    const int n_kb = (Kf + BKE - 1) / BKE;

    if (tid == 0) misc[1] = 0u;
    if (warp == NPW) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, NPW + 1); mbar_init(empty0 + 8 * s, 1); }
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
    if (warp == 0) {   // ordered list of the k-blocks that touch at least one active tap
        const uint32_t tap_mask = misc[1];
        int cnt = 0;
        for (int base = 0; base < n_kb; base += 32) {
            const int kb = base + lane;
            bool act = false;
            if (kb < n_kb) {
                const int t0 = (kb * BKE) / a.cin;
                int t1 = (kb * BKE + BKE - 1) / a.cin;
                if (t1 > a.K - 1) t1 = a.K - 1;
                const uint32_t upto = t1 >= 31 ? 0xffffffffu : ((1u << (t1 + 1)) - 1u);
                act = (tap_mask & upto & ~((1u << t0) - 1u)) != 0u;
            }
            const uint32_t b = __ballot_sync(0xffffffffu, act);
            if (act) kb_list[cnt + __popc(b & ((1u << lane) - 1u))] = (uint16_t)kb;
            cnt += __popc(b);
        }
        if (lane == 0) misc[2] = (uint32_t)cnt;
    }
    asm volatile("bar.sync 2, %0;" ::"n"(NTHREADS) : "memory");
    const int n_iters = (int)misc[2];

    if (warp < NPW) {
        // ================= producers =================
        const int c = tid & 7, r_base = tid >> 3;   // 16-byte smem chunk (8 channels), first row (rows r_base + RSTEP j)
        uint32_t soff[A_V];                         // swizzled byte offsets of this thread's chunks (loop invariant)
#pragma unroll
        for (int j = 0; j < A_V; ++j) soff[j] = swz(r_base + RSTEP * j, c);
        const uint32_t tiles_u32 = smem_u32(tiles);
        const size_t row_bytes = (size_t)a.cin * 4;                // image row = hi(cin) | lo(cin) bf16
        const uint32_t lo_off = (uint32_t)a.cin * 2;
        auto issue = [&](int it) {
            const int s = it % STAGES;
            mbar_wait(empty0 + 8 * s, ((it / STAGES) & 1) ^ 1);
            const int f = (int)kb_list[it] * BKE + c * 8;            // flattened (tap, channel) index of this thread's chunk
            const int k = f / a.cin, ch = f - k * a.cin;
            const bool k_ok = k < a.K;
            const int32_t *nb = nbr_s + (k_ok ? k : 0) * BM + r_base;
            const uint32_t dst = tiles_u32 + (uint32_t)(s * STAGE);
            const uint8_t *col = a.xs + (size_t)ch * 2;
#pragma unroll
            for (int j = 0; j < A_V; ++j) {
                const int32_t idx = k_ok ? nb[RSTEP * j] : -1;
                const uint8_t *src = col + (size_t)(idx >= 0 ? idx : 0) * row_bytes;
                const uint32_t sz = idx >= 0 ? 16u : 0u;                 // 0 -> the copy writes 16 zero bytes
                cp_async16(dst + soff[j], src, sz);
                cp_async16(dst + A_BYTES + soff[j], src + lo_off, sz);
            }
        };
        constexpr int D = STAGES - 1;               // k-blocks in flight per thread
        for (int it = 0; it < n_iters + D; ++it) {
            if (it < n_iters) issue(it);
            cp_async_commit();
            if (it >= D) {
                cp_async_wait<D>();                  // group it - D has landed
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(full0 + 8 * ((it - D) % STAGES));      // one arrival per producer warp
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
                tmem_ld16(taddr, u);
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
    } else if (warp == NPW) {
        // ================= MMA issuer =================
        for (int it = 0; it < n_iters; ++it) {
            const int s = it % STAGES;
            mbar_wait(full0 + 8 * s, (it / STAGES) & 1);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t st = smem_u32(tiles + s * STAGE);
                const uint64_t a_hi = make_desc(st), a_lo = make_desc(st + A_BYTES);
                const uint64_t b_hi = make_desc(st + 2 * A_BYTES), b_lo = make_desc(st + 2 * A_BYTES + B_BYTES);
                const int rem = Kf - (int)kb_list[it] * BKE;                  // contraction elements left from this k-block on
                const int k16n = rem >= BKE ? BKE / 16 : (rem + 15) / 16;
                const uint32_t d_main = tmem_base + (uint32_t)((it % NMAIN) * BN), d_corr = tmem_base + (uint32_t)(NMAIN * BN);
                for (int k16 = 0; k16 < k16n; ++k16) {
                    const uint64_t adv = (uint64_t)((k16 * 32) >> 4);   // +32 bytes along K inside the swizzle row
                    umma_bf16(d_main, a_hi + adv, b_hi + adv, IDESC, (it >= NMAIN || k16) ? 1u : 0u);
                    umma_bf16(d_corr, a_lo + adv, b_hi + adv, IDESC, (it | k16) ? 1u : 0u);
                    umma_bf16(d_corr, a_hi + adv, b_lo + adv, IDESC, 1u);
                }
                umma_commit(empty0 + 8 * s);            // frees the stage once these MMAs have read it
                if (it == n_iters - 1) umma_commit(accum_bar);
            }
            __syncwarp();
        }
        tc_fence_before();
    } else {
        // ================= B loader: one bulk copy of the pre-swizzled [hi | lo] weight tile per k-block =================
        const size_t tile_bytes = (size_t)(2 * B_BYTES);
        const int ntiles = a.cout / BN;
        for (int it = 0; it < n_iters; ++it) {
            const int s = it % STAGES;
            if (lane == 0) {
                mbar_wait(empty0 + 8 * s, ((it / STAGES) & 1) ^ 1);
                const uint8_t *src = a.wsplit + ((size_t)kb_list[it] * ntiles + blockIdx.y) * tile_bytes;
                mbar_arrive_expect_tx(full0 + 8 * s, (uint32_t)tile_bytes);
                bulk_copy_g2s(smem_u32(tiles + s * STAGE + 2 * A_BYTES), src, (uint32_t)tile_bytes, full0 + 8 * s);
            }
            __syncwarp();
        }
    }
    __syncthreads();
    if (warp == NPW) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols(BN)));
    }
}

template <int BN>
size_t smem_bytes(int K, int n_kb)
{
    return 1024 + (size_t)stages_for(BN) * stage_bytes(BN) + (size_t)K * BM * 4 + (2 * stages_for(BN) + 1) * 8 + 4 * 4 +
           (size_t)((n_kb + 7) / 8 * 8) * 2;
}

template <int BN>
int32_t launch_tc(const TcArgs &a, int n_kb, cudaStream_t stream)
{
    static bool configured = false;
    if (!configured) {
        CPD_CUDA(cudaFuncSetAttribute(gather_gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem_bytes<BN>(MAX_TAPS, MAX_KB)));
        configured = true;
    }
    dim3 grid((unsigned)div_up(a.m_out, BM), (unsigned)(a.cout / BN));
    gather_gemm_tc_kernel<BN><<<grid, NTHREADS, smem_bytes<BN>(a.K, n_kb), stream>>>(a);
    count_launch();
    return launch_status("cpd_gather_gemm[tcgen05]");
}

inline int bn_for(int cout) { return cout >= 256 ? 256 : cout; }

}  // namespace

bool gather_gemm_tc_supported(int32_t cin, int32_t K, int32_t cout)
{
    return cin % 8 == 0 && cin >= 8 && K <= MAX_TAPS && (long long)K * cin <= (long long)MAX_KB * BKE &&
           (cout == 16 || cout == 32 || cout == 64 || cout == 128 || (cout >= 256 && cout % 256 == 0 && cout <= 2048));
}

// workspace = the pre-swizzled bf16 hi / lo weight image
size_t gather_gemm_tc_workspace(int32_t cin, int32_t K, int32_t cout)
{
    const size_t n_kb = (size_t)div_up((long long)K * cin, BKE);
    return 256 + n_kb * (size_t)cout * 256;
}

int32_t gather_gemm_tc(const void *xs, int32_t cin, const float *w, int32_t K, int32_t cout, const int32_t *nbr,
                       int64_t m_out, const float *bias, const float *scale, const float *shift, const float *residual,
                       int32_t relu, float *stats, float *y, void *ws, size_t ws_bytes, cudaStream_t stream)
{
    CPD_REQUIRE(gather_gemm_tc_supported(cin, K, cout), CPD_ERR_UNSUPPORTED, "tcgen05 gather-GEMM: unsupported shape");
    CPD_REQUIRE((((uintptr_t)xs | (uintptr_t)w | (uintptr_t)y | (uintptr_t)bias | (uintptr_t)scale | (uintptr_t)shift |
                  (uintptr_t)residual) & 15) == 0, CPD_ERR_MISALIGNED, "tcgen05 gather-GEMM: pointers must be 16-byte aligned");
    CPD_REQUIRE(ws && ws_bytes >= gather_gemm_tc_workspace(cin, K, cout), CPD_ERR_WORKSPACE, "tcgen05 gather-GEMM: workspace too small");
    uint8_t *wsplit = reinterpret_cast<uint8_t *>(((uintptr_t)ws + 255) & ~(uintptr_t)255);
    const int Kf = K * cin, n_kb = (int)div_up(Kf, BKE), bn = bn_for(cout);
    const long long chunks = (long long)n_kb * cout * 8;
    weight_split_kernel<<<(unsigned)div_up(chunks, 256), 256, 0, stream>>>(w, cout, Kf, n_kb, bn, wsplit);
    count_launch();
    TcArgs a{bias, scale, shift, residual, reinterpret_cast<const uint8_t *>(xs), wsplit, nbr, stats, y, m_out, cin, K, cout, relu};
    switch (cout) {
        case 16: return launch_tc<16>(a, n_kb, stream);
        case 32: return launch_tc<32>(a, n_kb, stream);
        case 64: return launch_tc<64>(a, n_kb, stream);
        case 128: return launch_tc<128>(a, n_kb, stream);
        default: return launch_tc<256>(a, n_kb, stream);
    }
}

}  // namespace cpd
