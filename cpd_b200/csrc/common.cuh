// Shared helpers for libcpd_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/cpd_b200.h"

namespace cpd {

void set_error(const char *fmt, ...);
void count_launch(int n = 1);

// Status of the last launch on this thread; never synchronises.
static inline int32_t launch_status(const char *what)
{
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("%s: %s", what, cudaGetErrorString(e));
        return CPD_ERR_CUDA;
    }
    return CPD_OK;
}

#define CPD_REQUIRE(cond, code, ...)                 \
    do {                                             \
        if (!(cond)) {                               \
            ::cpd::set_error(__VA_ARGS__);           \
            return (code);                           \
        }                                            \
    } while (0)

#define CPD_CUDA(call)                                                            \
    do {                                                                          \
        cudaError_t e_ = (call);                                                  \
        if (e_ != cudaSuccess) {                                                  \
            ::cpd::set_error("%s: %s", #call, cudaGetErrorString(e_));            \
            return CPD_ERR_CUDA;                                                  \
        }                                                                         \
    } while (0)

// Function attributes (dynamic shared-memory opt-in) and the SM count are per DEVICE: a process that drives several
// GPUs must configure each kernel on each of them.  One bit per device ordinal (< 64).
struct PerDevice {
    unsigned long long configured = 0ull;
    int sms[64] = {0};
};
static inline int current_device()
{
    int dev = 0;
    cudaGetDevice(&dev);
    return dev & 63;
}
static inline int device_sms(PerDevice &pd)
{
    const int dev = current_device();
    if (!pd.sms[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        pd.sms[dev] = n;
    }
    return pd.sms[dev];
}
template <typename Kernel>
static inline cudaError_t opt_in_smem(PerDevice &pd, Kernel k, size_t bytes)
{
    const int dev = current_device();
    if (pd.configured >> dev & 1ull) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) pd.configured |= 1ull << dev;
    return e;
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int64_t div_up(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- open-addressing coordinate hash: uint32 linear cell key -> int32 row ------------
constexpr uint32_t HASH_EMPTY = 0xFFFFFFFFu;

struct HashView {
    uint32_t *keys;
    int32_t *vals;
    uint32_t mask;  // capacity - 1 (capacity is a power of two)
};

static inline uint64_t hash_capacity(int64_t m)
{
    uint64_t cap = 1024;
    while (cap < (uint64_t)m * 2 + 2) cap <<= 1;
    return cap;
}

__device__ __forceinline__ uint32_t hash_mix(uint32_t k)
{
    k ^= k >> 16; k *= 0x85ebca6bu; k ^= k >> 13; k *= 0xc2b2ae35u; k ^= k >> 16;
    return k;
}

__device__ __forceinline__ int32_t hash_find(const HashView &h, uint32_t key)
{
    uint32_t s = hash_mix(key) & h.mask;
    for (;;) {
        uint32_t k = __ldg(h.keys + s);
        if (k == key) return __ldg(h.vals + s);
        if (k == HASH_EMPTY) return -1;
        s = (s + 1) & h.mask;
    }
}

// block-wide exclusive scan of one int per thread (blockDim.x <= 1024, multiple of 32).
// returns the exclusive prefix; *total receives the block sum (valid for all threads).
__device__ __forceinline__ int block_exclusive_scan(int v, int *total)
{
    __shared__ int warp_sums[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    __syncthreads();  // protect warp_sums against a previous call
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int nw = (blockDim.x + 31) >> 5;
        int s = lane < nw ? warp_sums[lane] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= d) s += t;
        }
        warp_sums[lane] = s;  // inclusive
    }
    __syncthreads();
    int base = wid ? warp_sums[wid - 1] : 0;
    *total = warp_sums[((blockDim.x + 31) >> 5) - 1];
    return base + inc - v;
}

}  // namespace cpd
