// scan.cuh -- device-wide exclusive prefix sum over uint32 (hand-written; three launches: tile sums, scan of sums, apply).
#pragma once
#include "common.cuh"

namespace dge
{

constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v)
{
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
        uint32_t t = __shfl_up_sync(0xFFFFFFFFu, v, d);
        if (lane >= unsigned(d)) v += t;
    }
    return v;
}

// Exclusive scan of one value per thread across the block; returns the exclusive prefix, *total gets the block sum.
// `warp_sums` is shared scratch of >= 33 uint32.  Ends with a __syncthreads().
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *warp_sums, uint32_t *total)
{
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
    uint32_t inc = warp_inclusive_scan(v);
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0)
    {
        uint32_t w = lane < nwarps ? warp_sums[lane] : 0;
        uint32_t winc = warp_inclusive_scan(w);
        warp_sums[lane] = winc - w;
        if (lane == 31) warp_sums[32] = winc;
    }
    __syncthreads();
    uint32_t res = inc - v + warp_sums[warp];
    *total = warp_sums[32];
    __syncthreads();
    return res;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tile_sums(const uint32_t *__restrict__ in, size_t n, uint32_t *__restrict__ sums)
{
    __shared__ uint32_t ws[33];
    const size_t base = size_t(blockIdx.x) * SCAN_TILE;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
    {
        size_t i = base + size_t(k) * SCAN_THREADS + threadIdx.x;
        if (i < n) s += in[i];
    }
    uint32_t total;
    block_exclusive_scan(s, ws, &total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// Single block: exclusive scan of sums[0..nt) in place; sums[nt] = grand total.
__global__ void __launch_bounds__(1024) k_scan_sums(uint32_t *sums, size_t nt)
{
    __shared__ uint32_t ws[33];
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (size_t base = 0; base < nt; base += blockDim.x)
    {
        size_t i = base + threadIdx.x;
        uint32_t v = i < nt ? sums[i] : 0;
        uint32_t total;
        uint32_t ex = block_exclusive_scan(v, ws, &total);
        uint32_t carry = carry_s;
        if (i < nt) sums[i] = ex + carry;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) sums[nt] = carry_s;
}

// out[i] = exclusive prefix; each thread owns SCAN_ITEMS CONSECUTIVE elements so the scan order is the array order.
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const uint32_t *in, size_t n, const uint32_t *__restrict__ sums, uint32_t *out)
{
    __shared__ uint32_t ws[33];
    const size_t base = size_t(blockIdx.x) * SCAN_TILE + size_t(threadIdx.x) * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
    {
        size_t i = base + k;
        v[k] = i < n ? in[i] : 0;
        s += v[k];
    }
    uint32_t total;
    uint32_t ex = block_exclusive_scan(s, ws, &total) + sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
    {
        size_t i = base + k;
        if (i < n) out[i] = ex;
        ex += v[k];
    }
}

// Exclusive scan of in[0..n) into out[0..n) (in == out allowed).  `scratch` needs scan_scratch_elems(n) uint32.
// After completion scratch[n_tiles] holds the grand total (device memory); returned pointer addresses it.
inline size_t scan_scratch_elems(size_t n) { return div_up(n, size_t(SCAN_TILE)) + 2; }

inline const uint32_t *device_exclusive_scan(const uint32_t *in, uint32_t *out, size_t n, uint32_t *scratch, cudaStream_t st, unsigned *launches = nullptr)
{
    size_t nt = div_up(n, size_t(SCAN_TILE));
    if (nt == 0)
    {
        DGE_CUDA(cudaMemsetAsync(scratch, 0, sizeof(uint32_t), st));
        return scratch;
    }
    k_scan_tile_sums<<<unsigned(nt), SCAN_THREADS, 0, st>>>(in, n, scratch);
    k_scan_sums<<<1, 1024, 0, st>>>(scratch, nt);
    k_scan_apply<<<unsigned(nt), SCAN_THREADS, 0, st>>>(in, n, scratch, out);
    DGE_LAUNCH_CHECK();
    if (launches) *launches += 3;
    return scratch + nt;
}

} // namespace dge
