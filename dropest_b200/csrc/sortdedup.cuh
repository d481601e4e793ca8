// sortdedup.cuh -- L3 of the grouping pipeline: one thread block sorts one sub-bucket (<= THREADS*16 packed keys) and collapses
// equal (cell, gene, UMI) keys into one entry with a read count and the OR of the read marks.
// This is the per-read hot loop of the reference -- Gene::add_umi (Gene.cpp:17-24) + UMI::add_read (UMI.cpp:21-34): there it is
// two std::map walks per read; here it is a comparison sort in registers / shared memory with NO atomics:
//   1. coalesced load of the sub-bucket into shared memory, every thread takes 16 consecutive keys
//   2. per-thread bitonic network in registers (80 compare-exchanges, fully unrolled)
//   3. log2(THREADS) merge levels: merge-path binary search + 16-step serial merge out of shared memory
//   4. run detection on the sorted sequence: count = run length, mark = OR of the low 3 bits; block scan gives the output slot
// The keys carry the mark in their low 3 bits, so equal ukeys are adjacent after sorting the raw 64-bit words.
#pragma once
#include "common.cuh"
#include "scan.cuh"

namespace dge
{

constexpr int MS_ITEMS = 16;

// shared-memory index of logical element i: one pad word per 16 keys makes "thread t touches element 16*t + j" conflict-free
__device__ __forceinline__ int ms_phys(int i) { return i + (i >> 4); }

__device__ __forceinline__ void ms_ce(uint64_t &a, uint64_t &b)
{
    const bool sw = b < a;
    const uint64_t lo = sw ? b : a, hi = sw ? a : b;
    a = lo; b = hi;
}

template <int N> __device__ __forceinline__ void ms_thread_sort(uint64_t (&k)[N])
{
#pragma unroll
    for (int size = 2; size <= N; size <<= 1)
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1)
#pragma unroll
            for (int i = 0; i < N; ++i)
            {
                const int j = i ^ stride;
                if (j > i)
                {
                    if ((i & size) == 0) ms_ce(k[i], k[j]); else ms_ce(k[j], k[i]);
                }
            }
}

// keys[s..e) of sub-bucket blockIdx.x -> distinct ukeys ascending, IN PLACE at keys[s..s+m), values (count | mark<<29) at
// uvals[s..s+m), ucount[sb] = m.  Only sub-buckets with n_min < size <= n_max (<= THREADS*16) are handled by this launch.
template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_sort_dedup(uint64_t *__restrict__ keys, uint32_t *__restrict__ uvals,
                                                        const uint32_t *__restrict__ sub_off, const uint32_t *__restrict__ n_sub_ptr,
                                                        uint32_t *__restrict__ ucount, const uint32_t n_min, const uint32_t n_max)
{
    constexpr int CAP = THREADS * MS_ITEMS;
    __shared__ uint64_t sk[CAP + CAP / 16 + 1];
    __shared__ uint32_t ws[33];
    const uint32_t sb = blockIdx.x;
    if (sb >= *n_sub_ptr) return;
    const uint32_t s = sub_off[sb];
    const int n = int(sub_off[sb + 1] - s);
    if (uint32_t(n) <= n_min || uint32_t(n) > n_max) return;
    const int t = threadIdx.x;

#pragma unroll
    for (int j = 0; j < MS_ITEMS; ++j)
    {
        const int i = j * THREADS + t;
        sk[ms_phys(i)] = i < n ? keys[s + i] : EMPTY64;
    }
    __syncthreads();
    uint64_t k[MS_ITEMS];
    const int p0 = t * MS_ITEMS;
    const int row = ms_phys(p0); // 17 * t: the 16 elements of a thread are contiguous in shared memory
#pragma unroll
    for (int j = 0; j < MS_ITEMS; ++j) k[j] = sk[row + j];
    if (p0 < n) ms_thread_sort(k);

    // threads needed to cover n keys, rounded up to a power of two: levels above it have nothing to merge
    int need = 1;
    while (need * MS_ITEMS < n) need <<= 1;
    for (int w = 1; w < need; w <<= 1)
    {
        __syncthreads();
#pragma unroll
        for (int j = 0; j < MS_ITEMS; ++j) sk[row + j] = k[j];
        __syncthreads();
        const int first = t & ~(2 * w - 1);
        const int a_beg = first * MS_ITEMS, cnt = w * MS_ITEMS, b_beg = a_beg + cnt;
        const int a_cnt = min(max(n - a_beg, 0), cnt), b_cnt = min(max(n - b_beg, 0), cnt);
        const int diag = (t - first) * MS_ITEMS;
        if (diag >= a_cnt + b_cnt) continue;           // this thread's output range holds padding only (k[] is not read again)
        if (b_cnt == 0) continue;                      // nothing to merge with: the A run stays where it is
        int lo = max(0, diag - b_cnt), hi = min(diag, a_cnt);
        while (lo < hi)
        {
            const int mid = (lo + hi) >> 1;
            const uint64_t a = sk[ms_phys(a_beg + mid)], b = sk[ms_phys(b_beg + diag - 1 - mid)];
            if (a <= b) lo = mid + 1; else hi = mid;
        }
        int ai = lo, bi = diag - lo;
        uint64_t ka = ai < a_cnt ? sk[ms_phys(a_beg + ai)] : EMPTY64;
        uint64_t kb = bi < b_cnt ? sk[ms_phys(b_beg + bi)] : EMPTY64;
#pragma unroll
        for (int j = 0; j < MS_ITEMS; ++j)
        {
            const bool ta = bi >= b_cnt || (ai < a_cnt && ka <= kb);
            k[j] = ta ? ka : kb;
            if (ta) { ++ai; ka = ai < a_cnt ? sk[ms_phys(a_beg + ai)] : EMPTY64; }
            else { ++bi; kb = bi < b_cnt ? sk[ms_phys(b_beg + bi)] : EMPTY64; }
        }
    }

    // ---- runs of equal ukey
    __syncthreads();
#pragma unroll
    for (int j = 0; j < MS_ITEMS; ++j) sk[row + j] = k[j];
    __syncthreads();
    const int nv = min(max(n - p0, 0), MS_ITEMS);
    uint32_t head_mask = 0;
    {
        uint64_t prev = p0 > 0 && nv > 0 ? (sk[ms_phys(p0 - 1)] >> 3) : EMPTY64; // a ukey has 61 bits: never equal to EMPTY64
#pragma unroll
        for (int j = 0; j < MS_ITEMS; ++j)
        {
            const uint64_t uk = k[j] >> 3;
            if (j < nv && uk != prev) head_mask |= 1u << j;
            prev = uk;
        }
    }
    uint32_t total;
    const uint32_t base = block_exclusive_scan(uint32_t(__popc(head_mask)), ws, &total);
    if (nv > 0)
    {
        // the run of the last key may continue in the following threads' elements
        uint32_t run_cnt = 0, run_mark = 0;
        uint64_t cur = EMPTY64;
        if (nv == MS_ITEMS)
        {
            cur = k[MS_ITEMS - 1] >> 3;
            for (int q = p0 + MS_ITEMS; q < n; ++q)
            {
                const uint64_t x = sk[ms_phys(q)];
                if ((x >> 3) != cur) break;
                ++run_cnt; run_mark |= uint32_t(x) & 7u;
            }
        }
        uint64_t *ok = keys + s + base;
        uint32_t *ov = uvals + s + base;
#pragma unroll
        for (int j = MS_ITEMS - 1; j >= 0; --j)
        {
            if (j < nv)
            {
                const uint64_t uk = k[j] >> 3;
                if (uk != cur) { cur = uk; run_cnt = 0; run_mark = 0; }
                ++run_cnt; run_mark |= uint32_t(k[j]) & 7u;
                if (head_mask & (1u << j))
                {
                    const int o = __popc(head_mask & ((1u << j) - 1u));
                    ok[o] = uk;
                    ov[o] = run_cnt | (run_mark << VAL_MARK_SHIFT);
                }
            }
        }
    }
    if (t == 0) ucount[sb] = total;
}

} // namespace dge
