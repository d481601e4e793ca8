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
#include "tma.cuh"

namespace dge
{

// shared-memory index of logical element i: one pad word per ITEMS keys makes "thread t touches element ITEMS*t + j" conflict-free
template <int ITEMS> __device__ __forceinline__ int ms_phys(int i)
{
    static_assert(ITEMS == 8 || ITEMS == 16, "ITEMS");
    return i + (i >> (ITEMS == 16 ? 4 : 3));
}

constexpr int MS_CLASSES = 5; // warp class + 3 block size classes + the hash-table tail
#ifndef DGE_MS_WARPS
#define DGE_MS_WARPS 24
#endif
constexpr int MS_WARPS_PER_SM = DGE_MS_WARPS; // occupancy target of the sort kernels (caps registers per thread)

// Bins the sub-buckets by size into per-class work lists (order inside a list is irrelevant).
__global__ void __launch_bounds__(256) k_classify_sub(const uint32_t *__restrict__ sub_off, const uint32_t *__restrict__ n_sub_ptr, uint32_t cw, uint32_t c0,
                                                      uint32_t c1, uint32_t c2, uint32_t *__restrict__ cls_count, uint32_t *__restrict__ cls_list,
                                                      size_t list_stride)
{
    const uint32_t nsb = *n_sub_ptr;
    for (uint32_t sb = blockIdx.x * blockDim.x + threadIdx.x; sb < nsb; sb += gridDim.x * blockDim.x)
    {
        const uint32_t n = sub_off[sb + 1] - sub_off[sb];
        if (n == 0) continue;
        const int c = n <= cw ? 0 : n <= c0 ? 1 : n <= c1 ? 2 : n <= c2 ? 3 : 4;
        cls_list[size_t(c) * list_stride + atomicAdd(&cls_count[c], 1u)] = sb;
    }
}

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

// Work item = one sub-bucket from `list`: keys[s..e) -> distinct ukeys ascending, IN PLACE at keys[s..s+m), values
// (count | mark<<29) at uvals[s..s+m), ucount[sb] = m.  Persistent blocks stride over the list; sizes must be <= THREADS*ITEMS.
// BULK = true: the sub-bucket arrives in shared memory by ONE bulk asynchronous copy (cp.async.bulk -> UBLKCP, completion on an mbarrier)
// issued by thread 0, and the copy of the NEXT work item is issued as soon as every thread holds its keys of the current one in
// registers: the global-memory latency of a work item hides behind the sort of the previous one, and the load costs the threads no
// LDG / STS issue slots (the kernel is issue-bound).  The staging buffer is linear; a thread takes its ITEMS consecutive keys in a rotated
// order (element (j + t) mod ITEMS at step j), which spreads a warp's 8-byte loads over all banks.
template <int THREADS, int ITEMS, int MERGE, bool BULK>
__global__ void __launch_bounds__(THREADS, MS_WARPS_PER_SM * 32 / THREADS) k_sort_dedup(uint64_t *__restrict__ keys, uint32_t *__restrict__ uvals,
                                                        const uint32_t *__restrict__ sub_off, const uint32_t *__restrict__ list,
                                                        const uint32_t *__restrict__ list_count, uint32_t *__restrict__ ucount)
{
    constexpr int CAP = THREADS * ITEMS;
    __shared__ uint64_t sk[CAP + CAP / ITEMS + 1];
    __shared__ __align__(16) uint64_t raw[BULK ? CAP + 2 : 2];
    __shared__ uint64_t bar;
    __shared__ uint32_t ws[33];
    const int t = threadIdx.x;
    const int p0 = t * ITEMS;
    const int row = ms_phys<ITEMS>(p0); // the ITEMS elements of a thread are contiguous in shared memory
    const uint32_t n_items = *list_count;
    if (t == 0) sk[CAP + CAP / ITEMS] = EMPTY64; // sentinel slot (never overwritten: element indices stop one word short of it)
    uint32_t parity = 0;
    // bulk copies need 16-byte aligned addresses and sizes: start at the even key index at or below the sub-bucket's first key
    auto issue_load = [&](uint32_t it) {
        const uint32_t sb_n = list[it];
        const uint32_t s_n = sub_off[sb_n], e_n = sub_off[sb_n + 1];
        const uint32_t s_al = s_n & ~1u;
        const uint32_t bytes = (((e_n - s_al) + 1u) & ~1u) * 8u;
        mbar_arrive_expect_tx(&bar, bytes);
        bulk_g2s(raw, keys + s_al, bytes, &bar);
    };
    if (BULK)
    {
        if (t == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
        __syncthreads();
        if (t == 0 && blockIdx.x < n_items) issue_load(blockIdx.x);
    }
    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x)
    {
        const uint32_t sb = list[item];
        const uint32_t s = sub_off[sb];
        const int n = int(sub_off[sb + 1] - s);
        uint64_t k[ITEMS];
        if (BULK)
        {
            mbar_wait(&bar, parity);
            parity ^= 1u;
            const int shift = int(s & 1u);
#pragma unroll
            for (int j = 0; j < ITEMS; ++j)
            {
                const int i = p0 + ((j + t) & (ITEMS - 1));
                k[j] = i < n ? raw[shift + i] : EMPTY64;
            }
            __syncthreads(); // every thread holds its keys: the staging buffer may take the next work item
            if (t == 0 && item + gridDim.x < n_items) issue_load(item + gridDim.x);
        }
        else
        {
#pragma unroll
            for (int j = 0; j < ITEMS; ++j)
            {
                const int i = j * THREADS + t;
                sk[ms_phys<ITEMS>(i)] = i < n ? keys[s + i] : EMPTY64;
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < ITEMS; ++j) k[j] = sk[row + j];
        }
        if (p0 < n) ms_thread_sort(k);

        // threads needed to cover n keys, rounded up to a power of two: levels above it have nothing to merge
        int need = 1;
        while (need * ITEMS < n) need <<= 1;
        for (int w = 1; w < need; w <<= 1)
        {
            __syncthreads();
#pragma unroll
            for (int j = 0; j < ITEMS; ++j) sk[row + j] = k[j];
            __syncthreads();
            const int first = t & ~(2 * w - 1);
            const int a_beg = first * ITEMS, cnt = w * ITEMS, b_beg = a_beg + cnt;
            const int a_cnt = min(max(n - a_beg, 0), cnt), b_cnt = min(max(n - b_beg, 0), cnt);
            const int diag = (t - first) * ITEMS;
            if (diag >= a_cnt + b_cnt) continue;           // this thread's output range holds padding only (k[] is not read again)
            if (b_cnt == 0) continue;                      // nothing to merge with: the A run stays where it is
            int lo = max(0, diag - b_cnt), hi = min(diag, a_cnt);
            while (lo < hi)
            {
                const int mid = (lo + hi) >> 1;
                const uint64_t a = sk[ms_phys<ITEMS>(a_beg + mid)], b = sk[ms_phys<ITEMS>(b_beg + diag - 1 - mid)];
                if (a <= b) lo = mid + 1; else hi = mid;
            }
            // serial merge; an exhausted side reads as EMPTY64 = the largest word, so plain <= picks the live side
            int ai = lo, bi = diag - lo;
            auto ld_a = [&](int i) { return i < a_cnt ? sk[ms_phys<ITEMS>(a_beg + i)] : EMPTY64; };
            auto ld_b = [&](int i) { return i < b_cnt ? sk[ms_phys<ITEMS>(b_beg + i)] : EMPTY64; };
            if (MERGE == 0)
            {   // one element of look-ahead per side: the load issued in a step is consumed a step later
                uint64_t a0 = ld_a(ai), a1 = ld_a(ai + 1), b0 = ld_b(bi), b1 = ld_b(bi + 1);
#pragma unroll
                for (int j = 0; j < ITEMS; ++j)
                {
                    const bool ta = a0 <= b0;
                    k[j] = ta ? a0 : b0;
                    ai += ta ? 1 : 0; bi += ta ? 0 : 1;
                    const int nx = ta ? ai + 1 : bi + 1;
                    const bool ok = nx < (ta ? a_cnt : b_cnt);
                    const uint64_t v = ok ? sk[ms_phys<ITEMS>((ta ? a_beg : b_beg) + nx)] : EMPTY64;
                    a0 = ta ? a1 : a0; a1 = ta ? v : a1;
                    b0 = ta ? b0 : b1; b1 = ta ? b1 : v;
                }
            }
            else if (MERGE == 1)
            {   // no look-ahead, fully predicated: fewer ALU instructions per merged key, one dependent shared-memory load per step
                int ia = a_beg + ai, ib = b_beg + bi;
                const int ea = a_beg + a_cnt, eb = b_beg + b_cnt;
                uint64_t ka = ia < ea ? sk[ms_phys<ITEMS>(ia)] : EMPTY64, kb = ib < eb ? sk[ms_phys<ITEMS>(ib)] : EMPTY64;
#pragma unroll
                for (int j = 0; j < ITEMS; ++j)
                {
                    const bool ta = ka <= kb;
                    k[j] = ta ? ka : kb;
                    ia += ta ? 1 : 0; ib += ta ? 0 : 1;
                    const int idx = ta ? ia : ib;
                    const uint64_t v = idx < (ta ? ea : eb) ? sk[ms_phys<ITEMS>(idx)] : EMPTY64;
                    if (ta) ka = v; else kb = v;
                }
            }
            else
            {   // as MERGE == 1, but an exhausted side reads a sentinel slot (the spare last word of the buffer, kept at EMPTY64)
                // instead of predicating the load: address select, no default moves
                constexpr int SENT = CAP + CAP / ITEMS;
                int ia = a_beg + ai, ib = b_beg + bi;
                const int ea = a_beg + a_cnt, eb = b_beg + b_cnt;
                uint64_t ka = sk[ia < ea ? ms_phys<ITEMS>(ia) : SENT], kb = sk[ib < eb ? ms_phys<ITEMS>(ib) : SENT];
#pragma unroll
                for (int j = 0; j < ITEMS; ++j)
                {
                    const bool ta = ka <= kb;
                    k[j] = ta ? ka : kb;
                    ia += ta ? 1 : 0; ib += ta ? 0 : 1;
                    const int idx = ta ? ia : ib;
                    const uint64_t v = sk[idx < (ta ? ea : eb) ? ms_phys<ITEMS>(idx) : SENT];
                    if (ta) ka = v; else kb = v;
                }
            }
        }

        // ---- runs of equal ukey
        __syncthreads();
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) sk[row + j] = k[j];
        __syncthreads();
        const int nv = min(max(n - p0, 0), ITEMS);
        uint32_t head_mask = 0;
        {
            uint64_t prev = p0 > 0 && nv > 0 ? (sk[ms_phys<ITEMS>(p0 - 1)] >> 3) : EMPTY64; // a ukey has 61 bits: never equal to EMPTY64
#pragma unroll
            for (int j = 0; j < ITEMS; ++j)
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
            if (nv == ITEMS)
            {
                cur = k[ITEMS - 1] >> 3;
                for (int q = p0 + ITEMS; q < n; ++q)
                {
                    const uint64_t x = sk[ms_phys<ITEMS>(q)];
                    if ((x >> 3) != cur) break;
                    ++run_cnt; run_mark |= uint32_t(x) & 7u;
                }
            }
            uint64_t *ok = keys + s + base;
            uint32_t *ov = uvals + s + base;
#pragma unroll
            for (int j = ITEMS - 1; j >= 0; --j)
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
        __syncthreads(); // sk is reloaded by the next item
    }
}

// Warp-synchronous variant for sub-buckets of <= 32*ITEMS keys: every warp of the block owns a private shared-memory region and walks
// the work list on its own -- no block barriers at all (the merge levels of the block version spend a third of their stall cycles
// in __syncthreads skew), one merge level less than the 64-thread class.
template <int WARPS, int ITEMS>
__global__ void __launch_bounds__(WARPS * 32, MS_WARPS_PER_SM / WARPS) k_sort_dedup_warp(uint64_t *__restrict__ keys, uint32_t *__restrict__ uvals,
                                                                                     const uint32_t *__restrict__ sub_off, const uint32_t *__restrict__ list,
                                                                                     const uint32_t *__restrict__ list_count, uint32_t *__restrict__ ucount)
{
    constexpr int CAP = 32 * ITEMS;
    constexpr int REGION = CAP + CAP / ITEMS + 1;
    __shared__ uint64_t sk_all[WARPS * REGION];
    const int t = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    uint64_t *sk = sk_all + warp * REGION;
    const int p0 = t * ITEMS;
    const int row = ms_phys<ITEMS>(p0);
    const uint32_t n_items = *list_count;
    for (uint32_t item = blockIdx.x * WARPS + warp; item < n_items; item += gridDim.x * WARPS)
    {
        const uint32_t sb = list[item];
        const uint32_t s = sub_off[sb];
        const int n = int(sub_off[sb + 1] - s);
#pragma unroll
        for (int j = 0; j < ITEMS; ++j)
        {
            const int i = j * 32 + t;
            sk[ms_phys<ITEMS>(i)] = i < n ? keys[s + i] : EMPTY64;
        }
        __syncwarp();
        uint64_t k[ITEMS];
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) k[j] = sk[row + j];
        if (p0 < n) ms_thread_sort(k);
        int need = 1;
        while (need * ITEMS < n) need <<= 1;
        for (int w = 1; w < need; w <<= 1)
        {
            __syncwarp();
#pragma unroll
            for (int j = 0; j < ITEMS; ++j) sk[row + j] = k[j];
            __syncwarp();
            const int first = t & ~(2 * w - 1);
            const int a_beg = first * ITEMS, cnt = w * ITEMS, b_beg = a_beg + cnt;
            const int a_cnt = min(max(n - a_beg, 0), cnt), b_cnt = min(max(n - b_beg, 0), cnt);
            const int diag = (t - first) * ITEMS;
            if (diag >= a_cnt + b_cnt) continue;
            if (b_cnt == 0) continue;
            int lo = max(0, diag - b_cnt), hi = min(diag, a_cnt);
            while (lo < hi)
            {
                const int mid = (lo + hi) >> 1;
                const uint64_t a = sk[ms_phys<ITEMS>(a_beg + mid)], b = sk[ms_phys<ITEMS>(b_beg + diag - 1 - mid)];
                if (a <= b) lo = mid + 1; else hi = mid;
            }
            int ai = lo, bi = diag - lo;
            auto ld_a = [&](int i) { return i < a_cnt ? sk[ms_phys<ITEMS>(a_beg + i)] : EMPTY64; };
            auto ld_b = [&](int i) { return i < b_cnt ? sk[ms_phys<ITEMS>(b_beg + i)] : EMPTY64; };
            uint64_t a0 = ld_a(ai), a1 = ld_a(ai + 1), b0 = ld_b(bi), b1 = ld_b(bi + 1);
#pragma unroll
            for (int j = 0; j < ITEMS; ++j)
            {
                const bool ta = a0 <= b0;
                k[j] = ta ? a0 : b0;
                ai += ta ? 1 : 0; bi += ta ? 0 : 1;
                const int nx = ta ? ai + 1 : bi + 1;
                const bool ok = nx < (ta ? a_cnt : b_cnt);
                const uint64_t v = ok ? sk[ms_phys<ITEMS>((ta ? a_beg : b_beg) + nx)] : EMPTY64;
                a0 = ta ? a1 : a0; a1 = ta ? v : a1;
                b0 = ta ? b0 : b1; b1 = ta ? b1 : v;
            }
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) sk[row + j] = k[j];
        __syncwarp();
        const int nv = min(max(n - p0, 0), ITEMS);
        uint32_t head_mask = 0;
        {
            uint64_t prev = p0 > 0 && nv > 0 ? (sk[ms_phys<ITEMS>(p0 - 1)] >> 3) : EMPTY64;
#pragma unroll
            for (int j = 0; j < ITEMS; ++j)
            {
                const uint64_t uk = k[j] >> 3;
                if (j < nv && uk != prev) head_mask |= 1u << j;
                prev = uk;
            }
        }
        const uint32_t heads = uint32_t(__popc(head_mask));
        const uint32_t inc = warp_inclusive_scan(heads);
        const uint32_t base = inc - heads;
        const uint32_t total = __shfl_sync(0xFFFFFFFFu, inc, 31);
        if (nv > 0)
        {
            uint32_t run_cnt = 0, run_mark = 0;
            uint64_t cur = EMPTY64;
            if (nv == ITEMS)
            {
                cur = k[ITEMS - 1] >> 3;
                for (int q = p0 + ITEMS; q < n; ++q)
                {
                    const uint64_t x = sk[ms_phys<ITEMS>(q)];
                    if ((x >> 3) != cur) break;
                    ++run_cnt; run_mark |= uint32_t(x) & 7u;
                }
            }
            uint64_t *ok = keys + s + base;
            uint32_t *ov = uvals + s + base;
#pragma unroll
            for (int j = ITEMS - 1; j >= 0; --j)
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
        __syncwarp();
    }
}

} // namespace dge
