// sortcombine.cuh -- device-wide "group identical keys, combine their values, emit sorted" for 64-bit keys.
//
// This is the engine under the per-read grouping (reference Gene::add_umi / UMI::add_read, Gene.cpp:17-24, UMI.cpp:21-34)
// and under cell merges (Gene::merge, Gene.cpp:26-36).  Input: n keys laid out as [ukey : kb-3 | mark : 3] and optionally
// n values (count | mark<<29).  Output: the DISTINCT ukeys in ascending order with combined values (counts added, marks ORed).
//
// Shape (two-level sample sort, then shared-memory hashing):
//   L1  fixed-width radix partition on the top l1_bits of the key (exact histogram -> exclusive scan -> scatter)
//   L2  per L1 bucket: sampled splitters (sorted in shared memory) -> exact sub-histogram -> scan -> scatter
//   L3  per sub-bucket (~1.5 k records): stream through a shared-memory hash table (atomicCAS insert, atomicAdd/Or combine),
//       compact, bitonic sort in shared memory, write back in place; then one scan + gather makes the output dense.
// Range partitioning (not hashing) at L1/L2 is what makes the concatenated sub-bucket outputs globally sorted; sampling makes
// sub-bucket sizes independent of how skewed cells / genes are.  Everything is HBM-bound integer work: no tensor cores.
#pragma once
#include "common.cuh"
#include "scan.cuh"
#include "sortdedup.cuh"
#include <chrono>
#include <cooperative_groups.h>
#include <cub/device/device_segmented_radix_sort.cuh>
#include <cstdlib>

namespace dge
{

constexpr int SC_THREADS = 512;
constexpr int SC_ITEMS = 16;
constexpr int SC_TILE = SC_THREADS * SC_ITEMS; // 8192 keys per block tile
constexpr int SC_MAX_NB1 = 4096;               // max L1 buckets
constexpr int SC_MAX_P2 = 2048;                // max sub-buckets per L1 bucket
constexpr int SC_SAMPLE = 8192;                // max sample size per L1 bucket
constexpr int SC_HT_MAX = 4096;                // largest shared-memory hash table (slots) per sub-bucket
constexpr int SC_DEDUP_THREADS = 256;

// Tunables (environment overrides are for experiments; defaults are what the benchmarks use).
struct SortCombineTuning
{
    int l1_target = 200000; // records per L1 bucket
    int target = 832;       // records per sub-bucket (most then fit the 1024-key class of the merge sort)
    int ht = 2048;          // hash-table slots per sub-bucket (power of two, >= 2 * expected distinct keys)
    int dedup_threads = 256; // threads per dedup block (32..256, power of two)
    int sample = 8192;       // keys sampled per L1 bucket for its sub-bucket splitters (<= SC_SAMPLE)
    SortCombineTuning()
    {
        if (const char *e = std::getenv("DGE_L1_TARGET")) l1_target = std::max(1000, atoi(e));
        if (const char *e = std::getenv("DGE_SC_TARGET")) target = std::max(64, atoi(e));
        if (const char *e = std::getenv("DGE_SC_HT")) ht = atoi(e);
        if (const char *e = std::getenv("DGE_SC_SAMPLE")) sample = std::min(SC_SAMPLE, std::max(256, atoi(e)));
        if (const char *e = std::getenv("DGE_DEDUP_THREADS")) dedup_threads = atoi(e);
        if (dedup_threads != 512 && dedup_threads != 1024) dedup_threads = 256;
        int p = 256;
        while (p < ht && p < SC_HT_MAX) p <<= 1;
        ht = p;
    }
};
inline const SortCombineTuning &sc_tuning() { static SortCombineTuning t; return t; }

// ---------------------------------------------------------------------------------------------------------------------
template <bool HAS_VAL> __device__ __forceinline__ void bitonic_sort_smem(uint64_t *k, uint32_t *v, int P)
{
    for (int size = 2; size <= P; size <<= 1)
    {
        for (int stride = size >> 1; stride > 0; stride >>= 1)
        {
            __syncthreads();
            for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x)
            {
                int lo = 2 * t - (t & (stride - 1));
                int hi = lo + stride;
                bool asc = (lo & size) == 0;
                uint64_t a = k[lo], b = k[hi];
                if ((a > b) == asc)
                {
                    k[lo] = b; k[hi] = a;
                    if (HAS_VAL) { uint32_t va = v[lo]; v[lo] = v[hi]; v[hi] = va; }
                }
            }
        }
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SC_THREADS) k_l1_hist(const uint64_t *__restrict__ keys, size_t n, int shift, int nb1, uint32_t *__restrict__ hist)
{
    __shared__ uint32_t h[SC_MAX_NB1];
    for (int i = threadIdx.x; i < nb1; i += blockDim.x) h[i] = 0;
    __syncthreads();
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
        atomicAdd(&h[keys[i] >> shift], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < nb1; i += blockDim.x)
        if (h[i]) atomicAdd(&hist[i], h[i]);
}

// One tile of SC_TILE keys per block: rank inside the block with shared-memory atomics, reserve one run per (block, bucket)
// with a single global atomicAdd, then write.
template <bool HAS_VAL, int THREADS, int ITEMS>
__global__ void __launch_bounds__(THREADS) k_l1_scatter(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals, size_t n,
                                                        int shift, int nb1, uint32_t *__restrict__ cursor,
                                                        uint64_t *__restrict__ out_keys, uint32_t *__restrict__ out_vals)
{
    constexpr int SC_ITEMS = ITEMS, SC_THREADS = THREADS, SC_TILE = THREADS * ITEMS;
    __shared__ uint32_t cnt[SC_MAX_NB1];
    for (int i = threadIdx.x; i < nb1; i += blockDim.x) cnt[i] = 0;
    __syncthreads();
    const size_t base = size_t(blockIdx.x) * SC_TILE;
    uint64_t k[SC_ITEMS];
    uint32_t r[SC_ITEMS];
    // all loads of the tile first (the shared-memory atomics below would otherwise serialise them: one DRAM round trip per item)
#pragma unroll
    for (int j = 0; j < SC_ITEMS; ++j)
    {
        size_t i = base + size_t(j) * SC_THREADS + threadIdx.x;
        k[j] = i < n ? __ldg(keys + i) : EMPTY64;
    }
#pragma unroll
    for (int j = 0; j < SC_ITEMS; ++j)
    {
        size_t i = base + size_t(j) * SC_THREADS + threadIdx.x;
        if (i < n) r[j] = atomicAdd(&cnt[k[j] >> shift], 1u);
    }
    __syncthreads();
    {   // one global atomicAdd per non-empty (tile, bucket); all of a thread's atomics are issued before any result is consumed
        constexpr int PER = SC_MAX_NB1 / THREADS;
        uint32_t c[PER], o[PER];
#pragma unroll
        for (int q = 0; q < PER; ++q)
        {
            const int i = q * THREADS + int(threadIdx.x);
            c[q] = i < nb1 ? cnt[i] : 0u;
        }
#pragma unroll
        for (int q = 0; q < PER; ++q)
            if (c[q]) o[q] = atomicAdd(&cursor[q * THREADS + int(threadIdx.x)], c[q]);
#pragma unroll
        for (int q = 0; q < PER; ++q)
            if (c[q]) cnt[q * THREADS + int(threadIdx.x)] = o[q];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < SC_ITEMS; ++j)
    {
        size_t i = base + size_t(j) * SC_THREADS + threadIdx.x;
        if (i < n)
        {
            uint32_t pos = cnt[k[j] >> shift] + r[j];
            out_keys[pos] = k[j];
            if (HAS_VAL) out_vals[pos] = vals[i];
        }
    }
}

// Single block.  From the L1 offsets derive, per bucket, the number of sub-buckets p2 and of block tiles, and their
// exclusive scans (sb_base, tile_base; entry [nb1] = totals).
__global__ void __launch_bounds__(1024) k_l1_plan(const uint32_t *__restrict__ l1_off, int nb1, uint32_t SC_TARGET, uint32_t SC_TILE, uint32_t *__restrict__ p2,
                                                  uint32_t *__restrict__ sb_base, uint32_t *__restrict__ tile_base)
{
    __shared__ uint32_t ws[33];
    __shared__ uint32_t carry[2];
    if (threadIdx.x == 0) carry[0] = carry[1] = 0;
    __syncthreads();
    for (int base = 0; base < nb1; base += blockDim.x)
    {
        int b = base + threadIdx.x;
        uint32_t nb = 0, np = 0, nt = 0;
        if (b < nb1)
        {
            nb = l1_off[b + 1] - l1_off[b];
            np = nb == 0 ? 0u : min(uint32_t(SC_MAX_P2), (nb + SC_TARGET - 1) / SC_TARGET);
            nt = (nb + SC_TILE - 1) / SC_TILE;
            p2[b] = np;
        }
        uint32_t tot_p, tot_t;
        uint32_t ex_p = block_exclusive_scan(np, ws, &tot_p);
        uint32_t ex_t = block_exclusive_scan(nt, ws, &tot_t);
        if (b < nb1)
        {
            sb_base[b] = carry[0] + ex_p;
            tile_base[b] = carry[1] + ex_t;
        }
        __syncthreads();
        if (threadIdx.x == 0) { carry[0] += tot_p; carry[1] += tot_t; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { sb_base[nb1] = carry[0]; tile_base[nb1] = carry[1]; }
}

// One block per L1 bucket: strided sample of ukeys, bitonic sort in shared memory, pick p2-1 splitters.
__global__ void __launch_bounds__(SC_THREADS) k_splitters(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ l1_off,
                                                          const uint32_t *__restrict__ p2, uint64_t *__restrict__ splitters, uint32_t max_sample)
{
    extern __shared__ uint64_t samp[];
    const int b = blockIdx.x;
    const uint32_t np = p2[b];
    if (np <= 1) return;
    const uint32_t off = l1_off[b], nb = l1_off[b + 1] - off;
    const uint32_t S = min(nb, max_sample);
    int P = 2;
    while (uint32_t(P) < S) P <<= 1;
    for (int i = threadIdx.x; i < P; i += blockDim.x)
        samp[i] = uint32_t(i) < S ? (keys[off + uint32_t((uint64_t(i) * nb) / S)] >> 3) : EMPTY64;
    bitonic_sort_smem<false>(samp, nullptr, P);
    for (uint32_t j = 1 + threadIdx.x; j < np; j += blockDim.x)
        splitters[size_t(b) * SC_MAX_P2 + (j - 1)] = samp[(uint64_t(j) * S) / np];
}

// number of splitters <= ukey  (splitters ascending, count = np-1)
__device__ __forceinline__ uint32_t sub_bucket_of(const uint64_t *spl, uint32_t nspl, uint64_t ukey)
{
    uint32_t lo = 0, hi = nspl;
    while (lo < hi)
    {
        uint32_t mid = (lo + hi) >> 1;
        if (spl[mid] <= ukey) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ void block_to_bucket_tile(const uint32_t *__restrict__ tile_base, int nb1, uint32_t blk, int *bucket, uint32_t *tile)
{
    // largest b with tile_base[b] <= blk  (tile_base non-decreasing; empty buckets repeat the value)
    int lo = 0, hi = nb1;
    while (lo < hi)
    {
        int mid = (lo + hi + 1) >> 1;
        if (tile_base[mid] <= blk) lo = mid; else hi = mid - 1;
    }
    // skip to the last bucket sharing this base that actually owns tiles: buckets with zero tiles have tile_base[b+1]==tile_base[b]
    *bucket = lo;
    *tile = blk - tile_base[lo];
}

template <bool SCATTER, bool HAS_VAL, int THREADS, int ITEMS>
__global__ void __launch_bounds__(THREADS) k_l2_pass(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals,
                                                     const uint32_t *__restrict__ l1_off, int nb1, const uint32_t *__restrict__ p2,
                                                     const uint32_t *__restrict__ sb_base, const uint32_t *__restrict__ tile_base,
                                                     const uint64_t *__restrict__ splitters, uint32_t *__restrict__ sub_cnt_or_cursor,
                                                     uint64_t *__restrict__ out_keys, uint32_t *__restrict__ out_vals)
{
    constexpr int SC_ITEMS = ITEMS, SC_THREADS = THREADS, SC_TILE = THREADS * ITEMS;
    __shared__ uint64_t spl[SC_MAX_P2];
    __shared__ uint32_t cnt[SC_MAX_P2];
    const uint32_t blk = blockIdx.x;
    if (blk >= tile_base[nb1]) return;
    int b; uint32_t tile;
    block_to_bucket_tile(tile_base, nb1, blk, &b, &tile);
    const uint32_t np = p2[b];
    const uint32_t off = l1_off[b], end = l1_off[b + 1];
    const uint32_t t0 = off + tile * SC_TILE;
    const uint32_t t1 = min(end, t0 + uint32_t(SC_TILE));
    const uint32_t sb0 = sb_base[b];
    for (uint32_t i = threadIdx.x; i < np; i += blockDim.x)
    {
        cnt[i] = 0;
        if (i + 1 < np) spl[i] = splitters[size_t(b) * SC_MAX_P2 + i];
    }
    __syncthreads();
    uint64_t k[SC_ITEMS];
    uint32_t r[SC_ITEMS];
    uint16_t sbk[SC_ITEMS];
    // all loads of the tile first, then the splitter searches, then the shared-memory atomics
#pragma unroll
    for (int j = 0; j < SC_ITEMS; ++j)
    {
        uint32_t i = t0 + uint32_t(j) * SC_THREADS + threadIdx.x;
        k[j] = i < t1 ? __ldg(keys + i) : EMPTY64;
    }
#pragma unroll
    for (int j = 0; j < SC_ITEMS; ++j)
        sbk[j] = uint16_t(np > 1 ? sub_bucket_of(spl, np - 1, k[j] >> 3) : 0u);
#pragma unroll
    for (int j = 0; j < SC_ITEMS; ++j)
    {
        uint32_t i = t0 + uint32_t(j) * SC_THREADS + threadIdx.x;
        if (i < t1) r[j] = atomicAdd(&cnt[sbk[j]], 1u);
    }
    __syncthreads();
    {   // all of a thread's global atomics in flight together
        constexpr int PER = SC_MAX_P2 / THREADS;
        uint32_t c[PER], o[PER];
#pragma unroll
        for (int q = 0; q < PER; ++q)
        {
            const uint32_t i = uint32_t(q * THREADS) + threadIdx.x;
            c[q] = i < np ? cnt[i] : 0u;
        }
#pragma unroll
        for (int q = 0; q < PER; ++q)
            if (c[q]) o[q] = atomicAdd(&sub_cnt_or_cursor[sb0 + uint32_t(q * THREADS) + threadIdx.x], c[q]);
        if (SCATTER)
        {
#pragma unroll
            for (int q = 0; q < PER; ++q)
                if (c[q]) cnt[uint32_t(q * THREADS) + threadIdx.x] = o[q];
        }
    }
    if (!SCATTER) return;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < SC_ITEMS; ++j)
    {
        uint32_t i = t0 + uint32_t(j) * SC_THREADS + threadIdx.x;
        if (i < t1)
        {
            uint32_t pos = cnt[sbk[j]] + r[j];
            out_keys[pos] = k[j];
            if (HAS_VAL) out_vals[pos] = vals[i];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Staged scatter: the tile is reordered by bucket in shared memory before it leaves the SM, so a bucket's keys of one tile go
// out as ONE contiguous run (consecutive threads -> consecutive addresses) instead of one 8-byte partial-sector write per key.
// What bounds a partition pass on B200 is the number of L2 write transactions, not bytes: runs cut them by the run length.

// cnt[0..nb) = per-bucket counts of the tile  ->  cnt[b] = tile-local start of bucket b, gd[b] = global run start - cnt[b]
// (one global atomicAdd per non-empty bucket, all of a thread's atomics in flight together).  PER * blockDim.x >= nb.
template <int PER>
__device__ __forceinline__ void tile_bucket_offsets(uint32_t *cnt, uint32_t *gd, uint32_t nb, uint32_t *__restrict__ cursor, uint32_t *ws)
{
    uint32_t c[PER], o[PER], sum = 0;
#pragma unroll
    for (int q = 0; q < PER; ++q)
    {
        const uint32_t i = threadIdx.x * PER + q;
        c[q] = i < nb ? cnt[i] : 0u;
        sum += c[q];
    }
#pragma unroll
    for (int q = 0; q < PER; ++q)
    {
        o[q] = 0;
        if (c[q]) o[q] = atomicAdd(&cursor[threadIdx.x * PER + q], c[q]);
    }
    uint32_t tot;
    uint32_t ex = block_exclusive_scan(sum, ws, &tot);
#pragma unroll
    for (int q = 0; q < PER; ++q)
    {
        const uint32_t i = threadIdx.x * PER + q;
        if (i < nb) { cnt[i] = ex; gd[i] = o[q] - ex; }
        ex += c[q];
    }
    __syncthreads();
}

// L1: bucket = key >> shift.  Dynamic shared memory: cnt[nb1] | gd[nb1] | keys[THREADS*ITEMS].
template <int THREADS, int ITEMS>
__global__ void __launch_bounds__(THREADS) k_l1_scatter_staged(const uint64_t *__restrict__ keys, size_t n, int shift, int nb1,
                                                               uint32_t *__restrict__ cursor, uint64_t *__restrict__ out_keys)
{
    constexpr int TILE = THREADS * ITEMS;
    constexpr int PER = SC_MAX_NB1 / THREADS;
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint32_t ws[33];
    uint64_t *sk = reinterpret_cast<uint64_t *>(smem_raw);
    uint32_t *cnt = reinterpret_cast<uint32_t *>(sk + TILE);
    uint32_t *gd = cnt + nb1;
    for (int i = threadIdx.x; i < nb1; i += THREADS) cnt[i] = 0;
    __syncthreads();
    const size_t base = size_t(blockIdx.x) * TILE;
    const uint32_t n_tile = uint32_t(min(size_t(TILE), n - base));
    uint64_t k[ITEMS];
    uint32_t r[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j)
    {
        const uint32_t i = uint32_t(j) * THREADS + threadIdx.x;
        k[j] = i < n_tile ? __ldg(keys + base + i) : EMPTY64;
    }
#pragma unroll
    for (int j = 0; j < ITEMS; ++j)
    {
        const uint32_t i = uint32_t(j) * THREADS + threadIdx.x;
        if (i < n_tile) r[j] = atomicAdd(&cnt[k[j] >> shift], 1u);
    }
    __syncthreads();
    tile_bucket_offsets<PER>(cnt, gd, uint32_t(nb1), cursor, ws);
#pragma unroll
    for (int j = 0; j < ITEMS; ++j)
    {
        const uint32_t i = uint32_t(j) * THREADS + threadIdx.x;
        if (i < n_tile) sk[cnt[k[j] >> shift] + r[j]] = k[j];
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n_tile; i += THREADS)
    {
        const uint64_t key = sk[i];
        out_keys[gd[key >> shift] + i] = key;
    }
}

// L1 scatter straight from the fill kernel's per-block key regions (no dense key array, no histogram pass): block = one tile of one
// region.  Dynamic shared memory as k_l1_scatter_staged.
template <int THREADS, int ITEMS>
__global__ void __launch_bounds__(THREADS) k_l1_scatter_regions(const KeyTile *__restrict__ tiles, const uint32_t *__restrict__ n_tiles_ptr,
                                                                int shift, int nb1, uint32_t *__restrict__ cursor, uint64_t *__restrict__ out_keys)
{
    constexpr int TILE = THREADS * ITEMS;
    constexpr int PER = SC_MAX_NB1 / THREADS;
    if (blockIdx.x >= *n_tiles_ptr) return;
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint32_t ws[33];
    uint64_t *sk = reinterpret_cast<uint64_t *>(smem_raw);
    uint32_t *cnt = reinterpret_cast<uint32_t *>(sk + TILE);
    uint32_t *gd = cnt + nb1;
    const KeyTile td = tiles[blockIdx.x];
    for (int i = threadIdx.x; i < nb1; i += THREADS) cnt[i] = 0;
    __syncthreads();
    const uint64_t *__restrict__ keys = td.keys;
    const uint32_t n_tile = td.count;
    uint64_t k[ITEMS];
    uint32_t r[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j)
    {
        const uint32_t i = uint32_t(j) * THREADS + threadIdx.x;
        k[j] = i < n_tile ? __ldg(keys + i) : EMPTY64;
    }
#pragma unroll
    for (int j = 0; j < ITEMS; ++j)
    {
        const uint32_t i = uint32_t(j) * THREADS + threadIdx.x;
        if (i < n_tile) r[j] = atomicAdd(&cnt[k[j] >> shift], 1u);
    }
    __syncthreads();
    tile_bucket_offsets<PER>(cnt, gd, uint32_t(nb1), cursor, ws);
#pragma unroll
    for (int j = 0; j < ITEMS; ++j)
    {
        const uint32_t i = uint32_t(j) * THREADS + threadIdx.x;
        if (i < n_tile) sk[cnt[k[j] >> shift] + r[j]] = k[j];
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n_tile; i += THREADS)
    {
        const uint64_t key = sk[i];
        out_keys[gd[key >> shift] + i] = key;
    }
}

// fold the fill kernel's 12-bit histogram to the chosen L1 width
__global__ void k_hist_fold(const uint32_t *__restrict__ hist12, int l1_bits, uint32_t *__restrict__ out)
{
    const int nb1 = 1 << l1_bits, per = 1 << (12 - l1_bits);
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nb1; b += gridDim.x * blockDim.x)
    {
        uint32_t sum = 0;
        for (int q = 0; q < per; ++q) sum += hist12[b * per + q];
        out[b] = sum;
    }
}

// L2: bucket = position among the L1 bucket's sampled splitters.  Dynamic shared memory:
// spl[SC_MAX_P2] (u64) | keys[TILE] (u64) | cnt[SC_MAX_P2] | gd[SC_MAX_P2] | ids[TILE] (u16).
template <int THREADS, int ITEMS>
__global__ void __launch_bounds__(THREADS) k_l2_scatter_staged(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ l1_off, int nb1,
                                                               const uint32_t *__restrict__ p2, const uint32_t *__restrict__ sb_base,
                                                               const uint32_t *__restrict__ tile_base, const uint64_t *__restrict__ splitters,
                                                               uint32_t *__restrict__ cursor, uint64_t *__restrict__ out_keys)
{
    constexpr int TILE = THREADS * ITEMS;
    constexpr int PER = SC_MAX_P2 / THREADS;
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint32_t ws[33];
    uint64_t *spl = reinterpret_cast<uint64_t *>(smem_raw);
    uint64_t *sk = spl + SC_MAX_P2;
    uint32_t *cnt = reinterpret_cast<uint32_t *>(sk + TILE);
    uint32_t *gd = cnt + SC_MAX_P2;
    uint16_t *ids = reinterpret_cast<uint16_t *>(gd + SC_MAX_P2);
    const uint32_t blk = blockIdx.x;
    if (blk >= tile_base[nb1]) return;
    int b; uint32_t tile;
    block_to_bucket_tile(tile_base, nb1, blk, &b, &tile);
    const uint32_t np = p2[b];
    const uint32_t off = l1_off[b], end = l1_off[b + 1];
    const uint32_t t0 = off + tile * TILE;
    const uint32_t n_tile = min(end - t0, uint32_t(TILE));
    for (uint32_t i = threadIdx.x; i < np; i += THREADS)
    {
        cnt[i] = 0;
        if (i + 1 < np) spl[i] = splitters[size_t(b) * SC_MAX_P2 + i];
    }
    __syncthreads();
    uint64_t k[ITEMS];
    uint32_t r[ITEMS];
    uint16_t sbk[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j)
    {
        const uint32_t i = uint32_t(j) * THREADS + threadIdx.x;
        k[j] = i < n_tile ? __ldg(keys + t0 + i) : EMPTY64;
    }
#pragma unroll
    for (int j = 0; j < ITEMS; ++j)
        sbk[j] = uint16_t(np > 1 ? sub_bucket_of(spl, np - 1, k[j] >> 3) : 0u);
#pragma unroll
    for (int j = 0; j < ITEMS; ++j)
    {
        const uint32_t i = uint32_t(j) * THREADS + threadIdx.x;
        if (i < n_tile) r[j] = atomicAdd(&cnt[sbk[j]], 1u);
    }
    __syncthreads();
    tile_bucket_offsets<PER>(cnt, gd, np, cursor + sb_base[b], ws);
#pragma unroll
    for (int j = 0; j < ITEMS; ++j)
    {
        const uint32_t i = uint32_t(j) * THREADS + threadIdx.x;
        if (i < n_tile)
        {
            const uint32_t p = cnt[sbk[j]] + r[j];
            sk[p] = k[j];
            ids[p] = sbk[j];
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n_tile; i += THREADS) out_keys[gd[ids[i]] + i] = sk[i];
}

// L2 in ONE launch: a block owns a whole L1 bucket.  Phase A walks the bucket once (splitter search, shared-memory histogram, the
// sub-bucket id of every key parked in a 2-byte side array), an in-block scan turns the histogram into the bucket's sub-bucket
// offsets (written straight into the global offset table: sub-buckets of a bucket are contiguous inside the bucket's range), phase B
// walks the bucket again and scatters with shared-memory cursors.  No per-tile global atomics, no exact-histogram launch, no
// device-wide scan, one splitter search per key.
template <int THREADS, int ITEMS>
__global__ void __launch_bounds__(THREADS) k_l2_bucket(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ l1_off,
                                                       const uint32_t *__restrict__ p2, const uint32_t *__restrict__ sb_base,
                                                       const uint64_t *__restrict__ splitters, uint16_t *__restrict__ ids,
                                                       uint32_t *__restrict__ sub_off, uint64_t *__restrict__ out_keys,
                                                       const uint32_t *__restrict__ order, uint32_t giant = 0xFFFFFFFFu)
{
    constexpr int TILE = THREADS * ITEMS;
    constexpr int PER = SC_MAX_P2 / THREADS;
    __shared__ uint64_t spl[SC_MAX_P2];
    __shared__ uint32_t cnt[SC_MAX_P2];
    __shared__ uint32_t ws[33];
    const int b = order ? int(order[blockIdx.x]) : int(blockIdx.x);
    const uint32_t np = p2[b];
    const uint32_t off = l1_off[b], nbk = l1_off[b + 1] - off;
    if (nbk == 0 || np == 0 || nbk > giant) return; // buckets above `giant` keys belong to k_l2_bucket_cluster
    const uint32_t sb0 = sb_base[b];
    for (uint32_t i = threadIdx.x; i < np; i += THREADS)
    {
        cnt[i] = 0;
        if (i + 1 < np) spl[i] = splitters[size_t(b) * SC_MAX_P2 + i];
    }
    __syncthreads();
    for (uint32_t base = 0; base < nbk; base += TILE)
    {
        uint64_t k[ITEMS];
#pragma unroll
        for (int j = 0; j < ITEMS; ++j)
        {
            const uint32_t i = base + uint32_t(j) * THREADS + threadIdx.x;
            k[j] = i < nbk ? __ldg(keys + off + i) : EMPTY64;
        }
#pragma unroll
        for (int j = 0; j < ITEMS; ++j)
        {
            const uint32_t i = base + uint32_t(j) * THREADS + threadIdx.x;
            if (i < nbk)
            {
                const uint32_t sid = np > 1 ? sub_bucket_of(spl, np - 1, k[j] >> 3) : 0u;
                ids[off + i] = uint16_t(sid);
                atomicAdd(&cnt[sid], 1u);
            }
        }
    }
    __syncthreads();
    {   // exclusive scan of the histogram -> shared-memory cursors + global sub-bucket offsets
        uint32_t c[PER], sum = 0;
#pragma unroll
        for (int q = 0; q < PER; ++q)
        {
            const uint32_t i = threadIdx.x * PER + q;
            c[q] = i < np ? cnt[i] : 0u;
            sum += c[q];
        }
        uint32_t tot;
        uint32_t ex = block_exclusive_scan(sum, ws, &tot);
#pragma unroll
        for (int q = 0; q < PER; ++q)
        {
            const uint32_t i = threadIdx.x * PER + q;
            if (i < np) { cnt[i] = ex; sub_off[sb0 + i] = off + ex; }
            ex += c[q];
        }
    }
    __syncthreads();
    for (uint32_t base = 0; base < nbk; base += TILE)
    {
        uint64_t k[ITEMS];
        uint16_t sid[ITEMS];
#pragma unroll
        for (int j = 0; j < ITEMS; ++j)
        {
            const uint32_t i = base + uint32_t(j) * THREADS + threadIdx.x;
            k[j] = i < nbk ? __ldg(keys + off + i) : EMPTY64;
            sid[j] = i < nbk ? ids[off + i] : uint16_t(0);
        }
#pragma unroll
        for (int j = 0; j < ITEMS; ++j)
        {
            const uint32_t i = base + uint32_t(j) * THREADS + threadIdx.x;
            if (i < nbk) out_keys[off + atomicAdd(&cnt[sid[j]], 1u)] = k[j];
        }
    }
}

// The same pass for the LONG L1 buckets (one barcode with a few per cent of all reads lands in a single bucket, whatever the radix width:
// the top key bits are its table slot).  A block per bucket would make that block the critical path of the whole grouping
// (0.3 M keys per ms and block: a 3.6 M-key bucket = 12 ms against 4 ms for everything else), so a bucket longer than `giant` keys is
// shared by a thread-block CLUSTER of 8 CTAs: every CTA takes an eighth of the keys, the per-CTA sub-bucket histograms stay in shared
// memory and are combined through DISTRIBUTED SHARED MEMORY (each CTA reads the 8 histograms: totals -> the bucket's offsets, the
// counts of the lower-ranked CTAs -> its own cursors), then every CTA scatters its slice.  Clusters walk the size-ordered bucket list
// and stop at the first bucket that is not long.
template <int THREADS, int ITEMS>
__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(THREADS)
    k_l2_bucket_cluster(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ l1_off, const uint32_t *__restrict__ p2,
                        const uint32_t *__restrict__ sb_base, const uint64_t *__restrict__ splitters, uint16_t *__restrict__ ids,
                        uint32_t *__restrict__ sub_off, uint64_t *__restrict__ out_keys, const uint32_t *__restrict__ order, int nb1, uint32_t giant)
{
    namespace cg = cooperative_groups;
    constexpr int TILE = THREADS * ITEMS;
    constexpr int PER = SC_MAX_P2 / THREADS;
    constexpr uint32_t CL = 8;
    __shared__ uint64_t spl[SC_MAX_P2];
    __shared__ uint32_t cnt[SC_MAX_P2], cur[SC_MAX_P2];
    __shared__ uint32_t ws[33];
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t crank = cluster.block_rank();
    const uint32_t n_clusters = gridDim.x / CL;
    for (uint32_t g = blockIdx.x / CL; g < uint32_t(nb1); g += n_clusters)
    {
        const int b = int(order[g]);
        const uint32_t off = l1_off[b], nbk = l1_off[b + 1] - off;
        if (nbk <= giant) break; // the list is sorted by decreasing size; uniform over the cluster
        const uint32_t np = p2[b];
        const uint32_t sb0 = sb_base[b];
        for (uint32_t i = threadIdx.x; i < np; i += THREADS)
        {
            cnt[i] = 0;
            if (i + 1 < np) spl[i] = splitters[size_t(b) * SC_MAX_P2 + i];
        }
        __syncthreads();
        // this CTA's slice of the bucket, in whole tiles
        const uint32_t tiles = (nbk + TILE - 1) / TILE;
        const uint32_t lo = uint32_t((uint64_t(tiles) * crank) / CL) * TILE, hi = min(nbk, uint32_t((uint64_t(tiles) * (crank + 1)) / CL) * TILE);
        for (uint32_t base = lo; base < hi; base += TILE)
        {
            uint64_t k[ITEMS];
#pragma unroll
            for (int j = 0; j < ITEMS; ++j)
            {
                const uint32_t i = base + uint32_t(j) * THREADS + threadIdx.x;
                k[j] = i < hi ? __ldg(keys + off + i) : EMPTY64;
            }
#pragma unroll
            for (int j = 0; j < ITEMS; ++j)
            {
                const uint32_t i = base + uint32_t(j) * THREADS + threadIdx.x;
                if (i < hi)
                {
                    const uint32_t sid = np > 1 ? sub_bucket_of(spl, np - 1, k[j] >> 3) : 0u;
                    ids[off + i] = uint16_t(sid);
                    atomicAdd(&cnt[sid], 1u);
                }
            }
        }
        cluster.sync(); // every CTA's histogram is complete and visible
        {
            uint32_t c[PER], pre[PER], sum = 0;
#pragma unroll
            for (int q = 0; q < PER; ++q)
            {
                const uint32_t i = threadIdx.x * PER + q;
                c[q] = 0; pre[q] = 0;
                if (i < np)
                    for (uint32_t r = 0; r < CL; ++r)
                    {
                        const uint32_t v = cluster.map_shared_rank(cnt, r)[i];
                        c[q] += v;
                        if (r < crank) pre[q] += v;
                    }
                sum += c[q];
            }
            uint32_t tot;
            uint32_t ex = block_exclusive_scan(sum, ws, &tot);
#pragma unroll
            for (int q = 0; q < PER; ++q)
            {
                const uint32_t i = threadIdx.x * PER + q;
                if (i < np)
                {
                    cur[i] = ex + pre[q];
                    if (crank == 0) sub_off[sb0 + i] = off + ex;
                }
                ex += c[q];
            }
        }
        cluster.sync(); // nobody reads the histograms any more (they are zeroed for the next bucket); cursors are in place
        for (uint32_t base = lo; base < hi; base += TILE)
        {
            uint64_t k[ITEMS];
            uint16_t sid[ITEMS];
#pragma unroll
            for (int j = 0; j < ITEMS; ++j)
            {
                const uint32_t i = base + uint32_t(j) * THREADS + threadIdx.x;
                k[j] = i < hi ? __ldg(keys + off + i) : EMPTY64;
                sid[j] = i < hi ? ids[off + i] : uint16_t(0);
            }
#pragma unroll
            for (int j = 0; j < ITEMS; ++j)
            {
                const uint32_t i = base + uint32_t(j) * THREADS + threadIdx.x;
                if (i < hi) out_keys[off + atomicAdd(&cur[sid[j]], 1u)] = k[j];
            }
        }
        __syncthreads();
    }
}

__global__ void k_set_u32_at(uint32_t *p, const uint32_t *idx, uint32_t v) { p[*idx] = v; }

// L1 buckets by decreasing size (longest first: the one-block-per-bucket pass then ends with the small ones). nb1 <= 4096.
__global__ void __launch_bounds__(1024) k_bucket_order(const uint32_t *__restrict__ l1_off, int nb1, uint32_t *__restrict__ order)
{
    __shared__ uint64_t kk[SC_MAX_NB1];
    __shared__ uint32_t vv[SC_MAX_NB1];
    int P = 2;
    while (P < nb1) P <<= 1;
    for (int i = threadIdx.x; i < P; i += blockDim.x)
    {
        kk[i] = i < nb1 ? uint64_t(0xFFFFFFFFu - (l1_off[i + 1] - l1_off[i])) : EMPTY64;
        vv[i] = uint32_t(i);
    }
    bitonic_sort_smem<true>(kk, vv, P);
    for (int i = threadIdx.x; i < nb1; i += blockDim.x) order[i] = vv[i];
}

// One block per sub-bucket.  keys[s..e) -> distinct ukeys, ascending, written back IN PLACE at keys[s..s+m), values at
// uvals[s..s+m); ucount[sb] = m.
//   1. stream the records through a shared-memory hash table (atomicCAS claims a slot, atomicAdd/atomicOr combine values)
//   2. compact the occupied slots
//   3. stable LSD radix sort (8-bit digits) of the m distinct keys on the bits that actually vary inside the sub-bucket;
//      ranking is atomic-free: every warp owns a contiguous slice and a private digit histogram, equal digits inside a
//      32-key group are resolved with __match_any_sync
constexpr int SC_RADIX_BITS = 8;
constexpr int SC_RADIX = 1 << SC_RADIX_BITS;
constexpr int SC_DEDUP_WARPS = SC_DEDUP_THREADS / 32;

// shared memory: table ht x (8 B key + 4 B value) + three u16 arrays of `cap` entries (index list ping-pong + ranks) + histograms
inline size_t dedup_smem_bytes(int ht, int cap, int threads = SC_DEDUP_THREADS)
{
    return size_t(ht) * 12 + size_t(cap) * 6 + size_t(threads / 32) * SC_RADIX * 2 + 64;
}

template <bool HAS_VAL>
__global__ void __launch_bounds__(1024) k_dedup_sort(uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals_in,
                                                                 uint32_t *__restrict__ uvals, const uint32_t *__restrict__ sub_off,
                                                                 const uint32_t *__restrict__ n_sub_ptr, uint32_t *__restrict__ ucount,
                                                                 int *__restrict__ overflow, const int SC_HT, const int SC_CAP,
                                                                 const uint32_t n_min, const uint32_t n_max,
                                                                 const uint32_t *__restrict__ list = nullptr, const uint32_t *__restrict__ list_count = nullptr)
{
    // size classes: this launch only handles sub-buckets with n_min < records <= n_max (records <= 0.85 * SC_HT can never
    // overflow the table whatever the duplication rate; the last class takes everything larger and reports overflow).
    // Keys and values never leave the hash table: what gets compacted and radix-sorted is the list of occupied SLOT INDICES
    // (2 bytes each), so a pass moves 2 B per key instead of 12 B and five blocks fit per SM.
    extern __shared__ unsigned char smem_raw[];
    unsigned long long *ht_key = reinterpret_cast<unsigned long long *>(smem_raw);
    uint32_t *ht_val = reinterpret_cast<uint32_t *>(ht_key + SC_HT);
    uint16_t *idxA = reinterpret_cast<uint16_t *>(ht_val + SC_HT);   // slot index list (SC_CAP entries)
    uint16_t *idxB = idxA + SC_CAP;                                  // ping-pong partner
    uint16_t *rank_s = idxB + SC_CAP;                                // per element rank inside (warp, digit)
    uint16_t *hist = rank_s + SC_CAP;                                // [warp][digit]
    __shared__ uint32_t m_s;
    __shared__ unsigned long long red_or, red_and;
    __shared__ uint32_t ws[33];

    // work items: either every sub-bucket (filtered by size class) or the entries of a work list (persistent blocks)
    const uint32_t n_items = list ? *list_count : *n_sub_ptr;
    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x)
    {
    const uint32_t sb = list ? list[item] : item;
    const uint32_t s = sub_off[sb], e = sub_off[sb + 1];
    if (e - s <= n_min || e - s > n_max) continue; // another size class (ucount was zeroed by the host; empty buckets stay 0)
    __syncthreads(); // previous item fully written out before the table is cleared
    for (int i = threadIdx.x; i < SC_HT; i += blockDim.x) { ht_key[i] = EMPTY64; ht_val[i] = 0; }
    if (threadIdx.x == 0) { m_s = 0; red_or = 0; red_and = EMPTY64; }
    __syncthreads();

    // ---- 1. hash-combine; the thread that claims a slot appends the slot index to the list (warp-aggregated cursor)
    unsigned long long t_or = 0, t_and = EMPTY64;
    for (uint32_t i0 = s; i0 < e; i0 += blockDim.x)
    {
        const uint32_t i = i0 + threadIdx.x;
        bool is_new = false;
        uint32_t slot = NONE32;
        if (i < e)
        {
            const uint64_t key = keys[i];
            const uint64_t uk = key >> 3;
            const uint32_t v = HAS_VAL ? vals_in[i] : (1u | (uint32_t(key & 7) << VAL_MARK_SHIFT));
            slot = uint32_t((uk * 0x9E3779B97F4A7C15ull) >> 40) & (SC_HT - 1);
            int probes = 0;
            while (true)
            {
                unsigned long long cur = ht_key[slot];
                if (cur == EMPTY64)
                {
                    cur = atomicCAS(&ht_key[slot], EMPTY64, (unsigned long long)uk);
                    if (cur == EMPTY64) { cur = uk; is_new = true; t_or |= uk; t_and &= uk; }
                }
                if (cur == uk) break;
                slot = (slot + 1) & (SC_HT - 1);
                if (++probes >= SC_HT) { atomicExch(overflow, 1); slot = NONE32; break; }
            }
            if (slot != NONE32)
            {
                const uint32_t old = atomicAdd(&ht_val[slot], v & VAL_COUNT_MASK);
                const uint32_t mk = v & ~VAL_COUNT_MASK;
                if ((old & mk) != mk) atomicOr(&ht_val[slot], mk);
            }
        }
        const unsigned newmask = __ballot_sync(0xFFFFFFFFu, is_new);
        if (newmask)
        {
            uint32_t basepos = 0;
            if ((threadIdx.x & 31) == 0) basepos = atomicAdd(&m_s, uint32_t(__popc(newmask)));
            basepos = __shfl_sync(0xFFFFFFFFu, basepos, 0);
            if (is_new)
            {
                const uint32_t pos = basepos + __popc(newmask & ((1u << (threadIdx.x & 31)) - 1));
                if (pos < uint32_t(SC_CAP)) idxA[pos] = uint16_t(slot); else atomicExch(overflow, 1);
            }
        }
    }
    // which key bits vary inside the sub-bucket (only those digits need a pass)
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
    {
        t_or |= __shfl_xor_sync(0xFFFFFFFFu, t_or, d);
        t_and &= __shfl_xor_sync(0xFFFFFFFFu, t_and, d);
    }
    if ((threadIdx.x & 31) == 0) { atomicOr(&red_or, t_or); atomicAnd(&red_and, t_and); }
    __syncthreads();
    const uint32_t m = min(m_s, uint32_t(SC_CAP));
    const unsigned long long varying = red_or & ~red_and; // bits that differ between at least two keys
    const int top_bit = varying ? 63 - __clzll((long long)varying) : -1;

    // ---- 2. stable LSD radix sort of the index list by the keys they point to; ping-pong idxA <-> idxB
    uint16_t *src = idxA, *dst = idxB;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t n_warps = blockDim.x >> 5;
    const uint32_t chunk = ((m + n_warps * 32 - 1) / (n_warps * 32)) * 32; // per-warp slice, multiple of 32
    const uint32_t w_begin = min(m, warp * chunk), w_end = min(m, w_begin + chunk);
    uint16_t *my_hist = hist + warp * SC_RADIX;
    for (int shift = 0; shift <= top_bit; shift += SC_RADIX_BITS)
    {
        if (((varying >> shift) & (SC_RADIX - 1)) == 0) continue; // this digit is constant: nothing to do (uniform branch)
        for (int i = threadIdx.x; i < int(n_warps) * SC_RADIX / 2; i += blockDim.x) reinterpret_cast<uint32_t *>(hist)[i] = 0;
        __syncthreads();
        for (uint32_t g = w_begin; g < w_end; g += 32)
        {
            const uint32_t i = g + lane;
            const bool valid = i < w_end;
            const unsigned vmask = __ballot_sync(0xFFFFFFFFu, valid);
            if (valid)
            {
                const uint32_t d = uint32_t(ht_key[src[i]] >> shift) & (SC_RADIX - 1);
                const unsigned peers = __match_any_sync(vmask, d);
                const int leader = __ffs(peers) - 1;
                uint32_t old = 0;
                if (int(lane) == leader) { old = my_hist[d]; my_hist[d] = uint16_t(old + __popc(peers)); }
                old = __shfl_sync(peers, old, leader);
                rank_s[i] = uint16_t(old + __popc(peers & ((1u << lane) - 1)));
            }
            __syncwarp();
        }
        __syncthreads();
        {   // exclusive scan over (digit major, warp minor): thread d < SC_RADIX owns digit d (blockDim >= SC_RADIX)
            const uint32_t d = threadIdx.x;
            uint32_t run = 0;
            if (d < SC_RADIX)
                for (uint32_t w = 0; w < n_warps; ++w) run += hist[w * SC_RADIX + d];
            uint32_t total;
            uint32_t basev = block_exclusive_scan(run, ws, &total);
            if (d < SC_RADIX)
                for (uint32_t w = 0; w < n_warps; ++w)
                {
                    const uint32_t c = hist[w * SC_RADIX + d];
                    hist[w * SC_RADIX + d] = uint16_t(basev);
                    basev += c;
                }
        }
        __syncthreads();
        for (uint32_t g = w_begin; g < w_end; g += 32)
        {
            const uint32_t i = g + lane;
            if (i < w_end)
            {
                const uint16_t sl = src[i];
                const uint32_t d = uint32_t(ht_key[sl] >> shift) & (SC_RADIX - 1);
                dst[uint32_t(my_hist[d]) + rank_s[i]] = sl;
            }
        }
        __syncthreads();
        uint16_t *t = src; src = dst; dst = t;
    }
    for (uint32_t i = threadIdx.x; i < m; i += blockDim.x)
    {
        const uint16_t sl = src[i];
        keys[s + i] = ht_key[sl];
        uvals[s + i] = ht_val[sl];
    }
    if (threadIdx.x == 0) ucount[sb] = m;
    }
}

__global__ void __launch_bounds__(256) k_compact_uniques(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ uvals,
                                                         const uint32_t *__restrict__ sub_off, const uint32_t *__restrict__ u_off,
                                                         const uint32_t *__restrict__ ucount, const uint32_t *__restrict__ n_sub_ptr,
                                                         uint64_t *__restrict__ out_keys, uint32_t *__restrict__ out_vals)
{
    const uint32_t nsb = *n_sub_ptr;
    for (uint32_t sb = blockIdx.x; sb < nsb; sb += gridDim.x)
    {
        const uint32_t m = ucount[sb], src = sub_off[sb], dst = u_off[sb];
        for (uint32_t i = threadIdx.x; i < m; i += blockDim.x)
        {
            out_keys[dst + i] = keys[src + i];
            out_vals[dst + i] = uvals[src + i];
        }
    }
}

// ---- oversized sub-buckets (more keys than the largest sort class: the sampling tail of a very long L1 bucket, or one heavily duplicated
// key).  They used to stream through a shared-memory hash table, whose capacity made the whole run fail -- rarely, and depending on the
// order in which the keys happened to arrive -- when such a sub-bucket held more than ~4 k distinct keys.  Now: no capacity anywhere.  The
// listed sub-buckets are sorted as segments in global memory (cub::DeviceSegmentedRadixSort, library code, a few hundred k keys at most in
// practice) and run-length encoded by k_tail_dedup.
__global__ void __launch_bounds__(256) k_tail_offsets(const uint32_t *__restrict__ list, const uint32_t *__restrict__ count, const uint32_t *__restrict__ sub_off,
                                                      int *__restrict__ seg_begin, int *__restrict__ seg_end)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; // the grid covers the list rounded up to 256
    const uint32_t n = *count;
    int b = 0, e = 0;
    if (i < n) { const uint32_t sb = list[i]; b = int(sub_off[sb]); e = int(sub_off[sb + 1]); }
    seg_begin[i] = b; seg_end[i] = e;
}

// sorted[s..e) (full keys, ascending) -> distinct ukeys at keys[s..s+m), (count | mark << 29) at uvals[s..s+m), ucount[sb] = m
__global__ void __launch_bounds__(256) k_tail_dedup(const uint64_t *__restrict__ sorted, uint64_t *__restrict__ keys, uint32_t *__restrict__ uvals,
                                                    const uint32_t *__restrict__ sub_off, const uint32_t *__restrict__ list, const uint32_t *__restrict__ count,
                                                    uint32_t *__restrict__ ucount)
{
    __shared__ uint32_t ws[33];
    const uint32_t n_seg = *count;
    for (uint32_t it = blockIdx.x; it < n_seg; it += gridDim.x)
    {
        const uint32_t sb = list[it], s = sub_off[sb], n = sub_off[sb + 1] - s;
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) uvals[s + i] = 0;
        __syncthreads();
        uint32_t done = 0; // runs emitted by the chunks before this one
        for (uint32_t c = 0; c < n; c += blockDim.x)
        {
            const uint32_t i = c + threadIdx.x;
            uint64_t x = EMPTY64;
            uint32_t head = 0;
            if (i < n)
            {
                x = sorted[s + i];
                head = i == 0 || (sorted[s + i - 1] >> 3) != (x >> 3);
            }
            uint32_t tot;
            const uint32_t ex = block_exclusive_scan(head, ws, &tot);
            if (i < n)
            {
                const uint32_t r = done + ex + head - 1; // index of the run this key belongs to
                if (head) keys[s + r] = x >> 3;
                atomicAdd(&uvals[s + r], 1u);
                if (uint32_t(x) & 7u) atomicOr(&uvals[s + r], (uint32_t(x) & 7u) << VAL_MARK_SHIFT);
            }
            done += tot;
            __syncthreads();
        }
        if (threadIdx.x == 0) ucount[sb] = done;
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------------
struct SortCombineWorkspace
{
    DevBuf keysA, valsA, valsB, uvals_sparse, small, splitters, sub_cnt, sub_off, ucount, u_off, scan_scratch, cls_list, cls_count, ids, order;
    DevBuf tail_begin, tail_end, tail_tmp; // oversized sub-buckets: segment offsets + cub temp storage
    // side stream of the long-bucket cluster kernel (runs beside the block-per-bucket kernel: disjoint buckets)
    cudaStream_t side = nullptr;
    cudaEvent_t fork_ev = nullptr, join_ev = nullptr;
    int side_dev = -1, max_clusters = 0;
};

struct SortCombineStats
{
    unsigned launches = 0;
    unsigned dedup_launches = 0;
    float dedup_ms = 0;
};

inline int choose_l1_bits(size_t n)
{
    int b = ceil_log2_u64(div_up<uint64_t>(n ? n : 1, uint64_t(sc_tuning().l1_target)));
    if (b < 6) b = 6;
    if (b > 12) b = 12;
    return b;
}

// keys_in is CONSUMED (used as the L2 scatter target is NOT allowed: it is read-only here) -- buffers:
//   keys_in --L1 scatter--> keysA --L2 scatter--> keys_tmp (caller-provided, n entries; may alias keys_in) --dedup in place-->
//   gather --> out_keys/out_vals (caller-provided, capacity >= n_u; sized n by the caller or after reading the count).
// l1_hist: optional precomputed L1 histogram (device, nb1 entries) -- the fill kernel fuses it.
// Returns device pointer to n_u (uint32).  Sets *overflow_flag (device int) on table overflow.
class SortCombine
{
public:
    SortCombineWorkspace ws;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;

    // Optional input of the NEXT run(): the keys live in per-block regions of the fill kernel instead of one dense array
    // (keys_in is ignored, l1_hist_pre must be given).  region_tiles: device scratch word.
    const KeyTile *src_tiles = nullptr;
    uint32_t n_src_regions = 0;
    uint32_t *src_region_tiles = nullptr;
    void set_regions(const KeyTile *tiles, uint32_t n_regions, uint32_t *region_tiles)
    {
        src_tiles = tiles; n_src_regions = n_regions; src_region_tiles = region_tiles;
    }

    ~SortCombine()
    {
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
    }

    static void configure()
    {
        // the attribute is per device: keep one flag per device ordinal (a process may drive several GPUs)
        static bool done[64] = {false};
        int dev = 0;
        DGE_CUDA(cudaGetDevice(&dev));
        if (dev < 64 && done[dev]) return;
        DGE_CUDA(cudaFuncSetAttribute(k_splitters, cudaFuncAttributeMaxDynamicSharedMemorySize, SC_SAMPLE * 8));
        DGE_CUDA(cudaFuncSetAttribute(k_dedup_sort<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(dedup_smem_bytes(SC_HT_MAX, SC_HT_MAX, 1024))));
        DGE_CUDA(cudaFuncSetAttribute(k_dedup_sort<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(dedup_smem_bytes(SC_HT_MAX, SC_HT_MAX, 1024))));
        if (dev < 64) done[dev] = true;
    }

    // Layout of ws.small (uint32): [hist nb1+1][l1_off nb1+1][cursor nb1+1][p2 nb1+1][sb_base nb1+1][tile_base nb1+1]
    const uint32_t *run(const uint64_t *keys_in, const uint32_t *vals_in, size_t n, int key_bits, int l1_bits,
                        const uint32_t *l1_hist_pre, uint64_t *keys_tmp, uint64_t *out_keys, uint32_t *out_vals,
                        int *overflow_flag, cudaStream_t st, SortCombineStats *stats)
    {
        configure();
        int cur_dev = 0; // cudaFuncAttributeMaxDynamicSharedMemorySize is per device: one flag per ordinal
        DGE_CUDA(cudaGetDevice(&cur_dev));
        cur_dev &= 63;
        if (!ev0) { DGE_CUDA(cudaEventCreate(&ev0)); DGE_CUDA(cudaEventCreate(&ev1)); }
        const bool has_val = vals_in != nullptr;
        const int nb1 = 1 << l1_bits;
        const int shift = key_bits - l1_bits;
        const size_t stride = size_t(nb1) + 1;
        ws.small.reserve(stride * 6 * sizeof(uint32_t));
        uint32_t *hist = ws.small.as<uint32_t>(), *l1_off = hist + stride, *cursor = l1_off + stride, *p2 = cursor + stride,
                 *sb_base = p2 + stride, *tile_base = sb_base + stride;
        const int SC_TARGET = sc_tuning().target, SC_HT = sc_tuning().ht;
        const size_t nsb_bound = n / SC_TARGET + size_t(nb1) + 1;
        ws.keysA.reserve(n * 8);
        if (has_val) { ws.valsA.reserve(n * 4); ws.valsB.reserve(n * 4); }
        ws.uvals_sparse.reserve(n * 4);
        ws.splitters.reserve(size_t(nb1) * SC_MAX_P2 * 8);
        ws.sub_cnt.reserve((nsb_bound + 1) * 4);
        ws.sub_off.reserve((nsb_bound + 1) * 4);
        ws.ucount.reserve((nsb_bound + 1) * 4);
        ws.u_off.reserve((nsb_bound + 1) * 4);
        ws.scan_scratch.reserve(scan_scratch_elems(nsb_bound + 1) * 4);
        ws.cls_list.reserve((nsb_bound + 1) * 4 * MS_CLASSES);
        ws.cls_count.reserve(MS_CLASSES * 4);
        unsigned &L = stats->launches;
        const bool trace = std::getenv("DGE_TRACE") != nullptr;
        auto t_prev = std::chrono::steady_clock::now();
        auto mark = [&](const char *what) {
            if (!trace) return;
            cudaStreamSynchronize(st);
            auto t1 = std::chrono::steady_clock::now();
            fprintf(stderr, "[dge]   sc(n=%zu) %-18s %8.3f ms\n", n, what, std::chrono::duration<double, std::milli>(t1 - t_prev).count());
            t_prev = t1;
        };
        mark("alloc");

        // ---- L1
        if (l1_hist_pre)
            DGE_CUDA(cudaMemcpyAsync(hist, l1_hist_pre, nb1 * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
        else
        {
            DGE_CUDA(cudaMemsetAsync(hist, 0, stride * sizeof(uint32_t), st));
            unsigned g = unsigned(std::min<size_t>(div_up(n, size_t(SC_THREADS * 8)), 148 * 8));
            k_l1_hist<<<g ? g : 1, SC_THREADS, 0, st>>>(keys_in, n, shift, nb1, hist); ++L;
        }
        DGE_CUDA(cudaMemsetAsync(hist + nb1, 0, sizeof(uint32_t), st));
        device_exclusive_scan(hist, l1_off, stride, ws.scan_scratch.as<uint32_t>(), st, &L);
        uint64_t *keysA = ws.keysA.as<uint64_t>();
        DGE_CUDA(cudaMemcpyAsync(cursor, l1_off, stride * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
        static const int shape = std::getenv("DGE_TILE") ? atoi(std::getenv("DGE_TILE")) : 2;
        static const int staged = std::getenv("DGE_STAGED") ? atoi(std::getenv("DGE_STAGED")) : 1; // bit0: L1, bit1: L2 (measured: staging pays at L1 only)
        static const int stile = std::getenv("DGE_STILE") ? atoi(std::getenv("DGE_STILE")) : 2;    // staged tile shape (1024 x 8 measured best)
        size_t tile = 0;
        if (src_tiles)
        {   // keys straight from the fill kernel's regions (1024 x 8 staged tiles)
            if (has_val || !l1_hist_pre) throw std::runtime_error("region input needs a precomputed histogram and no values");
            constexpr int T = 1024, I = 8;
            const size_t smem = size_t(T) * I * 8 + size_t(nb1) * 8;
            static bool attr_done[64] = {};
            if (!attr_done[cur_dev]) { DGE_CUDA(cudaFuncSetAttribute(k_l1_scatter_regions<T, I>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr_done[cur_dev] = true; }
            const unsigned g_tiles = unsigned(n / (size_t(T) * I) + n_src_regions + 1); // upper bound; blocks beyond the real count exit
            k_l1_scatter_regions<T, I><<<g_tiles, T, smem, st>>>(src_tiles, src_region_tiles, shift, nb1, cursor, keysA);
            src_tiles = nullptr;
        }
        else if (!has_val && (staged & 1))
        {
#define DGE_L1S(IDX, T, I)                                                                                                         \
            if (stile == IDX)                                                                                                      \
            {                                                                                                                      \
                const size_t smem = size_t(T) * I * 8 + size_t(nb1) * 8;                                                           \
                static bool attr_done[64] = {};                                                                                    \
                if (!attr_done[cur_dev]) { DGE_CUDA(cudaFuncSetAttribute(k_l1_scatter_staged<T, I>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr_done[cur_dev] = true; } \
                k_l1_scatter_staged<T, I><<<unsigned(div_up(n, size_t(T) * I)), T, smem, st>>>(keys_in, n, shift, nb1, cursor, keysA); \
            }
            DGE_L1S(0, 512, 16) DGE_L1S(1, 256, 16) DGE_L1S(2, 1024, 8) DGE_L1S(3, 256, 8)
#undef DGE_L1S
        }
        else
        {
#define DGE_L1(IDX, T, I)                                                                                                          \
        if (shape == IDX)                                                                                                          \
        {                                                                                                                          \
            const unsigned g_tiles = unsigned(div_up(n, size_t(T) * I));                                                           \
            if (has_val) k_l1_scatter<true, T, I><<<g_tiles, T, 0, st>>>(keys_in, vals_in, n, shift, nb1, cursor, keysA, ws.valsA.as<uint32_t>()); \
            else k_l1_scatter<false, T, I><<<g_tiles, T, 0, st>>>(keys_in, nullptr, n, shift, nb1, cursor, keysA, nullptr);        \
        }
        DGE_L1(0, 512, 16) DGE_L1(1, 256, 16) DGE_L1(2, 256, 8) DGE_L1(3, 128, 16) DGE_L1(4, 512, 8) DGE_L1(5, 1024, 8)
#undef DGE_L1
        }
        // the L2 passes walk the L1 buckets in tiles of this many keys
        const bool l2_staged = !has_val && (staged & 2);
        static const int s2tile = std::getenv("DGE_S2TILE") ? atoi(std::getenv("DGE_S2TILE")) : 0;
        if (l2_staged) tile = s2tile == 0 ? 256 * 16 : s2tile == 1 ? 256 * 8 : 512 * 16;
        else
        {
            const int T[6] = {512, 256, 256, 128, 512, 1024}, I[6] = {16, 16, 8, 16, 8, 8};
            if (shape < 0 || shape > 5) throw std::runtime_error("DGE_TILE out of range");
            tile = size_t(T[shape]) * I[shape];
        }
        ++L;
        mark("l1 hist+scatter");

        // ---- L2
        const size_t tiles_bound = n / tile + size_t(nb1) + 1;
        k_l1_plan<<<1, 1024, 0, st>>>(l1_off, nb1, uint32_t(SC_TARGET), uint32_t(tile), p2, sb_base, tile_base); ++L;
        k_splitters<<<nb1, SC_THREADS, SC_SAMPLE * 8, st>>>(keysA, l1_off, p2, ws.splitters.as<uint64_t>(), uint32_t(sc_tuning().sample)); ++L;
        mark("plan+splitters");
        uint32_t *sub_cnt = ws.sub_cnt.as<uint32_t>(), *sub_off = ws.sub_off.as<uint32_t>();
        static const int l2_bucket = std::getenv("DGE_L2_BUCKET") ? atoi(std::getenv("DGE_L2_BUCKET")) : 3; // 0 = the two-launch tile path
        if (!has_val && l2_bucket)
        {
            ws.ids.reserve(n * 2 + 64);
            ws.order.reserve(size_t(nb1) * 4);
            uint32_t *order = ws.order.as<uint32_t>();
            k_bucket_order<<<1, 1024, 0, st>>>(l1_off, nb1, order);
            ++L;
            if (l2_bucket == 1) k_l2_bucket<512, 8><<<nb1, 512, 0, st>>>(keysA, l1_off, p2, sb_base, ws.splitters.as<uint64_t>(), ws.ids.as<uint16_t>(), sub_off, keys_tmp, order);
            else if (l2_bucket == 2) k_l2_bucket<1024, 4><<<nb1, 1024, 0, st>>>(keysA, l1_off, p2, sb_base, ws.splitters.as<uint64_t>(), ws.ids.as<uint16_t>(), sub_off, keys_tmp, order);
            else if (l2_bucket == 3)
            {
                // long buckets: shared by 8-CTA clusters on a side stream, beside the block-per-bucket kernel (disjoint buckets, disjoint outputs)
                const uint32_t giant = std::getenv("DGE_L2_GIANT") ? uint32_t(atoll(std::getenv("DGE_L2_GIANT"))) : 786432u; // read per run: the tests lower it
                const bool clusters = giant != 0xFFFFFFFFu && giant != 0;
                if (clusters)
                {
                    int dev = 0;
                    DGE_CUDA(cudaGetDevice(&dev));
                    if (ws.side_dev != dev)
                    {
                        DGE_CUDA(cudaStreamCreateWithFlags(&ws.side, cudaStreamNonBlocking));
                        DGE_CUDA(cudaEventCreateWithFlags(&ws.fork_ev, cudaEventDisableTiming));
                        DGE_CUDA(cudaEventCreateWithFlags(&ws.join_ev, cudaEventDisableTiming));
                        // clusters that can be resident together (8 CTAs must share a GPC): more would only queue behind them
                        cudaLaunchConfig_t lc = {};
                        lc.gridDim = dim3(8 * 64); lc.blockDim = dim3(1024); lc.dynamicSmemBytes = 0;
                        cudaLaunchAttribute at[1];
                        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 8; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                        lc.attrs = at; lc.numAttrs = 1;
                        int mc = 0;
                        if (cudaOccupancyMaxActiveClusters(&mc, k_l2_bucket_cluster<1024, 8>, &lc) != cudaSuccess || mc < 1) { (void)cudaGetLastError(); mc = 16; }
                        ws.max_clusters = std::min(mc, 64);
                        ws.side_dev = dev;
                    }
                    DGE_CUDA(cudaEventRecord(ws.fork_ev, st));
                    DGE_CUDA(cudaStreamWaitEvent(ws.side, ws.fork_ev, 0));
                    k_l2_bucket_cluster<1024, 8><<<ws.max_clusters * 8, 1024, 0, ws.side>>>(keysA, l1_off, p2, sb_base, ws.splitters.as<uint64_t>(), ws.ids.as<uint16_t>(), sub_off, keys_tmp, order, nb1, giant);
                    DGE_CUDA(cudaEventRecord(ws.join_ev, ws.side));
                    ++L;
                }
                k_l2_bucket<1024, 8><<<nb1, 1024, 0, st>>>(keysA, l1_off, p2, sb_base, ws.splitters.as<uint64_t>(), ws.ids.as<uint16_t>(), sub_off, keys_tmp, order, clusters ? giant : 0xFFFFFFFFu);
                if (clusters) DGE_CUDA(cudaStreamWaitEvent(st, ws.join_ev, 0));
            }
            else k_l2_bucket<1024, 4><<<nb1, 1024, 0, st>>>(keysA, l1_off, p2, sb_base, ws.splitters.as<uint64_t>(), ws.ids.as<uint16_t>(), sub_off, keys_tmp, nullptr);
            k_set_u32_at<<<1, 1, 0, st>>>(sub_off, sb_base + nb1, uint32_t(n));
            L += 2;
            mark("l2 bucket pass");
        }
        else
        {
        DGE_CUDA(cudaMemsetAsync(sub_cnt, 0, (nsb_bound + 1) * 4, st));
#define DGE_L2H(COND, T, I)                                                                                                        \
        if (COND) k_l2_pass<false, false, T, I><<<unsigned(tiles_bound), T, 0, st>>>(keysA, nullptr, l1_off, nb1, p2, sb_base, tile_base, \
                                                                                      ws.splitters.as<uint64_t>(), sub_cnt, nullptr, nullptr);
        if (l2_staged) { DGE_L2H(s2tile == 0, 256, 16) DGE_L2H(s2tile == 1, 256, 8) DGE_L2H(s2tile == 2, 512, 16) }
        else { DGE_L2H(shape == 0, 512, 16) DGE_L2H(shape == 1, 256, 16) DGE_L2H(shape == 2, 256, 8) DGE_L2H(shape == 3, 128, 16) DGE_L2H(shape == 4, 512, 8) DGE_L2H(shape == 5, 1024, 8) }
#undef DGE_L2H
        ++L;
        mark("l2 hist");
        device_exclusive_scan(sub_cnt, sub_off, nsb_bound + 1, ws.scan_scratch.as<uint32_t>(), st, &L);
        // cursors = copy of offsets (sub_cnt reused)
        DGE_CUDA(cudaMemcpyAsync(sub_cnt, sub_off, (nsb_bound + 1) * 4, cudaMemcpyDeviceToDevice, st));
        if (l2_staged)
        {
#define DGE_L2SS(IDX, T, I)                                                                                                        \
            if (s2tile == IDX)                                                                                                     \
            {                                                                                                                      \
                const size_t smem = size_t(SC_MAX_P2) * 16 + size_t(T) * I * 10;                                                   \
                static bool attr_done[64] = {};                                                                                    \
                if (!attr_done[cur_dev]) { DGE_CUDA(cudaFuncSetAttribute(k_l2_scatter_staged<T, I>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr_done[cur_dev] = true; } \
                k_l2_scatter_staged<T, I><<<unsigned(tiles_bound), T, smem, st>>>(keysA, l1_off, nb1, p2, sb_base, tile_base, ws.splitters.as<uint64_t>(), sub_cnt, keys_tmp); \
            }
            DGE_L2SS(0, 256, 16) DGE_L2SS(1, 256, 8) DGE_L2SS(2, 512, 16)
#undef DGE_L2SS
        }
        else
        {
#define DGE_L2S(IDX, T, I)                                                                                                         \
        if (shape == IDX)                                                                                                          \
        {                                                                                                                          \
            if (has_val) k_l2_pass<true, true, T, I><<<unsigned(tiles_bound), T, 0, st>>>(keysA, ws.valsA.as<uint32_t>(), l1_off, nb1, p2, sb_base, tile_base, \
                                                                                           ws.splitters.as<uint64_t>(), sub_cnt, keys_tmp, ws.valsB.as<uint32_t>()); \
            else k_l2_pass<true, false, T, I><<<unsigned(tiles_bound), T, 0, st>>>(keysA, nullptr, l1_off, nb1, p2, sb_base, tile_base, \
                                                                                    ws.splitters.as<uint64_t>(), sub_cnt, keys_tmp, nullptr); \
        }
        DGE_L2S(0, 512, 16) DGE_L2S(1, 256, 16) DGE_L2S(2, 256, 8) DGE_L2S(3, 128, 16) DGE_L2S(4, 512, 8) DGE_L2S(5, 1024, 8)
#undef DGE_L2S
        }
        ++L;
        mark("l2 scan+scatter");

        }
        // ---- L3: dedup + sort per sub-bucket
        const uint32_t *n_sub_ptr = sb_base + nb1;
        uint32_t *ucount = ws.ucount.as<uint32_t>(), *u_off = ws.u_off.as<uint32_t>();
        DGE_CUDA(cudaMemsetAsync(ucount, 0, (nsb_bound + 1) * 4, st));
        DGE_CUDA(cudaEventRecord(ev0, st));
        {
            auto launch = [&](int ht, uint32_t lo, uint32_t hi) {
                const int thr = sc_tuning().dedup_threads;
                const int cap = hi == 0xFFFFFFFFu ? ht : int(((hi + 31) / 32) * 32); // distinct keys <= records <= hi in a bounded class
                if (has_val)
                    k_dedup_sort<true><<<unsigned(nsb_bound), thr, dedup_smem_bytes(ht, cap, thr), st>>>(keys_tmp, ws.valsB.as<uint32_t>(), ws.uvals_sparse.as<uint32_t>(),
                                                                                                               sub_off, n_sub_ptr, ucount, overflow_flag, ht, cap, lo, hi);
                else
                    k_dedup_sort<false><<<unsigned(nsb_bound), thr, dedup_smem_bytes(ht, cap, thr), st>>>(keys_tmp, nullptr, ws.uvals_sparse.as<uint32_t>(),
                                                                                                                sub_off, n_sub_ptr, ucount, overflow_flag, ht, cap, lo, hi);
                ++L;
            };
            static const bool use_hash = std::getenv("DGE_HASH_DEDUP") != nullptr;
            static const int ms_items = std::getenv("DGE_MS_ITEMS") ? atoi(std::getenv("DGE_MS_ITEMS")) : 16;
            static const int ms_bps = std::getenv("DGE_MS_BPS") ? atoi(std::getenv("DGE_MS_BPS")) : 128; // grid >> resident blocks: late blocks balance the tail
            if (!has_val && !use_hash)
            {   // comparison sort + run detection on per-size-class work lists (persistent blocks); anything larger than the
                // biggest class (sampling tail, one heavily duplicated key) streams through the hash-table kernel
                uint32_t *uv = ws.uvals_sparse.as<uint32_t>();
                uint32_t *cc = ws.cls_count.as<uint32_t>(), *cl = ws.cls_list.as<uint32_t>();
                const size_t ls = nsb_bound + 1;
                DGE_CUDA(cudaMemsetAsync(cc, 0, MS_CLASSES * 4, st));
                const uint32_t c0 = 64u * ms_items, c1 = 2 * c0, c2 = 4 * c0;
                static const bool warp_class = std::getenv("DGE_MS_WARP") != nullptr; // measured neutral at C2 (the sort is ALU-bound, not barrier-bound): off
                const uint32_t cw = warp_class && ms_items == 16 ? 32u * 16u : 0u; // sub-buckets of <= 512 keys: one warp each
                k_classify_sub<<<148 * 4, 256, 0, st>>>(sub_off, n_sub_ptr, cw, c0, c1, c2, cc, cl, ls);
                auto grid = [&](int threads, int dflt) { return unsigned(148 * (ms_bps > 0 ? std::max(1, ms_bps * 64 / threads) : dflt)); };
                // merge-step formulation: 0 = one element of look-ahead per side, 1 = predicated without look-ahead (default: fewest
                // ALU instructions among the proven ones; the sort is ALU-pipe bound), 2 = 1 + sentinel slot instead of a predicated load
                static const int mv = std::getenv("DGE_MS_VARIANT") ? atoi(std::getenv("DGE_MS_VARIANT")) : 1;
                if (cw) k_sort_dedup_warp<4, 16><<<grid(128, 6), 128, 0, st>>>(keys_tmp, uv, sub_off, cl, cc, ucount);
                // bulk-async staging of the sub-bucket + prefetch of the next work item (k_sort_dedup<.., BULK = true>): measured on B200 at C2
                // 6.74 ms against 6.53 ms for the plain coalesced loads (profiles/r2_sort_dedup_bulk_ab.txt) -- with 12 resident blocks per SM the
                // load latency is already hidden and the kernel is ALU-bound, so the extra barrier and shared memory cost more than the
                // saved LDG/STS issue slots.  Kept selectable (DGE_MS_BULK=1), off by default.
                static const bool ms_bulk = std::getenv("DGE_MS_BULK") && atoi(std::getenv("DGE_MS_BULK")) == 1;
#define DGE_MS_NB(T, I, V, C, DFLT) k_sort_dedup<T, I, V, false><<<grid(T, DFLT), T, 0, st>>>(keys_tmp, uv, sub_off, cl + (C + 1) * ls, cc + (C + 1), ucount)
#define DGE_MS(T, I, V, C, DFLT)                                                                                                        \
                if (ms_bulk) k_sort_dedup<T, I, V, true><<<grid(T, DFLT), T, 0, st>>>(keys_tmp, uv, sub_off, cl + (C + 1) * ls, cc + (C + 1), ucount); \
                else DGE_MS_NB(T, I, V, C, DFLT)
                // (the 4096-key class is a handful of sub-buckets and would need 66 KB of staging: plain loads)
                if (ms_items == 16 && mv == 0) { DGE_MS(64, 16, 0, 0, 12); DGE_MS(128, 16, 0, 1, 6); DGE_MS_NB(256, 16, 0, 2, 3); }
                else if (ms_items == 16 && mv == 2) { DGE_MS(64, 16, 2, 0, 12); DGE_MS(128, 16, 2, 1, 6); DGE_MS_NB(256, 16, 2, 2, 3); }
                else if (ms_items == 16) { DGE_MS(64, 16, 1, 0, 12); DGE_MS(128, 16, 1, 1, 6); DGE_MS_NB(256, 16, 1, 2, 3); }
                else { DGE_MS(64, 8, 1, 0, 24); DGE_MS(128, 8, 1, 1, 12); DGE_MS_NB(256, 8, 1, 2, 6); }
#undef DGE_MS
#undef DGE_MS_NB
                L += 5;
                {   // oversized sub-buckets: segmented sort in global memory (the L1 buffer is free by now) + run-length encoding; no capacity limit.
                    // Their number comes back to the host here (one 4-byte read behind k_classify_sub, which finished long ago): the segmented
                    // sort is sized exactly and skipped when there is nothing to do (8 passes over empty segments cost 0.4 ms per step).
                    uint32_t n_tail = 0;
                    DGE_CUDA(cudaMemcpyAsync(&n_tail, cc + 4, 4, cudaMemcpyDeviceToHost, st));
                    DGE_CUDA(cudaStreamSynchronize(st));
                    if (n_tail)
                    {
                        const size_t padded = div_up(size_t(n_tail), size_t(256)) * 256;
                        ws.tail_begin.reserve(padded * 4); ws.tail_end.reserve(padded * 4);
                        int *tb = ws.tail_begin.as<int>(), *te = ws.tail_end.as<int>();
                        k_tail_offsets<<<unsigned(padded / 256), 256, 0, st>>>(cl + 4 * ls, cc + 4, sub_off, tb, te);
                        if (n >= (size_t(1) << 31)) throw std::runtime_error("SortCombine: more than 2^31 keys in one run");
                        size_t bytes = 0;
                        DGE_CUDA(cub::DeviceSegmentedRadixSort::SortKeys(nullptr, bytes, keys_tmp, keysA, int(n), int(n_tail), tb, te, 0, 64, st));
                        ws.tail_tmp.reserve(bytes + 16);
                        DGE_CUDA(cub::DeviceSegmentedRadixSort::SortKeys(ws.tail_tmp.p, bytes, keys_tmp, keysA, int(n), int(n_tail), tb, te, 0, 64, st));
                        k_tail_dedup<<<unsigned(std::min<uint32_t>(n_tail, 148 * 4)), 256, 0, st>>>(keysA, keys_tmp, uv, sub_off, cl + 4 * ls, cc + 4, ucount);
                        L += 4;
                    }
                }
            }
            else
            {
                const uint32_t cut = uint32_t(SC_HT * 0.85);
                if (SC_HT < SC_HT_MAX)
                {
                    launch(SC_HT, 0u, cut);                    // the common class
                    launch(SC_HT_MAX, cut, 0xFFFFFFFFu);       // oversized sub-buckets (sampling tail, heavy duplicates)
                }
                else launch(SC_HT, 0u, 0xFFFFFFFFu);
            }
        }
        DGE_CUDA(cudaEventRecord(ev1, st));
        ++stats->dedup_launches;
        pending_dedup_event = true;
        mark("dedup_sort");
        const uint32_t *n_u_ptr = device_exclusive_scan(ucount, u_off, nsb_bound + 1, ws.scan_scratch.as<uint32_t>(), st, &L);
        k_compact_uniques<<<148 * 8, 256, 0, st>>>(keys_tmp, ws.uvals_sparse.as<uint32_t>(), sub_off, u_off, ucount, n_sub_ptr, out_keys, out_vals);
        ++L;
        DGE_LAUNCH_CHECK();
        mark("scan+compact");
        last_stats = stats;
        return n_u_ptr;
    }

    // call after the stream has been synchronised
    void collect_timing()
    {
        if (pending_dedup_event && last_stats)
        {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, ev0, ev1) == cudaSuccess) last_stats->dedup_ms += ms;
        }
        pending_dedup_event = false;
    }

private:
    bool pending_dedup_event = false;
    SortCombineStats *last_stats = nullptr;
};

} // namespace dge
