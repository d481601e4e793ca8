// synth.cu -- counter-based synthetic read generator and the barcode-hash router (multi-GPU partition step).
// Record i is a pure function of (seed, i) and the tables, so host (dropest_b200/synth.py) and device produce the same
// stream bit for bit; the oracle consumes the host copy, the device pipeline the device copy.
#include "../../include/dropest_b200.h"
#include "common.cuh"

#include <vector>

using namespace dge;

namespace
{

__host__ __device__ inline uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

struct SynthDev
{
    uint64_t seed;
    uint32_t n_cells, n_genes, cb_len, umi_len;
    const uint64_t *cell_cdf, *cell_barcode, *cell_reads, *gene_cdf;
    const uint32_t *gene_weight;
    uint32_t cb_error_ppm, intergenic_ppm, intron_ppm, not_annotated_ppm, reads_per_umi;
};

// first index with cdf[idx] >= r  (cdf inclusive, last entry UINT64_MAX)
__device__ __forceinline__ uint32_t cdf_search(const uint64_t *__restrict__ cdf, uint32_t n, uint64_t r)
{
    uint32_t lo = 0, hi = n - 1;
    while (lo < hi)
    {
        uint32_t mid = (lo + hi) >> 1;
        if (cdf[mid] < r) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256) k_synth(SynthDev p, uint64_t first, uint64_t count, dge_record16 *__restrict__ out)
{
    for (uint64_t t = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < count; t += uint64_t(gridDim.x) * blockDim.x)
    {
        const uint64_t i = first + t;
        const uint64_t base = splitmix64(p.seed ^ (i * 0xD6E8FEB86659FD93ull));
        const uint64_t r0 = splitmix64(base + 0), r1 = splitmix64(base + 1), r2 = splitmix64(base + 2), r3 = splitmix64(base + 3),
                       r4 = splitmix64(base + 4);
        const uint32_t c = cdf_search(p.cell_cdf, p.n_cells, r0);
        const uint32_t g = cdf_search(p.gene_cdf, p.n_genes, r1);
        uint64_t pool = ((p.cell_reads[c] * uint64_t(p.gene_weight[g])) >> 32) / p.reads_per_umi;
        if (pool < 1) pool = 1;
        const uint64_t u_index = r2 % pool;
        const uint32_t umi = uint32_t(splitmix64(splitmix64((uint64_t(c) << 32) | g) + u_index)) & uint32_t((1ull << (2 * p.umi_len)) - 1);
        uint64_t cb = p.cell_barcode[c];
        if (r3 % 1000000ull < p.cb_error_ppm)
        {
            const uint32_t pos = uint32_t(r4 % p.cb_len);
            const uint32_t delta = 1 + uint32_t((r4 >> 8) % 3);
            const uint32_t shift = 2 * (p.cb_len - 1 - pos);
            const uint64_t old = (cb >> shift) & 3;
            cb = (cb & ~(3ull << shift)) | (((old + delta) & 3) << shift);
        }
        const bool intergenic = ((r3 >> 20) % 1000000ull) < p.intergenic_ppm;
        const uint64_t z = (r3 >> 40) % 1000000ull;
        const uint32_t mark = z < p.intron_ppm ? DGE_MARK_INTRON : (z < uint64_t(p.intron_ppm) + p.not_annotated_ppm ? DGE_MARK_NOT_ANNOTATED : DGE_MARK_EXON);
        dge_record16 r;
        r.key = (cb << 24) | umi;
        r.gene = (intergenic ? DGE_NO_GENE : g) | (mark << 24);
        r.read_idx = uint32_t(i);
        out[t] = r;
    }
}

__host__ __device__ inline uint32_t rank_of(uint64_t cb, uint32_t n_ranks)
{
    // owner = floor(hash32 * n_ranks / 2^32).  Written with an explicit 32x32 multiply-high: ptxas 12.9 mis-folds the
    // 64-bit form ((h >> 32) * n) >> 32 into the shared-memory address computation (observed: every key -> rank 0).
    // low hash word: the barcode table takes its slot from the TOP bits of the same hash, and the two must stay independent
    const uint32_t hi = uint32_t(barcode_hash(cb));
#ifdef __CUDA_ARCH__
    return __umulhi(hi, n_ranks);
#else
    return uint32_t((uint64_t(hi) * n_ranks) >> 32);
#endif
}

__global__ void __launch_bounds__(256) k_route_count(const dge_record16 *__restrict__ in, size_t n, uint32_t n_ranks, unsigned long long *__restrict__ counts)
{
    __shared__ uint32_t h[64];
    if (threadIdx.x < 64) h[threadIdx.x] = 0;
    __syncthreads();
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
        atomicAdd(&h[rank_of(in[i].key >> 24, n_ranks)], 1u);
    __syncthreads();
    if (threadIdx.x < n_ranks && h[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)h[threadIdx.x]);
}

constexpr int ROUTE_ITEMS = 8;
// Position of every lane among the lanes of its warp that hold the same destination + ONE shared-memory atomic per (warp, destination):
// with 2-8 destinations a per-lane atomicAdd on the destination counter serialises 4-16 ways.
__device__ __forceinline__ uint32_t route_rank_in_tile(uint32_t rk, uint32_t *cnt, bool valid)
{
    const unsigned lane = threadIdx.x & 31u;
    const unsigned m = __match_any_sync(0xFFFFFFFFu, valid ? rk : 0xFFFFFFFFu);
    const int leader = __ffs(m) - 1;
    uint32_t p = 0;
    if (valid && int(lane) == leader) p = atomicAdd(&cnt[rk], uint32_t(__popc(m)));
    p = __shfl_sync(0xFFFFFFFFu, p, leader);
    return p + uint32_t(__popc(m & ((1u << lane) - 1u)));
}

// direct variant (default): every record stored straight to its slot of the destination run
__global__ void __launch_bounds__(256) k_route_scatter_direct(const dge_record16 *__restrict__ in, size_t n, uint32_t n_ranks,
                                                              unsigned long long *__restrict__ cursor, dge_record16 *__restrict__ out)
{
    __shared__ uint32_t cnt[64];
    __shared__ unsigned long long basep[64];
    if (threadIdx.x < 64) cnt[threadIdx.x] = 0;
    __syncthreads();
    const size_t base = size_t(blockIdx.x) * 256 * ROUTE_ITEMS;
    uint4 rec[ROUTE_ITEMS];
    uint32_t rk[ROUTE_ITEMS], pos[ROUTE_ITEMS];
#pragma unroll
    for (int j = 0; j < ROUTE_ITEMS; ++j)
    {
        size_t i = base + size_t(j) * 256 + threadIdx.x;
        if (i < n) rec[j] = reinterpret_cast<const uint4 *>(in)[i];
    }
#pragma unroll
    for (int j = 0; j < ROUTE_ITEMS; ++j)
    {
        size_t i = base + size_t(j) * 256 + threadIdx.x;
        const bool valid = i < n;
        rk[j] = valid ? rank_of(((uint64_t(rec[j].y) << 32) | rec[j].x) >> 24, n_ranks) : 0u;
        pos[j] = route_rank_in_tile(rk[j], cnt, valid);
    }
    __syncthreads();
    if (threadIdx.x < n_ranks && cnt[threadIdx.x]) basep[threadIdx.x] = atomicAdd(&cursor[threadIdx.x], (unsigned long long)cnt[threadIdx.x]);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < ROUTE_ITEMS; ++j)
    {
        size_t i = base + size_t(j) * 256 + threadIdx.x;
        if (i < n) reinterpret_cast<uint4 *>(out)[basep[rk[j]] + pos[j]] = rec[j];
    }
}

// Single-pass variant: no counting pass.  Destination d owns the fixed window out[d * cap, (d + 1) * cap) of THIS rank's buffer (the owner's
// fill kernel pulls it over NVLink later); tiles append to it through the global cursor (which ends up holding the segment's size).
// PUSH: `push16` of every 16 tiles store their records for a remote destination d straight into d's HBM instead (push_base[d] = this
// rank's receive window there, push_cursor[d] its fill level): NVLink is idle while the scatter streams through local HBM, so part of the
// exchange rides along for free and the fill has less left to pull.  A tile that would run past a window writes nothing and raises
// *overflow: the caller then repeats the routing with the exact two-pass scheme (counts first), so the result never depends on the window sizes.
__global__ void __launch_bounds__(256) k_route_scatter_bounded(const dge_record16 *__restrict__ in, size_t n, uint32_t n_ranks, uint32_t me,
                                                               unsigned long long cap, unsigned long long *__restrict__ cursor,
                                                               unsigned int *__restrict__ overflow, dge_record16 *__restrict__ out,
                                                               uint32_t push16, unsigned long long push_cap, unsigned long long *__restrict__ push_cursor,
                                                               dge_record16 *const *__restrict__ push_base)
{
    __shared__ uint32_t cnt[64];
    __shared__ dge_record16 *dst[64];
    if (threadIdx.x < 64) cnt[threadIdx.x] = 0;
    __syncthreads();
    const size_t base = size_t(blockIdx.x) * 256 * ROUTE_ITEMS;
    const bool pushing = push16 && (blockIdx.x & 15u) < push16;
    uint4 rec[ROUTE_ITEMS];
    uint32_t rk[ROUTE_ITEMS], pos[ROUTE_ITEMS];
#pragma unroll
    for (int j = 0; j < ROUTE_ITEMS; ++j)
    {
        size_t i = base + size_t(j) * 256 + threadIdx.x;
        if (i < n) rec[j] = __ldcs(reinterpret_cast<const uint4 *>(in) + i);
    }
#pragma unroll
    for (int j = 0; j < ROUTE_ITEMS; ++j)
    {
        size_t i = base + size_t(j) * 256 + threadIdx.x;
        const bool valid = i < n;
        rk[j] = valid ? rank_of(((uint64_t(rec[j].y) << 32) | rec[j].x) >> 24, n_ranks) : 0u;
        pos[j] = route_rank_in_tile(rk[j], cnt, valid);
    }
    __syncthreads();
    if (threadIdx.x < n_ranks)
    {
        const uint32_t d = threadIdx.x;
        dge_record16 *p = nullptr;
        if (cnt[d])
        {
            if (pushing && d != me)
            {
                const unsigned long long b = atomicAdd(&push_cursor[d], (unsigned long long)cnt[d]);
                if (b + cnt[d] > push_cap) atomicExch(overflow, 1u);
                else p = push_base[d] + b;
            }
            else
            {
                const unsigned long long b = atomicAdd(&cursor[d], (unsigned long long)cnt[d]);
                if (b + cnt[d] > cap) atomicExch(overflow, 1u);
                else p = out + (unsigned long long)d * cap + b;
            }
        }
        dst[d] = p;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < ROUTE_ITEMS; ++j)
    {
        size_t i = base + size_t(j) * 256 + threadIdx.x;
        if (i < n && dst[rk[j]]) reinterpret_cast<uint4 *>(dst[rk[j]])[pos[j]] = rec[j];
    }
}

// Staged variant (A/B only, DGE_ROUTE_STAGED=1): the 2048-record tile is REORDERED BY DESTINATION IN SHARED MEMORY and written as one contiguous run
// per destination (coalesced full-sector stores; a scattered 16-byte store costs an L2 write transaction of its own), one global
// atomicAdd per (tile, destination).  Order inside a destination segment is arbitrary: first-seen order travels in read_idx.
__global__ void __launch_bounds__(256) k_route_scatter(const dge_record16 *__restrict__ in, size_t n, uint32_t n_ranks,
                                                       unsigned long long *__restrict__ cursor, dge_record16 *__restrict__ out)
{
    __shared__ uint32_t cnt[64], start[65];
    __shared__ unsigned long long basep[64];
    __shared__ uint4 staged[256 * ROUTE_ITEMS];
    if (threadIdx.x < 64) cnt[threadIdx.x] = 0;
    __syncthreads();
    const size_t base = size_t(blockIdx.x) * 256 * ROUTE_ITEMS;
    const uint32_t in_tile = uint32_t(min(size_t(256 * ROUTE_ITEMS), n - base));
    uint4 rec[ROUTE_ITEMS];
    uint32_t rk[ROUTE_ITEMS], pos[ROUTE_ITEMS];
#pragma unroll
    for (int j = 0; j < ROUTE_ITEMS; ++j)
    {
        const uint32_t i = uint32_t(j) * 256 + threadIdx.x;
        if (i < in_tile) rec[j] = __ldcs(reinterpret_cast<const uint4 *>(in) + base + i); // read once
    }
#pragma unroll
    for (int j = 0; j < ROUTE_ITEMS; ++j)
    {
        const uint32_t i = uint32_t(j) * 256 + threadIdx.x;
        const bool valid = i < in_tile;
        rk[j] = 0;
        if (valid) rk[j] = rank_of(((uint64_t(rec[j].y) << 32) | rec[j].x) >> 24, n_ranks);
        pos[j] = route_rank_in_tile(rk[j], cnt, valid);
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        uint32_t acc = 0;
        for (uint32_t r = 0; r < n_ranks; ++r) { start[r] = acc; acc += cnt[r]; }
        start[n_ranks] = acc;
    }
    if (threadIdx.x < n_ranks && cnt[threadIdx.x]) basep[threadIdx.x] = atomicAdd(&cursor[threadIdx.x], (unsigned long long)cnt[threadIdx.x]);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < ROUTE_ITEMS; ++j)
    {
        const uint32_t i = uint32_t(j) * 256 + threadIdx.x;
        if (i < in_tile) staged[start[rk[j]] + pos[j]] = rec[j];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < ROUTE_ITEMS; ++j)
    {
        const uint32_t i = uint32_t(j) * 256 + threadIdx.x;
        if (i < in_tile)
        {
            uint32_t d = 0;
            while (i >= start[d + 1]) ++d;
            reinterpret_cast<uint4 *>(out)[basep[d] + (i - start[d])] = staged[i];
        }
    }
}

// per-slice destination histogram: counts[s * 64 + r] for slices of `slice_len` records (a multiple of the 2048-record tile)
__global__ void __launch_bounds__(256) k_route_count_slices(const dge_record16 *__restrict__ in, size_t n, uint32_t n_ranks, size_t slice_len,
                                                            unsigned long long *__restrict__ counts)
{
    __shared__ uint32_t h[64];
    const size_t n_tiles = (n + 2047) / 2048;
    for (size_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
    {
        if (threadIdx.x < 64) h[threadIdx.x] = 0;
        __syncthreads();
        const size_t base = tile * 2048;
        unsigned long long key[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
        {   // all loads of the tile in flight before the first shared-memory atomic
            const size_t i = base + size_t(j) * 256 + threadIdx.x;
            key[j] = i < n ? __ldcs(reinterpret_cast<const unsigned long long *>(in + i)) : 0ull;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
        {
            const bool valid = base + size_t(j) * 256 + threadIdx.x < n;
            (void)route_rank_in_tile(valid ? rank_of(key[j] >> 24, n_ranks) : 0u, h, valid);
        }
        __syncthreads();
        const size_t sl = base / slice_len;
        if (threadIdx.x < n_ranks && h[threadIdx.x]) atomicAdd(&counts[sl * 64 + threadIdx.x], (unsigned long long)h[threadIdx.x]);
        __syncthreads();
    }
}

thread_local std::string g_err;

} // namespace

extern "C" {

int dge_synth_generate_device(int device, const dge_synth_params *p, uint64_t first, uint64_t count, dge_record16 *out_device, void *cuda_stream)
{
    if (!p || (!out_device && count)) return DGE_ERR_INVALID;
    if (p->n_cells == 0 || p->n_genes == 0 || p->reads_per_umi == 0 || p->cb_len == 0 || p->cb_len > 20 || p->umi_len == 0 || p->umi_len > 12)
        return DGE_ERR_INVALID;
    try
    {
        DGE_CUDA(cudaSetDevice(device));
        cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
        DevBuf d_ccdf, d_cbc, d_creads, d_gcdf, d_gw;
        d_ccdf.reserve(size_t(p->n_cells) * 8); d_cbc.reserve(size_t(p->n_cells) * 8); d_creads.reserve(size_t(p->n_cells) * 8);
        d_gcdf.reserve(size_t(p->n_genes) * 8); d_gw.reserve(size_t(p->n_genes) * 4);
        DGE_CUDA(cudaMemcpyAsync(d_ccdf.p, p->cell_cdf, size_t(p->n_cells) * 8, cudaMemcpyHostToDevice, st));
        DGE_CUDA(cudaMemcpyAsync(d_cbc.p, p->cell_barcode, size_t(p->n_cells) * 8, cudaMemcpyHostToDevice, st));
        DGE_CUDA(cudaMemcpyAsync(d_creads.p, p->cell_reads, size_t(p->n_cells) * 8, cudaMemcpyHostToDevice, st));
        DGE_CUDA(cudaMemcpyAsync(d_gcdf.p, p->gene_cdf, size_t(p->n_genes) * 8, cudaMemcpyHostToDevice, st));
        DGE_CUDA(cudaMemcpyAsync(d_gw.p, p->gene_weight, size_t(p->n_genes) * 4, cudaMemcpyHostToDevice, st));
        SynthDev d;
        d.seed = p->seed; d.n_cells = p->n_cells; d.n_genes = p->n_genes; d.cb_len = p->cb_len; d.umi_len = p->umi_len;
        d.cell_cdf = d_ccdf.as<uint64_t>(); d.cell_barcode = d_cbc.as<uint64_t>(); d.cell_reads = d_creads.as<uint64_t>();
        d.gene_cdf = d_gcdf.as<uint64_t>(); d.gene_weight = d_gw.as<uint32_t>();
        d.cb_error_ppm = p->cb_error_ppm; d.intergenic_ppm = p->intergenic_ppm; d.intron_ppm = p->intron_ppm;
        d.not_annotated_ppm = p->not_annotated_ppm; d.reads_per_umi = p->reads_per_umi;
        if (count)
        {
            unsigned grid = unsigned(std::min<uint64_t>(div_up<uint64_t>(count, 256), 148 * 32));
            k_synth<<<grid, 256, 0, st>>>(d, first, count, out_device);
            DGE_LAUNCH_CHECK();
        }
        DGE_CUDA(cudaStreamSynchronize(st));
        return DGE_OK;
    }
    catch (std::exception &e) { g_err = e.what(); fprintf(stderr, "dge_synth_generate_device: %s\n", e.what()); return DGE_ERR_CUDA; }
}

int dge_route_by_barcode_device(int device, const dge_record16 *in, size_t n, uint32_t n_ranks, dge_record16 *out, uint64_t *counts, void *cuda_stream)
{
    if (!counts || n_ranks == 0 || n_ranks > 64 || (n && (!in || !out))) return DGE_ERR_INVALID;
    try
    {
        DGE_CUDA(cudaSetDevice(device));
        cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
        DevBuf d_counts;
        d_counts.reserve(128 * 8);
        DGE_CUDA(cudaMemsetAsync(d_counts.p, 0, 128 * 8, st));
        unsigned long long *cnt = d_counts.as<unsigned long long>(), *cursor = cnt + 64;
        std::vector<unsigned long long> hc(64, 0);
        if (n)
        {
            unsigned grid = unsigned(std::min<size_t>(div_up(n, size_t(256 * 8)), 148 * 16));
            k_route_count<<<grid, 256, 0, st>>>(in, n, n_ranks, cnt);
            DGE_LAUNCH_CHECK();
            DGE_CUDA(cudaMemcpyAsync(hc.data(), cnt, 64 * 8, cudaMemcpyDeviceToHost, st));
            DGE_CUDA(cudaStreamSynchronize(st));
            std::vector<unsigned long long> off(64, 0);
            for (uint32_t r = 1; r < n_ranks; ++r) off[r] = off[r - 1] + hc[r - 1];
            DGE_CUDA(cudaMemcpyAsync(cursor, off.data(), 64 * 8, cudaMemcpyHostToDevice, st));
            k_route_scatter<<<unsigned(div_up(n, size_t(256 * ROUTE_ITEMS))), 256, 0, st>>>(in, n, n_ranks, cursor, out);
            DGE_LAUNCH_CHECK();
            DGE_CUDA(cudaStreamSynchronize(st));
        }
        for (uint32_t r = 0; r < n_ranks; ++r) counts[r] = hc[r];
        return DGE_OK;
    }
    catch (std::exception &e) { g_err = e.what(); fprintf(stderr, "dge_route_by_barcode_device: %s\n", e.what()); return DGE_ERR_CUDA; }
}

/* Routing for a PIPELINED exchange, step 1: the input is cut into n_slices slices of slice_len records (a multiple of 2048; the last one
 * may be shorter).  One pass counts the destinations of every slice: counts[s * n_ranks + r] (HOST) and, in `cursors_device`
 * (n_slices * 64 uint64, DEVICE, caller-owned), the exclusive prefix of every slice's segments.  Synchronises the stream once. */
int dge_route_count_slices_device(int device, const dge_record16 *in, size_t n, uint32_t n_ranks, size_t slice_len, uint32_t n_slices,
                                  uint64_t *counts, uint64_t *cursors_device, void *cuda_stream)
{
    if (!counts || !cursors_device || n_ranks == 0 || n_ranks > 64 || n_slices == 0 || n_slices > 256 || slice_len == 0 || (slice_len % 2048) ||
        (n && !in) || size_t(n_slices) * slice_len < n)
        return DGE_ERR_INVALID;
    try
    {
        DGE_CUDA(cudaSetDevice(device));
        cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
        const size_t words = size_t(n_slices) * 64;
        unsigned long long *cnt = reinterpret_cast<unsigned long long *>(cursors_device);
        DGE_CUDA(cudaMemsetAsync(cnt, 0, words * 8, st));
        std::vector<unsigned long long> hc(words, 0), off(words, 0);
        if (n)
        {
            unsigned grid = unsigned(std::min<size_t>(div_up(n, size_t(2048)), 148 * 16));
            k_route_count_slices<<<grid, 256, 0, st>>>(in, n, n_ranks, slice_len, cnt);
            DGE_LAUNCH_CHECK();
            DGE_CUDA(cudaMemcpyAsync(hc.data(), cnt, words * 8, cudaMemcpyDeviceToHost, st));
            DGE_CUDA(cudaStreamSynchronize(st));
            for (uint32_t sl = 0; sl < n_slices; ++sl)
                for (uint32_t r = 1; r < n_ranks; ++r) off[size_t(sl) * 64 + r] = off[size_t(sl) * 64 + r - 1] + hc[size_t(sl) * 64 + r - 1];
            DGE_CUDA(cudaMemcpyAsync(cnt, off.data(), words * 8, cudaMemcpyHostToDevice, st));
            DGE_CUDA(cudaStreamSynchronize(st)); // `off` is a local
        }
        for (uint32_t sl = 0; sl < n_slices; ++sl)
            for (uint32_t r = 0; r < n_ranks; ++r) counts[size_t(sl) * n_ranks + r] = hc[size_t(sl) * 64 + r];
        return DGE_OK;
    }
    catch (std::exception &e) { g_err = e.what(); fprintf(stderr, "dge_route_count_slices_device: %s\n", e.what()); return DGE_ERR_CUDA; }
}

/* Step 2, once per slice and fully asynchronous: the slice's records grouped by destination rank into out[0 .. n_slice), segment r at the
 * prefix prepared by step 1 (`slice_cursors_device` = cursors_device + 64 * slice; consumed by the launch).  The caller queues the
 * all-to-all of the slice right behind it and goes on with the next slice. */
int dge_route_scatter_slice_device(int device, const dge_record16 *in_slice, size_t n_slice, uint32_t n_ranks, uint64_t *slice_cursors_device,
                                   dge_record16 *out_slice, void *cuda_stream)
{
    if (n_ranks == 0 || n_ranks > 64 || !slice_cursors_device || (n_slice && (!in_slice || !out_slice))) return DGE_ERR_INVALID;
    try
    {
        DGE_CUDA(cudaSetDevice(device));
        cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
        if (n_slice)
        {
            // staging the tile by destination in shared memory LOSES here (3.46 vs 2.79 ms at 400 M reads, 2 destinations; profiles/r2_route_kernels_ab.txt):
            // with the warp-aggregated ranking a warp already writes at most n_ranks contiguous runs
            static const bool staged = std::getenv("DGE_ROUTE_STAGED") && atoi(std::getenv("DGE_ROUTE_STAGED")) == 1;
            if (staged)
                k_route_scatter<<<unsigned(div_up(n_slice, size_t(256 * ROUTE_ITEMS))), 256, 0, st>>>(in_slice, n_slice, n_ranks,
                                                                                                     reinterpret_cast<unsigned long long *>(slice_cursors_device), out_slice);
            else
                k_route_scatter_direct<<<unsigned(div_up(n_slice, size_t(256 * ROUTE_ITEMS))), 256, 0, st>>>(in_slice, n_slice, n_ranks,
                                                                                                            reinterpret_cast<unsigned long long *>(slice_cursors_device), out_slice);
            DGE_LAUNCH_CHECK();
        }
        return DGE_OK;
    }
    catch (std::exception &e) { g_err = e.what(); fprintf(stderr, "dge_route_scatter_slice_device: %s\n", e.what()); return DGE_ERR_CUDA; }
}

/* Single-pass routing (no counting pass).  Destination d gets the window out[d * seg_capacity, (d + 1) * seg_capacity) of this rank's buffer.
 * push16 > 0: that many of every 16 tiles store their records for a remote destination d directly into d's memory at push_base_device[d]
 * (room for push_capacity records; pointers from dge_peer_open).  `state_device` = 2 * n_ranks + 1 uint64 words, zeroed by the call:
 * [0, n_ranks) the sizes of the local windows, [n_ranks, 2 n_ranks) the records pushed to each destination, word 2 n_ranks != 0 = a window
 * overflowed -- the output is then unusable and the caller falls back to the exact two-pass routing.  Asynchronous. */
int dge_route_scatter_bounded_device(int device, const dge_record16 *in, size_t n, uint32_t n_ranks, uint32_t my_rank, size_t seg_capacity,
                                     uint64_t *state_device, dge_record16 *out, uint32_t push16, size_t push_capacity,
                                     dge_record16 *const *push_base_device, void *cuda_stream)
{
    if (n_ranks == 0 || n_ranks > 64 || my_rank >= n_ranks || !state_device || seg_capacity == 0 || (n && (!in || !out)) || push16 > 16 ||
        (push16 && (!push_base_device || push_capacity == 0)))
        return DGE_ERR_INVALID;
    try
    {
        DGE_CUDA(cudaSetDevice(device));
        cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
        DGE_CUDA(cudaMemsetAsync(state_device, 0, (size_t(n_ranks) * 2 + 1) * 8, st));
        if (n)
        {
            unsigned long long *state = reinterpret_cast<unsigned long long *>(state_device);
            k_route_scatter_bounded<<<unsigned(div_up(n, size_t(256 * ROUTE_ITEMS))), 256, 0, st>>>(
                in, n, n_ranks, my_rank, (unsigned long long)seg_capacity, state, reinterpret_cast<unsigned int *>(state + 2 * n_ranks), out, push16,
                (unsigned long long)push_capacity, state + n_ranks, push_base_device);
            DGE_LAUNCH_CHECK();
        }
        return DGE_OK;
    }
    catch (std::exception &e) { g_err = e.what(); fprintf(stderr, "dge_route_scatter_bounded_device: %s\n", e.what()); return DGE_ERR_CUDA; }
}

/* ---- peer memory for the exchange: the routed records stay in the SOURCE rank's HBM and the owner's fill kernel pulls its segment
 * straight out of it over NVLink (bulk-async copies of k_fill_pipe's producer warp), so the all-to-all pass and its receive buffer do not exist.
 * One process per GPU: the buffer is a plain cudaMalloc allocation exported as a CUDA IPC handle (64 bytes, sent to the peers by any transport). */
int dge_peer_alloc(int device, size_t bytes, void **ptr, unsigned char handle[64])
{
    if (!ptr || !handle || bytes == 0) return DGE_ERR_INVALID;
    try
    {
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
        DGE_CUDA(cudaSetDevice(device));
        void *p = nullptr;
        DGE_CUDA(cudaMalloc(&p, bytes));
        cudaIpcMemHandle_t hd;
        cudaError_t e = cudaIpcGetMemHandle(&hd, p);
        if (e != cudaSuccess) { cudaFree(p); throw std::runtime_error(std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e)); }
        memcpy(handle, &hd, 64);
        *ptr = p;
        return DGE_OK;
    }
    catch (std::exception &e) { g_err = e.what(); fprintf(stderr, "dge_peer_alloc: %s\n", e.what()); return DGE_ERR_CUDA; }
}

int dge_peer_free(int device, void *ptr)
{
    if (!ptr) return DGE_OK;
    try { DGE_CUDA(cudaSetDevice(device)); DGE_CUDA(cudaFree(ptr)); return DGE_OK; }
    catch (std::exception &e) { g_err = e.what(); fprintf(stderr, "dge_peer_free: %s\n", e.what()); return DGE_ERR_CUDA; }
}

/* Maps a peer's buffer (handle from ITS dge_peer_alloc) into this process, for loads by kernels running on `device`. */
int dge_peer_open(int device, const unsigned char handle[64], void **ptr)
{
    if (!ptr || !handle) return DGE_ERR_INVALID;
    try
    {
        DGE_CUDA(cudaSetDevice(device));
        cudaIpcMemHandle_t hd;
        memcpy(&hd, handle, 64);
        void *p = nullptr;
        DGE_CUDA(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
        *ptr = p;
        return DGE_OK;
    }
    catch (std::exception &e) { g_err = e.what(); fprintf(stderr, "dge_peer_open: %s\n", e.what()); return DGE_ERR_CUDA; }
}

int dge_peer_close(int device, void *ptr)
{
    if (!ptr) return DGE_OK;
    try { DGE_CUDA(cudaSetDevice(device)); DGE_CUDA(cudaIpcCloseMemHandle(ptr)); return DGE_OK; }
    catch (std::exception &e) { g_err = e.what(); fprintf(stderr, "dge_peer_close: %s\n", e.what()); return DGE_ERR_CUDA; }
}

} // extern "C"
