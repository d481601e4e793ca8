// segments.cuh -- segmented reductions over the sorted distinct (cell, gene, UMI) list.
//   U   : ukey[n_u] ascending, uval[n_u] = reads | mark<<29
//   CG  : one row per (cell, gene): cg_key = ukey >> ub, cg_start (into U), cg_req (#UMIs whose mark matches the query),
//         cg_reads (all reads), cg_req_reads (reads of matching UMIs)           -> values of cm / cm_raw
//   PC  : one row per cell that owns at least one UMI ("present cell"), in slot order: pc_slot, pc_u_start, pc_cg_start,
//         pc_reads, pc_req_genes, pc_req_umis
// Reference semantics: Gene::number_of_requested_umis (Gene.cpp:60-79), UMI::Mark::match (UMI.cpp:76-85, exact equality with
// one of the query marks), Cell::update_requested_size (Cell.cpp:130-143), Cell::size() / Stats TOTAL_* (Cell.cpp:105-123).
#pragma once
#include "common.cuh"
#include "scan.cuh"

namespace dge
{

constexpr int SEG_THREADS = 256;
constexpr int SEG_ITEMS = 8;
constexpr int SEG_TILE = SEG_THREADS * SEG_ITEMS;

// pass 1: heads per tile
__global__ void __launch_bounds__(SEG_THREADS) k_seg_count(const uint64_t *__restrict__ ukey, uint32_t n_u, int ub, int gub,
                                                           uint32_t *__restrict__ tile_cg, uint32_t *__restrict__ tile_cell)
{
    __shared__ uint32_t ws[33];
    const uint32_t base = blockIdx.x * SEG_TILE + threadIdx.x * SEG_ITEMS;
    uint32_t ncg = 0, ncell = 0;
    uint64_t prev = (base > 0 && base <= n_u) ? ukey[base - 1] : 0;
#pragma unroll
    for (int k = 0; k < SEG_ITEMS; ++k)
    {
        uint32_t i = base + k;
        if (i < n_u)
        {
            uint64_t cur = ukey[i];
            bool first = i == 0;
            ncg += first || (cur >> ub) != (prev >> ub);
            ncell += first || (cur >> gub) != (prev >> gub);
            prev = cur;
        }
    }
    uint32_t tot;
    block_exclusive_scan(ncg, ws, &tot);
    if (threadIdx.x == 0) tile_cg[blockIdx.x] = tot;
    block_exclusive_scan(ncell, ws, &tot);
    if (threadIdx.x == 0) tile_cell[blockIdx.x] = tot;
}

// pass 2: write segment starts AND reduce values into the (cell,gene) / cell rows.  Every thread owns SEG_ITEMS consecutive
// distinct UMIs, accumulates per run in registers and flushes a run with a few no-return atomics (rows are zeroed by the host):
// work per thread is uniform however long a gene's or a cell's UMI list is.
__global__ void __launch_bounds__(SEG_THREADS) k_seg_write(const uint64_t *__restrict__ ukey, const uint32_t *__restrict__ uval, uint32_t n_u,
                                                           int ub, int gub, uint32_t query_mask,
                                                           const uint32_t *__restrict__ tile_cg_off, const uint32_t *__restrict__ tile_cell_off,
                                                           uint64_t *__restrict__ cg_key, uint32_t *__restrict__ cg_start, uint32_t *__restrict__ cg_pc,
                                                           uint32_t *__restrict__ cg_req, uint32_t *__restrict__ cg_reads, uint32_t *__restrict__ cg_req_reads,
                                                           uint32_t *__restrict__ pc_slot, uint32_t *__restrict__ pc_u_start, uint32_t *__restrict__ pc_cg_start,
                                                           uint32_t *__restrict__ pc_reads, uint32_t *__restrict__ pc_req_umis)
{
    __shared__ uint32_t ws[33];
    const uint32_t base = blockIdx.x * SEG_TILE + threadIdx.x * SEG_ITEMS;
    uint64_t cur[SEG_ITEMS];
    uint32_t val[SEG_ITEMS];
    uint8_t hcg[SEG_ITEMS], hcell[SEG_ITEMS];
    uint32_t ncg = 0, ncell = 0;
    uint64_t prev = (base > 0 && base <= n_u) ? ukey[base - 1] : 0;
#pragma unroll
    for (int k = 0; k < SEG_ITEMS; ++k)
    {
        uint32_t i = base + k;
        hcg[k] = hcell[k] = 0;
        if (i < n_u)
        {
            cur[k] = ukey[i];
            val[k] = uval[i];
            bool first = i == 0;
            hcg[k] = first || (cur[k] >> ub) != (prev >> ub);
            hcell[k] = first || (cur[k] >> gub) != (prev >> gub);
            ncg += hcg[k]; ncell += hcell[k];
            prev = cur[k];
        }
    }
    uint32_t tot;
    uint32_t rcg = block_exclusive_scan(ncg, ws, &tot) + tile_cg_off[blockIdx.x];
    uint32_t rcell = block_exclusive_scan(ncell, ws, &tot) + tile_cell_off[blockIdx.x];
    // rows being accumulated: the run that continues from the previous thread has rank (exclusive rank - 1)
    // A run that both starts and ends inside this thread's chunk is owned exclusively: plain stores.  Only the run continuing
    // from the previous thread and the run still open at the end of the chunk can be shared: atomics.
    uint32_t a_req = 0, a_reads = 0, a_rreads = 0, c_reads = 0, c_req = 0;
    bool cg_owned = false, cell_owned = false; // current run started in this chunk
    auto flush_cg = [&](uint32_t row, bool exclusive) {
        if (a_reads)
        {
            if (exclusive)
            {
                cg_req[row] = a_req; cg_reads[row] = a_reads;
                if (cg_req_reads) cg_req_reads[row] = a_rreads;
            }
            else
            {
                if (a_req) atomicAdd(&cg_req[row], a_req);
                atomicAdd(&cg_reads[row], a_reads);
                if (cg_req_reads && a_rreads) atomicAdd(&cg_req_reads[row], a_rreads);
            }
        }
        a_req = a_reads = a_rreads = 0;
    };
    auto flush_cell = [&](uint32_t row, bool exclusive) {
        if (c_reads)
        {
            if (exclusive) { pc_reads[row] = c_reads; pc_req_umis[row] = c_req; }
            else
            {
                atomicAdd(&pc_reads[row], c_reads);
                if (c_req) atomicAdd(&pc_req_umis[row], c_req);
            }
        }
        c_reads = c_req = 0;
    };
#pragma unroll
    for (int k = 0; k < SEG_ITEMS; ++k)
    {
        uint32_t i = base + k;
        if (i < n_u)
        {
            if (hcell[k])
            {
                flush_cell(rcell - 1, cell_owned);
                cell_owned = true;
                pc_slot[rcell] = uint32_t(cur[k] >> gub);
                pc_u_start[rcell] = i;
                pc_cg_start[rcell] = rcg; // a cell head is also a cg head: rcg is that cg's rank
                ++rcell;
            }
            if (hcg[k])
            {
                flush_cg(rcg - 1, cg_owned);
                cg_owned = true;
                cg_key[rcg] = cur[k] >> ub;
                cg_start[rcg] = i;
                cg_pc[rcg] = rcell - 1;
                ++rcg;
            }
            const uint32_t c = val[k] & VAL_COUNT_MASK, m = val[k] >> VAL_MARK_SHIFT;
            const bool match = (query_mask >> m) & 1u;
            a_req += match; a_reads += c; a_rreads += match ? c : 0u;
            c_reads += c; c_req += match;
        }
    }
    if (base < n_u) { flush_cg(rcg - 1, false); flush_cell(rcell - 1, false); }
}

// requested genes per cell = number of its (cell,gene) rows with at least one matching UMI (Cell.cpp:130-143)
__global__ void __launch_bounds__(256) k_pc_req_genes(const uint32_t *__restrict__ cg_req, const uint32_t *__restrict__ cg_pc, uint32_t n_cg,
                                                      uint32_t *__restrict__ pc_req_genes)
{
    constexpr int ITEMS = 8;
    for (uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) * ITEMS; base < n_cg; base += gridDim.x * blockDim.x * ITEMS)
    {
        uint32_t row = NONE32, acc = 0;
#pragma unroll
        for (int k = 0; k < ITEMS; ++k)
        {
            const uint32_t i = base + k;
            if (i >= n_cg) break;
            const uint32_t pc = cg_pc[i];
            if (pc != row) { if (acc) atomicAdd(&pc_req_genes[row], acc); row = pc; acc = 0; }
            acc += cg_req[i] > 0;
        }
        if (acc) atomicAdd(&pc_req_genes[row], acc);
    }
}

__global__ void k_seg_sentinels(uint32_t n_u, uint32_t n_cg, uint32_t n_pc, uint32_t *cg_start, uint32_t *pc_u_start, uint32_t *pc_cg_start)
{
    cg_start[n_cg] = n_u;
    pc_u_start[n_pc] = n_u;
    pc_u_start[n_pc + 1] = n_u;
    pc_cg_start[n_pc] = n_cg;
    pc_cg_start[n_pc + 1] = n_cg; // index n_pc is usable as an EMPTY cell (barcodes without any UMI)
}

// Per (cell, gene): requested UMIs / reads.  Lane per segment for short segments, whole warp for long ones.
__global__ void __launch_bounds__(256) k_cg_reduce(const uint32_t *__restrict__ uval, const uint32_t *__restrict__ cg_start, uint32_t n_cg,
                                                   uint32_t query_mask, uint32_t *__restrict__ cg_req, uint32_t *__restrict__ cg_reads,
                                                   uint32_t *__restrict__ cg_req_reads)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t wbase = warp_global * 32; wbase < n_cg; wbase += n_warps * 32)
    {
        const uint32_t j = wbase + lane;
        uint32_t s = 0, e = 0;
        if (j < n_cg) { s = cg_start[j]; e = cg_start[j + 1]; }
        const bool is_long = (e - s) > 32;
        if (j < n_cg && !is_long)
        {
            uint32_t req = 0, reads = 0, rreads = 0;
            for (uint32_t i = s; i < e; ++i)
            {
                uint32_t v = uval[i], c = v & VAL_COUNT_MASK, m = v >> VAL_MARK_SHIFT;
                bool match = (query_mask >> m) & 1u;
                req += match; reads += c; rreads += match ? c : 0u;
            }
            cg_req[j] = req; cg_reads[j] = reads;
            if (cg_req_reads) cg_req_reads[j] = rreads;
        }
        unsigned long_mask = __ballot_sync(0xFFFFFFFFu, is_long);
        while (long_mask)
        {
            int src = __ffs(long_mask) - 1;
            long_mask &= long_mask - 1;
            uint32_t ls = __shfl_sync(0xFFFFFFFFu, s, src), le = __shfl_sync(0xFFFFFFFFu, e, src);
            uint32_t req = 0, reads = 0, rreads = 0;
            for (uint32_t i = ls + lane; i < le; i += 32)
            {
                uint32_t v = uval[i], c = v & VAL_COUNT_MASK, m = v >> VAL_MARK_SHIFT;
                bool match = (query_mask >> m) & 1u;
                req += match; reads += c; rreads += match ? c : 0u;
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1)
            {
                req += __shfl_down_sync(0xFFFFFFFFu, req, d);
                reads += __shfl_down_sync(0xFFFFFFFFu, reads, d);
                rreads += __shfl_down_sync(0xFFFFFFFFu, rreads, d);
            }
            if (lane == 0)
            {
                cg_req[wbase + src] = req; cg_reads[wbase + src] = reads;
                if (cg_req_reads) cg_req_reads[wbase + src] = rreads;
            }
        }
    }
}

// Per present cell: total reads, requested genes, requested UMIs.
__global__ void __launch_bounds__(256) k_pc_reduce(const uint32_t *__restrict__ cg_req, const uint32_t *__restrict__ cg_reads,
                                                   const uint32_t *__restrict__ pc_cg_start, uint32_t n_pc, uint32_t *__restrict__ pc_reads,
                                                   uint32_t *__restrict__ pc_req_genes, uint32_t *__restrict__ pc_req_umis)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t wbase = warp_global * 32; wbase < n_pc; wbase += n_warps * 32)
    {
        const uint32_t j = wbase + lane;
        uint32_t s = 0, e = 0;
        if (j < n_pc) { s = pc_cg_start[j]; e = pc_cg_start[j + 1]; }
        const bool is_long = (e - s) > 32;
        if (j < n_pc && !is_long)
        {
            uint32_t reads = 0, rg = 0, ru = 0;
            for (uint32_t i = s; i < e; ++i) { uint32_t q = cg_req[i]; reads += cg_reads[i]; rg += q > 0; ru += q; }
            pc_reads[j] = reads; pc_req_genes[j] = rg; pc_req_umis[j] = ru;
        }
        unsigned long_mask = __ballot_sync(0xFFFFFFFFu, is_long);
        while (long_mask)
        {
            int src = __ffs(long_mask) - 1;
            long_mask &= long_mask - 1;
            uint32_t ls = __shfl_sync(0xFFFFFFFFu, s, src), le = __shfl_sync(0xFFFFFFFFu, e, src);
            uint32_t reads = 0, rg = 0, ru = 0;
            for (uint32_t i = ls + lane; i < le; i += 32) { uint32_t q = cg_req[i]; reads += cg_reads[i]; rg += q > 0; ru += q; }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1)
            {
                reads += __shfl_down_sync(0xFFFFFFFFu, reads, d);
                rg += __shfl_down_sync(0xFFFFFFFFu, rg, d);
                ru += __shfl_down_sync(0xFFFFFFFFu, ru, d);
            }
            if (lane == 0) { pc_reads[wbase + src] = reads; pc_req_genes[wbase + src] = rg; pc_req_umis[wbase + src] = ru; }
        }
    }
}

} // namespace dge
