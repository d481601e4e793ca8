// segments.cuh -- segmented reductions over the sorted distinct (cell, gene, UMI) list.
//   U   : ukey[n_u] ascending, uval[n_u] = reads | mark<<29
//   CG  : one row per (cell, gene): cg_gene = gene id, cg_start (into U), cg_req (#UMIs whose mark matches the query),
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
constexpr int SEG_ITEMS = 4;
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
// distinct UMIs and accumulates per run in registers.  The (cell,gene) rows of a tile are assembled in shared memory (runs shared
// by several threads are combined with shared-memory atomics) and copied out fully coalesced; only the run continuing from
// the previous tile and the last row of the tile (which may continue in the next one) touch global memory with atomics
// (the rows are zeroed by the host).  Cell rows are few: warp-reduced, then one global atomic per warp and cell.
__global__ void __launch_bounds__(SEG_THREADS) k_seg_write(const uint64_t *__restrict__ ukey, const uint32_t *__restrict__ uval, uint32_t n_u,
                                                           int ub, int gub, uint32_t query_mask, uint32_t gene_mask,
                                                           const uint32_t *__restrict__ tile_cg_off, const uint32_t *__restrict__ tile_cell_off,
                                                           uint32_t *__restrict__ cg_gene, uint32_t *__restrict__ cg_start, uint32_t *__restrict__ cg_pc,
                                                           uint32_t *__restrict__ cg_req, uint32_t *__restrict__ cg_reads, uint32_t *__restrict__ cg_req_reads,
                                                           uint32_t *__restrict__ pc_slot, uint32_t *__restrict__ pc_u_start, uint32_t *__restrict__ pc_cg_start,
                                                           uint32_t *__restrict__ pc_reads, uint32_t *__restrict__ pc_req_umis)
{
    __shared__ uint32_t ws[33];
    __shared__ uint32_t s_gene[SEG_TILE], s_start[SEG_TILE], s_pc[SEG_TILE], s_req[SEG_TILE], s_reads[SEG_TILE], s_rreads[SEG_TILE];
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
    uint32_t tot_cg, tot_cell;
    int lcg = int(block_exclusive_scan(ncg, ws, &tot_cg));          // local rank of this thread's first head inside the tile
    const uint32_t cg_off = tile_cg_off[blockIdx.x];
    uint32_t rcell = block_exclusive_scan(ncell, ws, &tot_cell) + tile_cell_off[blockIdx.x];
    for (uint32_t i = threadIdx.x; i < tot_cg; i += SEG_THREADS) { s_req[i] = 0; s_reads[i] = 0; s_rreads[i] = 0; }
    __syncthreads();

    uint32_t a_req = 0, a_reads = 0, a_rreads = 0, c_reads = 0, c_req = 0;
    bool cg_owned = false; // current run started in this chunk
    // local row -1 = the run continuing from the previous tile: straight to global memory
    auto flush_cg = [&](int lrow, bool exclusive) {
        if (a_reads)
        {
            if (lrow < 0)
            {
                const uint32_t row = cg_off - 1;
                if (a_req) atomicAdd(&cg_req[row], a_req);
                atomicAdd(&cg_reads[row], a_reads);
                if (cg_req_reads && a_rreads) atomicAdd(&cg_req_reads[row], a_rreads);
            }
            else if (exclusive) { s_req[lrow] = a_req; s_reads[lrow] = a_reads; s_rreads[lrow] = a_rreads; }
            else
            {
                if (a_req) atomicAdd(&s_req[lrow], a_req);
                atomicAdd(&s_reads[lrow], a_reads);
                if (a_rreads) atomicAdd(&s_rreads[lrow], a_rreads);
            }
        }
        a_req = a_reads = a_rreads = 0;
    };
    auto flush_cell = [&](uint32_t row) {
        if (c_reads)
        {
            atomicAdd(&pc_reads[row], c_reads);
            if (c_req) atomicAdd(&pc_req_umis[row], c_req);
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
                flush_cell(rcell - 1);
                pc_slot[rcell] = uint32_t(cur[k] >> gub);
                pc_u_start[rcell] = i;
                pc_cg_start[rcell] = cg_off + uint32_t(lcg); // a cell head is also a cg head: lcg is that cg's local rank
                ++rcell;
            }
            if (hcg[k])
            {
                flush_cg(lcg - 1, cg_owned);
                cg_owned = true;
                s_gene[lcg] = uint32_t(cur[k] >> ub) & gene_mask;
                s_start[lcg] = i;
                s_pc[lcg] = rcell - 1;
                ++lcg;
            }
            const uint32_t c = val[k] & VAL_COUNT_MASK, m = val[k] >> VAL_MARK_SHIFT;
            const bool match = (query_mask >> m) & 1u;
            a_req += match; a_reads += c; a_rreads += match ? c : 0u;
            c_reads += c; c_req += match;
        }
    }
    if (base < n_u) flush_cg(lcg - 1, false);
    // open cell run at the end of the chunk: when no thread of the warp saw a cell head, the whole warp sits inside ONE cell
    {
        const bool warp_one_cell = __all_sync(0xFFFFFFFFu, ncell == 0);
        if (warp_one_cell)
        {
#pragma unroll
            for (int d = 16; d > 0; d >>= 1)
            {
                c_reads += __shfl_down_sync(0xFFFFFFFFu, c_reads, d);
                c_req += __shfl_down_sync(0xFFFFFFFFu, c_req, d);
            }
            if ((threadIdx.x & 31) == 0 && rcell > 0) flush_cell(rcell - 1);
        }
        else if (base < n_u) flush_cell(rcell - 1);
    }
    __syncthreads();
    for (uint32_t l = threadIdx.x; l < tot_cg; l += SEG_THREADS)
    {
        const uint32_t g = cg_off + l;
        cg_gene[g] = s_gene[l]; cg_start[g] = s_start[l]; cg_pc[g] = s_pc[l];
        if (l + 1 < tot_cg)
        {
            cg_req[g] = s_req[l]; cg_reads[g] = s_reads[l];
            if (cg_req_reads) cg_req_reads[g] = s_rreads[l];
        }
        else
        {   // the last row of the tile may continue in the next tile
            if (s_req[l]) atomicAdd(&cg_req[g], s_req[l]);
            if (s_reads[l]) atomicAdd(&cg_reads[g], s_reads[l]);
            if (cg_req_reads && s_rreads[l]) atomicAdd(&cg_req_reads[g], s_rreads[l]);
        }
    }
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
