// chrstats.cuh -- per-(cell, chromosome) read counters of Stats (reference Estimation/Stats.cpp:23-28, CellsDataContainer.cpp:73-78, 309-327):
//   read without a gene            -> INTERGENIC_READS_PER_CHR_PER_CELL[chr]++
//   read with a gene, mark & EXON  -> EXON_READS_PER_CHR_PER_CELL[chr]++
//   read with a gene, mark & INTRON-> INTRON_READS_PER_CHR_PER_CELL[chr]++   (a read can count for both)
// The chromosome id travels beside the 16-byte record as a 1-byte side array (SURVEY.md 8a, row a-R).  Counters live in an open-addressing
// table keyed by (barcode-table slot, chromosome): almost all reads belong to the few thousand true cells, whose ~25 entries each stay in L2,
// so a read costs one L2 probe and one or two L2 atomics.  Stats::merge (adds the source cell's counters to the merge target) is applied when
// the table is read out (dge_get_chr_stats): a counter belongs to the FINAL target of its cell.
#pragma once
#include "common.cuh"
#include "fill.cuh"

namespace dge
{

struct ChrEntry { uint32_t key; uint32_t exon, intron, intergenic; }; // key = ((slot << 8) | chr) + 1; 0 = empty
static_assert(sizeof(ChrEntry) == 16, "ChrEntry layout");

struct ChrCounters { unsigned int overflow, max_chr, missing_cell, pad; };

__device__ __forceinline__ uint32_t chr_hash(uint32_t k)
{
    k ^= k >> 16; k *= 0x7FEB352Du; k ^= k >> 15; k *= 0x846CA68Bu; k ^= k >> 16;
    return k;
}

template <bool SOA>
__global__ void __launch_bounds__(256) k_chr_stats(const Rec16 *__restrict__ recs, const unsigned long long *__restrict__ soa_keys,
                                                   const uint32_t *__restrict__ soa_genes, const uint8_t *__restrict__ chr, size_t n,
                                                   const CellSlot *__restrict__ tab, KeyLayout kl, ChrEntry *__restrict__ ct, uint32_t mask,
                                                   ChrCounters *__restrict__ cc)
{
    uint32_t max_chr = 0;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
    {
        uint64_t kk;
        uint32_t gw;
        if (SOA) { kk = soa_keys[i]; gw = soa_genes[i]; }
        else
        {
            const uint4 r = __ldcs(reinterpret_cast<const uint4 *>(recs) + i);
            kk = (uint64_t(r.y) << 32) | r.x; gw = r.z;
        }
        const uint32_t gene = gw & 0xFFFFFFu, mark = (gw >> 24) & 7u;
        const bool inter = gene == NO_GENE;
        if (!inter && !(mark & 6u)) continue; // nothing is counted per chromosome for this read
        uint64_t cb = kk >> 24;
        uint32_t umi = uint32_t(kk) & 0xFFFFFFu;
        if (!decode_n_flags(kl, gw, cb, umi)) continue; // malformed: the fill kernel has reported it
        const uint32_t slot = table_find(tab, kl.tb, cb);
        if (slot == NONE32) { cc->missing_cell = 1; continue; }
        const uint32_t c = chr[i];
        max_chr = max(max_chr, c);
        const uint32_t key = ((slot << 8) | c) + 1u;
        uint32_t s = chr_hash(key) & mask;
        bool found = false;
        for (int probes = 0; probes < 4096; ++probes)
        {
            uint32_t cur = __ldcg(&ct[s].key);
            if (cur == 0u)
            {
                cur = atomicCAS(&ct[s].key, 0u, key);
                if (cur == 0u) cur = key; // claimed
            }
            if (cur == key) { found = true; break; }
            s = (s + 1) & mask;
        }
        if (!found) { cc->overflow = 1; continue; }
        if (inter) atomicAdd(&ct[s].intergenic, 1u);
        else
        {
            if (mark & 2u) atomicAdd(&ct[s].exon, 1u);
            if (mark & 4u) atomicAdd(&ct[s].intron, 1u);
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) max_chr = max(max_chr, __shfl_xor_sync(0xFFFFFFFFu, max_chr, d));
    if ((threadIdx.x & 31) == 0 && max_chr) atomicMax(&cc->max_chr, max_chr);
}

// occupied entries, densely (order irrelevant)
__global__ void __launch_bounds__(256) k_chr_export(const ChrEntry *__restrict__ ct, size_t cap, ChrEntry *__restrict__ out, unsigned long long *__restrict__ n_out)
{
    const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    ChrEntry e;
    e.key = 0;
    if (i < cap) e = ct[i];
    const unsigned m = __ballot_sync(0xFFFFFFFFu, e.key != 0u);
    if (!m) return;
    const unsigned lane = threadIdx.x & 31u;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(n_out, (unsigned long long)__popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (e.key != 0u) out[base + __popc(m & ((1u << lane) - 1u))] = e;
}

} // namespace dge
