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

// ITEMS reads per thread, every stage issued for all of them before the next one (record loads -> barcode-table probes -> counter-table
// probes -> atomics): the kernel is a chain of three dependent memory round trips per read, so its speed is the number of chains in flight
// (one read per thread: 17 ms at 400 M reads, long_scoreboard 150 cycles per issue; profiles/r2_chr_stats_notes.txt).
template <bool SOA, int ITEMS>
__global__ void __launch_bounds__(256) k_chr_stats(const Rec16 *__restrict__ recs, const unsigned long long *__restrict__ soa_keys,
                                                   const uint32_t *__restrict__ soa_genes, const uint8_t *__restrict__ chr, size_t n,
                                                   const CellSlot *__restrict__ tab, KeyLayout kl, ChrEntry *__restrict__ ct, uint32_t mask,
                                                   ChrCounters *__restrict__ cc)
{
    uint32_t max_chr = 0;
    const uint32_t tmask = (1u << kl.tb) - 1;
    const size_t n_tiles = (n + size_t(256) * ITEMS - 1) / (size_t(256) * ITEMS);
    for (size_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
    {
        const size_t base = tile * 256 * ITEMS;
        uint64_t cb[ITEMS];
        uint32_t what[ITEMS], c[ITEMS], slot[ITEMS]; // what: bit0 exon, bit1 intron, bit2 intergenic; 0 = nothing to count
#pragma unroll
        for (int j = 0; j < ITEMS; ++j)
        {
            const size_t i = base + size_t(j) * 256 + threadIdx.x;
            what[j] = 0; cb[j] = 0; c[j] = 0;
            if (i >= n) continue;
            uint64_t kk;
            uint32_t gw;
            if (SOA) { kk = soa_keys[i]; gw = soa_genes[i]; }
            else
            {
                const uint4 r = __ldcs(reinterpret_cast<const uint4 *>(recs) + i);
                kk = (uint64_t(r.y) << 32) | r.x; gw = r.z;
            }
            c[j] = chr[i];
            const uint32_t gene = gw & 0xFFFFFFu, mark = (gw >> 24) & 7u;
            uint32_t w = gene == NO_GENE ? 4u : ((mark >> 1) & 3u);
            uint64_t b = kk >> 24;
            uint32_t umi = uint32_t(kk) & 0xFFFFFFu;
            if (w && !decode_n_flags(kl, gw, b, umi)) w = 0; // malformed: the fill kernel has reported it
            what[j] = w; cb[j] = b;
        }
        // barcode table: first probe of all reads in flight together, the rare displaced barcode walks on
        unsigned long long first[ITEMS];
#pragma unroll
        for (int j = 0; j < ITEMS; ++j)
        {
            slot[j] = uint32_t(barcode_hash(cb[j]) >> (64 - kl.tb));
            first[j] = what[j] ? __ldcg(&tab[slot[j]].cb) : 0ull;
        }
#pragma unroll
        for (int j = 0; j < ITEMS; ++j)
        {
            if (!what[j]) continue;
            unsigned long long cur = first[j];
            int probes = 0;
            while (cur != cb[j] && cur != EMPTY64 && ++probes < 8192) { slot[j] = (slot[j] + 1) & tmask; cur = __ldcg(&tab[slot[j]].cb); }
            if (cur != cb[j]) { cc->missing_cell = 1; what[j] = 0; }
        }
        // counter table: same scheme
        uint32_t key[ITEMS], s[ITEMS], k0[ITEMS];
#pragma unroll
        for (int j = 0; j < ITEMS; ++j)
        {
            key[j] = ((slot[j] << 8) | c[j]) + 1u;
            s[j] = chr_hash(key[j]) & mask;
            k0[j] = what[j] ? __ldcg(&ct[s[j]].key) : 0u;
        }
#pragma unroll
        for (int j = 0; j < ITEMS; ++j)
        {
            if (!what[j]) continue;
            max_chr = max(max_chr, c[j]);
            uint32_t cur = k0[j];
            bool found = false;
            for (int probes = 0; probes < 4096; ++probes)
            {
                if (cur == 0u)
                {
                    cur = atomicCAS(&ct[s[j]].key, 0u, key[j]);
                    if (cur == 0u) cur = key[j]; // claimed
                }
                if (cur == key[j]) { found = true; break; }
                s[j] = (s[j] + 1) & mask;
                cur = __ldcg(&ct[s[j]].key);
            }
            if (!found) { cc->overflow = 1; continue; }
            if (what[j] & 4u) atomicAdd(&ct[s[j]].intergenic, 1u);
            if (what[j] & 1u) atomicAdd(&ct[s[j]].exon, 1u);
            if (what[j] & 2u) atomicAdd(&ct[s[j]].intron, 1u);
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
