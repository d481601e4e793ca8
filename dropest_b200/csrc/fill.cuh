// fill.cuh -- per-read front end: barcode table insert, first-seen tracking, intergenic short-circuit, 64-bit key packing.
// Replaces, per read, CellsDataContainer::add_record (reference CellsDataContainer.cpp:59-88) up to the point where the
// read is handed to Gene::add_umi; the grouping itself happens in sortcombine.cuh.
#pragma once
#include "common.cuh"
#include "scan.cuh"
#include "sortcombine.cuh"
#include "tma.cuh"

namespace dge
{

struct Rec16 { unsigned long long key; uint32_t gene; uint32_t read_idx; };


struct FillCounters
{
    unsigned long long n_keys;       // reads with a gene (compact keys written)
    unsigned long long intergenic;   // CellsDataContainer::_intergenic_reads
    unsigned long long has_exon;     // _has_exon_reads
    unsigned long long has_intron;   // _has_intron_reads
    unsigned long long has_not_annotated;
    int table_overflow;
    int bad_gene;
    unsigned long long n_flagged; // reads carrying DGE_FLAG_UMI_N / DGE_FLAG_CB_N
    int bad_record; // key bits beyond the configured barcode / UMI lengths, reserved gene-word bits, or read_idx == 0xFFFFFFFF
};

__global__ void k_table_init(CellSlot *tab, size_t cap)
{
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < cap; i += size_t(gridDim.x) * blockDim.x)
    {
        CellSlot s; s.cb = EMPTY64; s.first_idx = NONE32; s.n_intergenic = 0;
        tab[i] = s;
    }
}

__global__ void k_fill_u32(uint32_t *p, size_t n, uint32_t v)
{
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) p[i] = v;
}

// Validates the barcode / UMI fields of a record against the configured lengths and turns N-list indices (DGE_FLAG_UMI_N / DGE_FLAG_CB_N)
// into their internal form: UMI -> [1 : index] inside the ub-bit field, barcode -> CB_N_BIT | index.  false = malformed record.
__device__ __forceinline__ bool decode_n_flags(const KeyLayout &kl, uint32_t gene_word, uint64_t &cb, uint32_t &umi)
{
    if (gene_word >> 29) return false;
    if (gene_word & (FLAG_UMI_N | FLAG_CB_N))
    {
        if (!kl.ne) return false;
        if (gene_word & FLAG_UMI_N)
        {
            if (umi >> (kl.ub - 1)) return false;
            umi |= 1u << (kl.ub - 1);
        }
        else if (umi >> kl.ul) return false;
        if (gene_word & FLAG_CB_N) cb |= CB_N_BIT; // 40 index bits: any value fits
        else if (cb >> kl.cbb) return false;
        return true;
    }
    return (cb >> kl.cbb) == 0 && (umi >> kl.ul) == 0;
}

// Insert-or-find; returns the slot index (= internal cell id) or NONE32 when the table is full.
__device__ __forceinline__ uint32_t table_insert(CellSlot *tab, int tb, uint64_t cb)
{
    const uint32_t mask = (1u << tb) - 1;
    uint32_t s = uint32_t(barcode_hash(cb) >> (64 - tb));
    for (int probes = 0; probes < 8192; ++probes)
    {
        unsigned long long cur = tab[s].cb;
        if (cur == EMPTY64)
        {
            cur = atomicCAS(&tab[s].cb, EMPTY64, (unsigned long long)cb);
            if (cur == EMPTY64) return s;
        }
        if (cur == cb) return s;
        s = (s + 1) & mask;
    }
    return NONE32;
}

__device__ __forceinline__ uint32_t table_find(const CellSlot *tab, int tb, uint64_t cb)
{
    const uint32_t mask = (1u << tb) - 1;
    uint32_t s = uint32_t(barcode_hash(cb) >> (64 - tb));
    for (int probes = 0; probes < 8192; ++probes)
    {
        unsigned long long cur = tab[s].cb;
        if (cur == cb) return s;
        if (cur == EMPTY64) return NONE32;
        s = (s + 1) & mask;
    }
    return NONE32;
}

// records -> compact keys (dense, order irrelevant) + L1 histogram of the keys + global read counters.
// SMEM_GENES: the gene first-seen words are served from a per-block shared-memory copy (a stale UPPER bound of the global value: the
// global word only ever decreases, so `idx < copy` is necessary for `idx < global` and no update can be missed).  A random 4-byte
// gather costs one L1 wavefront per lane in global memory and a few bank cycles in shared memory.
// SOA: the batch arrives as two arrays (64-bit key words, 32-bit gene|mark words) and read_idx = soa_first_idx + position: 12 bytes per
// read cross PCIe instead of 16 (dge_add_batch_soa).
template <int FILL_THREADS, int FILL_ITEMS, int MINB, bool SMEM_GENES, bool SOA>
__global__ void __launch_bounds__(FILL_THREADS, MINB) k_fill_compact(const Rec16 *__restrict__ recs, size_t n, CellSlot *__restrict__ tab, KeyLayout kl,
                                                               uint32_t n_genes, uint32_t *__restrict__ gene_first, uint64_t *__restrict__ out_keys,
                                                               FillCounters *__restrict__ ctr, uint32_t *__restrict__ umi_first,
                                                               const unsigned long long *__restrict__ soa_keys = nullptr,
                                                               const uint32_t *__restrict__ soa_genes = nullptr, uint32_t soa_first_idx = 0)
{
    constexpr int FILL_TILE = FILL_THREADS * FILL_ITEMS;
    extern __shared__ uint32_t gfirst_s[];
    __shared__ uint32_t ws[33];
    __shared__ unsigned long long out_base_s;
    if (SMEM_GENES)
    {
        for (uint32_t i = threadIdx.x; i < n_genes; i += FILL_THREADS) gfirst_s[i] = gene_first[i];
        __syncthreads();
    }

    uint32_t c_inter = 0, c_exon = 0, c_intron = 0, c_na = 0;
    const size_t n_tiles = (n + FILL_TILE - 1) / FILL_TILE;
    const uint32_t tmask = (1u << kl.tb) - 1;
    for (size_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
    {
        const size_t base = tile * FILL_TILE;
        uint64_t keys[FILL_ITEMS];
        uint4 raw[FILL_ITEMS];
        uint4 probe[FILL_ITEMS];
        uint32_t slot0[FILL_ITEMS], gfirst[FILL_ITEMS];
        uint32_t n_valid = 0;
        // phase 1: all record loads in flight
#pragma unroll
        for (int j = 0; j < FILL_ITEMS; ++j)
        {
            const size_t i = base + size_t(j) * FILL_THREADS + threadIdx.x;
            raw[j] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
            if (i < n)
            {
                if (SOA)
                {
                    const unsigned long long kw = __ldg(soa_keys + i);
                    raw[j] = make_uint4(uint32_t(kw), uint32_t(kw >> 32), __ldg(soa_genes + i), soa_first_idx + uint32_t(i));
                }
                else raw[j] = __ldg(reinterpret_cast<const uint4 *>(recs) + i);
            }
        }
        // phase 2: first probe of the barcode table + gene first-seen word, all in flight (random L2 accesses)
#pragma unroll
        for (int j = 0; j < FILL_ITEMS; ++j)
        {
            const uint64_t k = (uint64_t(raw[j].y) << 32) | raw[j].x;
            slot0[j] = uint32_t(barcode_hash(k >> 24) >> (64 - kl.tb));
            probe[j] = __ldcg(reinterpret_cast<const uint4 *>(&tab[slot0[j]]));
            const uint32_t gene = raw[j].z & 0xFFFFFFu;
            gfirst[j] = gene < n_genes ? (SMEM_GENES ? gfirst_s[gene] : gene_first[gene]) : 0u;
        }
        // phase 3: resolve
#pragma unroll
        for (int j = 0; j < FILL_ITEMS; ++j)
        {
            const size_t i = base + size_t(j) * FILL_THREADS + threadIdx.x;
            keys[j] = EMPTY64;
            if (i >= n) continue;
            const uint64_t k = (uint64_t(raw[j].y) << 32) | raw[j].x;
            uint64_t cb = k >> 24;
            uint32_t umi = uint32_t(k) & 0xFFFFFFu;
            const uint32_t gene = raw[j].z & 0xFFFFFFu;
            const uint32_t mark = (raw[j].z >> 24) & 7u;
            const uint32_t idx = raw[j].w;
            if (!decode_n_flags(kl, raw[j].z, cb, umi) || idx == NONE32) { ctr->bad_record = 1; continue; }
                if (raw[j].z & (FLAG_UMI_N | FLAG_CB_N)) atomicAdd(&ctr->n_flagged, 1ull);
            uint32_t slot = slot0[j];
            uint32_t seen_first = probe[j].z;
            if (((uint64_t(probe[j].y) << 32) | probe[j].x) != cb)
            {
                slot = table_insert(tab, kl.tb, cb);
                if (slot == NONE32) { ctr->table_overflow = 1; continue; }
                seen_first = tab[slot].first_idx;
            }
            (void)tmask;
            if (idx < seen_first) atomicMin(&tab[slot].first_idx, idx);
            if (gene == NO_GENE)
            {
                atomicAdd(&tab[slot].n_intergenic, 1u);
                ++c_inter;
                continue;
            }
            if (gene >= n_genes) { ctr->bad_gene = 1; continue; }
            if (idx < gfirst[j])
            {
                atomicMin(&gene_first[gene], idx);
                if (SMEM_GENES) atomicMin(&gfirst_s[gene], idx);
            }
            if (umi_first) atomicMin(&umi_first[umi], idx); // UMI ids are first-seen ranks too (StringIndexer via Gene::add_umi, Gene.cpp:17-24)
            c_exon += (mark >> 1) & 1u; c_intron += (mark >> 2) & 1u; c_na += mark & 1u;
            keys[j] = kl.compose(slot, gene, umi, mark);
            ++n_valid;
        }
        uint32_t total;
        uint32_t ex = block_exclusive_scan(n_valid, ws, &total);
        if (threadIdx.x == 0) out_base_s = total ? atomicAdd(&ctr->n_keys, (unsigned long long)total) : 0ull;
        __syncthreads();
        unsigned long long pos = out_base_s + ex;
#pragma unroll
        for (int j = 0; j < FILL_ITEMS; ++j)
            if (keys[j] != EMPTY64) out_keys[pos++] = keys[j];
        __syncthreads();
    }

    // block-level reduction of the global counters
    auto reduce_add = [&](uint32_t v, unsigned long long *dst) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, d);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(dst, (unsigned long long)v);
    };
    reduce_add(c_inter, &ctr->intergenic);
    reduce_add(c_exon, &ctr->has_exon);
    reduce_add(c_intron, &ctr->has_intron);
    reduce_add(c_na, &ctr->has_not_annotated);
}

__global__ void k_count_occupied(const CellSlot *__restrict__ tab, size_t cap, unsigned long long *out)
{
    uint32_t c = 0;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < cap; i += size_t(gridDim.x) * blockDim.x)
        c += tab[i].cb != EMPTY64;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_down_sync(0xFFFFFFFFu, c, d);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, (unsigned long long)c);
}

// genes that occur in the stream (= gene_indexer().values().size())
__global__ void k_count_seen(const uint32_t *__restrict__ first, size_t n, unsigned long long *out)
{
    uint32_t c = 0;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) c += first[i] != NONE32;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_down_sync(0xFFFFFFFFu, c, d);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, (unsigned long long)c);
}

// ---------------------------------------------------------------------------------------------------------------------
// k_fill_pipe -- the same per-read work as k_fill_compact, restructured for Blackwell's asynchronous data movement:
//   * one persistent block per SM; a PRODUCER warp streams record tiles from HBM into a ring of shared-memory stages with bulk
//     asynchronous copies (cp.async.bulk -> UBLKCP, completion on an mbarrier); CONS_WARPS consumer warps take the records out of
//     shared memory.  The record loads therefore cost the consumers no issue slots and no scoreboard stalls, and because the only
//     synchronisation is the per-stage full/empty mbarrier pair (no __syncthreads in the loop) the warps drift apart: while some wait for
//     their barcode-table probes (random 16-byte L2 accesses), others pack keys or write results.
//   * output: every block appends to its OWN region of the key buffer through a shared-memory cursor (one warp-aggregated shared atomic
//     per warp and tile, no global atomics, no block scan); the regions are consumed as a list by the L1 partition pass
//     (k_l1_scatter_regions), so the keys are never made dense.
//   * the L1 histogram (top 12 key bits) is accumulated in shared memory on the way and flushed once per block: the partition pass
//     needs no histogram pass of its own.
//   * gene first-seen: a 16-bit upper bound of (first read index >> 16) per gene in shared memory filters out all but the reads of the
//     64 Ki-read granule in which a gene first occurs; those go to the global atomicMin.  The bound is lowered with a shared-memory
//     atomic and READ without synchronisation: a stale value is only ever too large, which costs an extra trip to the global word,
//     never a missed update (compute-sanitizer racecheck reports exactly this read, see profiles/).
// 16-bit atomic minimum in shared memory (CAS on the containing word); updates are rare: a few per gene and block
__device__ __forceinline__ void smem_min_u16(uint16_t *arr, uint32_t i, uint32_t v)
{
    uint32_t *w = reinterpret_cast<uint32_t *>(arr) + (i >> 1);
    const int sh = int(i & 1u) * 16;
    uint32_t old = *reinterpret_cast<volatile uint32_t *>(w);
    while (true)
    {
        if (v >= ((old >> sh) & 0xFFFFu)) return;
        const uint32_t prev = atomicCAS(w, old, (old & ~(0xFFFFu << sh)) | (v << sh));
        if (prev == old) return;
        old = prev;
    }
}

// The input of one k_fill_pipe launch: up to FILL_MAX_SEGS record arrays (local HBM or a peer's, mapped with dge_peer_open).  Tiles never
// straddle segments.  Tile order: `rr_rounds` rounds that take one tile of every segment in turn (so that tiles pulled over NVLink and
// local tiles alternate in every block and the link is busy for the whole launch), then what is left of the longer segments, one after
// the other.  v -> (segment, tile in segment) is a bijection; blocks walk v = j * grid + (block + j) % grid, a rotation per round, because
// with a plain block-stride walk v % n_segs would be the same in every round whenever n_segs divides the grid size.
constexpr int FILL_MAX_SEGS = 64;
struct FillSegs
{
    const Rec16 *base[FILL_MAX_SEGS];
    unsigned long long count[FILL_MAX_SEGS];
    uint32_t rem_start[FILL_MAX_SEGS + 1]; // exclusive prefix of (tiles of the segment - rr_rounds)
    uint32_t n_segs, rr_rounds, n_tiles, pad;
};

template <int TILE>
__device__ __forceinline__ void fill_seg_tile(const FillSegs &fs, uint32_t v, uint32_t &seg, size_t &base, uint32_t &cnt)
{
    uint32_t i;
    const uint32_t rr = fs.rr_rounds * fs.n_segs;
    if (v < rr) { seg = v % fs.n_segs; i = v / fs.n_segs; }
    else
    {
        const uint32_t w = v - rr;
        seg = 0;
        while (w >= fs.rem_start[seg + 1]) ++seg;
        i = fs.rr_rounds + (w - fs.rem_start[seg]);
    }
    base = size_t(i) * TILE;
    cnt = uint32_t(min(size_t(TILE), size_t(fs.count[seg]) - base));
}

template <int CONS_WARPS, int ITEMS, int STAGES, bool SOA>
__global__ void __launch_bounds__((CONS_WARPS + 1) * 32, 1)
    k_fill_pipe(const __grid_constant__ FillSegs fs, CellSlot *__restrict__ tab, KeyLayout kl, uint32_t n_genes, uint32_t *__restrict__ gene_first,
                uint64_t *__restrict__ out_keys, size_t region_cap, KeyRegion *__restrict__ regions, FillCounters *__restrict__ ctr,
                uint32_t *__restrict__ umi_first, uint32_t *__restrict__ hist12, const unsigned long long *__restrict__ soa_keys,
                const uint32_t *__restrict__ soa_genes, uint32_t soa_first_idx)
{
    constexpr int CONS = CONS_WARPS * 32;
    constexpr int TILE = CONS * ITEMS;
    constexpr int STAGE_BYTES = TILE * 16;
    extern __shared__ __align__(128) unsigned char fp_smem[];
    uint16_t *g16 = reinterpret_cast<uint16_t *>(fp_smem + size_t(STAGES) * STAGE_BYTES);
    uint32_t *hist_s = reinterpret_cast<uint32_t *>(g16 + ((n_genes + 7u) & ~7u));
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES];
    __shared__ uint32_t cursor_s;
    __shared__ uint32_t stage_cnt[STAGES], stage_base[STAGES]; // records in the stage / its first record's position in the segment: written by the
                                                               // producer before it arms the stage's barrier, read by the consumers after their wait
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (uint32_t i = threadIdx.x; i < n_genes; i += blockDim.x)
    {
        const uint32_t f = gene_first[i];
        g16[i] = f == NONE32 ? uint16_t(0xFFFF) : uint16_t(f >> 16);
    }
    for (uint32_t i = threadIdx.x; i < 4096; i += blockDim.x) hist_s[i] = 0;
    if (threadIdx.x == 0)
    {
        cursor_s = 0;
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], CONS_WARPS); }
        mbar_fence_init();
    }
    __syncthreads();

    const uint32_t n_tiles = fs.n_tiles;
    uint32_t c_inter = 0, c_exon = 0, c_intron = 0, c_na = 0;
    if (warp == CONS_WARPS)
    {   // ---- producer
        if (lane == 0)
        {
            const uint64_t pol = l2_policy_evict_first();
            for (uint32_t k = 0; k * gridDim.x < n_tiles; ++k)
            {
                const uint32_t v = k * gridDim.x + (blockIdx.x + k) % gridDim.x;
                if (v >= n_tiles) break; // only in the last round
                const int s = int(k % STAGES);
                const uint32_t round = k / STAGES;
                mbar_wait(&empty_bar[s], (round & 1u) ^ 1u);
                uint32_t seg, cnt;
                size_t base;
                fill_seg_tile<TILE>(fs, v, seg, base, cnt);
                const Rec16 *recs = fs.base[seg];
                unsigned char *stage = fp_smem + size_t(s) * STAGE_BYTES;
                stage_cnt[s] = cnt; stage_base[s] = uint32_t(base);
                if (SOA)
                {
                    const uint32_t kbytes = (cnt * 8u + 15u) & ~15u, gbytes = (cnt * 4u + 15u) & ~15u;
                    mbar_arrive_expect_tx(&full_bar[s], kbytes + gbytes);
                    bulk_g2s_hint(stage, soa_keys + base, kbytes, &full_bar[s], pol);
                    bulk_g2s_hint(stage + TILE * 8, soa_genes + base, gbytes, &full_bar[s], pol);
                }
                else
                {
                    mbar_arrive_expect_tx(&full_bar[s], cnt * 16u);
                    bulk_g2s_hint(stage, recs + base, cnt * 16u, &full_bar[s], pol);
                }
            }
        }
        __syncwarp();
    }
    else
    {   // ---- consumers
        const int ctid = threadIdx.x; // consumer warps are warps 0 .. CONS_WARPS-1
        uint64_t *my_out = out_keys + size_t(blockIdx.x) * region_cap;
        const int hshift = kl.kb - 12;
        for (uint32_t k = 0; k * gridDim.x < n_tiles; ++k)
        {
            const uint32_t v = k * gridDim.x + (blockIdx.x + k) % gridDim.x;
            if (v >= n_tiles) break;
            const int s = int(k % STAGES);
            const uint32_t round = k / STAGES;
            const unsigned char *stage = fp_smem + size_t(s) * STAGE_BYTES;
            uint4 raw[ITEMS];
            mbar_wait(&full_bar[s], round & 1u);
            const uint32_t cnt = stage_cnt[s], base = stage_base[s];
#pragma unroll
            for (int j = 0; j < ITEMS; ++j)
            {
                const uint32_t i = uint32_t(j) * CONS + ctid;
                raw[j] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
                if (i < cnt)
                {
                    if (SOA)
                    {
                        const uint2 kw = reinterpret_cast<const uint2 *>(stage)[i];
                        raw[j] = make_uint4(kw.x, kw.y, reinterpret_cast<const uint32_t *>(stage + TILE * 8)[i], soa_first_idx + uint32_t(base) + i);
                    }
                    else raw[j] = reinterpret_cast<const uint4 *>(stage)[i];
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]); // the stage may be refilled: this warp holds its records in registers
            // first probe of the barcode table for all records of the thread (random L2 accesses, all in flight together)
            uint4 probe[ITEMS];
            uint32_t slot0[ITEMS];
#pragma unroll
            for (int j = 0; j < ITEMS; ++j)
            {
                const uint64_t kk = (uint64_t(raw[j].y) << 32) | raw[j].x;
                slot0[j] = uint32_t(barcode_hash(kk >> 24) >> (64 - kl.tb));
                probe[j] = __ldcg(reinterpret_cast<const uint4 *>(&tab[slot0[j]])); // (through L1 instead: measured neutral -- ~200 KB of the SM's 256 KB are shared memory here)
            }
            uint64_t keys[ITEMS];
#pragma unroll
            for (int j = 0; j < ITEMS; ++j)
            {
                const uint32_t i = uint32_t(j) * CONS + ctid;
                keys[j] = EMPTY64;
                if (i >= cnt) continue;
                const uint64_t kk = (uint64_t(raw[j].y) << 32) | raw[j].x;
                uint64_t cb = kk >> 24;
                uint32_t umi = uint32_t(kk) & 0xFFFFFFu;
                const uint32_t gene = raw[j].z & 0xFFFFFFu;
                const uint32_t mark = (raw[j].z >> 24) & 7u;
                const uint32_t idx = raw[j].w;
                if (!decode_n_flags(kl, raw[j].z, cb, umi) || idx == NONE32) { ctr->bad_record = 1; continue; }
                if (raw[j].z & (FLAG_UMI_N | FLAG_CB_N)) atomicAdd(&ctr->n_flagged, 1ull);
                uint32_t slot = slot0[j];
                uint32_t seen_first = probe[j].z;
                if (((uint64_t(probe[j].y) << 32) | probe[j].x) != cb)
                {
                    slot = table_insert(tab, kl.tb, cb);
                    if (slot == NONE32) { ctr->table_overflow = 1; continue; }
                    seen_first = tab[slot].first_idx;
                }
                if (idx < seen_first) atomicMin(&tab[slot].first_idx, idx);
                if (gene == NO_GENE)
                {
                    atomicAdd(&tab[slot].n_intergenic, 1u);
                    ++c_inter;
                    continue;
                }
                if (gene >= n_genes) { ctr->bad_gene = 1; continue; }
                const uint32_t hi = idx >> 16;
                const uint32_t bound = g16[gene];
                if (hi <= bound)
                {
                    if (idx < __ldcg(&gene_first[gene])) atomicMin(&gene_first[gene], idx);
                    if (hi < bound) smem_min_u16(g16, gene, hi);
                }
                if (umi_first) atomicMin(&umi_first[umi], idx);
                c_exon += (mark >> 1) & 1u; c_intron += (mark >> 2) & 1u; c_na += mark & 1u;
                keys[j] = kl.compose(slot, gene, umi, mark);
            }
            // append to the block's region: one shared-memory atomic per warp and tile
            unsigned vm[ITEMS];
            uint32_t nv = 0;
#pragma unroll
            for (int j = 0; j < ITEMS; ++j) { vm[j] = __ballot_sync(0xFFFFFFFFu, keys[j] != EMPTY64); nv += __popc(vm[j]); }
            uint32_t pos = 0;
            if (lane == 0 && nv) pos = atomicAdd(&cursor_s, nv);
            pos = __shfl_sync(0xFFFFFFFFu, pos, 0);
            const unsigned lt = (1u << lane) - 1u;
#pragma unroll
            for (int j = 0; j < ITEMS; ++j)
            {
                if (keys[j] != EMPTY64)
                {
                    __stcs(reinterpret_cast<unsigned long long *>(my_out + pos + __popc(vm[j] & lt)), (unsigned long long)keys[j]); // written once, read once much later
                    atomicAdd(&hist_s[uint32_t(keys[j] >> hshift)], 1u);
                }
                pos += __popc(vm[j]);
            }
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < 4096; i += blockDim.x)
        if (hist_s[i]) atomicAdd(&hist12[i], hist_s[i]);
    if (threadIdx.x == 0)
    {
        KeyRegion r;
        r.keys = out_keys + size_t(blockIdx.x) * region_cap; r.count = cursor_s; r.tile0 = 0;
        regions[blockIdx.x] = r;
        if (cursor_s) atomicAdd(&ctr->n_keys, (unsigned long long)cursor_s);
    }
    auto reduce_add = [&](uint32_t v, unsigned long long *dst) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, d);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(dst, (unsigned long long)v);
    };
    reduce_add(c_inter, &ctr->intergenic);
    reduce_add(c_exon, &ctr->has_exon);
    reduce_add(c_intron, &ctr->has_intron);
    reduce_add(c_na, &ctr->has_not_annotated);
}

// tile0 of every region = exclusive prefix of ceil(count / tile) (single block; a few thousand regions at most), one descriptor per
// tile for the partition pass (its blocks then start with ONE load instead of a search through the region table); total in *n_tiles
__global__ void __launch_bounds__(1024) k_region_tiles(KeyRegion *__restrict__ regions, uint32_t n_regions, uint32_t tile, uint32_t *__restrict__ n_tiles,
                                                       KeyTile *__restrict__ tiles)
{
    __shared__ uint32_t ws[33];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_regions; base += blockDim.x)
    {
        const uint32_t r = base + threadIdx.x;
        const uint32_t cnt = r < n_regions ? regions[r].count : 0u;
        const uint32_t t = (cnt + tile - 1) / tile;
        uint32_t tot;
        const uint32_t ex = block_exclusive_scan(t, ws, &tot);
        if (r < n_regions)
        {
            regions[r].tile0 = carry + ex;
            const uint64_t *k = regions[r].keys;
            for (uint32_t q = 0; q < t; ++q)
            {
                KeyTile d;
                d.keys = k + size_t(q) * tile; d.count = min(tile, cnt - q * tile); d.pad = 0;
                tiles[carry + ex + q] = d;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_tiles = carry;
}

} // namespace dge
