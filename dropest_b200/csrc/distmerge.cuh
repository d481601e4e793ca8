// distmerge.cuh -- kernels of the cross-rank whitelist merge for barcode-hash sharded runs, driven by
// dge_dist_step (engine.cu).  Every rank works on ITS OWN children only:
//   stage 0  real cells that are whitelist barcodes ("self" cells: the only possible merge targets, RealBarcodesMergeStrategy.cpp:63-109)
//            are exported as 16-byte summaries and all-gathered
//   stage 1  the gathered summaries go into a hash table; every local child enumerates its distance-class-1 whitelist neighbours
//            against it (same walk as k_wl_class01); (child, neighbour) pairs with a remote neighbour are packed per destination rank
//            together with the child's (gene|umi, value) list  -> all-to-all
//   stage 2  the owner of a neighbour counts the common (gene, umi) entries (k_intersect_foreign)                    -> all-to-all (u32 per pair)
//   stage 3  the child's owner picks the target exactly like get_best_merge_target (.cpp:31-61) and commits          -> all-to-all (16 B per merge)
//   stage 4  the target's owner adds the child's Stats counters and folds the list it already holds into its U
#pragma once
#include "common.cuh"
#include "merge.cuh"

namespace dge
{

struct SelfRec { unsigned long long cb; uint32_t umis; uint32_t index; };                        // index = position in the owner's self list
struct DistPairRec { uint32_t child_ref, nb_index, n_entries, entry_off; };                     // one (child, remote neighbour) pair inside a blob
struct DistCommitRec { uint32_t pair_pos; int32_t umis_stat, reads_stat; uint32_t n_intergenic; };
struct DistBlobLayout { unsigned long long base; uint32_t n_pairs, n_entries; };                // byte offset of a blob + its counts
static_assert(sizeof(SelfRec) == 16 && sizeof(DistPairRec) == 16 && sizeof(DistCommitRec) == 16 && sizeof(DistBlobLayout) == 16, "wire formats");

constexpr int DIST_MAX_WORLD = 64;
constexpr uint32_t GI_NONE = 0xFFFFFFFFu;

__host__ __device__ inline size_t dist_blob_bytes(uint32_t n_pairs, uint32_t n_entries)
{
    return 16 + size_t(n_pairs) * 16 + ((size_t(n_entries) * 8 + 15) & ~size_t(15)) + ((size_t(n_entries) * 4 + 15) & ~size_t(15));
}
// blob = [n_pairs u64, n_entries u64][DistPairRec x n_pairs][keys u64 x n_entries, padded to 16 B][values u32 x n_entries, padded to 16 B]
__host__ __device__ inline size_t dist_keys_off(uint32_t n_pairs) { return 16 + size_t(n_pairs) * 16; }
__host__ __device__ inline size_t dist_vals_off(uint32_t n_pairs, uint32_t n_entries)
{
    return 16 + size_t(n_pairs) * 16 + ((size_t(n_entries) * 8 + 15) & ~size_t(15));
}

// 1 = every part of the barcode equals exactly one whitelist token (the cell is its own merge target), 2 = a part matches several
// (duplicated tokens: the exact host enumeration decides), 0 = child
__global__ void __launch_bounds__(256) k_wl_self(const CellRow *__restrict__ rows, uint32_t n, WhitelistDev wl, uint32_t *__restrict__ self_flag)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t cell = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (cell >= n) return;
    const uint64_t cb = rows[cell].cb;
    if (cb & CB_N_BIT) { if (lane == 0) self_flag[cell] = 0u; return; } // barcode with N: never a whitelist barcode itself
    bool all_exact = true, multi = false;
    for (int k = 0; k < wl.n_parts; ++k)
    {
        const uint32_t pv = uint32_t(cb >> wl.part_shift[k]) & uint32_t((1ull << (2 * wl.part_len[k])) - 1);
        int exact = 0;
        for (uint32_t t = lane; t < wl.part_size[k]; t += 32) exact += wl.tokens[k][t] == pv;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) exact += __shfl_xor_sync(0xFFFFFFFFu, exact, d);
        if (exact > 1) multi = true;
        if (exact == 0) all_exact = false;
    }
    if (lane == 0) self_flag[cell] = multi ? 2u : (all_exact ? 1u : 0u);
}

__global__ void k_self_is_one(const uint32_t *__restrict__ self_flag, uint32_t n, uint32_t *__restrict__ one)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) one[i] = self_flag[i] != 0u;
}

__global__ void k_self_export(const CellRow *__restrict__ rows, const uint32_t *__restrict__ self_flag, const uint32_t *__restrict__ self_off, uint32_t n,
                              SelfRec *__restrict__ out, uint32_t *__restrict__ self_idx)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        if (self_flag[i])
        {
            const uint32_t p = self_off[i];
            SelfRec r;
            r.cb = rows[i].cb; r.umis = rows[i].n_umis; r.index = p;
            out[p] = r;
            self_idx[p] = i;
        }
}

// gathered summaries -> open-addressing table barcode -> index into the gathered array
__global__ void k_g_build(const SelfRec *__restrict__ all, uint32_t n_all, unsigned long long *__restrict__ g_cb, uint32_t *__restrict__ g_gi, uint32_t mask)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_all; i += gridDim.x * blockDim.x)
    {
        const unsigned long long cb = all[i].cb;
        uint32_t s = uint32_t(barcode_hash(cb) >> 20) & mask;
        while (true)
        {
            const unsigned long long cur = atomicCAS(&g_cb[s], EMPTY64, cb);
            if (cur == EMPTY64 || cur == cb) { g_gi[s] = i; break; }
            s = (s + 1) & mask;
        }
    }
}

__device__ __forceinline__ uint32_t g_find(const unsigned long long *__restrict__ g_cb, const uint32_t *__restrict__ g_gi, uint32_t mask, uint64_t cb)
{
    uint32_t s = uint32_t(barcode_hash(cb) >> 20) & mask;
    while (true)
    {
        const unsigned long long cur = g_cb[s];
        if (cur == cb) return g_gi[s];
        if (cur == EMPTY64) return GI_NONE;
        s = (s + 1) & mask;
    }
}

// Distance-class-1 neighbours of every local child among the gathered self cells of ALL ranks (one warp per real cell; the walk of
// k_wl_class01 with the global table instead of the local barcode table).  nb_count: NB_SELF for self cells, NB_SLOW when no
// eligible class-1 neighbour exists anywhere (or more than WL_K do): the exact host enumeration then takes the cell.
__global__ void __launch_bounds__(256) k_wl_class01_g(const CellRow *__restrict__ rows, const uint32_t *__restrict__ self_flag, uint32_t n, WhitelistDev wl,
                                                      const unsigned long long *__restrict__ g_cb, const uint32_t *__restrict__ g_gi, uint32_t mask,
                                                      const SelfRec *__restrict__ all, int *__restrict__ nb_count, uint32_t *__restrict__ nb_gi)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t cell = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (cell >= n) return;
    const uint32_t sf = self_flag[cell];
    if (sf) { if (lane == 0) nb_count[cell] = NB_SELF; return; } // a whitelist barcode (also with duplicated tokens: its first neighbour is itself)
    const uint64_t cb = rows[cell].cb;
    const uint32_t base_umis = rows[cell].n_umis;
    if (cb & CB_N_BIT) { if (lane == 0) nb_count[cell] = NB_SLOW; return; } // exact enumeration with N wildcards on the host
    int n_exact_parts = 0, missing_part = -1;
    uint32_t part_vals[WL_MAX_PARTS];
    for (int k = 0; k < wl.n_parts; ++k)
    {
        const uint32_t pv = uint32_t(cb >> wl.part_shift[k]) & uint32_t((1ull << (2 * wl.part_len[k])) - 1);
        part_vals[k] = pv;
        int exact = 0;
        for (uint32_t t = lane; t < wl.part_size[k]; t += 32) exact += wl.tokens[k][t] == pv;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) exact += __shfl_xor_sync(0xFFFFFFFFu, exact, d);
        if (exact >= 1) ++n_exact_parts; else missing_part = k;
    }
    int found = 0;
    bool overflow = false;
    if (n_exact_parts == wl.n_parts - 1)
    {
        const int k = missing_part;
        const uint64_t part_mask = ((1ull << (2 * wl.part_len[k])) - 1) << wl.part_shift[k];
        for (uint32_t t0 = 0; t0 < wl.part_size[k]; t0 += 32)
        {
            const uint32_t t = t0 + lane;
            bool eligible = false;
            uint32_t gi = GI_NONE;
            if (t < wl.part_size[k])
            {
                const uint32_t tok = wl.tokens[k][t];
                if (hamming2bit(tok, part_vals[k]) == 1)
                {
                    const uint64_t cand = (cb & ~part_mask) | (uint64_t(tok) << wl.part_shift[k]);
                    gi = g_find(g_cb, g_gi, mask, cand);
                    eligible = gi != GI_NONE && all[gi].umis >= base_umis; // self cells are real: the size condition holds by construction
                }
            }
            const unsigned m = __ballot_sync(0xFFFFFFFFu, eligible);
            if (eligible)
            {
                const int pos = found + __popc(m & ((1u << lane) - 1));
                if (pos < WL_K) nb_gi[size_t(cell) * WL_K + pos] = gi; else overflow = true;
            }
            found += __popc(m);
        }
        overflow = __any_sync(0xFFFFFFFFu, overflow);
    }
    if (lane == 0) nb_count[cell] = (found == 0 || overflow) ? NB_SLOW : found;
}

// (child, neighbour) pairs in CSR order of the children
__global__ void k_dist_pairs(const int *__restrict__ nb_count, const uint32_t *__restrict__ nb_gi, const uint32_t *__restrict__ pair_off, uint32_t n,
                             uint32_t *__restrict__ pair_child, uint32_t *__restrict__ pair_gi)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const int c = nb_count[i];
        for (int k = 0; k < c; ++k) { pair_child[pair_off[i] + k] = i; pair_gi[pair_off[i] + k] = nb_gi[size_t(i) * WL_K + k]; }
    }
}

__device__ __forceinline__ uint32_t rank_of_gi(const uint32_t *__restrict__ rank_off, uint32_t world, uint32_t gi)
{
    uint32_t r = 0;
    while (r + 1 < world && gi >= rank_off[r + 1]) ++r;
    return r;
}

// per destination rank: number of remote pairs and of list entries they carry
__global__ void k_dist_count(const uint32_t *__restrict__ pair_child, const uint32_t *__restrict__ pair_gi, uint32_t n_pairs, const CellRow *__restrict__ rows,
                             const uint32_t *__restrict__ rank_off, uint32_t world, uint32_t my_rank, uint32_t *__restrict__ dest_pairs,
                             uint32_t *__restrict__ dest_entries)
{
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n_pairs; p += gridDim.x * blockDim.x)
    {
        const uint32_t d = rank_of_gi(rank_off, world, pair_gi[p]);
        if (d == my_rank) continue;
        const CellRow r = rows[pair_child[p]];
        atomicAdd(&dest_pairs[d], 1u);
        atomicAdd(&dest_entries[d], r.pc == NONE32 ? 0u : r.n_umis);
    }
}

// headers of the remote pairs (claims a position in the destination's blob) and local intersection jobs
__global__ void k_dist_pack_heads(const uint32_t *__restrict__ pair_child, const uint32_t *__restrict__ pair_gi, uint32_t n_pairs, const CellRow *__restrict__ rows,
                                  const SelfRec *__restrict__ all, const uint32_t *__restrict__ rank_off, uint32_t world, uint32_t my_rank,
                                  const uint32_t *__restrict__ self_idx, uint32_t empty_pc, const DistBlobLayout *__restrict__ lay,
                                  uint32_t *__restrict__ cur_pairs, uint32_t *__restrict__ cur_entries, unsigned char *__restrict__ send,
                                  uint32_t *__restrict__ pair_pos, uint32_t *__restrict__ pair_eoff, PairJob *__restrict__ local_jobs)
{
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n_pairs; p += gridDim.x * blockDim.x)
    {
        const uint32_t gi = pair_gi[p];
        const uint32_t d = rank_of_gi(rank_off, world, gi);
        const CellRow r = rows[pair_child[p]];
        const uint32_t a_pc = r.pc == NONE32 ? empty_pc : r.pc;
        if (d == my_rank)
        {
            const uint32_t nb_pc = rows[self_idx[all[gi].index]].pc;
            local_jobs[p] = PairJob{a_pc, nb_pc == NONE32 ? empty_pc : nb_pc};
            pair_pos[p] = GI_NONE;
            continue;
        }
        local_jobs[p] = PairJob{empty_pc, empty_pc};
        const uint32_t ne = r.pc == NONE32 ? 0u : r.n_umis;
        const uint32_t pos = atomicAdd(&cur_pairs[d], 1u);
        const uint32_t eoff = atomicAdd(&cur_entries[d], ne);
        DistPairRec rec{p, all[gi].index, ne, eoff};
        reinterpret_cast<DistPairRec *>(send + lay[d].base + 16)[pos] = rec;
        pair_pos[p] = pos;
        pair_eoff[p] = eoff;
    }
}

// blob headers {n_pairs, n_entries}
__global__ void k_dist_blob_headers(const DistBlobLayout *__restrict__ lay, uint32_t world, unsigned char *__restrict__ send)
{
    const uint32_t d = threadIdx.x;
    if (d >= world || (lay[d].n_pairs == 0 && lay[d].n_entries == 0 && lay[d].base == ~0ull)) return;
    unsigned long long *hdr = reinterpret_cast<unsigned long long *>(send + lay[d].base);
    hdr[0] = lay[d].n_pairs; hdr[1] = lay[d].n_entries;
}

// the child's (gene|umi) keys and values behind the pair headers of its destination blob (one block per pair)
__global__ void __launch_bounds__(128) k_dist_pack_lists(const uint32_t *__restrict__ pair_child, const uint32_t *__restrict__ pair_gi, uint32_t n_pairs,
                                                         const CellRow *__restrict__ rows, const uint32_t *__restrict__ rank_off, uint32_t world, uint32_t my_rank,
                                                         const DistBlobLayout *__restrict__ lay, const uint32_t *__restrict__ pair_eoff,
                                                         const uint64_t *__restrict__ ukey, const uint32_t *__restrict__ uval, const uint32_t *__restrict__ pc_u_start,
                                                         int gub, unsigned char *__restrict__ send)
{
    const uint64_t gu_mask = (1ull << gub) - 1;
    for (uint32_t p = blockIdx.x; p < n_pairs; p += gridDim.x)
    {
        const uint32_t d = rank_of_gi(rank_off, world, pair_gi[p]);
        if (d == my_rank) continue;
        const CellRow r = rows[pair_child[p]];
        if (r.pc == NONE32) continue;
        const uint32_t s = pc_u_start[r.pc], e = pc_u_start[r.pc + 1];
        const DistBlobLayout L = lay[d];
        uint64_t *keys = reinterpret_cast<uint64_t *>(send + L.base + dist_keys_off(L.n_pairs)) + pair_eoff[p];
        uint32_t *vals = reinterpret_cast<uint32_t *>(send + L.base + dist_vals_off(L.n_pairs, L.n_entries)) + pair_eoff[p];
        for (uint32_t i = s + threadIdx.x; i < e; i += blockDim.x) { keys[i - s] = ukey[i] & gu_mask; vals[i - s] = uval[i]; }
    }
}

// received pair headers of every source -> intersection jobs against this rank's cells.  rl[s] = layout of the blob received from s;
// job_base[s] = index of its first pair in the concatenated job list.
__global__ void k_dist_recv_jobs(const unsigned char *__restrict__ recv, const DistBlobLayout *__restrict__ rl, const uint32_t *__restrict__ job_base, uint32_t world,
                                 const CellRow *__restrict__ rows, const uint32_t *__restrict__ self_idx, uint32_t n_self, uint32_t empty_pc,
                                 ForeignJob *__restrict__ jobs, int *__restrict__ bad)
{
    const uint32_t total = job_base[world];
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < total; j += gridDim.x * blockDim.x)
    {
        uint32_t s = 0;
        while (s + 1 < world && j >= job_base[s + 1]) ++s;
        const DistBlobLayout L = rl[s];
        const DistPairRec rec = reinterpret_cast<const DistPairRec *>(recv + L.base + 16)[j - job_base[s]];
        if (rec.nb_index >= n_self || rec.entry_off + rec.n_entries > L.n_entries) { *bad = 1; jobs[j] = ForeignJob{0, 0, empty_pc}; continue; }
        const uint32_t nb_pc = rows[self_idx[rec.nb_index]].pc;
        const unsigned long long key_off = (L.base + dist_keys_off(L.n_pairs)) / 8 + rec.entry_off; // in 8-byte units from the start of recv
        jobs[j] = ForeignJob{uint32_t(key_off), rec.n_entries, nb_pc == NONE32 ? empty_pc : nb_pc};
    }
}

// replies: one u32 per received pair, grouped by source (segments padded to 16 bytes: reply_base in u32 units)
__global__ void k_dist_reply(const uint32_t *__restrict__ isect, const uint32_t *__restrict__ job_base, const uint32_t *__restrict__ reply_base, uint32_t world,
                             uint32_t *__restrict__ reply)
{
    const uint32_t total = job_base[world];
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < total; j += gridDim.x * blockDim.x)
    {
        uint32_t s = 0;
        while (s + 1 < world && j >= job_base[s + 1]) ++s;
        reply[reply_base[s] + (j - job_base[s])] = isect[j];
    }
}

// intersection size of every pair: local ones from k_intersect, remote ones from the destination's reply
__global__ void k_dist_collect(const uint32_t *__restrict__ pair_gi, const uint32_t *__restrict__ pair_pos, uint32_t n_pairs, const uint32_t *__restrict__ rank_off,
                               uint32_t world, const uint32_t *__restrict__ local_isect, const uint32_t *__restrict__ reply, const uint32_t *__restrict__ reply_base,
                               uint32_t *__restrict__ pair_isect)
{
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n_pairs; p += gridDim.x * blockDim.x)
    {
        if (pair_pos[p] == GI_NONE) { pair_isect[p] = local_isect[p]; continue; }
        const uint32_t d = rank_of_gi(rank_off, world, pair_gi[p]);
        pair_isect[p] = reply[reply_base[d] + pair_pos[p]];
    }
}

// RealBarcodesMergeStrategy::get_best_merge_target (.cpp:31-61) per child over its pair list: best pair (first maximum, strict <),
// -1 below min_merge_fraction; tie = the maximum is reached more than once and is admissible (the reference's neighbour order decides)
__global__ void k_dist_best2(const uint32_t *__restrict__ pair_off, const uint32_t *__restrict__ pair_cnt, const uint32_t *__restrict__ pair_gi,
                             const uint32_t *__restrict__ pair_isect, const CellRow *__restrict__ rows, const SelfRec *__restrict__ all, uint32_t n,
                             double min_frac, int *__restrict__ best_pair, uint32_t *__restrict__ tie)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const uint32_t c = pair_cnt[i];
        best_pair[i] = -2; tie[i] = 0;
        if (!c) continue;
        const double inv_base = __ddiv_rn(1., double(rows[i].n_umis));
        double max_frac = 0, top = -1;
        int best = int(pair_off[i]), n_top = 0;
        for (uint32_t k = 0; k < c; ++k)
        {
            const uint32_t p = pair_off[i] + k;
            const double frac = __dmul_rn(__dmul_rn(0.5, double(pair_isect[p])), __dadd_rn(inv_base, __ddiv_rn(1., double(all[pair_gi[p]].umis))));
            if (max_frac < frac) { max_frac = frac; best = int(p); }
            if (frac > top) { top = frac; n_top = 1; } else if (frac == top) ++n_top;
        }
        if (c > 1 && n_top > 1 && !(top < min_frac)) tie[i] = 1;
        best_pair[i] = max_frac < min_frac ? -1 : best;
    }
}

constexpr int32_t DF_REMOTE = -4; // CellState::target of a child merged into a cell of another rank

// Outcome per child: excluded, merged into a local cell (target = its row) or into a remote one (commit record for its owner).
__global__ void k_dist_decide(const int *__restrict__ nb_count, const int *__restrict__ best_pair, const uint32_t *__restrict__ pair_gi,
                              const uint32_t *__restrict__ pair_pos, const CellRow *__restrict__ rows, const SelfRec *__restrict__ all,
                              const uint32_t *__restrict__ rank_off, uint32_t world, uint32_t my_rank, const uint32_t *__restrict__ self_idx, uint32_t n,
                              CellState *__restrict__ st, unsigned long long *__restrict__ merged_cb, const DistBlobLayout *__restrict__ clay,
                              uint32_t *__restrict__ commit_cur, unsigned char *__restrict__ commits)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        merged_cb[i] = EMPTY64;
        if (nb_count[i] == NB_SELF) { st[i].target = int32_t(i); continue; }
        const int bp = best_pair[i];
        if (bp < 0) { st[i].target = -1; continue; } // no neighbour at all (-2) or best fraction below the threshold (-1): excluded
        const uint32_t gi = pair_gi[bp];
        const uint32_t d = rank_of_gi(rank_off, world, gi);
        if (d == my_rank) { st[i].target = int32_t(self_idx[all[gi].index]); continue; }
        st[i].target = DF_REMOTE;
        merged_cb[i] = all[gi].cb;
        const CellRow r = rows[i];
        DistCommitRec rec{pair_pos[bp], int32_t(r.n_umis), int32_t(r.n_reads), r.n_intergenic};
        const uint32_t pos = atomicAdd(&commit_cur[d], 1u);
        reinterpret_cast<DistCommitRec *>(commits + clay[d].base)[pos] = rec;
    }
}

// children merged into a remote cell: flagged merged here (their content stays, like the reference's merged cells)
__global__ void k_dist_flag_remote(CellState *__restrict__ st, uint32_t n, DevFlowCounters *__restrict__ ctr)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        if (st[i].target == DF_REMOTE) { st[i].flags |= 2u; atomicAdd(&ctr->n_merged, 1u); }
}

// received commits: Stats::merge into the local target + one move job per merged foreign list (kept blob of stage 2)
__global__ void k_dist_apply_commits(const unsigned char *__restrict__ crecv, const DistBlobLayout *__restrict__ cl, const uint32_t *__restrict__ c_base, uint32_t world,
                                     const unsigned char *__restrict__ kept, const DistBlobLayout *__restrict__ rl, const CellRow *__restrict__ rows,
                                     const uint32_t *__restrict__ self_idx, CellState *__restrict__ st, uint32_t *__restrict__ move_size, ForeignMove *__restrict__ moves,
                                     int *__restrict__ bad)
{
    const uint32_t total = c_base[world];
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < total; j += gridDim.x * blockDim.x)
    {
        uint32_t s = 0;
        while (s + 1 < world && j >= c_base[s + 1]) ++s;
        const DistCommitRec c = reinterpret_cast<const DistCommitRec *>(crecv + cl[s].base)[j - c_base[s]];
        const DistBlobLayout L = rl[s];
        if (c.pair_pos >= L.n_pairs) { *bad = 1; move_size[j] = 0; moves[j] = ForeignMove{0, 0, 0, 0}; continue; }
        const DistPairRec rec = reinterpret_cast<const DistPairRec *>(kept + L.base + 16)[c.pair_pos];
        const uint32_t t = self_idx[rec.nb_index];
        atomicAdd(&st[t].umis_stat, c.umis_stat);
        atomicAdd(&st[t].reads_stat, c.reads_stat);
        atomicAdd(&st[t].n_intergenic, c.n_intergenic);
        st[t].is_target = 1u;
        const unsigned long long key_off = (L.base + dist_keys_off(L.n_pairs)) / 8 + rec.entry_off;
        move_size[j] = rec.n_entries;
        moves[j] = ForeignMove{uint32_t(key_off), rec.n_entries, rows[t].slot, uint32_t((L.base + dist_vals_off(L.n_pairs, L.n_entries)) / 4 + rec.entry_off)};
    }
}

// foreign lists re-labelled to their local destination slot (sortcombine input).  ForeignMove::off = key offset (8-byte units),
// ::out_off = value offset (4-byte units) inside the kept blob buffer; out positions from the exclusive scan of the sizes.
__global__ void __launch_bounds__(256) k_dist_gather_foreign(const ForeignMove *__restrict__ jobs, const uint32_t *__restrict__ out_off, uint32_t n_jobs, uint32_t out_base,
                                                             const unsigned char *__restrict__ kept, int gub, uint64_t *__restrict__ out_keys, uint32_t *__restrict__ out_vals)
{
    const uint64_t *ckeys = reinterpret_cast<const uint64_t *>(kept);
    const uint32_t *cvals = reinterpret_cast<const uint32_t *>(kept);
    for (uint32_t j = blockIdx.x; j < n_jobs; j += gridDim.x)
    {
        const ForeignMove job = jobs[j];
        const uint64_t prefix = uint64_t(job.dst_slot) << gub;
        const uint32_t o = out_base + out_off[j];
        for (uint32_t i = threadIdx.x; i < job.n; i += blockDim.x)
        {
            out_keys[o + i] = (prefix | ckeys[job.off + i]) << 3;
            out_vals[o + i] = cvals[job.out_off + i];
        }
    }
}

} // namespace dge
