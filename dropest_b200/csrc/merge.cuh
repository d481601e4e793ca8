// merge.cuh -- device kernels of the cell-barcode merge stage (reference Estimation/Merge/*).
//   k_wl_class01   nearest-whitelist neighbours in distance classes 0 and 1 (RealBarcodesMergeStrategy::get_real_neighbour_cbs,
//                  RealBarcodesMergeStrategy.cpp:63-109 + BarcodesParser.cpp:21-74) for equal-length, N-free whitelist parts, where
//                  Levenshtein <= 1  <=>  Hamming <= 1.  Cells whose nearest eligible class is >= 2 are flagged for the exact path.
//   k_intersect    |{(gene,umi)} of A  ∩  {(gene,umi)} of B|  (MergeStrategyBase::get_umigs_intersect_size, MergeStrategyBase.cpp:100-147)
//   k_gather_relabel / k_probe_merge / k_merge_path : apply cell merges (CellsDataContainer::merge_cells, CellsDataContainer.cpp:90-104;
//                  Gene::merge, Gene.cpp:26-36; UMI::merge, UMI.cpp:15-19: counts add, marks OR).
#pragma once
#include "common.cuh"
#include "fill.cuh"
#include "scan.cuh"

namespace dge
{

constexpr int WL_MAX_PARTS = 4;
constexpr int WL_K = 8;          // max neighbours reported by the fast path
constexpr int NB_SELF = -1;      // base barcode is itself a whitelist barcode -> target = base
constexpr int NB_SLOW = -2;      // needs the exact (distance class >= 2 / overflow) path

struct WhitelistDev
{
    int n_parts;
    int part_len[WL_MAX_PARTS];     // bases
    int part_shift[WL_MAX_PARTS];   // bit offset of the part inside the packed barcode
    uint32_t part_size[WL_MAX_PARTS];
    const uint32_t *tokens[WL_MAX_PARTS]; // 2-bit packed tokens
};

__device__ __forceinline__ int hamming2bit(uint32_t a, uint32_t b)
{
    uint32_t x = a ^ b;
    return __popc((x | (x >> 1)) & 0x55555555u);
}

// One warp per real cell.
__global__ void __launch_bounds__(256) k_wl_class01(const uint64_t *__restrict__ cell_cb, const uint32_t *__restrict__ cell_umis, uint32_t n_cells,
                                                    WhitelistDev wl, const CellSlot *__restrict__ tab, int tb,
                                                    const uint32_t *__restrict__ slot_pc, const uint32_t *__restrict__ pc_cg_start,
                                                    const uint32_t *__restrict__ pc_u_start, uint32_t min_genes,
                                                    int *__restrict__ nb_count, uint32_t *__restrict__ nb_pc)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t cell = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (cell >= n_cells) return;
    const uint64_t cb = cell_cb[cell];
    const uint32_t base_umis = cell_umis[cell];
    if (cb & CB_N_BIT)
    {   // a barcode containing N (an index into the caller's list): the exact enumeration with N wildcards runs on the host
        if (lane == 0) nb_count[cell] = NB_SLOW;
        return;
    }

    // per part: count of exact tokens, and up to 32 distance-1 tokens kept as a per-lane register (one per lane) + overflow flag
    int n_exact_parts = 0;
    int missing_part = -1;          // the single part without an exact token (class 1 requires exactly one)
    bool multi_exact = false;
    uint32_t part_vals[WL_MAX_PARTS];
    for (int k = 0; k < wl.n_parts; ++k)
    {
        const uint32_t pv = uint32_t(cb >> wl.part_shift[k]) & uint32_t((1ull << (2 * wl.part_len[k])) - 1);
        part_vals[k] = pv;
        int exact = 0;
        for (uint32_t t = lane; t < wl.part_size[k]; t += 32) exact += wl.tokens[k][t] == pv;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) exact += __shfl_xor_sync(0xFFFFFFFFu, exact, d);
        if (exact > 1) multi_exact = true;
        if (exact >= 1) ++n_exact_parts; else missing_part = k;
    }
    if (multi_exact)
    {   // duplicated tokens inside a part: leave the bookkeeping of duplicate leaves to the exact path
        if (lane == 0) nb_count[cell] = NB_SLOW;
        return;
    }
    if (n_exact_parts == wl.n_parts)
    {
        if (lane == 0) nb_count[cell] = NB_SELF;
        return;
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
            uint32_t pc = NONE32;
            if (t < wl.part_size[k])
            {
                const uint32_t tok = wl.tokens[k][t];
                if (hamming2bit(tok, part_vals[k]) == 1)
                {
                    const uint64_t cand = (cb & ~part_mask) | (uint64_t(tok) << wl.part_shift[k]);
                    const uint32_t slot = table_find(tab, tb, cand);
                    if (slot != NONE32)
                    {
                        pc = slot_pc[slot];
                        if (pc != NONE32)
                        {
                            const uint32_t ng = pc_cg_start[pc + 1] - pc_cg_start[pc];
                            const uint32_t nu = pc_u_start[pc + 1] - pc_u_start[pc];
                            eligible = ng >= min_genes && nu >= base_umis;
                        }
                    }
                }
            }
            const unsigned m = __ballot_sync(0xFFFFFFFFu, eligible);
            if (eligible)
            {
                const int pos = found + __popc(m & ((1u << lane) - 1));
                if (pos < WL_K) nb_pc[size_t(cell) * WL_K + pos] = pc; else overflow = true;
            }
            found += __popc(m);
        }
        overflow = __any_sync(0xFFFFFFFFu, overflow);
    }
    if (lane == 0) nb_count[cell] = (found == 0 || overflow) ? NB_SLOW : found;
}

struct PairJob { uint32_t a_pc, b_pc; };
struct MoveJobLite { uint32_t src_pc, out_off; };

// One block per pair: every (gene, umi) of A is searched in B's sorted range.
__global__ void __launch_bounds__(128) k_intersect(const PairJob *__restrict__ jobs, uint32_t n_jobs, const uint64_t *__restrict__ ukey,
                                                   const uint32_t *__restrict__ pc_u_start, const uint32_t *__restrict__ pc_slot, int gub,
                                                   uint32_t *__restrict__ out)
{
    __shared__ uint32_t red[4];
    const uint32_t job = blockIdx.x;
    if (job >= n_jobs) return;
    uint32_t a = jobs[job].a_pc, b = jobs[job].b_pc;
    uint32_t as = pc_u_start[a], ae = pc_u_start[a + 1], bs = pc_u_start[b], be = pc_u_start[b + 1];
    if (ae - as > be - bs) { uint32_t t; t = a; a = b; b = t; t = as; as = bs; bs = t; t = ae; ae = be; be = t; }
    const uint64_t gu_mask = (1ull << gub) - 1;
    const uint64_t b_prefix = uint64_t(pc_slot[b]) << gub;
    uint32_t cnt = 0;
    for (uint32_t i = as + threadIdx.x; i < ae; i += blockDim.x)
    {
        const uint64_t want = b_prefix | (ukey[i] & gu_mask);
        uint32_t lo = bs, hi = be;
        while (lo < hi)
        {
            uint32_t mid = (lo + hi) >> 1;
            if (ukey[mid] < want) lo = mid + 1; else hi = mid;
        }
        cnt += lo < be && ukey[lo] == want;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) cnt += __shfl_down_sync(0xFFFFFFFFu, cnt, d);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) out[job] = red[0] + red[1] + red[2] + red[3];
}

// ---- cross-rank merge: candidate cells ("children") travel as (gene|umi, value) lists --------------------------------
struct ForeignJob { uint32_t off, n, nb_pc; };          // child entries [off, off+n) of the gathered list vs local cell nb_pc
struct ForeignMove { uint32_t off, n, dst_slot, out_off; };

// |child ∩ local cell|: every (gene,umi) of the child is searched in the local cell's sorted range
__global__ void __launch_bounds__(128) k_intersect_foreign(const ForeignJob *__restrict__ jobs, uint32_t n_jobs, const uint64_t *__restrict__ ckeys,
                                                           const uint64_t *__restrict__ ukey, const uint32_t *__restrict__ pc_u_start,
                                                           const uint32_t *__restrict__ pc_slot, int gub, uint32_t *__restrict__ out)
{
    __shared__ uint32_t red[4];
    const uint32_t job = blockIdx.x;
    if (job >= n_jobs) return;
    const ForeignJob j = jobs[job];
    const uint32_t bs = pc_u_start[j.nb_pc], be = pc_u_start[j.nb_pc + 1];
    const uint64_t b_prefix = uint64_t(pc_slot[j.nb_pc]) << gub;
    uint32_t cnt = 0;
    for (uint32_t i = threadIdx.x; i < j.n; i += blockDim.x)
    {
        const uint64_t want = b_prefix | ckeys[j.off + i];
        uint32_t lo = bs, hi = be;
        while (lo < hi)
        {
            uint32_t mid = (lo + hi) >> 1;
            if (ukey[mid] < want) lo = mid + 1; else hi = mid;
        }
        cnt += lo < be && ukey[lo] == want;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) cnt += __shfl_down_sync(0xFFFFFFFFu, cnt, d);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) out[job] = red[0] + red[1] + red[2] + red[3];
}

// export: (gene|umi) part + value of the listed local cells, concatenated
__global__ void __launch_bounds__(256) k_export_cells(const MoveJobLite *__restrict__ jobs, uint32_t n_jobs, const uint64_t *__restrict__ ukey,
                                                      const uint32_t *__restrict__ uval, const uint32_t *__restrict__ pc_u_start, int gub,
                                                      uint64_t *__restrict__ out_keys, uint32_t *__restrict__ out_vals)
{
    const uint64_t gu_mask = (1ull << gub) - 1;
    for (uint32_t j = blockIdx.x; j < n_jobs; j += gridDim.x)
    {
        const MoveJobLite job = jobs[j];
        const uint32_t s = pc_u_start[job.src_pc], e = pc_u_start[job.src_pc + 1];
        for (uint32_t i = s + threadIdx.x; i < e; i += blockDim.x)
        {
            out_keys[job.out_off + (i - s)] = ukey[i] & gu_mask;
            out_vals[job.out_off + (i - s)] = uval[i];
        }
    }
}

// foreign child lists re-labelled to a local destination slot, as sortcombine input
__global__ void __launch_bounds__(256) k_gather_relabel_foreign(const ForeignMove *__restrict__ jobs, uint32_t n_jobs, const uint64_t *__restrict__ ckeys,
                                                                const uint32_t *__restrict__ cvals, int gub, uint64_t *__restrict__ out_keys,
                                                                uint32_t *__restrict__ out_vals)
{
    for (uint32_t j = blockIdx.x; j < n_jobs; j += gridDim.x)
    {
        const ForeignMove job = jobs[j];
        const uint64_t prefix = uint64_t(job.dst_slot) << gub;
        for (uint32_t i = threadIdx.x; i < job.n; i += blockDim.x)
        {
            out_keys[job.out_off + i] = (prefix | ckeys[job.off + i]) << 3;
            out_vals[job.out_off + i] = cvals[job.off + i];
        }
    }
}

// ---- phase 1 of the whitelist merge on the device (RealBarcodesMergeStrategy::get_best_merge_target, .cpp:31-61) ----------------
// columns of the real-cell rows (cell-id order) + present-cell -> real-cell map
struct CellRow;
__global__ void k_p1_columns(const CellRow *__restrict__ rows, uint32_t n, uint64_t *__restrict__ cb, uint32_t *__restrict__ umis,
                             uint32_t *__restrict__ pc, uint32_t *__restrict__ pc_to_real);

__global__ void k_p1_counts(const int *__restrict__ nb_count, uint32_t n, uint32_t *__restrict__ cnt)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) cnt[i] = nb_count[i] > 0 ? uint32_t(nb_count[i]) : 0u;
}

__global__ void k_p1_jobs(const int *__restrict__ nb_count, const uint32_t *__restrict__ nb_pc, const uint32_t *__restrict__ pc, const uint32_t *__restrict__ off,
                          uint32_t n, PairJob *__restrict__ jobs)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const int c = nb_count[i];
        for (int k = 0; k < c; ++k) jobs[off[i] + k] = PairJob{pc[i], nb_pc[size_t(i) * WL_K + k]};
    }
}

// target[i] = real index of the neighbour with the largest fraction 0.5 * I * (1/U_base + 1/U_nb) (first one on ties of the running
// maximum, like the reference's strict <), -1 when that maximum is below min_merge_fraction.  needs_host[i] is set when the best
// fraction is reached by more than one neighbour and is admissible: the reference's neighbour order then decides (host replay).
__global__ void k_p1_best(const int *__restrict__ nb_count, const uint32_t *__restrict__ nb_pc, const uint32_t *__restrict__ off,
                          const uint32_t *__restrict__ isect, const uint32_t *__restrict__ umis, const uint32_t *__restrict__ pc_to_real, uint32_t n,
                          double min_frac, int *__restrict__ target, uint32_t *__restrict__ needs_host)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const int c = nb_count[i];
        target[i] = -2; needs_host[i] = 0;
        if (c <= 0) continue;
        const double inv_base = __ddiv_rn(1., double(umis[i]));
        double max_frac = 0;
        uint32_t best = pc_to_real[nb_pc[size_t(i) * WL_K]];
        double top = -1; int n_top = 0;
        for (int k = 0; k < c; ++k)
        {
            const uint32_t nb = pc_to_real[nb_pc[size_t(i) * WL_K + k]];
            const double frac = __dmul_rn(__dmul_rn(0.5, double(isect[off[i] + k])), __dadd_rn(inv_base, __ddiv_rn(1., double(umis[nb]))));
            if (max_frac < frac) { max_frac = frac; best = nb; }
            if (frac > top) { top = frac; n_top = 1; } else if (frac == top) ++n_top;
        }
        if (c > 1 && n_top > 1 && !(top < min_frac)) needs_host[i] = 1;
        if (pc_to_real[nb_pc[size_t(i) * WL_K]] == i) target[i] = int(i); // first neighbour is the base itself (.cpp:33-34)
        else target[i] = max_frac < min_frac ? -1 : int(best);
    }
}

// ---- sharded runs: combine the per-rank candidates of every child (dge_dist_apply) --------------------------------------------
struct DistResult { double best_fraction; unsigned long long best_barcode; uint32_t n_neighbours, n_best; }; // = dge_dist_result

// all[r * n + c] = rank r's best local candidate of child c  ->  red[c] = the combination over ranks in rank order, with the rule of
// RealBarcodesMergeStrategy::get_best_merge_target (strict < keeps the first maximum; exact ties add up and keep the smallest barcode)
__global__ void k_dist_reduce(const DistResult *__restrict__ all, uint32_t world, size_t n, DistResult *__restrict__ red)
{
    for (size_t c = size_t(blockIdx.x) * blockDim.x + threadIdx.x; c < n; c += size_t(gridDim.x) * blockDim.x)
    {
        uint32_t n_nb = 0, n_best = 0;
        double best = 0;
        unsigned long long best_cb = EMPTY64;
        for (uint32_t r = 0; r < world; ++r)
        {
            const DistResult res = all[size_t(r) * n + c];
            if (!res.n_neighbours) continue;
            n_nb += res.n_neighbours;
            if (n_best == 0 || best < res.best_fraction) { best = res.best_fraction; best_cb = res.best_barcode; n_best = res.n_best; }
            else if (res.best_fraction == best) { n_best += res.n_best; best_cb = min(best_cb, res.best_barcode); }
        }
        DistResult o;
        o.best_fraction = best; o.best_barcode = best_cb; o.n_neighbours = n_nb; o.n_best = n_best;
        red[c] = o;
    }
}

// candidate columns of the gathered children straight from their device-resident summaries
struct DistChildDev { unsigned long long barcode; int32_t umis_stat, reads_stat, n_genes; uint32_t n_intergenic, n_entries, local_index; }; // = dge_dist_child
__global__ void k_dist_child_columns(const DistChildDev *__restrict__ infos, size_t n, uint64_t *__restrict__ cb, uint32_t *__restrict__ umis)
{
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
    {
        cb[i] = infos[i].barcode; umis[i] = uint32_t(infos[i].umis_stat);
    }
}

// ---- sharded runs: candidates of the gathered children against THIS rank's cells, without the host (dge_dist_eval_children) ---------
__global__ void k_dist_entry_counts(const DistChildDev *__restrict__ infos, size_t n, uint32_t *__restrict__ cnt)
{
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) cnt[i] = infos[i].n_entries;
}

__global__ void k_dist_jobs(const int *__restrict__ nb_count, const uint32_t *__restrict__ nb_pc, const DistChildDev *__restrict__ infos,
                            const uint32_t *__restrict__ entry_off, const uint32_t *__restrict__ job_off, size_t n, ForeignJob *__restrict__ jobs)
{
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
    {
        const int c = nb_count[i];
        for (int k = 0; k < c; ++k) jobs[job_off[i] + k] = ForeignJob{entry_off[i], infos[i].n_entries, nb_pc[i * WL_K + size_t(k)]};
    }
}

// best local candidate of every child: largest 0.5 * I * (1/U_child + 1/U_nb) (RealBarcodesMergeStrategy.cpp:46-47, same operation order),
// exact ties counted and resolved towards the smallest barcode (the combination over ranks applies the same rule)
__global__ void k_dist_best(const int *__restrict__ nb_count, const uint32_t *__restrict__ nb_pc, const uint32_t *__restrict__ job_off,
                            const uint32_t *__restrict__ isect, const DistChildDev *__restrict__ infos, const uint32_t *__restrict__ pc_to_real,
                            const uint64_t *__restrict__ local_cb, const uint32_t *__restrict__ local_umis, size_t n,
                            DistResult *__restrict__ res, uint32_t *__restrict__ best_local)
{
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
    {
        const int c = nb_count[i] > 0 ? nb_count[i] : 0;
        DistResult r;
        r.best_fraction = 0; r.best_barcode = EMPTY64; r.n_neighbours = uint32_t(c); r.n_best = 0;
        uint32_t bl = NONE32;
        const double inv_child = __ddiv_rn(1., double(size_t(infos[i].umis_stat)));
        for (int k = 0; k < c; ++k)
        {
            const uint32_t nb = pc_to_real[nb_pc[i * WL_K + size_t(k)]];
            const double frac = __dmul_rn(__dmul_rn(0.5, double(isect[job_off[i] + k])), __dadd_rn(inv_child, __ddiv_rn(1., double(local_umis[nb]))));
            const unsigned long long cb = local_cb[nb];
            if (r.n_best == 0 || r.best_fraction < frac) { r.best_fraction = frac; r.best_barcode = cb; r.n_best = 1; bl = nb; }
            else if (frac == r.best_fraction)
            {
                ++r.n_best;
                if (cb < r.best_barcode) { r.best_barcode = cb; bl = nb; }
            }
        }
        res[i] = r;
        best_local[i] = bl;
    }
}

// ---- applying merges -------------------------------------------------------------------------------------------------
struct MoveJob { uint32_t src_pc, dst_slot, out_off; };   // out_off = exclusive prefix of the source sizes

// copies the source cell's (gene, umi) list re-labelled to the destination slot, as sortcombine input ([ukey|000], val)
__global__ void __launch_bounds__(256) k_gather_relabel(const MoveJob *__restrict__ jobs, uint32_t n_jobs, const uint64_t *__restrict__ ukey,
                                                        const uint32_t *__restrict__ uval, const uint32_t *__restrict__ pc_u_start, int gub,
                                                        uint64_t *__restrict__ out_keys, uint32_t *__restrict__ out_vals)
{
    const uint64_t gu_mask = (1ull << gub) - 1;
    for (uint32_t j = blockIdx.x; j < n_jobs; j += gridDim.x)
    {
        const MoveJob job = jobs[j];
        const uint32_t s = pc_u_start[job.src_pc], e = pc_u_start[job.src_pc + 1];
        const uint64_t prefix = uint64_t(job.dst_slot) << gub;
        for (uint32_t i = s + threadIdx.x; i < e; i += blockDim.x)
        {
            out_keys[job.out_off + (i - s)] = (prefix | (ukey[i] & gu_mask)) << 3;
            out_vals[job.out_off + (i - s)] = uval[i];
        }
    }
}

// For every combined moved entry: if the destination already holds that (gene, umi) add into it, else keep it as an "extra".
// keep[i] = 1 for extras.
__global__ void __launch_bounds__(256) k_probe_merge(const uint64_t *__restrict__ ekey, const uint32_t *__restrict__ eval, uint32_t n_e,
                                                     const uint64_t *__restrict__ ukey, uint32_t *__restrict__ uval,
                                                     const uint32_t *__restrict__ slot_pc, const uint32_t *__restrict__ pc_u_start, int gub,
                                                     uint32_t *__restrict__ keep)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_e; i += gridDim.x * blockDim.x)
    {
        const uint64_t want = ekey[i];
        const uint32_t pc = slot_pc[uint32_t(want >> gub)];
        uint32_t k = 1;
        if (pc != NONE32)
        {
            uint32_t lo = pc_u_start[pc], hi = pc_u_start[pc + 1];
            const uint32_t end = hi;
            while (lo < hi)
            {
                uint32_t mid = (lo + hi) >> 1;
                if (ukey[mid] < want) lo = mid + 1; else hi = mid;
            }
            if (lo < end && ukey[lo] == want)
            {
                const uint32_t v = eval[i];
                uint32_t old = atomicAdd(&uval[lo], v & VAL_COUNT_MASK);
                uint32_t mk = v & ~VAL_COUNT_MASK;
                if ((old & mk) != mk) atomicOr(&uval[lo], mk);
                k = 0;
            }
        }
        keep[i] = k;
    }
}

__global__ void __launch_bounds__(256) k_compact_keep(const uint64_t *__restrict__ ekey, const uint32_t *__restrict__ eval, uint32_t n_e,
                                                      const uint32_t *__restrict__ keep, const uint32_t *__restrict__ keep_off,
                                                      uint64_t *__restrict__ okey, uint32_t *__restrict__ oval)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_e; i += gridDim.x * blockDim.x)
        if (keep[i]) { okey[keep_off[i]] = ekey[i]; oval[keep_off[i]] = eval[i]; }
}

// Merge of two sorted key/value arrays with disjoint keys (A = current list, B = extras) by ranking:
//   position of A[i] = i + lower_bound(B, A[i]);  position of B[j] = j + lower_bound(A, B[j]).
// A is the long list (all of U), B the short one: A is walked in tiles of consecutive elements whose ranks in B are confined to
// the range spanned by the tile's first and last key (two full searches per tile, a handful of steps per element); every B
// element does one full search over A.
__global__ void __launch_bounds__(256) k_merge_rank(const uint64_t *__restrict__ akey, const uint32_t *__restrict__ aval, uint32_t na,
                                                    const uint64_t *__restrict__ bkey, const uint32_t *__restrict__ bval, uint32_t nb,
                                                    uint64_t *__restrict__ okey, uint32_t *__restrict__ oval)
{
    constexpr uint32_t TILE = 2048;
    __shared__ uint32_t rng[2];
    auto lower_bound = [](const uint64_t *__restrict__ v, uint32_t lo, uint32_t hi, uint64_t key) {
        while (lo < hi)
        {
            const uint32_t mid = (lo + hi) >> 1;
            if (__ldg(v + mid) < key) lo = mid + 1; else hi = mid;
        }
        return lo;
    };
    const uint32_t n_tiles = (na + TILE - 1) / TILE;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
    {
        const uint32_t i0 = tile * TILE, i1 = min(na, i0 + TILE);
        if (threadIdx.x < 2) rng[threadIdx.x] = lower_bound(bkey, 0, nb, akey[threadIdx.x == 0 ? i0 : i1 - 1]);
        __syncthreads();
        const uint32_t lo0 = rng[0], hi0 = rng[1];
        for (uint32_t i = i0 + threadIdx.x; i < i1; i += blockDim.x)
        {
            const uint64_t key = akey[i];
            const uint32_t lo = lower_bound(bkey, lo0, hi0, key);
            okey[i + lo] = key;
            oval[i + lo] = aval[i];
        }
        __syncthreads();
    }
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < nb; j += gridDim.x * blockDim.x)
    {
        const uint64_t key = bkey[j];
        const uint32_t lo = lower_bound(akey, 0, na, key);
        okey[j + lo] = key;
        oval[j + lo] = bval[j];
    }
}

// slot -> present-cell index (NONE32 when the barcode owns no UMI)
__global__ void k_build_slot_pc(const uint32_t *__restrict__ pc_slot, uint32_t n_pc, uint32_t *__restrict__ slot_pc)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pc; i += gridDim.x * blockDim.x) slot_pc[pc_slot[i]] = i;
}

// ---- per-cell summaries for the host --------------------------------------------------------------------------------
struct CellRow
{
    unsigned long long cb;
    uint32_t slot, pc, first_idx, n_intergenic;
    uint32_t n_genes, n_umis, n_reads, req_genes, req_umis;
    uint32_t pad;
};

__global__ void k_p1_columns(const CellRow *__restrict__ rows, uint32_t n, uint64_t *__restrict__ cb, uint32_t *__restrict__ umis,
                             uint32_t *__restrict__ pc, uint32_t *__restrict__ pc_to_real)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const CellRow r = rows[i];
        cb[i] = r.cb; umis[i] = r.n_umis; pc[i] = r.pc;
        if (r.pc != NONE32) pc_to_real[r.pc] = i;
    }
}

__global__ void k_real_flags(const uint32_t *__restrict__ pc_cg_start, uint32_t n_pc, uint32_t min_genes, uint32_t *__restrict__ flags)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pc; i += gridDim.x * blockDim.x)
        flags[i] = (pc_cg_start[i + 1] - pc_cg_start[i]) >= min_genes;
}

__global__ void k_gather_rows_flagged(const uint32_t *__restrict__ flags, const uint32_t *__restrict__ off, uint32_t n_pc,
                                      const CellSlot *__restrict__ tab, const uint32_t *__restrict__ pc_slot,
                                      const uint32_t *__restrict__ pc_cg_start, const uint32_t *__restrict__ pc_u_start,
                                      const uint32_t *__restrict__ pc_reads, const uint32_t *__restrict__ pc_req_genes,
                                      const uint32_t *__restrict__ pc_req_umis, CellRow *__restrict__ rows)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pc; i += gridDim.x * blockDim.x)
    {
        if (!flags[i]) continue;
        CellRow r;
        r.slot = pc_slot[i]; r.pc = i;
        const CellSlot s = tab[r.slot];
        r.cb = s.cb; r.first_idx = s.first_idx; r.n_intergenic = s.n_intergenic;
        r.n_genes = pc_cg_start[i + 1] - pc_cg_start[i];
        r.n_umis = pc_u_start[i + 1] - pc_u_start[i];
        r.n_reads = pc_reads[i]; r.req_genes = pc_req_genes[i]; r.req_umis = pc_req_umis[i];
        r.pad = 0;
        rows[off[i]] = r;
    }
}

// ---- ordering of the real-cell rows on the device (cell-id order = first-seen order; compare_cells order) -------------
__global__ void k_rows_first_keys(const CellRow *__restrict__ rows, uint32_t n, uint64_t *__restrict__ key, uint32_t *__restrict__ idx)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { key[i] = rows[i].first_idx; idx[i] = i; }
}

__global__ void k_rows_permute(const CellRow *__restrict__ rows, const uint32_t *__restrict__ perm, uint32_t n, CellRow *__restrict__ out)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = rows[perm[i]];
}

// compare_cells (CellsDataContainer.cpp:329-344): ascending (requested genes, requested umis, TOTAL_UMIS stat, barcode).  Two stable
// radix sorts: first by barcode (k_rows_cb_keys), then by the packed counters genes : 16 | umis : 24 | stat : 24 (k_rows_filter_keys,
// values = the permutation left by the first sort).  *overflow is set when a counter does not fit its field (the host then sorts
// with full widths).
__global__ void k_rows_cb_keys(const CellRow *__restrict__ rows, uint32_t n, uint64_t *__restrict__ key, uint32_t *__restrict__ idx)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { key[i] = rows[i].cb; idx[i] = i; }
}

__global__ void k_rows_filter_keys(const CellRow *__restrict__ rows, const uint32_t *__restrict__ perm, uint32_t n, uint64_t *__restrict__ key,
                                   int *__restrict__ overflow)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const CellRow r = rows[perm[i]];
        if (r.req_genes >= (1u << 16) || r.req_umis >= (1u << 24) || r.n_umis >= (1u << 24)) *overflow = 1;
        key[i] = (uint64_t(r.req_genes) << 48) | (uint64_t(r.req_umis & 0xFFFFFFu) << 24) | uint64_t(r.n_umis & 0xFFFFFFu);
    }
}

// gene ids that occur, with their first read index as sort key (StringIndexer order, StringIndexer.cpp:10-18)
__global__ void k_gene_first_keys(const uint32_t *__restrict__ gene_first, uint32_t n_genes, uint64_t *__restrict__ key, uint32_t *__restrict__ idx)
{
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < n_genes; g += gridDim.x * blockDim.x) { key[g] = gene_first[g]; idx[g] = g; }
}

// rows for an explicit list of present cells (after merges)
__global__ void k_gather_rows_list(const uint32_t *__restrict__ pcs, uint32_t n, const CellSlot *__restrict__ tab,
                                   const uint32_t *__restrict__ pc_slot, const uint32_t *__restrict__ pc_cg_start,
                                   const uint32_t *__restrict__ pc_u_start, const uint32_t *__restrict__ pc_reads,
                                   const uint32_t *__restrict__ pc_req_genes, const uint32_t *__restrict__ pc_req_umis,
                                   CellRow *__restrict__ rows)
{
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x)
    {
        const uint32_t i = pcs[t];
        CellRow r;
        r.slot = pc_slot[i]; r.pc = i;
        const CellSlot s = tab[r.slot];
        r.cb = s.cb; r.first_idx = s.first_idx; r.n_intergenic = s.n_intergenic;
        r.n_genes = pc_cg_start[i + 1] - pc_cg_start[i];
        r.n_umis = pc_u_start[i + 1] - pc_u_start[i];
        r.n_reads = pc_reads[i]; r.req_genes = pc_req_genes[i]; r.req_umis = pc_req_umis[i];
        r.pad = 0;
        rows[t] = r;
    }
}

// ---- count matrices (ResultsPrinter.cpp:334-396) -----------------------------------------------------------------------
// column sizes: number of (cell, gene) rows of the column's cell with a non-zero value
__global__ void k_matrix_col_nnz(const uint32_t *__restrict__ col_pc, uint32_t n_cols, const uint32_t *__restrict__ pc_cg_start,
                                 const uint32_t *__restrict__ cg_req, int filtered, uint32_t *__restrict__ col_nnz)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t c = warp_global; c < n_cols; c += n_warps)
    {
        const uint32_t pc = col_pc[c];
        const uint32_t s = pc_cg_start[pc], e = pc_cg_start[pc + 1];
        uint32_t cnt = 0;
        if (filtered) { for (uint32_t i = s + lane; i < e; i += 32) cnt += cg_req[i] > 0; }
        else cnt = lane == 0 ? e - s : 0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) cnt += __shfl_down_sync(0xFFFFFFFFu, cnt, d);
        if (lane == 0) col_nnz[c] = cnt;
    }
}

// fill: one block per column, ordered compaction by block scan (genes stay ascending inside a column)
__global__ void __launch_bounds__(256) k_matrix_fill(const uint32_t *__restrict__ col_pc, uint32_t n_cols, const uint32_t *__restrict__ col_off,
                                                     const uint32_t *__restrict__ pc_cg_start, const uint32_t *__restrict__ cg_gene,
                                                     const uint32_t *__restrict__ values, const uint32_t *__restrict__ cg_start, int mode,
                                                     int32_t *__restrict__ out_gene, int32_t *__restrict__ out_val)
{
    // mode 0: value = values[i] (skip zeros); mode 1: value = cg_start[i+1]-cg_start[i] (all UMIs); mode 2: value = values[i], keep all
    __shared__ uint32_t ws[33];
    for (uint32_t c = blockIdx.x; c < n_cols; c += gridDim.x)
    {
        const uint32_t pc = col_pc[c];
        const uint32_t s = pc_cg_start[pc], e = pc_cg_start[pc + 1];
        uint32_t pos = col_off[c];
        for (uint32_t i0 = s; i0 < e; i0 += blockDim.x)
        {
            const uint32_t i = i0 + threadIdx.x;
            uint32_t v = 0;
            if (i < e) v = mode == 1 ? cg_start[i + 1] - cg_start[i] : values[i];
            const uint32_t keepit = (i < e && (mode != 0 || v > 0)) ? 1u : 0u;
            uint32_t total;
            const uint32_t ex = block_exclusive_scan(keepit, ws, &total);
            if (keepit)
            {
                out_gene[pos + ex] = int32_t(cg_gene[i]);
                out_val[pos + ex] = int32_t(v);
            }
            pos += total;
        }
    }
}

// Gene::number_of_requested_umis (Gene.cpp:60-79) for ANOTHER set of query marks than the container's: per (cell, gene) row, the UMIs (or
// their reads) whose accumulated mark is one of `mark_mask`'s.  Eight lanes share a row.  Feeds the matrices of `-V`
// (ResultsPrinter::save_intron_exon_matrices, ResultsPrinter.cpp:455-474).
__global__ void __launch_bounds__(256) k_cg_mark_values(const uint32_t *__restrict__ uval, const uint32_t *__restrict__ cg_start, uint32_t n_cg,
                                                         uint32_t mark_mask, int reads, uint32_t *__restrict__ out)
{
    const uint32_t sub = threadIdx.x & 7u;
    const uint32_t groups = (gridDim.x * blockDim.x) >> 3;
    for (uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) >> 3; base < ((n_cg + groups - 1) / groups) * groups; base += groups)
    {   // every lane of a warp runs the same number of rounds: the shuffles below need the whole warp
        uint32_t v = 0;
        if (base < n_cg)
        {
            const uint32_t s = cg_start[base], e = cg_start[base + 1];
            for (uint32_t i = s + sub; i < e; i += 8)
            {
                const uint32_t u = uval[i];
                if ((mark_mask >> (u >> VAL_MARK_SHIFT)) & 1u) v += reads ? (u & VAL_COUNT_MASK) : 1u;
            }
        }
        v += __shfl_xor_sync(0xFFFFFFFFu, v, 4);
        v += __shfl_xor_sync(0xFFFFFFFFu, v, 2);
        v += __shfl_xor_sync(0xFFFFFFFFu, v, 1);
        if (sub == 0 && base < n_cg) out[base] = v;
    }
}

// ---- device-resident merge flow (RealBarcodesMergeStrategy without host cell rows) ------------------------------------------------
// Per real cell (cell-id order, parallel to the CellRow table): the Stats counters that are counters and not set sizes
// (TOTAL_UMIS_PER_CB / TOTAL_READS_PER_CB are ADDED on merges, Stats.cpp:29-43), the merge target and the Cell flags.
struct CellState
{
    int32_t umis_stat, reads_stat;
    uint32_t n_intergenic;
    int32_t target;      // real-cell index, -1 = excluded by the merge, DF_TODO = the device pass could not decide
    uint32_t flags;      // DGE_CELL_REAL 1 | DGE_CELL_MERGED 2 | DGE_CELL_EXCLUDED 4
    uint32_t is_target;  // received at least one merged cell
};
constexpr int32_t DF_TODO = -3;

struct DevFlowCounters
{
    uint32_t n_todo, n_chain, n_merged, n_excluded, n_real, n_filtered, key_overflow, pad;
};

__global__ void k_state_init(const CellRow *__restrict__ rows, uint32_t n, CellState *__restrict__ st, DevFlowCounters *__restrict__ count_n_barcodes)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const CellRow r = rows[i];
        if (count_n_barcodes && (r.cb & CB_N_BIT)) atomicAdd(&count_n_barcodes->n_todo, 1u);
        CellState s;
        s.umis_stat = int32_t(r.n_umis); s.reads_stat = int32_t(r.n_reads); s.n_intergenic = r.n_intergenic;
        s.target = int32_t(i); s.flags = 1u; s.is_target = 0;
        st[i] = s;
    }
}

// Phase 1 outcome per cell: whitelist barcodes keep themselves, cells settled by k_p1_best take its target, everything else
// (far distance classes, order-dependent ties, neighbour overflow) is counted in n_todo: the exact host path then takes the run.
__global__ void k_p1_finalize(const int *__restrict__ nb_count, const int *__restrict__ p1_target, const uint32_t *__restrict__ p1_flag, uint32_t n,
                              CellState *__restrict__ st, DevFlowCounters *__restrict__ ctr)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const int c = nb_count[i];
        int32_t t;
        if (c == NB_SELF) t = int32_t(i);
        else if (c > 0 && !p1_flag[i]) t = p1_target[i];
        else { t = DF_TODO; atomicAdd(&ctr->n_todo, 1u); }
        st[i].target = t;
    }
}

// Sizes of the lists that move (phase 2, MergeStrategyBase.cpp:29-51).  With RealBarcodesMergeStrategy every target is a whitelist
// barcode that keeps itself, so no chain of merges exists and phase 2 is order-free; n_chain counts violations (-> host path).
__global__ void k_phase2_sizes(const CellRow *__restrict__ rows, const CellState *__restrict__ st, uint32_t n, uint32_t *__restrict__ move_size,
                               DevFlowCounters *__restrict__ ctr)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const int32_t t = st[i].target;
        uint32_t sz = 0;
        if (t >= 0 && uint32_t(t) != i)
        {
            if (st[t].target != t) atomicAdd(&ctr->n_chain, 1u);
            if (rows[i].pc != NONE32) sz = rows[i].n_umis;
        }
        move_size[i] = sz;
    }
}

// merge_cells (CellsDataContainer.cpp:90-104) for every cell with a foreign target: Stats::merge adds the source's counters, the
// source is flagged merged; target -1 excludes the cell (MergeStrategyBase.cpp:33-37).  One move job per cell (an empty one --
// the sentinel cell n_pc -- for cells that move nothing).
__global__ void k_phase2_apply(const CellRow *__restrict__ rows, CellState *__restrict__ st, uint32_t n, const uint32_t *__restrict__ move_off,
                               uint32_t empty_pc, MoveJob *__restrict__ moves, DevFlowCounters *__restrict__ ctr)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const int32_t t = st[i].target;
        MoveJob job{empty_pc, 0u, move_off[i]};
        if (t < -1) {} // merged into a cell of another rank (sharded runs): flagged by k_dist_flag_remote, nothing moves locally
        else if (t < 0)
        {
            st[i].flags = (st[i].flags & ~1u) | 4u;
            atomicAdd(&ctr->n_excluded, 1u);
        }
        else if (uint32_t(t) != i)
        {
            const CellRow r = rows[i];
            st[i].flags |= 2u;
            atomicAdd(&ctr->n_merged, 1u);
            atomicAdd(&st[t].umis_stat, int32_t(r.n_umis));
            atomicAdd(&st[t].reads_stat, int32_t(r.n_reads));
            atomicAdd(&st[t].n_intergenic, r.n_intergenic);
            st[t].is_target = 1u;
            if (r.pc != NONE32 && r.n_umis) job = MoveJob{r.pc, rows[t].slot, move_off[i]};
        }
        moves[i] = job;
    }
}

// After the lists were merged and the segment tables rebuilt: sizes of the merge targets (the only cells whose content changed),
// Cell::is_real (Cell.cpp:125-128) on the merged content, and the flag column for the cm_raw columns.
__global__ void k_refresh_rows(CellRow *__restrict__ rows, CellState *__restrict__ st, uint32_t n, const uint32_t *__restrict__ pc_cg_start,
                               const uint32_t *__restrict__ pc_u_start, const uint32_t *__restrict__ pc_req_genes, const uint32_t *__restrict__ pc_req_umis,
                               uint32_t min_genes, uint32_t *__restrict__ real_flag)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        CellState s = st[i];
        uint32_t n_genes = rows[i].n_genes;
        if (s.is_target)
        {
            const uint32_t pc = rows[i].pc;
            if (pc != NONE32)
            {
                n_genes = pc_cg_start[pc + 1] - pc_cg_start[pc];
                rows[i].n_genes = n_genes;
                rows[i].n_umis = pc_u_start[pc + 1] - pc_u_start[pc];
                rows[i].req_genes = pc_req_genes[pc];
                rows[i].req_umis = pc_req_umis[pc];
            }
        }
        const bool real = !(s.flags & 6u) && n_genes >= min_genes;
        s.flags = (s.flags & ~1u) | (real ? 1u : 0u);
        st[i].flags = s.flags;
        real_flag[i] = real ? 1u : 0u;
    }
}

// compare_cells keys (CellsDataContainer.cpp:329-344) of the final filter: `perm` is the barcode order (stable first sort); cells
// that are not real or have too few requested genes get the largest key and sort to the end; n_filtered counts the others.
__global__ void k_final_filter_keys(const CellRow *__restrict__ rows, const CellState *__restrict__ st, const uint32_t *__restrict__ perm, uint32_t n,
                                    uint32_t min_req_genes, uint64_t *__restrict__ key, DevFlowCounters *__restrict__ ctr)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const uint32_t c = perm[i];
        const CellRow r = rows[c];
        const CellState s = st[c];
        if (!(s.flags & 1u) || r.req_genes < min_req_genes) { key[i] = EMPTY64; continue; }
        const uint32_t stat = uint32_t(s.umis_stat);
        if (r.req_genes >= (1u << 16) - 1 || r.req_umis >= (1u << 24) || stat >= (1u << 24)) ctr->key_overflow = 1;
        key[i] = (uint64_t(r.req_genes) << 48) | (uint64_t(r.req_umis & 0xFFFFFFu) << 24) | uint64_t(stat & 0xFFFFFFu);
        atomicAdd(&ctr->n_filtered, 1u);
    }
}

// matrix columns: present-cell index of the listed real cells (cells without UMIs -> the empty sentinel cell)
__global__ void k_cols_from_list(const CellRow *__restrict__ rows, const uint32_t *__restrict__ list, uint32_t n, uint32_t empty_pc, uint32_t *__restrict__ cols)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const uint32_t pc = rows[list[i]].pc;
        cols[i] = pc == NONE32 ? empty_pc : pc;
    }
}

__global__ void k_cols_from_flags(const CellRow *__restrict__ rows, const uint32_t *__restrict__ flag, const uint32_t *__restrict__ off, uint32_t n,
                                  uint32_t empty_pc, uint32_t *__restrict__ cols)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        if (flag[i])
        {
            const uint32_t pc = rows[i].pc;
            cols[off[i]] = pc == NONE32 ? empty_pc : pc;
        }
}

} // namespace dge
