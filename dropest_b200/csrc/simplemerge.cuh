// simplemerge.cuh -- device side of SimpleMergeStrategy (reference Estimation/Merge/SimpleMergeStrategy.cpp:16-109), the cell-barcode
// merge used when no whitelist is given (Drop-seq style protocols).
//   reference                                              here
//   init(): unordered_map<(umi,gene), set<cell>> over      k_umig_keys + SortCombine: the (gene, umi, cell) triplets of all real cells
//           every UMI of every filtered cell (:90-103)     sorted by (gene, umi) -- the inverted index is a sorted array, a group = a run
//   get_cells_with_common_umigs (:16-41)                   k_pair_count / k_pair_write: ordered pairs (base, other) inside every run with
//                                                          size(other) >= size(base); SortCombine counts equal pairs = common UMI-genes
//   get_merge_target (:48-88): fraction, edit distance     k_pair_eval (fraction in FP64, Levenshtein on 2-bit barcodes) + k_base_best:
//           < max, EPS tie rule in hash-iteration order    best / runner-up per base; bases whose winner is not clear of the runner-up by
//                                                          more than 2*EPS are replayed on the host with the reference's own containers
#pragma once
#include "common.cuh"

namespace dge
{

struct UmigJob { uint32_t pc, real_idx, out_off; };

// Bijective scrambling of a (gene, umi) word inside its `bits` bits.  The inverted index only needs equal (gene, umi) to be adjacent, not
// any particular order of the groups; scrambling spreads the Zipf-distributed gene ids evenly over the radix buckets of the sort
// (without it a handful of L1 buckets would hold half of the keys and overflow the sub-bucket capacity).
__host__ __device__ inline uint64_t umig_scramble(uint64_t gu, int bits)
{
    const uint64_t mask = bits >= 64 ? ~0ull : ((1ull << bits) - 1);
    const int h = bits / 2 > 0 ? bits / 2 : 1;
    gu = (gu * 0x9E3779B97F4A7C15ull) & mask;
    gu ^= gu >> h;
    gu = (gu * 0xD6E8FEB86659FD93ull) & mask;
    gu ^= gu >> h;
    return gu;
}

// (gene, umi) of every UMI of the listed cells, tagged with the cell's index in the real-cell list: [scramble(gu) : gub | real_idx : rb] << 3
__global__ void __launch_bounds__(256) k_umig_keys(const UmigJob *__restrict__ jobs, uint32_t n_jobs, const uint64_t *__restrict__ ukey,
                                                   const uint32_t *__restrict__ pc_u_start, int gub, int rb, uint64_t *__restrict__ out_keys)
{
    const uint64_t gu_mask = (1ull << gub) - 1;
    for (uint32_t j = blockIdx.x; j < n_jobs; j += gridDim.x)
    {
        const UmigJob job = jobs[j];
        const uint32_t s = pc_u_start[job.pc], n = pc_u_start[job.pc + 1] - s;
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
            out_keys[job.out_off + i] = ((umig_scramble(ukey[s + i] & gu_mask, gub) << rb) | job.real_idx) << 3;
    }
}

// Element i of the sorted inverted index belongs to cell a = low rb bits; its run = all elements with the same (gene, umi).
// WRITE = false: pair_cnt[i] = number of cells b != a in the run with n_genes[b] >= n_genes[a]  (SimpleMergeStrategy.cpp:30-36)
// WRITE = true : the pairs themselves, [a : rb | b : rb] << 3, at pair_off[i]...
template <bool WRITE>
__global__ void __launch_bounds__(256) k_pairs(const uint64_t *__restrict__ ekey, uint32_t n_e, int rb, const uint32_t *__restrict__ n_genes,
                                               uint32_t *__restrict__ pair_cnt, const uint32_t *__restrict__ pair_off, uint64_t *__restrict__ out_pairs)
{
    const uint64_t rmask = (1ull << rb) - 1;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_e; i += gridDim.x * blockDim.x)
    {
        const uint64_t me = ekey[i];
        const uint64_t g = me >> rb;
        const uint32_t a = uint32_t(me & rmask);
        const uint32_t ga = n_genes[a];
        uint32_t cnt = 0;
        uint32_t pos = WRITE ? pair_off[i] : 0u;
        for (uint32_t j = i; j-- > 0;)
        {
            const uint64_t o = ekey[j];
            if ((o >> rb) != g) break;
            const uint32_t b = uint32_t(o & rmask);
            if (n_genes[b] >= ga) { if (WRITE) out_pairs[pos + cnt] = ((uint64_t(a) << rb) | b) << 3; ++cnt; }
        }
        for (uint32_t j = i + 1; j < n_e; ++j)
        {
            const uint64_t o = ekey[j];
            if ((o >> rb) != g) break;
            const uint32_t b = uint32_t(o & rmask);
            if (n_genes[b] >= ga) { if (WRITE) out_pairs[pos + cnt] = ((uint64_t(a) << rb) | b) << 3; ++cnt; }
        }
        if (!WRITE) pair_cnt[i] = cnt;
    }
}

// Levenshtein distance of two 2-bit packed sequences of `len` bases (Tools::edit_distance with its defaults, UtilFunctions.cpp:32-65:
// unbanded; N wildcards cannot occur in packed barcodes).
__device__ inline int packed_edit_distance(uint64_t x, uint64_t y, int len)
{
    uint8_t col[24];
    for (int i = 0; i <= len; ++i) col[i] = uint8_t(i);
    for (int j = 1; j <= len; ++j)
    {
        const uint32_t cy = uint32_t(y >> (2 * (len - j))) & 3u;
        uint8_t diag = col[0];
        col[0] = uint8_t(j);
        for (int i = 1; i <= len; ++i)
        {
            const uint32_t cx = uint32_t(x >> (2 * (len - i))) & 3u;
            const uint8_t up = col[i];
            const int sub = diag + (cx != cy);
            const int v = min(min(int(col[i - 1]) + 1, int(up) + 1), sub);
            diag = up;
            col[i] = uint8_t(v);
        }
    }
    return col[len];
}

// fraction and edit-distance test per distinct (base, other) pair; frac = -1 when the edit distance rules the pair out
__global__ void __launch_bounds__(256) k_pair_eval(const uint64_t *__restrict__ pkey, const uint32_t *__restrict__ pval, uint32_t n_p,
                                                   int rb, const uint64_t *__restrict__ cb, const uint32_t *__restrict__ umis, int cb_len, int max_ed,
                                                   double *__restrict__ frac)
{
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n_p; p += gridDim.x * blockDim.x)
    {
        const uint64_t k = pkey[p];
        const uint32_t a = uint32_t(k >> rb), b = uint32_t(k & ((1ull << rb) - 1));
        const double cnt = double(pval[p] & VAL_COUNT_MASK);
        const double f = __dmul_rn(__dmul_rn(0.5, cnt), __dadd_rn(__ddiv_rn(1., double(umis[a])), __ddiv_rn(1., double(umis[b]))));
        frac[p] = packed_edit_distance(cb[a], cb[b], cb_len) < max_ed ? f : -1.;
    }
}

struct BaseBest
{
    uint32_t best;      // real idx of the admissible candidate with the largest fraction, NONE32 when there is none
    uint32_t count;     // its number of common UMI-genes
    uint32_t ambiguous; // another admissible candidate lies within 2 * EPS of the best one: the reference's iteration order decides
    uint32_t pad;
};

// the head of every run of equal `base` in the sorted pair list reduces its run
__global__ void __launch_bounds__(256) k_base_best(const uint64_t *__restrict__ pkey, const uint32_t *__restrict__ pval, const double *__restrict__ frac,
                                                   uint32_t n_p, int rb, double eps2, BaseBest *__restrict__ out)
{
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n_p; p += gridDim.x * blockDim.x)
    {
        const uint32_t a = uint32_t(pkey[p] >> rb);
        if (p > 0 && uint32_t(pkey[p - 1] >> rb) == a) continue;
        double best = -1., second = -1.;
        uint32_t bi = NONE32, bc = 0;
        for (uint32_t q = p; q < n_p && uint32_t(pkey[q] >> rb) == a; ++q)
        {
            const double f = frac[q];
            if (f < 0) continue;
            if (f > best) { second = best; best = f; bi = uint32_t(pkey[q] & ((1ull << rb) - 1)); bc = pval[q] & VAL_COUNT_MASK; }
            else if (f > second) second = f;
        }
        BaseBest r;
        r.best = bi; r.count = bc; r.pad = 0;
        r.ambiguous = (bi != NONE32 && second >= 0 && second >= best - eps2) ? 1u : 0u;
        out[a] = r;
    }
}

// ---- MergeAllMergeStrategy (reference Estimation/Merge/MergeAllMergeStrategy.h:16-50) -------------------------------------
// Tools::edit_distance(a, b, skip_n = false, max_ed) on 2-bit packed barcodes: the banded single-column recurrence of
// UtilFunctions.cpp:32-65 kept literally (cells outside the band keep stale values, early return of the running row minimum).
__device__ inline unsigned packed_edit_distance_banded(uint64_t a, uint64_t b, int len, unsigned max_ed)
{
    int col[24];
    for (int i = 0; i <= len; ++i) col[i] = i;
    const int band = int(min(max_ed, 64u));
    for (int j = 1; j <= len; ++j)
    {
        const int first = max(0, j - band), last = min(len, j + band);
        int diag = col[first];
        col[first] = j;
        int row_best = j;
        const uint32_t cb = uint32_t(b >> (2 * (len - j))) & 3u;
        for (int i = first + 1; i <= last; ++i)
        {
            const int above = col[i];
            const uint32_t ca = uint32_t(a >> (2 * (len - i))) & 3u;
            const int v = min(min(above + 1, col[i - 1] + 1), diag + (ca == cb ? 0 : 1));
            row_best = min(row_best, v + abs(i - j));
            col[i] = v;
            diag = above;
        }
        if (unsigned(row_best) > max_ed) return unsigned(row_best);
    }
    return unsigned(col[len]);
}

// One block per base cell (cells in filtered_cells() order).  Candidates: cells with MORE UMIs within the edit distance; the winner has
// the smallest distance, then the most UMIs, then the earliest position (the reference's strict comparisons keep the first one).
__global__ void __launch_bounds__(256) k_merge_all_targets(const uint64_t *__restrict__ cb, const uint32_t *__restrict__ umis, uint32_t n, int cb_len,
                                                           unsigned max_ed, uint32_t *__restrict__ target)
{
    __shared__ unsigned long long best_s;
    for (uint32_t base = blockIdx.x; base < n; base += gridDim.x)
    {
        if (threadIdx.x == 0) best_s = ~0ull;
        __syncthreads();
        const uint64_t my_cb = cb[base];
        const uint32_t my_umis = umis[base];
        unsigned long long best = ~0ull;
        for (uint32_t j = threadIdx.x; j < n; j += blockDim.x)
        {
            const uint32_t u = umis[j];
            if (u <= my_umis) continue;
            const unsigned ed = packed_edit_distance_banded(my_cb, cb[j], cb_len, max_ed);
            if (ed > max_ed) continue;
            const unsigned long long key = ((unsigned long long)(ed & 63u) << 58) | ((unsigned long long)(0x7FFFFFFFu - u) << 27) | j;
            best = min(best, key);
        }
        atomicMin(&best_s, best);
        __syncthreads();
        if (threadIdx.x == 0) target[base] = best_s == ~0ull ? base : uint32_t(best_s & ((1ull << 27) - 1));
        __syncthreads();
    }
}

} // namespace dge
