// poisson.cuh -- device side of the "precise merge" (-M) strategies: PoissonTargetEstimator (reference
// Estimation/Merge/PoissonTargetEstimator.cpp:14-119) on top of Tools::CollisionsAdjuster (collisions.cuh), used by
// PoissonSimpleMergeStrategy (PoissonSimpleMergeStrategy.cpp:15-42) and PoissonRealBarcodesMergeStrategy (.cpp:20-51).
//   reference                                                        here
//   CellsDataContainer::umi_distribution (:182-197)                  k_umi_hist: per packed UMI, the number of (cell, gene) pairs of filtered
//                                                                    cells that hold it; p_i = count / sum (k_hist_to_prob)
//   estimate_genes_intersection_size (:92-119), memoised per        the distinct adjusted size pairs of ALL candidate pairs are collected
//     (adjusted size a <= b): sum_i (1-(1-p_i)^a)(1-(1-p_i)^a(1-p_i)^(b-a))   in a hash set (k_pp_shared) and evaluated once each, one block
//                                                                    per pair over the UMI space (k_pp_est; Tools::fpow's multiplication order)
//   estimate_intersection_prob (:68-90): lambda = sum over shared    k_pp_lambda: merge join of the two cells' (cell, gene) rows, lambda summed in
//     genes, p = ppois(I - 1, lambda, lower = FALSE) = P[X >= I]     gene-id order, upper Poisson tail from its definition
//   get_best_merge_target (:14-44)                                   k_pp_best: argmin p per base, threshold max_prob / #neighbours
// FP64.  The reference sums est over the UMIs in the iteration order of an unordered_map<string,...> and lambda over the genes in
// StringIndexer order; here both sums run in a fixed device order.  The values therefore agree to rounding (~1e-16 relative), and they
// only feed comparisons (argmin, threshold): a base whose two smallest probabilities or whose threshold test are closer than 1e-9
// relative is reported ambiguous and its candidate ORDER is replayed on the host (exact ties are the case that occurs in practice).
#pragma once
#include "collisions.cuh"
#include "common.cuh"
#include "simplemerge.cuh"

namespace dge
{

__global__ void k_umi_hist(const uint64_t *__restrict__ ukey, uint32_t n_u, int ub, int gub, const uint32_t *__restrict__ slot_pc,
                           const uint32_t *__restrict__ pc_real, unsigned long long *__restrict__ hist)
{
    const uint64_t umask = (1ull << ub) - 1;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_u; i += gridDim.x * blockDim.x)
    {
        const uint64_t k = ukey[i];
        const uint32_t pc = slot_pc[uint32_t(k >> gub)];
        if (pc != NONE32 && pc_real[pc]) atomicAdd(&hist[k & umask], 1ull);
    }
}

__global__ void k_hist_total(const unsigned long long *__restrict__ hist, size_t n, unsigned long long *__restrict__ total)
{
    unsigned long long c = 0;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) c += hist[i];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_down_sync(0xFFFFFFFFu, c, d);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(total, c);
}

__global__ void k_hist_to_prob(const unsigned long long *__restrict__ hist, size_t n, const unsigned long long *__restrict__ total, double *__restrict__ p)
{
    const double sum = double(*total);
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
        p[i] = hist[i] ? __ddiv_rn(double(hist[i]), sum) : 0.0;
}

// largest Gene::size() among the real cells (bounds the CollisionsAdjuster table)
__global__ void k_max_gene_size(const uint32_t *__restrict__ cg_start, const uint32_t *__restrict__ cg_pc, uint32_t n_cg, const uint32_t *__restrict__ pc_real,
                                uint32_t *__restrict__ out)
{
    uint32_t m = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_cg; i += gridDim.x * blockDim.x)
        if (pc_real[cg_pc[i]]) m = max(m, cg_start[i + 1] - cg_start[i]);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = max(m, __shfl_down_sync(0xFFFFFFFFu, m, d));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

// candidate pairs (base a, other b) of the sorted pair list whose barcodes are within the edit distance (PoissonSimpleMergeStrategy.cpp:27-30: <=)
__global__ void k_pp_admissible(const uint64_t *__restrict__ pkey, uint32_t n_p, int rb, const uint64_t *__restrict__ cb, int cb_len, int max_ed,
                                uint32_t *__restrict__ flag)
{
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n_p; p += gridDim.x * blockDim.x)
    {
        const uint64_t k = pkey[p];
        const uint32_t a = uint32_t(k >> rb), b = uint32_t(k & ((1ull << rb) - 1));
        flag[p] = packed_edit_distance(cb[a], cb[b], cb_len) <= max_ed ? 1u : 0u;
    }
}

// The distinct adjusted size pairs (a <= b) live in an open-addressing hash set in global memory: key = (a << 29) | b, EMPTY64 = free.
// A few pairs -- (1,1), (1,2), ... -- occur millions of times, so the set is filled by insert-if-absent (no sorting of duplicates).
__device__ __forceinline__ bool sp_insert(unsigned long long *__restrict__ set, uint32_t mask, unsigned long long key)
{
    uint32_t s = uint32_t(mix64(key)) & mask;
    for (uint32_t probes = 0; probes <= mask; ++probes)
    {
        unsigned long long cur = set[s];
        if (cur == key) return true;
        if (cur == EMPTY64)
        {
            cur = atomicCAS(&set[s], EMPTY64, key);
            if (cur == EMPTY64 || cur == key) return true;
        }
        s = (s + 1) & mask;
    }
    return false;
}

__device__ __forceinline__ uint32_t sp_find(const unsigned long long *__restrict__ set, uint32_t mask, unsigned long long key)
{
    uint32_t s = uint32_t(mix64(key)) & mask;
    while (set[s] != key) s = (s + 1) & mask; // every looked-up key was inserted before
    return s;
}

// merge join of the (cell, gene) rows of both cells of every admissible pair: the adjusted size pair of each shared gene goes into the set;
// n_inserted counts NEW keys (load factor check), *full is raised when the table is exhausted
__global__ void k_pp_shared(const uint64_t *__restrict__ pkey, const uint32_t *__restrict__ flag, uint32_t n_p, int rb, const uint32_t *__restrict__ real_pc,
                            const uint32_t *__restrict__ pc_cg_start, const uint32_t *__restrict__ cg_gene, const uint32_t *__restrict__ cg_start,
                            const unsigned long long *__restrict__ adj, unsigned long long *__restrict__ set, uint32_t mask, int *__restrict__ full)
{
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n_p; p += gridDim.x * blockDim.x)
    {
        if (!flag[p]) continue;
        const uint64_t k = pkey[p];
        const uint32_t pa = real_pc[uint32_t(k >> rb)], pb = real_pc[uint32_t(k & ((1ull << rb) - 1))];
        uint32_t i = pc_cg_start[pa], ie = pc_cg_start[pa + 1], j = pc_cg_start[pb], je = pc_cg_start[pb + 1];
        unsigned long long last = EMPTY64;
        while (i < ie && j < je)
        {
            const uint32_t gi = cg_gene[i], gj = cg_gene[j];
            if (gi == gj)
            {
                unsigned long long s1 = adj[cg_start[i + 1] - cg_start[i] - 1], s2 = adj[cg_start[j + 1] - cg_start[j] - 1];
                if (s1 > s2) { const unsigned long long t = s1; s1 = s2; s2 = t; }
                const unsigned long long key = (s1 << 29) | s2;
                if (key != last) { if (!sp_insert(set, mask, key)) *full = 1; last = key; }
                ++i; ++j;
            }
            else if (gi < gj) ++i; else ++j;
        }
    }
}

__global__ void k_sp_occupied(const unsigned long long *__restrict__ set, uint32_t cap, uint32_t *__restrict__ occ)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += gridDim.x * blockDim.x) occ[i] = set[i] != EMPTY64;
}

__global__ void k_sp_list(const unsigned long long *__restrict__ set, const uint32_t *__restrict__ occ, const uint32_t *__restrict__ off, uint32_t cap,
                          uint32_t *__restrict__ slots)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += gridDim.x * blockDim.x)
        if (occ[i]) slots[off[i]] = i;
}

// est(a, b) = sum_i (1 - q_i^a) (1 - q_i^a q_i^(b-a)), q_i = 1 - p_i (PoissonTargetEstimator.cpp:109-115), one block per distinct pair
__global__ void __launch_bounds__(256) k_pp_est(const unsigned long long *__restrict__ set, const uint32_t *__restrict__ slots, uint32_t n_pairs,
                                                const double *__restrict__ p, size_t n_umi, double *__restrict__ est)
{
    __shared__ double red[256];
    for (uint32_t u = blockIdx.x; u < n_pairs; u += gridDim.x)
    {
        const uint32_t slot = slots[u];
        const uint64_t k = set[slot];
        const long long a = (long long)(k >> 29), b = (long long)(k & ((1ull << 29) - 1));
        double acc = 0;
        for (size_t i = threadIdx.x; i < n_umi; i += 256)
        {
            const double pi = p[i];
            if (pi == 0.0) continue; // UMIs that never occur are not part of the distribution (and would add exact zeros)
            const double q = __dsub_rn(1.0, pi);
            const double min_prob = ca_fpow(q, a);
            const double max_prob = __dmul_rn(min_prob, ca_fpow(q, b - a));
            acc = __dadd_rn(acc, __dmul_rn(__dsub_rn(1.0, min_prob), __dsub_rn(1.0, max_prob)));
        }
        red[threadIdx.x] = acc;
        __syncthreads();
        for (int w = 128; w > 0; w >>= 1)
        {
            if (int(threadIdx.x) < w) red[threadIdx.x] = __dadd_rn(red[threadIdx.x], red[threadIdx.x + w]);
            __syncthreads();
        }
        if (threadIdx.x == 0) est[slot] = red[0];
        __syncthreads();
    }
}

// P[X > x], X ~ Poisson(lambda): R's ppois(x, lambda, lower.tail = FALSE) from its definition (terms exp(-l + k ln l - lgamma(k + 1)))
__device__ inline double poisson_upper(long long x, double lambda)
{
    if (x < 0) return 1.0;
    if (lambda <= 0) return 0.0;
    const double ll = log(lambda);
    if (double(x + 1) < lambda)
    {   // the lower tail is the short side: 1 - P[X <= x], summed from the largest term down
        double s = 0;
        for (long long k = x; k >= 0; --k)
        {
            const double t = exp(-lambda + double(k) * ll - lgamma(double(k) + 1.0));
            s += t;
            if (double(k) < lambda && t < s * 1e-17) break;
        }
        return 1.0 - (s > 1 ? 1.0 : s);
    }
    double s = 0;
    for (long long k = x + 1;; ++k)
    {
        const double t = exp(-lambda + double(k) * ll - lgamma(double(k) + 1.0));
        s += t;
        if (t <= s * 1e-17 || k > x + 100000) break;
    }
    return s > 1 ? 1.0 : s;
}

// lambda and merge probability of every admissible pair (estimate_intersection_prob, PoissonTargetEstimator.cpp:68-90)
__global__ void k_pp_lambda(const uint64_t *__restrict__ pkey, const uint32_t *__restrict__ pval, const uint32_t *__restrict__ flag, uint32_t n_p, int rb,
                            const uint32_t *__restrict__ real_pc, const uint32_t *__restrict__ pc_cg_start, const uint32_t *__restrict__ cg_gene,
                            const uint32_t *__restrict__ cg_start, const unsigned long long *__restrict__ adj, const unsigned long long *__restrict__ set,
                            uint32_t mask, const double *__restrict__ est, double *__restrict__ prob)
{
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n_p; p += gridDim.x * blockDim.x)
    {
        if (!flag[p]) { prob[p] = 2.0; continue; }
        const uint64_t k = pkey[p];
        const uint32_t pa = real_pc[uint32_t(k >> rb)], pb = real_pc[uint32_t(k & ((1ull << rb) - 1))];
        uint32_t i = pc_cg_start[pa], ie = pc_cg_start[pa + 1], j = pc_cg_start[pb], je = pc_cg_start[pb + 1];
        double lambda = 0;
        while (i < ie && j < je)
        {
            const uint32_t gi = cg_gene[i], gj = cg_gene[j];
            if (gi == gj)
            {
                unsigned long long s1 = adj[cg_start[i + 1] - cg_start[i] - 1], s2 = adj[cg_start[j + 1] - cg_start[j] - 1];
                if (s1 > s2) { const unsigned long long t = s1; s1 = s2; s2 = t; }
                lambda = __dadd_rn(lambda, est[sp_find(set, mask, (s1 << 29) | s2)]);
                ++i; ++j;
            }
            else if (gi < gj) ++i; else ++j;
        }
        const long long isect = (long long)(pval[p] & VAL_COUNT_MASK);
        prob[p] = isect == 0 ? 1.0 : poisson_upper(isect - 1, lambda);
    }
}

struct PoissonBest
{
    uint32_t best;       // real idx of the neighbour with the smallest probability (first one in list order), NONE32 without neighbours
    uint32_t n_nb;       // neighbours within the edit distance
    double min_prob;
    uint32_t ambiguous;  // near-tie of the two smallest probabilities, or of the threshold test: the host replays the reference's order
    uint32_t pad;
};

__global__ void k_pp_best(const uint64_t *__restrict__ pkey, const uint32_t *__restrict__ flag, const double *__restrict__ prob, uint32_t n_p, int rb,
                          double max_prob, PoissonBest *__restrict__ out)
{
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n_p; p += gridDim.x * blockDim.x)
    {
        const uint32_t a = uint32_t(pkey[p] >> rb);
        if (p > 0 && uint32_t(pkey[p - 1] >> rb) == a) continue;
        double best = 2.0, second = 2.0;
        uint32_t bi = NONE32, n_nb = 0;
        for (uint32_t q = p; q < n_p && uint32_t(pkey[q] >> rb) == a; ++q)
        {
            if (!flag[q]) continue;
            ++n_nb;
            const double v = prob[q];
            if (v < best) { second = best; best = v; bi = uint32_t(pkey[q] & ((1ull << rb) - 1)); }
            else if (v < second) second = v;
        }
        PoissonBest r;
        r.best = bi; r.n_nb = n_nb; r.min_prob = best; r.pad = 0;
        const double thr = n_nb ? max_prob / double(n_nb) : 0.0;
        const bool near_tie = n_nb > 1 && second <= best * (1.0 + 1e-9) + 1e-300;
        const bool near_thr = n_nb > 0 && fabs(best - thr) <= 1e-9 * thr;
        r.ambiguous = (near_tie || near_thr) ? 1u : 0u;
        out[a] = r;
    }
}

} // namespace dge
