// collisions.cuh -- Tools::CollisionsAdjuster on the device (reference Tools/CollisionsAdjuster.cpp:12-49, Tools::fpow
// UtilFunctions.cpp:13-30): adjusted[s] = lround(s + sum of expected UMI collisions) for s = 1..max, where every step updates, for
// EVERY UMI i, neg_prod[i] *= (1 - p_i)^(total - last_total) and sums q = sum_i p_i * (1 - neg_prod[i]).
// The recurrence is sequential in s and parallel over the UMI space (4^L entries): one launch per step, element-wise update in the
// reference's exact operation order (no FMA contraction), block partial sums, the last block finishes the step.
// FP64.  The sum over i is done in parallel, i.e. in another association order than the reference's sequential loop; only
// integers leave this module (adjusted sizes, and floor(sum_collisions) inside the recurrence), so a step is flagged as RISKY when
// the running sum comes closer to a rounding boundary than the worst-case reordering drift -- the caller then reruns in EXACT mode,
// where the terms are summed by one thread in index order (bit-identical to the reference by construction).
#pragma once
#include "common.cuh"

namespace dge
{

struct CollisionsState
{
    double sum_collisions;
    unsigned long long last_total;
    unsigned int ticket;
    unsigned int risky; // first risky step (0 = none)
};

__device__ __forceinline__ double ca_fpow(double base, long long exp) // UtilFunctions.cpp:13-30
{
    if (exp == 1) return base;
    double result = 1;
    while (exp)
    {
        if (exp & 1) result = __dmul_rn(result, base);
        exp >>= 1;
        base = __dmul_rn(base, base);
    }
    return result;
}

constexpr int CA_THREADS = 256;

// one step s of CollisionsAdjuster::update_adjusted_sizes (CollisionsAdjuster.cpp:21-39)
template <bool EXACT>
__global__ void __launch_bounds__(CA_THREADS) k_collisions_step(const double *__restrict__ p, double *__restrict__ neg_prod, double *__restrict__ terms,
                                                                size_t n, unsigned long long s, CollisionsState *__restrict__ st,
                                                                double *__restrict__ partial, unsigned long long *__restrict__ adjusted,
                                                                double drift_per_step)
{
    __shared__ double red[CA_THREADS];
    __shared__ bool last_block;
    const unsigned long long total = s + (unsigned long long)(st->sum_collisions);
    const long long d = (long long)(total - st->last_total);
    double acc = 0;
    for (size_t i = size_t(blockIdx.x) * CA_THREADS + threadIdx.x; i < n; i += size_t(gridDim.x) * CA_THREADS)
    {
        const double pi = p[i];
        const double np = __dmul_rn(neg_prod[i], ca_fpow(__dsub_rn(1.0, pi), d));
        neg_prod[i] = np;
        const double term = __dmul_rn(pi, __dsub_rn(1.0, np));
        if (EXACT) terms[i] = term; else acc = __dadd_rn(acc, term);
    }
    if (!EXACT)
    {
        red[threadIdx.x] = acc;
        __syncthreads();
        for (int w = CA_THREADS / 2; w > 0; w >>= 1)
        {
            if (int(threadIdx.x) < w) red[threadIdx.x] = __dadd_rn(red[threadIdx.x], red[threadIdx.x + w]);
            __syncthreads();
        }
        if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
    }
    __threadfence();
    __syncthreads(); // every thread of the block has read the step state and published its terms before the ticket is taken
    if (threadIdx.x == 0) last_block = atomicAdd(&st->ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last_block || threadIdx.x != 0) return;
    __threadfence();
    double q = 0;
    if (EXACT) { for (size_t i = 0; i < n; ++i) q = __dadd_rn(q, terms[i]); }   // the reference's own order
    else { for (unsigned b = 0; b < gridDim.x; ++b) q = __dadd_rn(q, partial[b]); }
    const double collision_num = __dsub_rn(__ddiv_rn(1.0, __dsub_rn(1.0, q)), 1.0);
    const double sum = __dadd_rn(st->sum_collisions, collision_num);
    st->sum_collisions = sum;
    st->last_total = total;
    st->ticket = 0;
    const double v = __dadd_rn(double(s), sum);
    adjusted[s - 1] = (unsigned long long)llround(v);
    if (!EXACT && st->risky == 0)
    {
        const double delta = drift_per_step * double(s);
        const double fv = v - floor(v), fs = sum - floor(sum);
        const bool near_half = fabs(fv - 0.5) < delta;
        const bool near_int = (fs < delta && sum >= 0.5) || (1.0 - fs) < delta;
        if (near_half || near_int) st->risky = (unsigned int)(s > 0xFFFFFFFFull ? 0xFFFFFFFFull : s);
    }
}

} // namespace dge
