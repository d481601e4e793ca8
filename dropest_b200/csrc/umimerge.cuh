// umimerge.cuh -- directional UMI merge (`-u`, MergeUMIsStrategyDirectional) on the sorted distinct list U.
//
// Reference semantics (MergeUMIsStrategyDirectional.cpp:18-116), per (real cell, gene) segment:
//   * the UMIs are listed in UMI-id order (= first-seen order of the UMI string, StringIndexer) and std::sort-ed by read count;
//   * a source merges into the first UMI, scanning from the largest, with reads_src * mult <= reads_dst and
//     edit_distance <= max (first one with the smallest distance; distance <= 1 stops the scan);
//   * chains are followed to their root (the descending second loop compresses completely);
//   * Gene::merge adds read counts and ORs marks (Gene.cpp:38-58), Cell::merge_umis decrements TOTAL_UMIS_PER_CB once per
//     merged UMI (Cell.cpp:31-42).
// The winner of a source is therefore argmin over (distance asc, reads desc, position desc).  Position among EQUAL read
// counts is whatever std::sort left: for n <= 16 libstdc++ runs a plain insertion sort (stable -> first-seen order), for
// larger n it is introsort.  The kernel resolves everything except exact (distance, reads) ties inside segments of more than 16
// UMIs; those segments (and oversized ones) are listed for an exact host replay with the same std::sort.
#pragma once
#include "common.cuh"

namespace dge
{

struct UmiDirParams
{
    int ub;          // bits of a packed UMI
    int len;         // bases
    unsigned max_ed; // MergeUMIsStrategyDirectional::_max_edit_distance
    double mult;     // _mult
    int n_bit;       // bit of the UMI field that marks an index into the N-UMI list (allow_n), -1 = no such UMIs
};

constexpr int UMI_WARP_CAP = 128;   // UMIs per segment handled by one warp
constexpr int UMI_BLOCK_CAP = 8192; // ... by one block (dynamic shared memory)
constexpr int UMI_STABLE_N = 16;    // std::sort is an insertion sort up to here (libstdc++ _S_threshold)

// Tools::edit_distance (UtilFunctions.cpp:32-65) on 2-bit packed sequences of equal length (no N): the same banded column
// recurrence, including the boundary cell it overwrites and the early exit (see whitelist.hpp:edit_distance_ref).
__device__ inline unsigned umi_edit_distance(uint32_t a, uint32_t b, int len, unsigned max_ed)
{
    if (max_ed <= 1)
    {   // equal-length, N-free, distinct strings: distance <= 1 <=> exactly one substitution (checked exhaustively on the host side)
        uint32_t x = a ^ b;
        x = (x | (x >> 1)) & 0x55555555u;
        const unsigned hd = unsigned(__popc(x));
        return hd <= max_ed ? hd : max_ed + 1;
    }
    int col[17];
    for (int i = 0; i <= len; ++i) col[i] = i;
    const int band = int(max_ed);
    for (int j = 1; j <= len; ++j)
    {
        const int first = max(0, j - band), last = min(len, j + band);
        int diag = col[first];
        col[first] = j;
        int row_best = j;
        const uint32_t cb = (b >> (2 * (len - j))) & 3u;
        for (int i = first + 1; i <= last; ++i)
        {
            const int above = col[i];
            const uint32_t ca = (a >> (2 * (len - i))) & 3u;
            const int v = min(min(above + 1, col[i - 1] + 1), diag + (ca == cb ? 0 : 1));
            row_best = min(row_best, v + abs(i - j));
            col[i] = v;
            diag = above;
        }
        if (unsigned(row_best) > max_ed) return unsigned(row_best);
    }
    return unsigned(col[len]);
}

struct UmiDirOut
{
    uint32_t *big_list, *big_count;   // segments too large for a warp (cg indices)
    uint32_t big_cap;
    uint32_t *host_list, *host_count; // segments that need the exact host replay
    uint32_t host_cap;
    uint32_t *pc_dec;                 // merged UMIs per present cell (TOTAL_UMIS_PER_CB decrement)
    unsigned long long *n_merged;
    uint32_t *u_target;               // optional (save_umi_merge_targets): per U entry, the U index of the root it was merged into
};

// One group (a warp, or a whole block) resolves one segment held in shared memory.
template <bool BLOCK>
__device__ __forceinline__ void umi_dir_segment(uint32_t *s_umi, uint32_t *s_cnt, uint32_t *s_first, uint32_t *s_tgt, const uint64_t *__restrict__ ukey,
                                                uint32_t *uval, const uint32_t *__restrict__ umi_first, uint32_t s, uint32_t n,
                                                uint32_t cg, uint32_t pc, const UmiDirParams &p, const UmiDirOut &o)
{
    const uint32_t G = BLOCK ? blockDim.x : 32u, t = BLOCK ? threadIdx.x : (threadIdx.x & 31u);
    auto sync = [&]() { if (BLOCK) __syncthreads(); else __syncwarp(); };
    const uint32_t umask = p.ub >= 32 ? 0xFFFFFFFFu : ((1u << p.ub) - 1);
    for (uint32_t i = t; i < n; i += G)
    {
        const uint32_t umi = uint32_t(ukey[s + i]) & umask;
        s_umi[i] = umi;
        s_cnt[i] = uval[s + i] & VAL_COUNT_MASK;
        s_first[i] = umi_first[umi];
    }
    sync();
    bool ambiguous = false;
    for (uint32_t i = t; i < n; i += G)
    {
        const uint32_t ui = s_umi[i], ci = s_cnt[i], fi = s_first[i];
        const double need = double(ci) * p.mult;
        uint32_t best = NONE32, best_ed = 0xFFFFFFFFu, best_cnt = 0, best_first = 0;
        bool tie = false;
        for (uint32_t j = 0; j < n; ++j)
        {
            const uint32_t cj = s_cnt[j];
            if (j == i || cj < ci || need > double(cj)) continue; // scan stops at the first UMI with too few reads
            if (cj == ci)
            {   // only reachable with mult <= 1: whether j sits above i depends on the order std::sort left among equals
                if (n > UMI_STABLE_N) { ambiguous = true; continue; }
                if (s_first[j] < fi) continue;
            }
            const unsigned ed = umi_edit_distance(ui, s_umi[j], p.len, p.max_ed);
            if (ed > p.max_ed) continue;
            if (best == NONE32 || ed < best_ed || (ed == best_ed && cj > best_cnt))
            {
                best = j; best_ed = ed; best_cnt = cj; best_first = s_first[j]; tie = false;
            }
            else if (ed == best_ed && cj == best_cnt)
            {
                if (n > UMI_STABLE_N) tie = true;
                else if (s_first[j] > best_first) { best = j; best_first = s_first[j]; } // stable order: later first-seen is scanned first
            }
        }
        s_tgt[i] = best;
        ambiguous |= tie;
    }
    const bool any_amb = BLOCK ? (__syncthreads_or(ambiguous) != 0) : (__any_sync(0xFFFFFFFFu, ambiguous) != 0);
    sync();
    if (any_amb)
    {
        if (t == 0)
        {
            const uint32_t at = atomicAdd(o.host_count, 1u);
            if (at < o.host_cap) o.host_list[at] = cg;
        }
        return;
    }
    for (uint32_t i = t; i < n; i += G)
    {
        uint32_t r = s_tgt[i];
        if (r == NONE32) continue;
        while (s_tgt[r] != NONE32) r = s_tgt[r]; // targets always sit strictly above their sources: no cycles
        const uint32_t v = uval[s + i];
        atomicAdd(&uval[s + r], v & VAL_COUNT_MASK);
        atomicOr(&uval[s + r], v & ~VAL_COUNT_MASK);
        if (o.u_target) o.u_target[s + i] = s + r;
    }
    sync(); // every read of a source value above happens before it is cleared below (roots are never sources)
    uint32_t m = 0;
    for (uint32_t i = t; i < n; i += G)
        if (s_tgt[i] != NONE32) { uval[s + i] = 0; ++m; } // tombstone, removed by the compaction that follows
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m += __shfl_xor_sync(0xFFFFFFFFu, m, d);
    if ((threadIdx.x & 31u) == 0 && m)
    {
        atomicAdd(&o.pc_dec[pc], m);
        atomicAdd(o.n_merged, (unsigned long long)m);
    }
    sync();
}

// Warp per segment over all (cell, gene) segments of real cells; larger segments are deferred.
__global__ void __launch_bounds__(256) k_umi_dir_warp(const uint64_t *__restrict__ ukey, uint32_t *uval, const uint32_t *__restrict__ cg_start,
                                                      const uint32_t *__restrict__ cg_pc, uint32_t n_cg, const uint32_t *__restrict__ pc_real,
                                                      const uint32_t *__restrict__ umi_first, UmiDirParams p, UmiDirOut o)
{
    __shared__ uint32_t sm[8][4][UMI_WARP_CAP];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t warps_total = gridDim.x * (blockDim.x >> 5);
    for (uint32_t base = (blockIdx.x * (blockDim.x >> 5) + warp) * 32u; base < n_cg; base += warps_total * 32u)
    {
        const uint32_t cg = base + lane;
        uint32_t s = 0, n = 0, pc = 0;
        if (cg < n_cg)
        {
            s = cg_start[cg];
            n = cg_start[cg + 1] - s;
            pc = cg_pc[cg];
            // UMIs with N sort last in their segment (the marker bit is the top bit of the UMI field): such a segment -- even a single
            // N-UMI, which the reference renames with random bases -- is decided on the host, where the strings and rand() are
            if (p.n_bit >= 0 && n >= 1 && pc_real[pc] && ((uint32_t(ukey[s + n - 1]) >> p.n_bit) & 1u))
            {
                const uint32_t at = atomicAdd(o.host_count, 1u);
                if (at < o.host_cap) o.host_list[at] = cg;
                n = 0;
            }
            if (n < 2 || !pc_real[pc]) n = 0;
        }
        unsigned todo = __ballot_sync(0xFFFFFFFFu, n >= 2);
        while (todo)
        {
            const int l = __ffs(todo) - 1;
            todo &= todo - 1;
            const uint32_t ss = __shfl_sync(0xFFFFFFFFu, s, l), nn = __shfl_sync(0xFFFFFFFFu, n, l), pp = __shfl_sync(0xFFFFFFFFu, pc, l);
            const uint32_t cc = base + uint32_t(l);
            if (nn > uint32_t(UMI_WARP_CAP))
            {
                if (lane == 0)
                {
                    if (nn <= uint32_t(UMI_BLOCK_CAP))
                    {
                        const uint32_t at = atomicAdd(o.big_count, 1u);
                        if (at < o.big_cap) o.big_list[at] = cc;
                    }
                    else
                    {
                        const uint32_t at = atomicAdd(o.host_count, 1u);
                        if (at < o.host_cap) o.host_list[at] = cc;
                    }
                }
                continue;
            }
            umi_dir_segment<false>(sm[warp][0], sm[warp][1], sm[warp][2], sm[warp][3], ukey, uval, umi_first, ss, nn, cc, pp, p, o);
        }
    }
}

// Block per listed segment (UMI_WARP_CAP < n <= UMI_BLOCK_CAP).
__global__ void __launch_bounds__(256) k_umi_dir_block(const uint64_t *__restrict__ ukey, uint32_t *uval, const uint32_t *__restrict__ cg_start,
                                                       const uint32_t *__restrict__ cg_pc, const uint32_t *__restrict__ list, uint32_t n_list,
                                                       const uint32_t *__restrict__ umi_first, UmiDirParams p, UmiDirOut o)
{
    extern __shared__ uint32_t sm_dyn[];
    for (uint32_t k = blockIdx.x; k < n_list; k += gridDim.x)
    {
        const uint32_t cg = list[k];
        const uint32_t s = cg_start[cg], n = cg_start[cg + 1] - s;
        umi_dir_segment<true>(sm_dyn, sm_dyn + UMI_BLOCK_CAP, sm_dyn + 2 * UMI_BLOCK_CAP, sm_dyn + 3 * UMI_BLOCK_CAP, ukey, uval, umi_first, s, n, cg,
                              cg_pc[cg], p, o);
    }
}

// ---- exact host replay support --------------------------------------------------------------------------------------
__global__ void k_umi_seg_meta(const uint32_t *__restrict__ list, uint32_t n_list, const uint32_t *__restrict__ cg_start, const uint32_t *__restrict__ cg_pc,
                               uint32_t *__restrict__ seg_start, uint32_t *__restrict__ seg_n, uint32_t *__restrict__ seg_pc)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_list) return;
    const uint32_t cg = list[k];
    seg_start[k] = cg_start[cg];
    seg_n[k] = cg_start[cg + 1] - cg_start[cg];
    seg_pc[k] = cg_pc[cg];
}

// flat copy of the listed segments: (umi, count|mark, first-seen read index) per entry
__global__ void k_umi_seg_gather(const uint32_t *__restrict__ seg_start, const uint32_t *__restrict__ seg_off, uint32_t n_list,
                                 const uint64_t *__restrict__ ukey, const uint32_t *__restrict__ uval, const uint32_t *__restrict__ umi_first, int ub,
                                 uint32_t *__restrict__ out_umi, uint32_t *__restrict__ out_val, uint32_t *__restrict__ out_first)
{
    const uint32_t umask = ub >= 32 ? 0xFFFFFFFFu : ((1u << ub) - 1);
    for (uint32_t k = blockIdx.x; k < n_list; k += gridDim.x)
    {
        const uint32_t s = seg_start[k], o = seg_off[k], n = seg_off[k + 1] - o;
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
        {
            const uint32_t umi = uint32_t(ukey[s + i]) & umask;
            out_umi[o + i] = umi;
            out_val[o + i] = uval[s + i];
            out_first[o + i] = umi_first[umi];
        }
    }
}

// (source index, root index) pairs in U decided on the host; two phases so that sources are read before they are cleared
__global__ void k_umi_apply_pairs(const uint2 *__restrict__ pairs, uint32_t n, uint32_t *uval, int phase, uint32_t *u_target = nullptr)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint2 pr = pairs[k];
    if (phase == 0)
    {
        const uint32_t v = uval[pr.x];
        atomicAdd(&uval[pr.y], v & VAL_COUNT_MASK);
        atomicOr(&uval[pr.y], v & ~VAL_COUNT_MASK);
        if (u_target) u_target[pr.x] = pr.y;
    }
    else uval[pr.x] = 0;
}

// Gene::_merge_targets (Gene.cpp:54-57) of the merges decided above: (key of the source entry, UMI of its root), in no particular order
__global__ void k_umi_targets_collect(const uint64_t *__restrict__ ukey, const uint32_t *__restrict__ u_target, uint32_t n_u, int ub,
                                      uint64_t *__restrict__ out_key, uint32_t *__restrict__ out_dst, uint32_t cap, uint32_t *__restrict__ count)
{
    const uint32_t umask = ub >= 32 ? 0xFFFFFFFFu : ((1u << ub) - 1);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_u; i += gridDim.x * blockDim.x)
    {
        const uint32_t t = u_target[i];
        if (t == NONE32) continue;
        const uint32_t at = atomicAdd(count, 1u);
        if (at < cap) { out_key[at] = ukey[i]; out_dst[at] = uint32_t(ukey[t]) & umask; }
    }
}

// (cell, gene) segments of real cells that hold a UMI with N: inside a segment U is sorted by the UMI field and N-UMIs carry its top
// bit, so the segment's last entry tells
__global__ void k_seg_has_n(const uint64_t *__restrict__ ukey, const uint32_t *__restrict__ cg_start, const uint32_t *__restrict__ cg_pc, uint32_t n_cg,
                            const uint32_t *__restrict__ pc_real, int ub, uint32_t *__restrict__ list, uint32_t *__restrict__ count)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_cg; i += gridDim.x * blockDim.x)
    {
        if (!pc_real[cg_pc[i]]) continue;
        const uint32_t e = cg_start[i + 1];
        if (e > cg_start[i] && ((ukey[e - 1] >> (ub - 1)) & 1ull)) list[atomicAdd(count, 1u)] = i;
    }
}

__global__ void k_gather_u32(const uint32_t *__restrict__ src, const uint32_t *__restrict__ idx, uint32_t n, uint32_t *__restrict__ dst)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[idx[i]];
}

__global__ void k_zero_list(const uint32_t *__restrict__ idx, uint32_t n, uint32_t *__restrict__ uval)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) uval[idx[i]] = 0;
}

__global__ void k_flag_live(const uint32_t *__restrict__ uval, uint32_t n, uint32_t *__restrict__ keep)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) keep[i] = uval[i] != 0;
}

__global__ void k_flag_list(const uint32_t *__restrict__ list, uint32_t n, uint32_t *__restrict__ flags)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) flags[list[k]] = 1u;
}

} // namespace dge
