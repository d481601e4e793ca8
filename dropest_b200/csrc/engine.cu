// engine.cu -- orchestration of the device pipeline behind the C ABI in include/dropest_b200.h.
//
// Stage map (reference call stack SURVEY.md 3.1-3.3):
//   dge_add_batch*        CellsDataContainer::add_record                 -> k_fill_compact (fill.cuh), one launch per batch
//   dge_set_initialized   (rest of add_record work) + set_initialized    -> SortCombine (sortcombine.cuh) + segments.cuh + real/filtered cells
//   dge_merge_and_filter  merge_and_filter: MergeStrategy::merge         -> merge.cuh kernels + host phase 2 (MergeStrategyBase.cpp:29-51)
//                         UMI merge, update_cell_sizes, matrices         -> segments.cuh + k_matrix_*
// There is no CPU implementation of the grouping here: without a CUDA device every entry point fails.
#include "../../include/dropest_b200.h"
#include "collisions.cuh"
#include "chrstats.cuh"
#include "distmerge.cuh"
#include "common.cuh"
#include "fill.cuh"
#include "merge.cuh"
#include "poisson.cuh"
#include "scan.cuh"
#include "segments.cuh"
#include "simplemerge.cuh"
#include "sortcombine.cuh"
#include "umimerge.cuh"
#include "whitelist.hpp"

#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <numeric>
#include <unordered_map>
#include <unordered_set>
#include <vector>

using namespace dge;

namespace
{
thread_local std::string g_create_error;

struct HostCell
{
    uint64_t cb = 0;
    uint32_t slot = 0, pc = 0, first_idx = 0, n_intergenic = 0;
    int32_t n_genes = 0, umis_stat = 0, reads_stat = 0, req_genes = 0, req_umis = 0;
    int32_t n_umis_distinct = 0;
    bool merged = false, excluded = false, real = true;
    int32_t target = -1; // index into `real`
    uint64_t merged_to_cb = EMPTY64; // sharded runs: barcode of a merge target living on another rank
};

struct KeyChunk
{
    DevBuf keys;
    size_t capacity = 0;
    size_t count = 0; // valid after counts were read back
    unsigned long long *d_count = nullptr;
};

struct MatrixDev
{
    DevBuf indptr, gene, val, cols;
    size_t n_cols = 0, nnz = 0;
    bool built = false;
    const uint32_t *cols_used = nullptr; // the column list the matrix was built from (device memory owned by the handle)
};
} // namespace

struct dge_handle
{
    dge_config cfg{};
    std::string barcodes_file;
    std::string err;
    std::string n_umi_strings;                // dge_set_n_strings: the strings behind DGE_FLAG_UMI_N indices (umi_len characters each)
    std::vector<std::string> n_cb_list;       // the strings behind DGE_FLAG_CB_N indices: barcodes with N and / or of another length than cb_len
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int state = 0; // 0 filling, 1 initialized, 2 merged
    bool device_ready = false;

    KeyLayout kl{};
    size_t table_cap = 0;
    uint32_t min_after_eff = 0;

    // fill-stage device state
    DevBuf tab, gene_first, umi_first, ctr, staging[2], ctr_counts;
    DevBuf regions, hist12, hist_fold, region_tiles, key_tiles; // k_fill_pipe output: per-block key regions of every batch + the 12-bit L1 histogram
    uint32_t n_regions = 0;
    bool pipe_fill = false;       // the batches of this run were filled by k_fill_pipe (regions instead of dense chunks)
    bool track_umi_first = false; // strategies that depend on the UMI indexer's first-seen order
    std::vector<std::unique_ptr<KeyChunk>> chunks, chunk_pool;
    std::vector<DevBuf *> chunk_counts;
    DevBuf chunk_count_pool;
    size_t n_chunk_counters = 0;
    uint64_t n_reads = 0;
    int staging_turn = 0;
    // per-(cell, chromosome) read counters (chrstats.cuh); allocated by the first batch that brings a chromosome array
    DevBuf chr_tab, chr_ctr, chr_export, chr_staging[2];
    size_t chr_cap = 0;
    bool chr_used = false;
    cudaEvent_t staging_ev[2] = {nullptr, nullptr};
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t copied_ev[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> fill_ev; // start/stop pairs around the fill kernel of every batch of the current run
    size_t n_fill_ev = 0;

    // grouped state
    DevBuf keys_all, ukey, uval, ukey2, uval2, overflow_flag;
    DevBuf cg_gene, cg_start, cg_pc, cg_req, cg_reads, cg_req_reads;
    DevBuf pc_slot, pc_u_start, pc_cg_start, pc_reads, pc_req_genes, pc_req_umis, slot_pc;
    DevBuf tile_a, tile_b, scan_scratch, flags, flags_off, rows_dev, rows_dev2, misc, cub_tmp, sort_k[2], sort_v[2], gsort_k[2], gsort_v[2];
    PinnedBuf pin_filtered, pin_fkeys, pin_gene_ids;
    uint32_t n_u = 0, n_cg = 0, n_pc = 0;
    size_t n_keys = 0;
    FillCounters counters{};
    uint64_t total_cells = 0;

    SortCombine sc, sc2; // sc2: the (much smaller) cell-merge pass, so neither resizes the other's workspaces
    SortCombineStats sc_stats;
    // merge-stage workspaces (grow only, reused across runs)
    DevBuf d_jobs, d_isect, d_cb, d_umis, d_count, d_nb, d_moves, mkeys, mvals, ekey, eval, keep, keep_off, xkey, xval, mat_nnz;
    uint32_t *isect_p = nullptr;
    std::vector<uint32_t> h_pc_to_real, h_nb_pc, h_umis, h_nb_off, h_nbs;
    std::vector<int> h_nb_count;
    std::vector<uint64_t> h_cbs;
    std::vector<PairJob> h_jobs;
    std::vector<MoveJob> h_moves;
    uint64_t moves_total = 0;
    bool moves_ready = false;
    std::vector<long> h_target;
    std::vector<uint64_t> h_sortkey;
    bool wl_uploaded = false;
    bool rows_on_device = false; // rows_dev2 holds the real cells in cell-id order (same order as `real`)
    DevBuf p1_pc, p1_map, p1_cnt, p1_off, p1_target, p1_flag;
    PinnedBuf pin_p1t, pin_p1f;
    // SimpleMergeStrategy workspaces
    DevBuf sm_jobs, sm_ikeys, sm_ekey, sm_eval, sm_ngenes, sm_umis, sm_cb, sm_pcnt, sm_poff, sm_pairs, sm_pkey, sm_pval, sm_frac, sm_best;
    PinnedBuf pin_best;
    uint64_t n_simple_replayed = 0;
    PinnedBuf pin_rows, pin_nbc, pin_nbp, pin_isect, pin_misc;
    // cross-rank merge state (sharded runs)
    DevBuf umi_lists, umi_ctr, umi_pc_dec, umi_seg, umi_flat, umi_pairs;
    PinnedBuf pin_umi;
    uint64_t n_umis_merged = 0, n_umi_segments_replayed = 0;
    // Gene::_merge_targets (Gene.cpp:54-57, kept when the container is built with save_umi_merge_targets): one row per UMI that the UMI
    // merge strategy moved into another one, codes as dge_get_umigs reports them
    struct UmiMergeTarget { uint64_t cb; uint32_t gene, src, dst; uint32_t created; }; // created: the target did not exist, the source's UMI object became it (Gene.cpp:48)
    std::vector<UmiMergeTarget> umi_mt;
    DevBuf u_target, mt_keys, mt_dst;
    DevBuf dist_infos, dist_keys, dist_vals, dist_jobs, dist_cb, dist_umis, dist_eoff;
    std::vector<uint32_t> g_off;
    const uint64_t *g_keys = nullptr;
    const uint32_t *g_vals = nullptr;
    std::vector<uint32_t> dist_targets;
    bool dist_done = false, slot_pc_built = false;
    uint64_t n_order_ties = 0;

    // host state
    std::vector<HostCell> real;            // cell-id (first-seen) order
    std::vector<uint32_t> filtered;        // indices into real, ascending compare_cells
    std::vector<int32_t> gene_order;       // gene ids in first-seen order
    std::vector<std::pair<uint32_t, uint32_t>> merge_events; // (src real idx, dst real idx) in application order
    uint64_t n_merged = 0, n_excluded = 0, n_unresolved = 0;
    Whitelist wl;
    bool wl_fast = false;
    DevBuf wl_tokens[WL_MAX_PARTS];
    WhitelistDev wl_dev{};

    // device-resident cell state: the host mirror (`real`, `filtered`, `gene_order`) is built lazily, only when the query surface
    // or a host-side strategy asks for it (materialize_host); the whitelist merge itself runs without it (merge_device_flow)
    DevBuf cell_state, df_ctr, move_size, move_off, real_flag, real_off, cols_f, cols_r, fsort_k[2], fsort_v[2];
    PinnedBuf pin_state, pin_ctr;
    uint32_t n_real_rows = 0;        // rows of rows_dev2 (real cells at set_initialized, cell-id order)
    uint32_t n_filtered_dev = 0;     // device flow: length of the final filtered list (fsort_v[..], see dev_filtered_buf)
    int dev_filtered_buf = 0;
    int init_filter_overflow = 0;    // packed compare_cells keys did not fit at set_initialized: the host re-sorts with full widths
    int host_stage = 0;              // 0 no host mirror yet, 1 mirror of set_initialized, 2 merge results applied
    bool lazy_rows = false;          // rows_dev2 / sort_v[0] / gsort_* hold everything the mirror needs
    bool dev_merged = false;         // the merge ran in the device flow
    uint64_t sum_real = 0, sum_filtered = 0, sum_genes_seen = 0, n_host_fallback = 0;

    // PoissonTargetEstimator state (-M): UMI distribution, CollisionsAdjuster table, per-pair work
    DevBuf pp_pc_real, pp_real_pc, pp_hist, pp_p, pp_adj, pp_cnt, pp_off, pp_spairs, pp_skey, pp_sval, pp_est, pp_prob, pp_flag, pp_best, pp_misc;
    PinnedBuf pin_pp;
    bool pp_ready = false;
    uint32_t pp_max_gene_size = 0;
    uint64_t n_poisson_replayed = 0;

    // cross-rank merge state machine (dge_dist_step)
    int dist_stage = 0;
    uint32_t dist_world = 0, dist_rank = 0, n_self = 0, n_all = 0, n_pairs = 0, g_mask = 0;
    uint64_t n_dist_slow = 0, n_dist_ties = 0;
    DevBuf x_self_flag, x_self_one, x_self_off, x_self_idx, x_self_send, x_all, x_gcb, x_ggi, x_rank_off, x_nb_count, x_nb_gi, x_pair_cnt, x_pair_off,
        x_pair_child, x_pair_gi, x_pair_pos, x_pair_eoff, x_pair_isect, x_dest, x_lay, x_send, x_local_jobs, x_local_isect, x_kept, x_rl, x_job_base,
        x_reply_base, x_fjobs, x_fisect, x_reply, x_best, x_tie, x_merged_cb, x_clay, x_ccur, x_commit, x_commit_send, x_cl, x_cbase, x_fmoves,
        x_fsize, x_foff, x_bad;
    std::vector<uint32_t> hx_rank_off, hx_reply_base_recv;
    std::vector<DistBlobLayout> hx_lay, hx_rl;
    std::vector<SelfRec> hx_all;
    std::vector<CellRow> hx_rows;

    MatrixDev cm, cm_raw, cm_marks; // cm_marks: the filtered matrix for another set of query marks (dge_get_matrix_marks), built on request
    uint32_t cm_marks_mask = 0;
    DevBuf cg_marks;
    dge_timings timings{};
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    unsigned launches = 0;

    ~dge_handle()
    {
        for (auto &e : ev) if (e) cudaEventDestroy(e);
        for (auto &e : staging_ev) if (e) cudaEventDestroy(e);
        for (auto &e : copied_ev) if (e) cudaEventDestroy(e);
        if (copy_stream) cudaStreamDestroy(copy_stream);
        for (auto &e : fill_ev) if (e) cudaEventDestroy(e);
        if (own_stream && stream) cudaStreamDestroy(stream);
    }
};

namespace
{

// DGE_TRACE=1 prints host-side wall-clock stage marks to stderr (replaces the reference's Tools::trace_time, Logs.cpp:63-71)
struct Tracer
{
    bool on;
    std::chrono::steady_clock::time_point t0;
    Tracer() : on(std::getenv("DGE_TRACE") != nullptr), t0(std::chrono::steady_clock::now()) {}
    cudaStream_t st = nullptr;
    void mark(const char *what)
    {
        if (!on) return;
        cudaStreamSynchronize(st); // tracing only (st may be the default stream): make the marks mean device time
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[dge] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

int fail(dge_handle *h, int code, const std::string &msg)
{
    if (h) h->err = msg; else g_create_error = msg;
    return code;
}

template <class F> int guarded(dge_handle *h, F &&f)
{
    try { return f(); }
    catch (CudaError &e) { return fail(h, DGE_ERR_CUDA, e.what()); }
    catch (InvalidInput &e) { return fail(h, DGE_ERR_INVALID, e.what()); }
    catch (CapacityError &e) { return fail(h, DGE_ERR_CAPACITY, e.what()); }
    catch (std::bad_alloc &) { return fail(h, DGE_ERR_INTERNAL, "out of host memory"); }
    catch (std::exception &e) { return fail(h, DGE_ERR_INTERNAL, e.what()); }
}

unsigned grid_for(size_t n, unsigned threads, unsigned max_blocks = 148 * 16)
{
    size_t g = div_up(n, size_t(threads));
    if (g < 1) g = 1;
    if (g > max_blocks) g = max_blocks;
    return unsigned(g);
}

template <class T> void d2h(std::vector<T> &dst, const void *src, size_t n, cudaStream_t st)
{
    dst.resize(n);
    if (n) DGE_CUDA(cudaMemcpyAsync(dst.data(), src, n * sizeof(T), cudaMemcpyDeviceToHost, st));
}

// device -> pinned host; valid after the stream is synchronised
template <class T> T *d2h_pinned(PinnedBuf &dst, const void *src, size_t n, cudaStream_t st)
{
    dst.reserve(std::max<size_t>(n, 1) * sizeof(T));
    if (n) DGE_CUDA(cudaMemcpyAsync(dst.p, src, n * sizeof(T), cudaMemcpyDeviceToHost, st));
    return dst.as<T>();
}

template <class T> T d2h_scalar(const void *src, cudaStream_t st)
{
    T v;
    DGE_CUDA(cudaMemcpyAsync(&v, src, sizeof(T), cudaMemcpyDeviceToHost, st));
    DGE_CUDA(cudaStreamSynchronize(st));
    return v;
}

void reset_fill_state(dge_handle *h);

// Barcode / UMI as the strings the reference sees: 2-bit unpacked, or -- entries made from DGE_FLAG_CB_N / DGE_FLAG_UMI_N records -- the
// caller's string with its N's (dge_set_n_strings).
std::string cb_string(const dge_handle *h, uint64_t cb)
{
    if (!(cb & CB_N_BIT)) return unpack_seq(cb, h->cfg.cb_len);
    const size_t idx = size_t(cb & (CB_N_BIT - 1));
    if (idx >= h->n_cb_list.size()) throw InvalidInput("barcode index beyond the escaped-barcode list (dge_set_n_strings / dge_set_cb_strings)");
    return h->n_cb_list[idx];
}

bool umi_is_n(const dge_handle *h, uint32_t umi) { return h->kl.ne && ((umi >> (h->kl.ub - 1)) & 1u); }
// internal UMI field -> the code of the query surface (DGE_UMI_N_BIT | index into the N-UMI list)
uint32_t umi_public_code(const dge_handle *h, uint32_t umi) { return umi_is_n(h, umi) ? (0x80000000u | (umi & ((1u << (h->kl.ub - 1)) - 1))) : umi; }

std::string umi_string(const dge_handle *h, uint32_t umi)
{
    if (!umi_is_n(h, umi)) return unpack_seq(umi, h->cfg.umi_len);
    const size_t idx = size_t(umi & ((1u << (h->kl.ub - 1)) - 1)), len = h->cfg.umi_len;
    if ((idx + 1) * len > h->n_umi_strings.size()) throw InvalidInput("UMI index beyond the N-UMI list (dge_set_n_strings)");
    return h->n_umi_strings.substr(idx * len, len);
}

// Radix sort of (key, value) pairs on the low `end_bit` key bits.  Used ONLY for per-cell / per-gene tables (~1e5 rows: cell-id
// order, compare_cells order, gene first-seen order) -- library code (CUB), not part of the per-read path.
void device_sort_pairs(dge_handle *h, const uint64_t *kin, uint64_t *kout, const uint32_t *vin, uint32_t *vout, size_t n, int end_bit)
{
    size_t bytes = 0;
    DGE_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, kin, kout, vin, vout, int(n), 0, end_bit, h->stream));
    h->cub_tmp.reserve(bytes + 16);
    DGE_CUDA(cub::DeviceRadixSort::SortPairs(h->cub_tmp.p, bytes, kin, kout, vin, vout, int(n), 0, end_bit, h->stream));
    h->launches += 4;
}

void ensure_device(dge_handle *h)
{
    DGE_CUDA(cudaSetDevice(h->cfg.device));
    if (h->device_ready) return;
    if (!h->stream)
    {
        DGE_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        h->own_stream = true;
    }
    for (auto &e : h->ev) DGE_CUDA(cudaEventCreate(&e));
    for (auto &e : h->staging_ev) DGE_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));

    // key layout: [slot:tb | gene:gb | umi:ub | mark:3]
    KeyLayout &kl = h->kl;
    kl.ul = int(2 * h->cfg.umi_len);
    kl.ne = h->cfg.allow_n ? 1 : 0;
    kl.ub = kl.ne ? std::max(kl.ul, 20) + 1 : kl.ul; // allow_n: [1 : index of an N-UMI string] needs a flag bit and room for the index
    kl.gb = std::max(1, ceil_log2_u64(h->cfg.n_genes));
    const int tb_max = 61 - kl.gb - kl.ub;
    int want = h->cfg.max_barcodes_hint ? ceil_log2_u64(2 * h->cfg.max_barcodes_hint) : 22;
    kl.tb = std::min(std::min(tb_max, 28), std::max(10, want));
    if (kl.tb < 10) throw std::runtime_error("key layout does not fit 64 bits: reduce n_genes or umi_len");
    kl.kb = kl.tb + kl.gb + kl.ub + 3;
    kl.cbb = int(2 * h->cfg.cb_len);
    h->table_cap = size_t(1) << kl.tb;
    // strategies whose tie rules depend on the UMI indexer's first-seen order (UMI ids): directional UMI merge, simple CB merge
    h->track_umi_first = h->cfg.umi_merge_type == DGE_UMI_MERGE_DIRECTIONAL || h->cfg.merge_type == DGE_MERGE_SIMPLE ||
                         h->cfg.merge_type == DGE_MERGE_POISSON_SIMPLE || h->cfg.allow_n; // N repair walks UMIs in StringIndexer order
    if (h->track_umi_first && kl.ub > 26) throw std::runtime_error("directional UMI merge / simple CB merge support UMIs of up to 13 bases");

    h->tab.reserve(h->table_cap * sizeof(CellSlot));
    h->device_ready = true;
    reset_fill_state(h);
}

void reset_fill_state(dge_handle *h)
{
    k_table_init<<<grid_for(h->table_cap, 256), 256, 0, h->stream>>>(h->tab.as<CellSlot>(), h->table_cap);
    h->gene_first.reserve(size_t(h->cfg.n_genes) * 4);
    k_fill_u32<<<grid_for(h->cfg.n_genes, 256), 256, 0, h->stream>>>(h->gene_first.as<uint32_t>(), h->cfg.n_genes, NONE32);
    if (h->track_umi_first)
    {
        const size_t n_umi = size_t(1) << h->kl.ub;
        h->umi_first.reserve(n_umi * 4);
        k_fill_u32<<<grid_for(n_umi, 256), 256, 0, h->stream>>>(h->umi_first.as<uint32_t>(), n_umi, NONE32);
        ++h->launches;
    }
    h->ctr.reserve(sizeof(FillCounters));
    DGE_CUDA(cudaMemsetAsync(h->ctr.p, 0, sizeof(FillCounters), h->stream));
    h->hist12.reserve(4096 * 4); h->hist_fold.reserve(4096 * 4); h->region_tiles.reserve(64);
    h->regions.reserve(size_t(4096) * 148 * sizeof(KeyRegion)); // every batch adds one region per block (<= 4096 batches)
    DGE_CUDA(cudaMemsetAsync(h->hist12.p, 0, 4096 * 4, h->stream));
    h->n_regions = 0; h->pipe_fill = false;
    if (h->chr_cap)
    {
        DGE_CUDA(cudaMemsetAsync(h->chr_tab.p, 0, h->chr_cap * sizeof(ChrEntry), h->stream));
        DGE_CUDA(cudaMemsetAsync(h->chr_ctr.p, 0, sizeof(ChrCounters), h->stream));
    }
    h->chr_used = false;
    h->overflow_flag.reserve(sizeof(int));
    DGE_CUDA(cudaMemsetAsync(h->overflow_flag.p, 0, sizeof(int), h->stream));
    h->scan_scratch.reserve(((size_t(1) << 20) + 64) * 4); // enough for any scan of < 2^32 elements
    h->chunk_count_pool.reserve(4096 * sizeof(unsigned long long));
    DGE_CUDA(cudaMemsetAsync(h->chunk_count_pool.p, 0, 4096 * sizeof(unsigned long long), h->stream));
    DGE_LAUNCH_CHECK();
    h->launches += 2;
}

// Experiment kept behind DGE_L2_WINDOW=1 (default off): pin the barcode table's lines with a persisting access-policy window on the
// launching stream for the duration of the fill kernel.  Measured on B200 at C2: the L2 set-aside it needs slows every later kernel
// of the step (L2 partition pass 4.2 -> 8.1 ms) and the fill kernel itself does not gain (4.1 -> 5.0 ms): rejected.
void set_table_l2_window(dge_handle *h, bool on)
{
    static const bool enabled = std::getenv("DGE_L2_WINDOW") && atoi(std::getenv("DGE_L2_WINDOW")) == 1;
    if (!enabled) return;
    static size_t persist_max[64] = {}, window_max[64] = {};
    static bool queried[64] = {};
    const int dev = h->cfg.device & 63;
    if (!queried[dev])
    {
        cudaDeviceProp prop;
        DGE_CUDA(cudaGetDeviceProperties(&prop, h->cfg.device));
        persist_max[dev] = size_t(prop.persistingL2CacheMaxSize);
        window_max[dev] = size_t(prop.accessPolicyMaxWindowSize);
        if (persist_max[dev]) DGE_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, persist_max[dev]));
        queried[dev] = true;
    }
    if (!persist_max[dev] || !window_max[dev]) return;
    cudaStreamAttrValue attr{};
    const size_t bytes = h->table_cap * sizeof(CellSlot);
    attr.accessPolicyWindow.base_ptr = on ? h->tab.p : nullptr;
    attr.accessPolicyWindow.num_bytes = on ? std::min(bytes, window_max[dev]) : 0;
    attr.accessPolicyWindow.hitRatio = on ? float(std::min(1.0, double(persist_max[dev]) / double(std::max<size_t>(bytes, 1)))) : 0.f;
    attr.accessPolicyWindow.hitProp = on ? cudaAccessPropertyPersisting : cudaAccessPropertyNormal;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    DGE_CUDA(cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &attr));
}

// One batch already resident on the device: barcode-table insert + key packing into a fresh chunk.
void fill_from_device(dge_handle *h, const dge_record16 *recs, size_t n, const unsigned long long *soa_keys = nullptr,
                      const uint32_t *soa_genes = nullptr, uint32_t soa_first = 0, const dge_record16 *const *segs = nullptr,
                      const uint64_t *seg_counts = nullptr, uint32_t n_segs = 0)
{
    // segs != nullptr: ONE launch over several record arrays (the exchange of a sharded run: this rank's own segment plus the segments
    // it pulls out of the peers' HBM); recs / n are then ignored
    const bool soa = soa_keys != nullptr;
    if (segs)
    {
        n = 0;
        for (uint32_t k = 0; k < n_segs; ++k) n += seg_counts[k];
    }
    if (n == 0) return;
    if (h->n_chunk_counters >= 4096) throw std::runtime_error("too many batches (max 4096); use larger batches");
    std::unique_ptr<KeyChunk> chunk;
    if (!h->chunk_pool.empty()) { chunk = std::move(h->chunk_pool.back()); h->chunk_pool.pop_back(); }
    else chunk.reset(new KeyChunk());
    chunk->capacity = n;
    chunk->d_count = h->chunk_count_pool.as<unsigned long long>() + h->n_chunk_counters++;
    // the kernel appends through FillCounters::n_keys; give every chunk its own cursor by pointing a private counter struct at it
    // (we keep one FillCounters per handle and move the cursor: n_keys is reset per chunk and accumulated on the host later)
    DGE_CUDA(cudaMemsetAsync(&h->ctr.as<FillCounters>()->n_keys, 0, sizeof(unsigned long long), h->stream));
    if (h->n_fill_ev + 2 > h->fill_ev.size())
    {
        h->fill_ev.resize(h->n_fill_ev + 2, nullptr);
        DGE_CUDA(cudaEventCreate(&h->fill_ev[h->n_fill_ev]));
        DGE_CUDA(cudaEventCreate(&h->fill_ev[h->n_fill_ev + 1]));
    }
    DGE_CUDA(cudaEventRecord(h->fill_ev[h->n_fill_ev], h->stream));
    // variant 0: 256 threads x 8 records, gene first-seen words gathered from global memory
    // variant 1: 1024 threads x 4 records, one block per SM, gene first-seen words served from a shared-memory copy (needs n_genes * 4 B of it)
    // variant 2 (default): k_fill_pipe -- bulk-async record tiles through a shared-memory ring, per-block output regions, fused L1 histogram
    static const int fill_variant = std::getenv("DGE_FILL_VARIANT") ? atoi(std::getenv("DGE_FILL_VARIANT")) : 2;
    // shape of the pipelined kernel: consumer warps x records per thread x ring stages (DGE_FILL_SHAPE picks among the instantiations)
    static const int fill_shape = std::getenv("DGE_FILL_SHAPE") ? atoi(std::getenv("DGE_FILL_SHAPE")) : 0;
    auto pipe_launch = [&](auto cw_c, auto it_c, auto stg_c) -> bool {
        constexpr int CW = decltype(cw_c)::value, IT = decltype(it_c)::value, STG = decltype(stg_c)::value, TILE = CW * 32 * IT;
        const size_t smem = size_t(STG) * TILE * 16 + size_t((h->cfg.n_genes + 7u) & ~7u) * 2 + 4096 * 4;
        const bool first_batch = h->chunks.empty();
        if (smem > 224 * 1024 || !(first_batch || h->pipe_fill)) return false;
        const void *src0 = soa ? static_cast<const void *>(soa_keys) : segs ? nullptr : static_cast<const void *>(recs);
        if ((reinterpret_cast<uintptr_t>(src0) & 15u) || (soa && (reinterpret_cast<uintptr_t>(soa_genes) & 15u)))
            throw InvalidInput("device record arrays must be 16-byte aligned");
        FillSegs fs;
        memset(&fs, 0, sizeof(fs));
        if (segs)
        {
            for (uint32_t k = 0; k < n_segs; ++k)
            {
                if (seg_counts[k] == 0) continue;
                if (fs.n_segs == FILL_MAX_SEGS) throw InvalidInput("too many segments in one batch (max 64)");
                if (reinterpret_cast<uintptr_t>(segs[k]) & 15u) throw InvalidInput("device record arrays must be 16-byte aligned");
                fs.base[fs.n_segs] = reinterpret_cast<const Rec16 *>(segs[k]);
                fs.count[fs.n_segs++] = seg_counts[k];
            }
        }
        else { fs.base[0] = reinterpret_cast<const Rec16 *>(recs); fs.count[0] = n; fs.n_segs = 1; }
        size_t n_tiles = 0;
        uint32_t min_tiles = ~0u;
        for (uint32_t k = 0; k < fs.n_segs; ++k)
        {
            const size_t t = div_up(size_t(fs.count[k]), size_t(TILE));
            n_tiles += t;
            min_tiles = std::min<uint32_t>(min_tiles, uint32_t(t));
        }
        if (n_tiles >= (size_t(1) << 31)) throw CapacityError("batch too large for one fill launch");
        fs.rr_rounds = min_tiles;
        fs.n_tiles = uint32_t(n_tiles);
        for (uint32_t k = 0; k < fs.n_segs; ++k)
            fs.rem_start[k + 1] = fs.rem_start[k] + uint32_t(div_up(size_t(fs.count[k]), size_t(TILE))) - min_tiles;
        for (uint32_t k = fs.n_segs; k < FILL_MAX_SEGS; ++k) fs.rem_start[k + 1] = ~0u; // the search of fill_seg_tile stops here at the latest
        const unsigned grid = unsigned(std::min<size_t>(n_tiles, 148));
        const size_t region_cap = div_up(n_tiles, size_t(grid)) * TILE;
        chunk->keys.reserve(size_t(grid) * region_cap * 8);
        KeyRegion *regs = h->regions.as<KeyRegion>() + h->n_regions;
        uint32_t *umi_first_p = h->track_umi_first ? h->umi_first.as<uint32_t>() : nullptr;
        set_table_l2_window(h, true);
        static bool done[64][2] = {};
        if (soa)
        {
            auto kern = k_fill_pipe<CW, IT, STG, true>;
            if (!done[h->cfg.device & 63][1]) { DGE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024)); done[h->cfg.device & 63][1] = true; }
            kern<<<grid, (CW + 1) * 32, smem, h->stream>>>(fs, h->tab.as<CellSlot>(), h->kl, h->cfg.n_genes, h->gene_first.as<uint32_t>(), chunk->keys.as<uint64_t>(),
                                                          region_cap, regs, h->ctr.as<FillCounters>(), umi_first_p, h->hist12.as<uint32_t>(), soa_keys, soa_genes, soa_first);
        }
        else
        {
            auto kern = k_fill_pipe<CW, IT, STG, false>;
            if (!done[h->cfg.device & 63][0]) { DGE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024)); done[h->cfg.device & 63][0] = true; }
            kern<<<grid, (CW + 1) * 32, smem, h->stream>>>(fs, h->tab.as<CellSlot>(), h->kl, h->cfg.n_genes, h->gene_first.as<uint32_t>(), chunk->keys.as<uint64_t>(),
                                                          region_cap, regs, h->ctr.as<FillCounters>(), umi_first_p, h->hist12.as<uint32_t>(), nullptr, nullptr, 0u);
        }
        DGE_LAUNCH_CHECK();
        set_table_l2_window(h, false);
        h->n_regions += grid;
        h->pipe_fill = true;
        DGE_CUDA(cudaEventRecord(h->fill_ev[h->n_fill_ev + 1], h->stream));
        h->n_fill_ev += 2;
        DGE_CUDA(cudaMemcpyAsync(chunk->d_count, &h->ctr.as<FillCounters>()->n_keys, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, h->stream));
        ++h->launches;
        h->chunks.push_back(std::move(chunk));
        h->n_reads += n;
        return true;
    };
    if (fill_variant == 2)
    {
        using std::integral_constant;
        bool ok;
        if (fill_shape == 1) ok = pipe_launch(integral_constant<int, 31>{}, integral_constant<int, 4>{}, integral_constant<int, 2>{});
        else if (fill_shape == 2) ok = pipe_launch(integral_constant<int, 16>{}, integral_constant<int, 8>{}, integral_constant<int, 2>{});
        else if (fill_shape == 3) ok = pipe_launch(integral_constant<int, 24>{}, integral_constant<int, 4>{}, integral_constant<int, 3>{});
        else if (fill_shape == 4) ok = pipe_launch(integral_constant<int, 12>{}, integral_constant<int, 8>{}, integral_constant<int, 3>{});
        else if (fill_shape == 5) ok = pipe_launch(integral_constant<int, 24>{}, integral_constant<int, 4>{}, integral_constant<int, 2>{});
        else ok = pipe_launch(integral_constant<int, 30>{}, integral_constant<int, 4>{}, integral_constant<int, 2>{}); // measured best at C2
        if (ok) return;
    }
    if (segs)
    {   // the older kernels take one array: one batch per segment (give the pooled objects back first)
        h->chunk_pool.push_back(std::move(chunk));
        --h->n_chunk_counters;
        for (uint32_t k = 0; k < n_segs; ++k) fill_from_device(h, segs[k], size_t(seg_counts[k]));
        return;
    }
    chunk->keys.reserve(n * 8);
    const size_t gene_smem = size_t(h->cfg.n_genes) * 4;
    uint32_t *umi_first = h->track_umi_first ? h->umi_first.as<uint32_t>() : nullptr;
    const Rec16 *r16 = reinterpret_cast<const Rec16 *>(recs);
#define DGE_FILL_ARGS r16, n, h->tab.as<CellSlot>(), h->kl, h->cfg.n_genes, h->gene_first.as<uint32_t>(), chunk->keys.as<uint64_t>(), \
                      h->ctr.as<FillCounters>(), umi_first, soa_keys, soa_genes, soa_first
    auto set_smem = [&](const void *kern) {
        DGE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    };
    if (fill_variant == 1 && gene_smem <= 160 * 1024)
    {
        unsigned grid = unsigned(std::min<size_t>(div_up(n, size_t(1024 * 4)), 148));
        if (soa)
        {
            auto kern = k_fill_compact<1024, 4, 1, true, true>;
            static bool done[64] = {};
            if (!done[h->cfg.device & 63]) { set_smem(reinterpret_cast<const void *>(kern)); done[h->cfg.device & 63] = true; }
            kern<<<grid, 1024, gene_smem, h->stream>>>(DGE_FILL_ARGS);
        }
        else
        {
            auto kern = k_fill_compact<1024, 4, 1, true, false>;
            static bool done[64] = {};
            if (!done[h->cfg.device & 63]) { set_smem(reinterpret_cast<const void *>(kern)); done[h->cfg.device & 63] = true; }
            kern<<<grid, 1024, gene_smem, h->stream>>>(DGE_FILL_ARGS);
        }
    }
    else
    {
        unsigned grid = unsigned(std::min<size_t>(div_up(n, size_t(256 * 8)), 148 * 8));
        if (soa) k_fill_compact<256, 8, 2, false, true><<<grid, 256, 0, h->stream>>>(DGE_FILL_ARGS);
        else k_fill_compact<256, 8, 2, false, false><<<grid, 256, 0, h->stream>>>(DGE_FILL_ARGS);
    }
#undef DGE_FILL_ARGS
    DGE_LAUNCH_CHECK();
    DGE_CUDA(cudaEventRecord(h->fill_ev[h->n_fill_ev + 1], h->stream));
    h->n_fill_ev += 2;
    DGE_CUDA(cudaMemcpyAsync(chunk->d_count, &h->ctr.as<FillCounters>()->n_keys, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, h->stream));
    ++h->launches;
    h->chunks.push_back(std::move(chunk));
    h->n_reads += n;
}

// CellsDataContainer::compare_cells (CellsDataContainer.cpp:329-344) orders by (requested genes, requested umis, TOTAL_UMIS stat,
// barcode); barcode strings of equal length over ACGT compare like their 2-bit packings.  update_filtered sorts by exactly that.

// Stable LSD radix sort of (key, idx) pairs by 8-bit digits (host; n ~ 1e5 cells, replaces std::sort with a comparator).
// Digits on which all keys agree are skipped; 256 write streams keep the scatter cache friendly.
struct HostPairSorter
{
    std::vector<uint64_t> key, key2;
    std::vector<uint32_t> idx, idx2;
    void sort(size_t n, int key_bits)
    {
        key2.resize(n); idx2.resize(n);
        if (n < 2) return;
        uint64_t all_or = 0, all_and = ~0ull;
        for (size_t i = 0; i < n; ++i) { all_or |= key[i]; all_and &= key[i]; }
        const uint64_t varying = all_or & ~all_and;
        for (int sh = 0; sh < key_bits; sh += 8)
        {
            if (((varying >> sh) & 0xFFu) == 0) continue;
            size_t cnt[257] = {0};
            for (size_t i = 0; i < n; ++i) ++cnt[((key[i] >> sh) & 0xFFu) + 1];
            for (int d = 0; d < 256; ++d) cnt[d + 1] += cnt[d];
            for (size_t i = 0; i < n; ++i)
            {
                const size_t o = cnt[(key[i] >> sh) & 0xFFu]++;
                key2[o] = key[i]; idx2[o] = idx[i];
            }
            key.swap(key2); idx.swap(idx2);
        }
    }
};

// update_filtered_gene_counts (CellsDataContainer.cpp:250-276): real cells with enough requested genes, ascending by
// compare_cells = (requested genes, requested umis, TOTAL_UMIS stat, barcode) -- a total order, so any sort gives the same list.
void update_filtered(dge_handle *h, uint32_t threshold, int cell_threshold)
{
    std::vector<uint32_t> &f = h->filtered;
    f.clear();
    const std::vector<HostCell> &R = h->real;
    static thread_local HostPairSorter ps;
    ps.key.clear(); ps.idx.clear();
    bool fits = true;
    for (uint32_t i = 0; i < R.size(); ++i)
    {
        const HostCell &c = R[i];
        if (!(c.real && uint32_t(c.req_genes) >= threshold)) continue;
        // one 64-bit composite key (genes:16 | umis:24 | stat:24); the barcode only breaks exact ties
        ps.key.push_back((uint64_t(uint32_t(c.req_genes)) << 48) | (uint64_t(uint32_t(c.req_umis)) << 24) | uint64_t(uint32_t(c.umis_stat)));
        ps.idx.push_back(i);
        fits &= uint32_t(c.req_genes) < (1u << 16) && uint32_t(c.req_umis) < (1u << 24) && uint32_t(c.umis_stat) < (1u << 24);
    }
    if (fits)
    {
        ps.sort(ps.key.size(), 64);
        f.assign(ps.idx.begin(), ps.idx.end());
        const std::vector<uint64_t> &key = ps.key;
        for (size_t a = 0; a < f.size();)
        {
            size_t b = a + 1;
            while (b < f.size() && key[b] == key[a]) ++b;
            if (b - a > 1)
                std::sort(f.begin() + long(a), f.begin() + long(b), [&](uint32_t x, uint32_t y) {
                    // barcode strings; N-free strings of one length compare like their 2-bit packings
                    if (!((R[x].cb | R[y].cb) & CB_N_BIT)) return R[x].cb < R[y].cb;
                    return cb_string(h, R[x].cb) < cb_string(h, R[y].cb);
                });
            a = b;
        }
    }
    else
    {   // counters beyond the packed widths: successive stable passes, least significant criterion first
        f.assign(ps.idx.begin(), ps.idx.end());
        auto pass = [&](int bits, auto keyfn) {
            ps.key.resize(f.size()); ps.idx = f;
            for (size_t i = 0; i < f.size(); ++i) ps.key[i] = keyfn(R[f[i]]);
            ps.sort(f.size(), bits);
            f.assign(ps.idx.begin(), ps.idx.end());
        };
        pass(48, [](const HostCell &c) { return uint64_t(c.cb); });
        pass(32, [](const HostCell &c) { return uint64_t(uint32_t(c.umis_stat)); });
        pass(32, [](const HostCell &c) { return uint64_t(uint32_t(c.req_umis)); });
        pass(32, [](const HostCell &c) { return uint64_t(uint32_t(c.req_genes)); });
    }
    if (cell_threshold > 0 && size_t(cell_threshold) < f.size()) f.erase(f.begin(), f.end() - cell_threshold);
}

// (Re)build CG / PC tables from the current sorted U list.
void build_segments(dge_handle *h)
{
    cudaStream_t st = h->stream;
    const uint32_t n_u = h->n_u;
    const int ub = h->kl.ub, gub = h->kl.gb + h->kl.ub;
    const size_t nt = div_up(size_t(n_u), size_t(SEG_TILE));
    h->tile_a.reserve((nt + 2) * 4); h->tile_b.reserve((nt + 2) * 4);
    uint32_t *ta = h->tile_a.as<uint32_t>(), *tb = h->tile_b.as<uint32_t>();
    if (nt == 0)
    {
        h->n_cg = h->n_pc = 0;
        h->cg_start.reserve(8); h->pc_u_start.reserve(8); h->pc_cg_start.reserve(8);
        DGE_CUDA(cudaMemsetAsync(h->cg_start.p, 0, 8, st));
        DGE_CUDA(cudaMemsetAsync(h->pc_u_start.p, 0, 8, st));
        DGE_CUDA(cudaMemsetAsync(h->pc_cg_start.p, 0, 8, st));
        return;
    }
    k_seg_count<<<unsigned(nt), SEG_THREADS, 0, st>>>(h->ukey.as<uint64_t>(), n_u, ub, gub, ta, tb);
    ++h->launches;
    // two scans share the scratch: run sequentially on the stream, read totals in between
    const uint32_t *tot_cg = device_exclusive_scan(ta, ta, nt, h->scan_scratch.as<uint32_t>(), st, &h->launches);
    h->n_cg = d2h_scalar<uint32_t>(tot_cg, st);
    const uint32_t *tot_pc = device_exclusive_scan(tb, tb, nt, h->scan_scratch.as<uint32_t>(), st, &h->launches);
    h->n_pc = d2h_scalar<uint32_t>(tot_pc, st);

    const size_t ncg = h->n_cg, npc = h->n_pc;
    h->cg_gene.reserve((ncg + 1) * 4); h->cg_start.reserve((ncg + 1) * 4); h->cg_pc.reserve((ncg + 1) * 4);
    h->cg_req.reserve((ncg + 1) * 4); h->cg_reads.reserve((ncg + 1) * 4);
    if (h->cfg.reads_output) h->cg_req_reads.reserve((ncg + 1) * 4);
    h->pc_slot.reserve((npc + 2) * 4); h->pc_u_start.reserve((npc + 2) * 4); h->pc_cg_start.reserve((npc + 2) * 4);
    h->pc_reads.reserve((npc + 2) * 4); h->pc_req_genes.reserve((npc + 2) * 4); h->pc_req_umis.reserve((npc + 2) * 4);
    DGE_CUDA(cudaMemsetAsync(h->cg_req.p, 0, (ncg + 1) * 4, st));
    DGE_CUDA(cudaMemsetAsync(h->cg_reads.p, 0, (ncg + 1) * 4, st));
    if (h->cfg.reads_output) DGE_CUDA(cudaMemsetAsync(h->cg_req_reads.p, 0, (ncg + 1) * 4, st));
    DGE_CUDA(cudaMemsetAsync(h->pc_reads.p, 0, (npc + 2) * 4, st));
    DGE_CUDA(cudaMemsetAsync(h->pc_req_genes.p, 0, (npc + 2) * 4, st));
    DGE_CUDA(cudaMemsetAsync(h->pc_req_umis.p, 0, (npc + 2) * 4, st));
    k_seg_write<<<unsigned(nt), SEG_THREADS, 0, st>>>(h->ukey.as<uint64_t>(), h->uval.as<uint32_t>(), n_u, ub, gub, h->cfg.query_mark_mask, (1u << h->kl.gb) - 1, ta, tb,
                                                       h->cg_gene.as<uint32_t>(), h->cg_start.as<uint32_t>(), h->cg_pc.as<uint32_t>(),
                                                       h->cg_req.as<uint32_t>(), h->cg_reads.as<uint32_t>(),
                                                       h->cfg.reads_output ? h->cg_req_reads.as<uint32_t>() : nullptr,
                                                       h->pc_slot.as<uint32_t>(), h->pc_u_start.as<uint32_t>(), h->pc_cg_start.as<uint32_t>(),
                                                       h->pc_reads.as<uint32_t>(), h->pc_req_umis.as<uint32_t>());
    k_seg_sentinels<<<1, 1, 0, st>>>(n_u, h->n_cg, h->n_pc, h->cg_start.as<uint32_t>(), h->pc_u_start.as<uint32_t>(), h->pc_cg_start.as<uint32_t>());
    k_pc_req_genes<<<grid_for(div_up(ncg, size_t(8)), 256), 256, 0, st>>>(h->cg_req.as<uint32_t>(), h->cg_pc.as<uint32_t>(), h->n_cg,
                                                                           h->pc_req_genes.as<uint32_t>());
    DGE_LAUNCH_CHECK();
    h->launches += 3;
}

void build_slot_pc(dge_handle *h)
{
    h->slot_pc.reserve(h->table_cap * 4);
    k_fill_u32<<<grid_for(h->table_cap, 256), 256, 0, h->stream>>>(h->slot_pc.as<uint32_t>(), h->table_cap, NONE32);
    if (h->n_pc)
        k_build_slot_pc<<<grid_for(h->n_pc, 256), 256, 0, h->stream>>>(h->pc_slot.as<uint32_t>(), h->n_pc, h->slot_pc.as<uint32_t>());
    DGE_LAUNCH_CHECK();
    h->launches += 2;
    h->slot_pc_built = true;
}

void gather_rows(dge_handle *h, const std::vector<uint32_t> &pcs, std::vector<CellRow> &rows)
{
    rows.clear();
    if (pcs.empty()) return;
    h->misc.reserve(pcs.size() * 4);
    h->rows_dev.reserve(pcs.size() * sizeof(CellRow));
    DGE_CUDA(cudaMemcpyAsync(h->misc.p, pcs.data(), pcs.size() * 4, cudaMemcpyHostToDevice, h->stream));
    k_gather_rows_list<<<grid_for(pcs.size(), 256), 256, 0, h->stream>>>(h->misc.as<uint32_t>(), uint32_t(pcs.size()), h->tab.as<CellSlot>(),
                                                                        h->pc_slot.as<uint32_t>(), h->pc_cg_start.as<uint32_t>(),
                                                                        h->pc_u_start.as<uint32_t>(), h->pc_reads.as<uint32_t>(),
                                                                        h->pc_req_genes.as<uint32_t>(), h->pc_req_umis.as<uint32_t>(),
                                                                        h->rows_dev.as<CellRow>());
    DGE_LAUNCH_CHECK();
    ++h->launches;
    const CellRow *p = d2h_pinned<CellRow>(h->pin_rows, h->rows_dev.p, pcs.size(), h->stream);
    DGE_CUDA(cudaStreamSynchronize(h->stream));
    rows.assign(p, p + pcs.size());
}

// Host mirror of the real cells (cell-id order), the gene indexer order and the set_initialized filtered list, from tables that
// were sorted on the device (rows_sorted) or gathered unsorted.
void build_host_mirror(dge_handle *h, const CellRow *rows_p, size_t n_rows, bool rows_sorted, const uint32_t *dev_filtered,
                       const uint64_t *dev_gene_keys, const uint32_t *dev_gene_ids, int dev_filter_overflow, Tracer &tr)
{
    cudaStream_t st = h->stream;
    static thread_local HostPairSorter order_sorter;
    if (!rows_sorted)
    {
        order_sorter.key.resize(n_rows); order_sorter.idx.resize(n_rows);
        for (size_t i = 0; i < n_rows; ++i) { order_sorter.key[i] = rows_p[i].first_idx; order_sorter.idx[i] = uint32_t(i); }
        order_sorter.sort(n_rows, 32);
    }
    const std::vector<uint32_t> &order = order_sorter.idx;
    tr.mark("init:  order sort");
    h->real.clear();
    h->real.reserve(n_rows);
    for (size_t k = 0; k < n_rows; ++k)
    {
        const CellRow &r = rows_p[rows_sorted ? k : size_t(order[k])];
        HostCell c;
        c.cb = r.cb; c.slot = r.slot; c.pc = r.pc; c.first_idx = r.first_idx; c.n_intergenic = r.n_intergenic;
        c.n_genes = int32_t(r.n_genes); c.umis_stat = int32_t(r.n_umis); c.n_umis_distinct = int32_t(r.n_umis);
        c.reads_stat = int32_t(r.n_reads); c.req_genes = int32_t(r.req_genes); c.req_umis = int32_t(r.req_umis);
        c.target = int32_t(h->real.size());
        h->real.push_back(c);
    }
    tr.mark("init: real cells -> host");
    // gene first-seen order (StringIndexer::add, StringIndexer.cpp:10-18)
    if (dev_gene_ids)
    {   // sorted on the device: genes never seen carry NONE32 and sort to the end
        size_t seen = 0;
        while (seen < h->cfg.n_genes && dev_gene_keys[seen] != uint64_t(NONE32)) ++seen;
        h->gene_order.assign(dev_gene_ids, dev_gene_ids + seen);
    }
    else
    {
        const uint32_t *gf = d2h_pinned<uint32_t>(h->pin_misc, h->gene_first.p, h->cfg.n_genes, st);
        DGE_CUDA(cudaStreamSynchronize(st));
        static thread_local HostPairSorter gs;
        gs.key.clear(); gs.idx.clear();
        for (uint32_t g = 0; g < h->cfg.n_genes; ++g)
            if (gf[g] != NONE32) { gs.key.push_back(gf[g]); gs.idx.push_back(g); }
        gs.sort(gs.key.size(), 32);
        h->gene_order.assign(gs.idx.begin(), gs.idx.end());
    }
    tr.mark("init:  gene order");
    // set_initialized: update_cell_sizes(query, 0, -1)  (CellsDataContainer.cpp:168) -- every real cell, ascending compare_cells
    bool any_n_cb = false;
    if (h->kl.ne) for (auto const &c : h->real) any_n_cb |= (c.cb & CB_N_BIT) != 0;
    // (the device sort orders barcodes by their packing: not the string order once a barcode contains N)
    if (rows_sorted && !dev_filter_overflow && !any_n_cb) h->filtered.assign(dev_filtered, dev_filtered + n_rows); // exact total order, ties included
    else update_filtered(h, 0, -1);
    tr.mark("init: gene order + filtered");
    h->host_stage = 1;
}

// Lazy part of set_initialized: copies the device-sorted tables to the host and builds the mirror.  After a device-flow merge the
// merge outcome (CellState, refreshed rows, final filtered list) is applied on top (stage 2).
void materialize_host(dge_handle *h)
{
    const int want = h->state >= 2 ? 2 : 1;
    if (h->host_stage >= want || h->state == 0) return;
    DGE_CUDA(cudaSetDevice(h->cfg.device));
    cudaStream_t st = h->stream;
    Tracer tr; tr.st = st;
    const size_t n = h->n_real_rows;
    if (h->host_stage == 0)
    {
        if (!h->lazy_rows) throw std::runtime_error("host mirror missing");
        const CellRow *rows_p = d2h_pinned<CellRow>(h->pin_rows, h->rows_dev2.p, n, st);
        const uint32_t *dev_filtered = d2h_pinned<uint32_t>(h->pin_filtered, h->sort_v[0].p, n, st);
        const uint64_t *gk = nullptr; const uint32_t *gi = nullptr;
        if (h->cfg.n_genes > 0)
        {
            gk = d2h_pinned<uint64_t>(h->pin_fkeys, h->gsort_k[1].p, h->cfg.n_genes, st);
            gi = d2h_pinned<uint32_t>(h->pin_gene_ids, h->gsort_v[1].p, h->cfg.n_genes, st);
        }
        DGE_CUDA(cudaStreamSynchronize(st));
        build_host_mirror(h, rows_p, n, true, dev_filtered, gk, gi, h->init_filter_overflow, tr);
    }
    if (want == 2 && h->host_stage < 2)
    {
        if (!h->dev_merged) throw std::runtime_error("merge results missing on the host");
        // refreshed sizes of the merge targets + counters, flags and targets of every cell + the final filtered list
        const CellRow *rows_p = d2h_pinned<CellRow>(h->pin_rows, h->rows_dev2.p, n, st);
        const CellState *cs = d2h_pinned<CellState>(h->pin_state, h->cell_state.p, n, st);
        const uint32_t *fl = d2h_pinned<uint32_t>(h->pin_filtered, h->fsort_v[h->dev_filtered_buf].p, h->n_filtered_dev, st);
        DGE_CUDA(cudaStreamSynchronize(st));
        for (size_t i = 0; i < n; ++i)
        {
            HostCell &c = h->real[i];
            const CellRow &r = rows_p[i];
            c.n_genes = int32_t(r.n_genes); c.req_genes = int32_t(r.req_genes); c.req_umis = int32_t(r.req_umis); c.n_umis_distinct = int32_t(r.n_umis);
            c.umis_stat = cs[i].umis_stat; c.reads_stat = cs[i].reads_stat; c.n_intergenic = cs[i].n_intergenic;
            c.real = (cs[i].flags & 1u) != 0; c.merged = (cs[i].flags & 2u) != 0; c.excluded = (cs[i].flags & 4u) != 0;
            c.target = cs[i].target < 0 ? int32_t(i) : cs[i].target;
        }
        if (h->dist_done && n)
        {   // children merged into a cell of another rank: the target's barcode (dge_get_merge_pairs)
            std::vector<unsigned long long> mcb;
            d2h(mcb, h->x_merged_cb.p, n, st);
            DGE_CUDA(cudaStreamSynchronize(st));
            for (size_t i = 0; i < n; ++i) h->real[i].merged_to_cb = mcb[i];
        }
        // merge_cells calls in application order: the walk of MergeStrategyBase::merge_inited over filtered_cells() as set_initialized left them
        // (MergeStrategyBase.cpp:29-51); the device flow has no chains, so every merged cell went straight to its final target
        h->merge_events.clear();
        for (uint32_t base : h->filtered)
            if (h->real[base].merged && uint32_t(h->real[base].target) != base) h->merge_events.emplace_back(base, uint32_t(h->real[base].target));
        const uint32_t nf = h->n_filtered_dev;
        const uint32_t skip = (h->cfg.max_cells > 0 && uint32_t(h->cfg.max_cells) < nf) ? nf - uint32_t(h->cfg.max_cells) : 0u;
        h->filtered.assign(fl + skip, fl + nf);
        h->host_stage = 2;
    }
}

void do_set_initialized(dge_handle *h)
{
    ensure_device(h);
    cudaStream_t st = h->stream;
    Tracer tr;
    tr.st = st;
    DGE_CUDA(cudaEventRecord(h->ev[0], st));

    // ---- collect compact keys of all batches into one array (chunks were produced at add time)
    std::vector<unsigned long long> counts(h->n_chunk_counters);
    if (!counts.empty())
        DGE_CUDA(cudaMemcpyAsync(counts.data(), h->chunk_count_pool.p, counts.size() * 8, cudaMemcpyDeviceToHost, st));
    DGE_CUDA(cudaMemcpyAsync(&h->counters, h->ctr.p, sizeof(FillCounters), cudaMemcpyDeviceToHost, st));
    DGE_CUDA(cudaStreamSynchronize(st));
    if (h->counters.table_overflow)
        throw CapacityError("barcode table overflow: more distinct barcodes than the table holds; set max_barcodes_hint (or shard across GPUs)");
    if (h->counters.bad_gene) throw InvalidInput("record with gene id >= n_genes");
    if (h->counters.bad_record)
        throw InvalidInput("malformed record: barcode / UMI bits beyond cb_len / umi_len, reserved bits of the gene word set, or read_idx == 0xFFFFFFFF");
    size_t n_keys = 0;
    for (size_t c = 0; c < h->chunks.size(); ++c) { h->chunks[c]->count = size_t(counts[c]); n_keys += h->chunks[c]->count; }
    if (n_keys >= 0xFFFFFFF0ull) throw CapacityError("more than 2^32 reads with genes on one device (32-bit positions): shard across GPUs");
    h->n_keys = n_keys;

    const uint64_t *keys_in = nullptr;
    if (h->pipe_fill) { /* the keys stay in the fill kernel's per-block regions: consumed as a list by the L1 partition pass */ }
    else if (h->chunks.size() == 1) keys_in = h->chunks[0]->keys.as<uint64_t>();
    else if (h->chunks.size() > 1)
    {
        h->keys_all.reserve(n_keys * 8);
        size_t off = 0;
        for (auto &c : h->chunks)
        {
            if (c->count)
                DGE_CUDA(cudaMemcpyAsync(h->keys_all.as<uint64_t>() + off, c->keys.p, c->count * 8, cudaMemcpyDeviceToDevice, st));
            off += c->count;
        }
        keys_in = h->keys_all.as<uint64_t>();
    }

    // ---- group: distinct (cell, gene, UMI), sorted
    h->ukey.reserve(std::max<size_t>(n_keys, 1) * 8);
    h->uval.reserve(std::max<size_t>(n_keys, 1) * 4);
    h->n_u = 0;
    if (n_keys)
    {
        const int l1_bits = std::min(choose_l1_bits(n_keys), h->kl.kb - 3);
        const uint32_t *l1_hist = nullptr;
        uint64_t *keys_tmp = const_cast<uint64_t *>(keys_in); // the compact-key array doubles as the L2 scatter target / in-place dedup buffer
        if (h->pipe_fill)
        {
            h->key_tiles.reserve((n_keys / (1024 * 8) + h->n_regions + 2) * sizeof(KeyTile));
            k_region_tiles<<<1, 1024, 0, st>>>(h->regions.as<KeyRegion>(), h->n_regions, 1024 * 8, h->region_tiles.as<uint32_t>(), h->key_tiles.as<KeyTile>());
            k_hist_fold<<<4, 256, 0, st>>>(h->hist12.as<uint32_t>(), l1_bits, h->hist_fold.as<uint32_t>());
            h->launches += 2;
            l1_hist = h->hist_fold.as<uint32_t>();
            h->keys_all.reserve(n_keys * 8);
            keys_tmp = h->keys_all.as<uint64_t>();
            h->sc.set_regions(h->key_tiles.as<KeyTile>(), h->n_regions, h->region_tiles.as<uint32_t>());
        }
        const uint32_t *n_u_ptr = h->sc.run(keys_in, nullptr, n_keys, h->kl.kb, l1_bits, l1_hist, keys_tmp,
                                            h->ukey.as<uint64_t>(), h->uval.as<uint32_t>(), h->overflow_flag.as<int>(), st, &h->sc_stats);
        h->n_u = d2h_scalar<uint32_t>(n_u_ptr, st);
        h->sc.collect_timing();
        int ovf = d2h_scalar<int>(h->overflow_flag.p, st);
        if (ovf) throw std::runtime_error("sub-bucket hash table overflow in umig_dedup_sort");
    }

    tr.mark("init: group (sortcombine)");
    build_segments(h);
    tr.mark("init: segments");
    h->misc.reserve(64);
    DGE_CUDA(cudaMemsetAsync(h->misc.p, 0, 16, st));
    k_count_occupied<<<grid_for(h->table_cap, 256), 256, 0, st>>>(h->tab.as<CellSlot>(), h->table_cap, h->misc.as<unsigned long long>());
    k_count_seen<<<grid_for(h->cfg.n_genes, 256), 256, 0, st>>>(h->gene_first.as<uint32_t>(), h->cfg.n_genes, h->misc.as<unsigned long long>() + 1);
    h->launches += 2;
    unsigned long long occ_seen[2] = {0, 0};
    DGE_CUDA(cudaMemcpyAsync(occ_seen, h->misc.p, 16, cudaMemcpyDeviceToHost, st));
    DGE_CUDA(cudaStreamSynchronize(st));
    h->total_cells = occ_seen[0];
    const uint64_t n_genes_seen = occ_seen[1];
    DGE_CUDA(cudaEventRecord(h->ev[1], st));
    tr.mark("init:  count occupied");

    // ---- real cells -> host (Cell::is_real, Cell.cpp:125-128: n_genes >= min_genes_before_merge)
    std::vector<CellRow> rows;
    const CellRow *rows_p = nullptr;
    size_t n_rows = 0;
    bool rows_sorted = false;            // rows_p already in cell-id order, dev_filtered = compare_cells order of all of them
    const uint32_t *dev_filtered = nullptr;
    const uint64_t *dev_gene_keys = nullptr;
    const uint32_t *dev_gene_ids = nullptr;
    int dev_filter_overflow = 0;
    bool lazy_ok = false;
    static const bool eager_host = std::getenv("DGE_EAGER_HOST") != nullptr; // experiments: build the host mirror at set_initialized
    h->n_real_rows = 0;
    if (h->n_pc)
    {
        h->flags.reserve((size_t(h->n_pc) + 1) * 4); h->flags_off.reserve((size_t(h->n_pc) + 1) * 4);
        k_real_flags<<<grid_for(h->n_pc, 256), 256, 0, st>>>(h->pc_cg_start.as<uint32_t>(), h->n_pc, h->cfg.min_genes_before_merge, h->flags.as<uint32_t>());
        ++h->launches;
        const uint32_t *tot = device_exclusive_scan(h->flags.as<uint32_t>(), h->flags_off.as<uint32_t>(), h->n_pc, h->scan_scratch.as<uint32_t>(), st, &h->launches);
        uint32_t n_real = d2h_scalar<uint32_t>(tot, st);
        if (n_real)
        {
            h->rows_dev.reserve(size_t(n_real) * sizeof(CellRow));
            k_gather_rows_flagged<<<grid_for(h->n_pc, 256), 256, 0, st>>>(h->flags.as<uint32_t>(), h->flags_off.as<uint32_t>(), h->n_pc, h->tab.as<CellSlot>(),
                                                                          h->pc_slot.as<uint32_t>(), h->pc_cg_start.as<uint32_t>(), h->pc_u_start.as<uint32_t>(),
                                                                          h->pc_reads.as<uint32_t>(), h->pc_req_genes.as<uint32_t>(),
                                                                          h->pc_req_umis.as<uint32_t>(), h->rows_dev.as<CellRow>());
            DGE_LAUNCH_CHECK();
            ++h->launches;
            // cell-id order (first-seen barcode order, CellsDataContainer.cpp:64-69) and the set_initialized filtered order
            // (all real cells, ascending compare_cells) are computed on the device: the host only receives sorted tables
            for (int b = 0; b < 2; ++b) { h->sort_k[b].reserve(size_t(n_real) * 8); h->sort_v[b].reserve(size_t(n_real) * 4); }
            h->rows_dev2.reserve(size_t(n_real) * sizeof(CellRow));
            const unsigned g = grid_for(n_real, 256);
            k_rows_first_keys<<<g, 256, 0, st>>>(h->rows_dev.as<CellRow>(), n_real, h->sort_k[0].as<uint64_t>(), h->sort_v[0].as<uint32_t>());
            device_sort_pairs(h, h->sort_k[0].as<uint64_t>(), h->sort_k[1].as<uint64_t>(), h->sort_v[0].as<uint32_t>(), h->sort_v[1].as<uint32_t>(), n_real, 32);
            k_rows_permute<<<g, 256, 0, st>>>(h->rows_dev.as<CellRow>(), h->sort_v[1].as<uint32_t>(), n_real, h->rows_dev2.as<CellRow>());
            DGE_CUDA(cudaMemsetAsync(h->overflow_flag.p, 0, sizeof(int), st));
            // compare_cells order: stable sort by barcode, then stable sort by the packed (genes, umis, stat) counters
            k_rows_cb_keys<<<g, 256, 0, st>>>(h->rows_dev2.as<CellRow>(), n_real, h->sort_k[0].as<uint64_t>(), h->sort_v[0].as<uint32_t>());
            device_sort_pairs(h, h->sort_k[0].as<uint64_t>(), h->sort_k[1].as<uint64_t>(), h->sort_v[0].as<uint32_t>(), h->sort_v[1].as<uint32_t>(), n_real,
                              int(2 * h->cfg.cb_len));
            k_rows_filter_keys<<<g, 256, 0, st>>>(h->rows_dev2.as<CellRow>(), h->sort_v[1].as<uint32_t>(), n_real, h->sort_k[0].as<uint64_t>(), h->overflow_flag.as<int>());
            device_sort_pairs(h, h->sort_k[0].as<uint64_t>(), h->sort_k[1].as<uint64_t>(), h->sort_v[1].as<uint32_t>(), h->sort_v[0].as<uint32_t>(), n_real, 64);
            DGE_LAUNCH_CHECK();
            h->launches += 4;
            DGE_CUDA(cudaMemcpyAsync(&dev_filter_overflow, h->overflow_flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
            // gene first-seen order on the device as well
            if (h->cfg.n_genes > 0)
            {
                const uint32_t ng = h->cfg.n_genes;
                for (int b = 0; b < 2; ++b) { h->gsort_k[b].reserve(size_t(ng) * 8); h->gsort_v[b].reserve(size_t(ng) * 4); }
                k_gene_first_keys<<<grid_for(ng, 256), 256, 0, st>>>(h->gene_first.as<uint32_t>(), ng, h->gsort_k[0].as<uint64_t>(), h->gsort_v[0].as<uint32_t>());
                device_sort_pairs(h, h->gsort_k[0].as<uint64_t>(), h->gsort_k[1].as<uint64_t>(), h->gsort_v[0].as<uint32_t>(), h->gsort_v[1].as<uint32_t>(), ng, 32);
                ++h->launches;
            }
            // the host copies are made lazily (materialize_host) unless something below needs them now
            lazy_ok = !(h->cfg.min_genes_before_merge == 0 && h->total_cells > h->n_pc) && !eager_host;
            if (!lazy_ok)
            {
                rows_p = d2h_pinned<CellRow>(h->pin_rows, h->rows_dev2.p, n_real, st);
                dev_filtered = d2h_pinned<uint32_t>(h->pin_filtered, h->sort_v[0].p, n_real, st);
                if (h->cfg.n_genes > 0)
                {
                    dev_gene_keys = d2h_pinned<uint64_t>(h->pin_fkeys, h->gsort_k[1].p, h->cfg.n_genes, st);
                    dev_gene_ids = d2h_pinned<uint32_t>(h->pin_gene_ids, h->gsort_v[1].p, h->cfg.n_genes, st);
                }
            }
            h->n_real_rows = n_real;
            n_rows = n_real;
            DGE_CUDA(cudaStreamSynchronize(st));
            rows_sorted = true;
        }
    }
    h->rows_on_device = rows_sorted;
    tr.mark("init:  rows d2h");
    // min_genes_before_merge == 0 makes every barcode real, including barcodes that only have intergenic reads
    // (they own no UMI and are not in the PC table): pick them up from the barcode table.
    if (h->cfg.min_genes_before_merge == 0 && h->total_cells > h->n_pc)
    {
        rows.assign(rows_p, rows_p + n_rows);
        std::vector<CellSlot> tab_host;
        d2h(tab_host, h->tab.p, h->table_cap, st);
        std::vector<uint32_t> pc_slots;
        d2h(pc_slots, h->pc_slot.p, h->n_pc, st);
        DGE_CUDA(cudaStreamSynchronize(st));
        std::vector<char> present(h->table_cap, 0);
        for (uint32_t s : pc_slots) present[s] = 1;
        for (size_t s = 0; s < h->table_cap; ++s)
            if (tab_host[s].cb != EMPTY64 && !present[s])
            {
                CellRow r{};
                r.cb = tab_host[s].cb; r.slot = uint32_t(s); r.pc = NONE32; r.first_idx = tab_host[s].first_idx; r.n_intergenic = tab_host[s].n_intergenic;
                rows.push_back(r);
            }
        rows_p = rows.data();
        n_rows = rows.size();
        rows_sorted = false;
        h->rows_on_device = false;
    }
    h->lazy_rows = rows_sorted && lazy_ok;
    h->init_filter_overflow = dev_filter_overflow;
    h->dev_merged = false;
    if (h->lazy_rows)
    {   // no host mirror yet: materialize_host builds it on demand
        h->real.clear(); h->filtered.clear(); h->gene_order.clear();
        h->host_stage = 0;
        h->sum_real = h->sum_filtered = n_rows;
        h->sum_genes_seen = n_genes_seen;
    }
    else
    {
        build_host_mirror(h, rows_p, n_rows, rows_sorted, dev_filtered, dev_gene_keys, dev_gene_ids, dev_filter_overflow, tr);
        h->sum_real = h->real.size(); h->sum_filtered = h->filtered.size(); h->sum_genes_seen = h->gene_order.size();
    }
    DGE_CUDA(cudaEventRecord(h->ev[2], st));
    DGE_CUDA(cudaStreamSynchronize(st));
    h->state = 1;
    cudaEventElapsedTime(&h->timings.ms_fill, h->ev[0], h->ev[1]);
    cudaEventElapsedTime(&h->timings.ms_init, h->ev[1], h->ev[2]);
    h->timings.ms_total = h->timings.ms_fill + h->timings.ms_init;
    h->timings.ms_dedup_kernel = h->sc_stats.dedup_ms;
    h->timings.n_kernel_launches = h->launches + h->sc_stats.launches;
    h->timings.n_dedup_launches = h->sc_stats.dedup_launches;
    h->timings.ms_fill_kernel = 0;
    for (size_t e = 0; e + 1 < h->n_fill_ev; e += 2)
    {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, h->fill_ev[e], h->fill_ev[e + 1]) == cudaSuccess) h->timings.ms_fill_kernel += ms;
    }
    h->timings.n_fill_launches = uint32_t(h->n_fill_ev / 2);
}

// Tools::CollisionsAdjuster table on the device: adj[s-1] = estimate_adjusted_gene_expression(s), s = 1..max_expr, for the UMI
// probabilities d_p (device).  Fast pass with a parallel sum; when a step comes too close to a rounding boundary the table is
// recomputed with the terms summed in index order (collisions.cuh).
void collisions_adjusted_device(cudaStream_t st, const double *d_p, size_t n, size_t max_expr, unsigned long long *d_adj, uint32_t *exact_rerun)
{
    DevBuf neg, terms, partial, state;
    neg.reserve(std::max<size_t>(n, 1) * 8);
    const unsigned blocks = unsigned(std::max<size_t>(1, std::min<size_t>(div_up(n, size_t(CA_THREADS)), 148 * 8)));
    partial.reserve(size_t(blocks) * 8); state.reserve(sizeof(CollisionsState));
    std::vector<double> ones(n, 1.0);
    // worst-case difference between two association orders of the sum, propagated through 1/(1-q): see collisions.cuh
    const double drift = std::max(1e-13, double(n) * 4e-15);
    const bool force_exact = std::getenv("DGE_CA_FORCE_EXACT") != nullptr; // tests: exercise the sequential-order path
    for (int exact = force_exact ? 1 : 0; exact < 2; ++exact)
    {
        DGE_CUDA(cudaMemcpyAsync(neg.p, ones.data(), n * 8, cudaMemcpyHostToDevice, st));
        DGE_CUDA(cudaMemsetAsync(state.p, 0, sizeof(CollisionsState), st));
        if (exact) terms.reserve(std::max<size_t>(n, 1) * 8);
        for (size_t s = 1; s <= max_expr; ++s)
        {
            if (exact)
                k_collisions_step<true><<<blocks, CA_THREADS, 0, st>>>(d_p, neg.as<double>(), terms.as<double>(), n, s, state.as<CollisionsState>(),
                                                                      partial.as<double>(), d_adj, drift);
            else
                k_collisions_step<false><<<blocks, CA_THREADS, 0, st>>>(d_p, neg.as<double>(), nullptr, n, s, state.as<CollisionsState>(),
                                                                       partial.as<double>(), d_adj, drift);
        }
        DGE_LAUNCH_CHECK();
        CollisionsState hs;
        DGE_CUDA(cudaMemcpyAsync(&hs, state.p, sizeof(hs), cudaMemcpyDeviceToHost, st));
        DGE_CUDA(cudaStreamSynchronize(st));
        if (!exact && hs.risky == 0) break;
        if (!exact && exact_rerun) *exact_rerun = hs.risky;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
void upload_whitelist(dge_handle *h)
{
    WhitelistDev &d = h->wl_dev;
    d.n_parts = int(h->wl.parts.size());
    int shift = int(2 * h->cfg.cb_len);
    for (int k = 0; k < d.n_parts; ++k)
    {
        const auto &p = h->wl.parts[size_t(k)];
        d.part_len[k] = int(p[0].size());
        shift -= 2 * d.part_len[k];
        d.part_shift[k] = shift;
        d.part_size[k] = uint32_t(p.size());
        std::vector<uint32_t> packed(p.size());
        for (size_t t = 0; t < p.size(); ++t) { uint64_t v; pack_seq(p[t], v); packed[t] = uint32_t(v); }
        h->wl_tokens[k].reserve(packed.size() * 4);
        DGE_CUDA(cudaMemcpyAsync(h->wl_tokens[k].p, packed.data(), packed.size() * 4, cudaMemcpyHostToDevice, h->stream));
        DGE_CUDA(cudaStreamSynchronize(h->stream));
        d.tokens[k] = h->wl_tokens[k].as<uint32_t>();
    }
}

// Device intersections for a flat job list; results land in h->h_isect (same order).
void run_intersections(dge_handle *h, const std::vector<PairJob> &jobs)
{
    h->pin_isect.reserve(std::max<size_t>(jobs.size(), 1) * 4);
    h->isect_p = h->pin_isect.as<uint32_t>();
    if (jobs.empty()) return;
    h->d_jobs.reserve(jobs.size() * sizeof(PairJob)); h->d_isect.reserve(jobs.size() * 4);
    DGE_CUDA(cudaMemcpyAsync(h->d_jobs.p, jobs.data(), jobs.size() * sizeof(PairJob), cudaMemcpyHostToDevice, h->stream));
    k_intersect<<<unsigned(jobs.size()), 128, 0, h->stream>>>(h->d_jobs.as<PairJob>(), uint32_t(jobs.size()), h->ukey.as<uint64_t>(),
                                                              h->pc_u_start.as<uint32_t>(), h->pc_slot.as<uint32_t>(), h->kl.gb + h->kl.ub,
                                                              h->d_isect.as<uint32_t>());
    DGE_LAUNCH_CHECK();
    ++h->launches;
    DGE_CUDA(cudaMemcpyAsync(h->isect_p, h->d_isect.p, jobs.size() * 4, cudaMemcpyDeviceToHost, h->stream));
    DGE_CUDA(cudaStreamSynchronize(h->stream));
}

// RealBarcodesMergeStrategy::get_best_merge_target (RealBarcodesMergeStrategy.cpp:31-61) over an ORDERED neighbour list.
long best_target_real(const dge_handle *h, uint32_t base, const uint32_t *nbs, const uint32_t *isect, size_t n_nb)
{
    if (nbs[0] == base) return long(base);
    double max_frac = 0;
    uint32_t best = nbs[0];
    for (size_t k = 0; k < n_nb; ++k)
    {
        double frac = 0.5 * isect[k] * (1. / h->real[base].umis_stat + 1. / h->real[nbs[k]].umis_stat);
        if (max_frac < frac) { max_frac = frac; best = nbs[k]; }
    }
    if (max_frac < h->cfg.min_merge_fraction) return -1;
    return long(best);
}

// Phase 1 of the whitelist merge for the n real cells of rows_dev2, entirely on the device: neighbour classes 0/1, pair jobs,
// (gene,UMI) intersections, best target per cell.  Leaves d_count (neighbour count / NB_SELF / NB_SLOW), p1_target and p1_flag
// (1 = order-dependent tie: the reference's neighbour order decides) in device memory.
void p1_device_pass(dge_handle *h, size_t n)
{
    cudaStream_t st = h->stream;
    if (!h->wl_uploaded) { upload_whitelist(h); h->wl_uploaded = true; }
    h->d_cb.reserve(n * 8); h->d_umis.reserve(n * 4); h->d_count.reserve(n * 4); h->d_nb.reserve(n * WL_K * 4);
    h->p1_pc.reserve(n * 4); h->p1_map.reserve((size_t(h->n_pc) + 2) * 4);
    k_fill_u32<<<grid_for(size_t(h->n_pc) + 1, 256), 256, 0, st>>>(h->p1_map.as<uint32_t>(), size_t(h->n_pc) + 1, NONE32);
    k_p1_columns<<<grid_for(n, 256), 256, 0, st>>>(h->rows_dev2.as<CellRow>(), uint32_t(n), h->d_cb.as<uint64_t>(), h->d_umis.as<uint32_t>(),
                                                   h->p1_pc.as<uint32_t>(), h->p1_map.as<uint32_t>());
    k_wl_class01<<<unsigned(div_up(n * 32, size_t(256))), 256, 0, st>>>(h->d_cb.as<uint64_t>(), h->d_umis.as<uint32_t>(), uint32_t(n), h->wl_dev,
                                                                        h->tab.as<CellSlot>(), h->kl.tb, h->slot_pc.as<uint32_t>(),
                                                                        h->pc_cg_start.as<uint32_t>(), h->pc_u_start.as<uint32_t>(),
                                                                        h->cfg.min_genes_before_merge, h->d_count.as<int>(), h->d_nb.as<uint32_t>());
    DGE_LAUNCH_CHECK();
    h->launches += 3;
    h->p1_cnt.reserve((n + 1) * 4); h->p1_off.reserve((n + 1) * 4); h->p1_target.reserve(n * 4); h->p1_flag.reserve(n * 4);
    k_p1_counts<<<grid_for(n, 256), 256, 0, st>>>(h->d_count.as<int>(), uint32_t(n), h->p1_cnt.as<uint32_t>());
    DGE_CUDA(cudaMemsetAsync(h->p1_cnt.as<uint32_t>() + n, 0, 4, st));
    device_exclusive_scan(h->p1_cnt.as<uint32_t>(), h->p1_off.as<uint32_t>(), n + 1, h->scan_scratch.as<uint32_t>(), st, &h->launches);
    const uint32_t n_jobs = d2h_scalar<uint32_t>(h->p1_off.as<uint32_t>() + n, st);
    h->d_jobs.reserve(std::max<size_t>(n_jobs, 1) * sizeof(PairJob)); h->d_isect.reserve(std::max<size_t>(n_jobs, 1) * 4);
    k_p1_jobs<<<grid_for(n, 256), 256, 0, st>>>(h->d_count.as<int>(), h->d_nb.as<uint32_t>(), h->p1_pc.as<uint32_t>(), h->p1_off.as<uint32_t>(), uint32_t(n),
                                                h->d_jobs.as<PairJob>());
    if (n_jobs)
        k_intersect<<<n_jobs, 128, 0, st>>>(h->d_jobs.as<PairJob>(), n_jobs, h->ukey.as<uint64_t>(), h->pc_u_start.as<uint32_t>(),
                                            h->pc_slot.as<uint32_t>(), h->kl.gb + h->kl.ub, h->d_isect.as<uint32_t>());
    k_p1_best<<<grid_for(n, 256), 256, 0, st>>>(h->d_count.as<int>(), h->d_nb.as<uint32_t>(), h->p1_off.as<uint32_t>(), h->d_isect.as<uint32_t>(),
                                                h->d_umis.as<uint32_t>(), h->p1_map.as<uint32_t>(), uint32_t(n), h->cfg.min_merge_fraction,
                                                h->p1_target.as<int>(), h->p1_flag.as<uint32_t>());
    DGE_LAUNCH_CHECK();
    h->launches += 4;
}

// Phase 1 for RealBarcodesMergeStrategy: target (index into real, or -1) for every real cell.
void phase1_real(dge_handle *h, std::vector<long> &target)
{
    cudaStream_t st = h->stream;
    Tracer tr; tr.st = st;
    const size_t n = h->real.size();
    target.assign(n, -2);
    h->n_unresolved = 0;
    if (n == 0) return;
    h->pin_nbc.reserve(n * 4); h->pin_nbp.reserve(n * WL_K * 4);
    int *nb_count = h->pin_nbc.as<int>();
    uint32_t *nb_pc = h->pin_nbp.as<uint32_t>();
    std::vector<char> todo(n, 1);        // cells whose target the host still has to work out
    const bool device_rows = h->rows_on_device && !h->cfg.sharded && h->wl_fast;
    if (h->wl_fast)
    {
        if (!h->wl_uploaded) { upload_whitelist(h); h->wl_uploaded = true; }
        h->d_cb.reserve(n * 8); h->d_umis.reserve(n * 4); h->d_count.reserve(n * 4); h->d_nb.reserve(n * WL_K * 4);
        if (device_rows)
        {   // the per-cell columns come straight from the rows gathered at set_initialized (nothing changed since); intersections and
            // the best neighbour per cell on the device; only ties / far classes come back to the host logic
            p1_device_pass(h, n);
            DGE_CUDA(cudaMemcpyAsync(nb_count, h->d_count.p, n * 4, cudaMemcpyDeviceToHost, st));
            const int *dt = d2h_pinned<int>(h->pin_p1t, h->p1_target.p, n, st);
            const uint32_t *df = d2h_pinned<uint32_t>(h->pin_p1f, h->p1_flag.p, n, st);
            DGE_CUDA(cudaStreamSynchronize(st));
            size_t n_todo = 0;
            for (size_t i = 0; i < n; ++i)
            {
                const int c = nb_count[i];
                if (c == NB_SELF) { target[i] = long(i); todo[i] = 0; }
                else if (c > 0 && !df[i]) { target[i] = long(dt[i]); todo[i] = 0; }
                else ++n_todo;
            }
            if (n_todo == 0) { tr.mark("merge:  p1 device pass"); return; }
        }
        else
        {
            h->h_cbs.resize(n); h->h_umis.resize(n);
            for (size_t i = 0; i < n; ++i) { h->h_cbs[i] = h->real[i].cb; h->h_umis[i] = uint32_t(h->real[i].umis_stat); }
            DGE_CUDA(cudaMemcpyAsync(h->d_cb.p, h->h_cbs.data(), n * 8, cudaMemcpyHostToDevice, st));
            DGE_CUDA(cudaMemcpyAsync(h->d_umis.p, h->h_umis.data(), n * 4, cudaMemcpyHostToDevice, st));
            k_wl_class01<<<unsigned(div_up(n * 32, size_t(256))), 256, 0, st>>>(h->d_cb.as<uint64_t>(), h->d_umis.as<uint32_t>(), uint32_t(n), h->wl_dev,
                                                                                h->tab.as<CellSlot>(), h->kl.tb, h->slot_pc.as<uint32_t>(),
                                                                                h->pc_cg_start.as<uint32_t>(), h->pc_u_start.as<uint32_t>(),
                                                                                h->cfg.min_genes_before_merge, h->d_count.as<int>(), h->d_nb.as<uint32_t>());
            DGE_LAUNCH_CHECK();
            ++h->launches;
            DGE_CUDA(cudaMemcpyAsync(nb_count, h->d_count.p, n * 4, cudaMemcpyDeviceToHost, st));
        }
        DGE_CUDA(cudaMemcpyAsync(nb_pc, h->d_nb.p, n * WL_K * 4, cudaMemcpyDeviceToHost, st));
        DGE_CUDA(cudaStreamSynchronize(st));
    }
    else std::fill(nb_count, nb_count + n, int(NB_SLOW));
    std::vector<uint32_t> &pc_to_real = h->h_pc_to_real;
    pc_to_real.assign(size_t(h->n_pc) + 1, NONE32);
    for (uint32_t i = 0; i < n; ++i) if (h->real[i].pc != NONE32) pc_to_real[h->real[i].pc] = i;
    tr.mark("merge:  p1 wl kernel + d2h");

    // exact host path (built lazily: most runs never need it)
    std::unordered_map<uint64_t, uint32_t> by_cb;
    auto exact_neighbours = [&](uint32_t i) {
        if (by_cb.empty())
        {
            by_cb.reserve(n * 2);
            for (uint32_t r = 0; r < n; ++r) by_cb.emplace(h->real[r].cb, r);
        }
        const std::string cb = cb_string(h, h->real[i].cb);
        auto lookup = [&](const std::string &s) -> long {
            uint64_t v;
            if (!pack_seq(s, v)) return -1;
            auto it = by_cb.find(v);
            return it == by_cb.end() ? -1 : long(it->second);
        };
        // a barcode known to the container but not real fails the size test exactly like in the reference
        auto eligible = [&](long id) {
            return uint32_t(h->real[size_t(id)].n_genes) >= h->cfg.min_genes_before_merge &&
                   h->real[size_t(id)].umis_stat >= h->real[i].umis_stat;
        };
        std::vector<long> ids = h->wl.neighbours(cb, false, lookup, eligible);
        return std::vector<uint32_t>(ids.begin(), ids.end());
    };

    // neighbour lists in CSR form
    std::vector<uint32_t> &off = h->h_nb_off, &nbs = h->h_nbs;
    off.assign(n + 1, 0);
    std::unordered_map<uint32_t, std::vector<uint32_t>> slow_lists;
    for (uint32_t i = 0; i < n; ++i)
    {
        uint32_t c = 0;
        if (!todo[i]) { off[i + 1] = off[i]; continue; } // settled by the device pass
        if (nb_count[i] == NB_SELF) target[i] = long(i);
        else if (nb_count[i] == NB_SLOW && h->cfg.sharded)
        {   // candidates may live on another shard: never guess, leave the cell as it is and report it
            target[i] = long(i);
            ++h->n_unresolved;
        }
        else if (nb_count[i] == NB_SLOW)
        {
            auto &lst = slow_lists[i];
            lst = exact_neighbours(i);
            c = uint32_t(lst.size());
            if (lst.empty()) target[i] = -1;
            else if (lst[0] == i) { target[i] = long(i); c = 0; }
        }
        else c = uint32_t(nb_count[i]);
        off[i + 1] = off[i] + c;
    }
    nbs.resize(off[n]);
    std::vector<PairJob> &jobs = h->h_jobs;
    jobs.resize(off[n]);
    const uint32_t empty_pc = h->n_pc; // the empty sentinel cell: intersects nothing
    for (uint32_t i = 0; i < n; ++i)
    {
        const uint32_t c = off[i + 1] - off[i];
        if (!c) continue;
        if (nb_count[i] == NB_SLOW) { auto const &lst = slow_lists[i]; std::copy(lst.begin(), lst.end(), nbs.begin() + off[i]); }
        else for (uint32_t k = 0; k < c; ++k) nbs[off[i] + k] = pc_to_real[nb_pc[size_t(i) * WL_K + k]];
        for (uint32_t k = 0; k < c; ++k)
        {
            const uint32_t a = h->real[i].pc, b = h->real[nbs[off[i] + k]].pc;
            jobs[off[i] + k] = PairJob{a == NONE32 ? empty_pc : a, b == NONE32 ? empty_pc : b};
        }
    }
    run_intersections(h, jobs);
    const uint32_t *isect = h->isect_p;
    tr.mark("merge:  p1 jobs + intersections");

    for (uint32_t i = 0; i < n; ++i)
    {
        if (target[i] != -2) continue;
        const uint32_t c = off[i + 1] - off[i];
        const uint32_t *my_nbs = nbs.data() + off[i];
        const uint32_t *my_is = isect + off[i];
        // The fast path reports the neighbour SET of the nearest class; the reference walks them in the order left by two
        // unstable sorts.  That order only matters on exact ties of the best fraction: replay it then.
        if (nb_count[i] > 1)
        {
            double best = -1; int n_best = 0;
            for (uint32_t k = 0; k < c; ++k)
            {
                double frac = 0.5 * my_is[k] * (1. / h->real[i].umis_stat + 1. / h->real[my_nbs[k]].umis_stat);
                if (frac > best) { best = frac; n_best = 1; } else if (frac == best) ++n_best;
            }
            if (n_best > 1 && !(best < h->cfg.min_merge_fraction))
            {
                std::vector<uint32_t> ordered = exact_neighbours(i);
                std::vector<uint32_t> isect_ord(ordered.size(), 0);
                for (size_t a = 0; a < ordered.size(); ++a)
                    for (uint32_t b = 0; b < c; ++b)
                        if (my_nbs[b] == ordered[a]) isect_ord[a] = my_is[b];
                target[i] = best_target_real(h, i, ordered.data(), isect_ord.data(), ordered.size());
                continue;
            }
        }
        target[i] = best_target_real(h, i, my_nbs, my_is, c);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// PoissonTargetEstimator (reference Merge/PoissonTargetEstimator.cpp).  init(): UMI distribution over the filtered cells + the
// CollisionsAdjuster table; then estimate_intersection_prob for a batch of (base, other) pairs of real-cell indices.
void poisson_init(dge_handle *h)
{
    if (h->pp_ready) return;
    cudaStream_t st = h->stream;
    const size_t n = h->real.size();
    if (!h->slot_pc_built) build_slot_pc(h);
    // real flag per present cell + real index -> present cell
    std::vector<uint32_t> pc_real(size_t(h->n_pc) + 2, 0), real_pc(std::max<size_t>(n, 1), h->n_pc);
    for (size_t i = 0; i < n; ++i)
    {
        const HostCell &c = h->real[i];
        if (c.pc == NONE32) continue;
        real_pc[i] = c.pc;
        if (c.real) pc_real[c.pc] = 1; // filtered_cells() at merge time = every real cell
    }
    h->pp_pc_real.reserve(pc_real.size() * 4); h->pp_real_pc.reserve(real_pc.size() * 4);
    DGE_CUDA(cudaMemcpyAsync(h->pp_pc_real.p, pc_real.data(), pc_real.size() * 4, cudaMemcpyHostToDevice, st));
    DGE_CUDA(cudaMemcpyAsync(h->pp_real_pc.p, real_pc.data(), real_pc.size() * 4, cudaMemcpyHostToDevice, st));
    DGE_CUDA(cudaStreamSynchronize(st));
    const size_t n_umi = size_t(1) << h->kl.ub;
    h->pp_hist.reserve(n_umi * 8); h->pp_p.reserve(n_umi * 8); h->pp_misc.reserve(64);
    DGE_CUDA(cudaMemsetAsync(h->pp_hist.p, 0, n_umi * 8, st));
    DGE_CUDA(cudaMemsetAsync(h->pp_misc.p, 0, 64, st));
    unsigned long long *total = h->pp_misc.as<unsigned long long>();
    uint32_t *max_size = reinterpret_cast<uint32_t *>(total + 1);
    if (h->n_u)
        k_umi_hist<<<grid_for(h->n_u, 256), 256, 0, st>>>(h->ukey.as<uint64_t>(), h->n_u, h->kl.ub, h->kl.gb + h->kl.ub, h->slot_pc.as<uint32_t>(),
                                                          h->pp_pc_real.as<uint32_t>(), h->pp_hist.as<unsigned long long>());
    k_hist_total<<<grid_for(n_umi, 256), 256, 0, st>>>(h->pp_hist.as<unsigned long long>(), n_umi, total);
    k_hist_to_prob<<<grid_for(n_umi, 256), 256, 0, st>>>(h->pp_hist.as<unsigned long long>(), n_umi, total, h->pp_p.as<double>());
    if (h->n_cg)
        k_max_gene_size<<<grid_for(h->n_cg, 256), 256, 0, st>>>(h->cg_start.as<uint32_t>(), h->cg_pc.as<uint32_t>(), h->n_cg, h->pp_pc_real.as<uint32_t>(), max_size);
    DGE_LAUNCH_CHECK();
    h->launches += 4;
    h->pp_max_gene_size = d2h_scalar<uint32_t>(max_size, st);
    const size_t m = std::max<uint32_t>(h->pp_max_gene_size, 1);
    h->pp_adj.reserve(m * 8);
    uint32_t rerun = 0;
    collisions_adjusted_device(st, h->pp_p.as<double>(), n_umi, m, h->pp_adj.as<unsigned long long>(), &rerun);
    h->pp_ready = true;
}

// prob[p] = PoissonTargetEstimator::estimate_intersection_prob(base, other).merge_probability for the flagged pairs (2.0 for the others).
// d_pkey[p] = (base << rb) | other (real-cell indices), d_pval[p] = intersection size, all DEVICE arrays of n_p entries.
void poisson_eval_pairs(dge_handle *h, const uint64_t *d_pkey, const uint32_t *d_pval, const uint32_t *d_flag, uint32_t n_p, int rb, double *d_prob)
{
    poisson_init(h);
    if (!n_p) return;
    cudaStream_t st = h->stream;
    const unsigned g = grid_for(n_p, 256);
    if (h->pp_max_gene_size && d2h_scalar<unsigned long long>(h->pp_adj.as<unsigned long long>() + (h->pp_max_gene_size - 1), st) >= (1ull << 29))
        throw CapacityError("adjusted gene sizes beyond 2^29");
    // distinct adjusted size pairs -> hash set (grown until the load factor stays below 1/2)
    uint32_t cap = 1u << 16, n_sp = 0;
    h->pp_misc.reserve(64);
    int *full = reinterpret_cast<int *>(h->pp_misc.as<unsigned long long>() + 4);
    while (true)
    {
        h->pp_skey.reserve(size_t(cap) * 8); h->pp_cnt.reserve((size_t(cap) + 1) * 4); h->pp_off.reserve((size_t(cap) + 1) * 4);
        DGE_CUDA(cudaMemsetAsync(h->pp_skey.p, 0xFF, size_t(cap) * 8, st));
        DGE_CUDA(cudaMemsetAsync(full, 0, 4, st));
        k_pp_shared<<<g, 256, 0, st>>>(d_pkey, d_flag, n_p, rb, h->pp_real_pc.as<uint32_t>(), h->pc_cg_start.as<uint32_t>(), h->cg_gene.as<uint32_t>(),
                                       h->cg_start.as<uint32_t>(), h->pp_adj.as<unsigned long long>(), h->pp_skey.as<unsigned long long>(), cap - 1, full);
        k_sp_occupied<<<grid_for(cap, 256), 256, 0, st>>>(h->pp_skey.as<unsigned long long>(), cap, h->pp_cnt.as<uint32_t>());
        DGE_CUDA(cudaMemsetAsync(h->pp_cnt.as<uint32_t>() + cap, 0, 4, st));
        device_exclusive_scan(h->pp_cnt.as<uint32_t>(), h->pp_off.as<uint32_t>(), size_t(cap) + 1, h->scan_scratch.as<uint32_t>(), st, &h->launches);
        DGE_LAUNCH_CHECK();
        n_sp = d2h_scalar<uint32_t>(h->pp_off.as<uint32_t>() + cap, st);
        const int is_full = d2h_scalar<int>(full, st);
        if (!is_full && n_sp <= cap / 2) break;
        if (cap >= (1u << 30)) throw CapacityError("too many distinct gene size pairs");
        cap <<= 2;
    }
    h->pp_sval.reserve(std::max<size_t>(n_sp, 1) * 4); h->pp_est.reserve(size_t(cap) * 8);
    if (n_sp)
    {
        k_sp_list<<<grid_for(cap, 256), 256, 0, st>>>(h->pp_skey.as<unsigned long long>(), h->pp_cnt.as<uint32_t>(), h->pp_off.as<uint32_t>(), cap, h->pp_sval.as<uint32_t>());
        k_pp_est<<<std::min<uint32_t>(n_sp, 148 * 16), 256, 0, st>>>(h->pp_skey.as<unsigned long long>(), h->pp_sval.as<uint32_t>(), n_sp, h->pp_p.as<double>(),
                                                                    size_t(1) << h->kl.ub, h->pp_est.as<double>());
    }
    k_pp_lambda<<<g, 256, 0, st>>>(d_pkey, d_pval, d_flag, n_p, rb, h->pp_real_pc.as<uint32_t>(), h->pc_cg_start.as<uint32_t>(), h->cg_gene.as<uint32_t>(),
                                   h->cg_start.as<uint32_t>(), h->pp_adj.as<unsigned long long>(), h->pp_skey.as<unsigned long long>(), cap - 1, h->pp_est.as<double>(), d_prob);
    DGE_LAUNCH_CHECK();
    h->launches += 4;
}

// ---------------------------------------------------------------------------------------------------------------------
// SimpleMergeStrategy (Merge/SimpleMergeStrategy.cpp).  Device: inverted index, common UMI-gene counts, best candidate per base
// (simplemerge.cuh).  Host: the final threshold compare in the reference's own expression, and an exact replay with the
// reference's containers for the bases whose outcome depends on hash-iteration order (near-ties of the top fraction).
struct SimpleReplay
{
    bool ready = false;
    std::vector<uint64_t> E;                 // sorted inverted index [gu : gub | real idx : rb]
    std::vector<uint32_t> gene_rank;         // gene id -> StringIndexer id
    std::vector<uint32_t> umi_first;         // UMI value -> first read index (orders UMI ids)
    std::vector<size_t> gid;                 // real idx -> reference cell id (first-seen rank among ALL barcodes)
    std::unordered_map<size_t, uint32_t> ridx_of_gid;
    std::vector<uint32_t> pos_in_filtered;   // real idx -> position in filtered_cells()
};

void simple_replay_prepare(dge_handle *h, SimpleReplay &R, uint32_t n_e)
{
    cudaStream_t st = h->stream;
    const size_t n = h->real.size();
    d2h(R.E, h->sm_ekey.p, n_e, st);
    d2h(R.umi_first, h->umi_first.p, size_t(1) << h->kl.ub, st);
    std::vector<CellSlot> tab;
    d2h(tab, h->tab.p, h->table_cap, st);
    DGE_CUDA(cudaStreamSynchronize(st));
    R.gene_rank.assign(h->cfg.n_genes, NONE32);
    for (size_t r = 0; r < h->gene_order.size(); ++r) R.gene_rank[size_t(h->gene_order[r])] = uint32_t(r);
    std::vector<uint32_t> firsts;
    firsts.reserve(size_t(h->total_cells));
    for (auto const &s : tab) if (s.cb != EMPTY64) firsts.push_back(s.first_idx);
    std::sort(firsts.begin(), firsts.end());
    R.gid.resize(n);
    R.ridx_of_gid.reserve(n * 2);
    for (uint32_t i = 0; i < n; ++i)
    {
        R.gid[i] = size_t(std::lower_bound(firsts.begin(), firsts.end(), h->real[i].first_idx) - firsts.begin());
        R.ridx_of_gid.emplace(R.gid[i], i);
    }
    R.pos_in_filtered.assign(n, NONE32);
    for (uint32_t p = 0; p < h->filtered.size(); ++p) R.pos_in_filtered[h->filtered[p]] = p;
    R.ready = true;
}

// SimpleMergeStrategy::get_cells_with_common_umigs for one base cell with the reference's containers and insertion sequences:
// (other real idx, common UMI-genes) in the iteration order of the reference's unordered_map.
std::vector<std::pair<uint32_t, size_t>> simple_replay_candidates(dge_handle *h, SimpleReplay &R, uint32_t base, int rb)
{
    cudaStream_t st = h->stream;
    const HostCell &bc = h->real[base];
    const int ub = h->kl.ub, gub = h->kl.gb + h->kl.ub;
    uint32_t range[2];
    DGE_CUDA(cudaMemcpyAsync(range, h->pc_u_start.as<uint32_t>() + bc.pc, 8, cudaMemcpyDeviceToHost, st));
    DGE_CUDA(cudaStreamSynchronize(st));
    std::vector<uint64_t> mine;
    d2h(mine, h->ukey.as<uint64_t>() + range[0], range[1] - range[0], st);
    DGE_CUDA(cudaStreamSynchronize(st));
    struct Item { uint32_t gene_rank, umi_first; uint64_t gu; };
    std::vector<Item> items;
    items.reserve(mine.size());
    const uint64_t gu_mask = (1ull << gub) - 1, rmask = (1ull << rb) - 1;
    for (uint64_t uk : mine)
    {
        const uint64_t gu = uk & gu_mask;
        items.push_back(Item{R.gene_rank[size_t(gu >> ub)], R.umi_first[size_t(gu & ((1ull << ub) - 1))], gu});
    }
    // Cell::genes() is a std::map over gene ids, Gene::umis() a std::map over UMI ids: ascending StringIndexer ids
    std::sort(items.begin(), items.end(), [](const Item &a, const Item &b) { return a.gene_rank != b.gene_rank ? a.gene_rank < b.gene_rank : a.umi_first < b.umi_first; });
    std::unordered_map<size_t, size_t> common; // u_u_hash_t
    std::vector<uint32_t> members;
    for (auto const &it : items)
    {
        const uint64_t sg = umig_scramble(it.gu, gub); // the index is sorted by the scrambled (gene, umi) word
        auto lo = std::lower_bound(R.E.begin(), R.E.end(), sg << rb);
        auto hi = std::lower_bound(lo, R.E.end(), (sg + 1) << rb);
        members.clear();
        for (auto e = lo; e != hi; ++e) members.push_back(uint32_t(*e & rmask));
        // init() walks filtered_cells() in order and emplaces every cell into the set of each of its UMI-genes
        std::sort(members.begin(), members.end(), [&](uint32_t a, uint32_t b) { return R.pos_in_filtered[a] < R.pos_in_filtered[b]; });
        std::unordered_set<size_t> cells; // sul_set_t
        for (uint32_t m : members) cells.emplace(R.gid[m]);
        for (size_t other : cells)
        {
            if (other == R.gid[base]) continue;
            if (h->real[R.ridx_of_gid.at(other)].n_genes >= bc.n_genes) common[other]++;
        }
    }
    std::vector<std::pair<uint32_t, size_t>> res;
    res.reserve(common.size());
    for (auto const &kv : common) res.emplace_back(R.ridx_of_gid.at(kv.first), kv.second);
    return res;
}

// SimpleMergeStrategy::get_merge_target for one base cell (SimpleMergeStrategy.cpp:48-88) over the replayed candidate order.
long simple_replay(dge_handle *h, SimpleReplay &R, uint32_t base, int rb)
{
    const HostCell &bc = h->real[base];
    long top = -1, top_genes = -1;
    double top_frac = -1;
    const std::string base_cb = cb_string(h, bc.cb);
    for (auto const &kv : simple_replay_candidates(h, R, base, rb))
    {
        const uint32_t o = kv.first;
        const HostCell &oc = h->real[o];
        const double frac = 0.5 * kv.second * (1. / size_t(bc.umis_stat) + 1. / size_t(oc.umis_stat));
        if (frac - top_frac > 0.00001 || (std::abs(frac - top_frac) < 0.00001 && long(oc.n_genes) > top_genes))
        {
            const int ed = int(edit_distance_ref(base_cb.c_str(), cb_string(h, oc.cb).c_str()));
            if (ed >= int(h->cfg.max_cb_merge_edit_distance)) continue;
            top = long(o); top_frac = frac; top_genes = long(oc.n_genes);
        }
    }
    if (top_frac < h->cfg.min_merge_fraction) return long(base);
    return top;
}

void phase1_simple(dge_handle *h, std::vector<long> &target, bool poisson = false)
{
    cudaStream_t st = h->stream;
    Tracer tr; tr.st = st;
    const size_t n = h->real.size();
    target.resize(n);
    for (size_t i = 0; i < n; ++i) target[i] = long(i);
    h->n_simple_replayed = 0;
    if (n < 2) return;
    const int gub = h->kl.gb + h->kl.ub;
    const int rb = std::max(1, ceil_log2_u64(n));
    if (gub + rb + 3 > 64) throw std::runtime_error("simple merge: key layout does not fit 64 bits");

    // ---- inverted index over the UMIs of every real (= filtered at this point) cell
    std::vector<UmigJob> jobs;
    jobs.reserve(n);
    uint64_t total = 0;
    for (uint32_t i = 0; i < n; ++i)
    {
        const HostCell &c = h->real[i];
        if (c.pc == NONE32 || c.n_umis_distinct == 0) continue;
        jobs.push_back(UmigJob{c.pc, i, uint32_t(total)});
        total += uint64_t(c.n_umis_distinct);
    }
    if (total == 0) return;
    if (total >= 0xFFFFFFF0ull) throw std::runtime_error("simple merge: more than 2^32 UMIs in real cells");
    h->sm_jobs.reserve(jobs.size() * sizeof(UmigJob));
    h->sm_ikeys.reserve(total * 8); h->sm_ekey.reserve(total * 8); h->sm_eval.reserve(total * 4);
    DGE_CUDA(cudaMemcpyAsync(h->sm_jobs.p, jobs.data(), jobs.size() * sizeof(UmigJob), cudaMemcpyHostToDevice, st));
    k_umig_keys<<<grid_for(jobs.size(), 1, 148 * 16), 256, 0, st>>>(h->sm_jobs.as<UmigJob>(), uint32_t(jobs.size()), h->ukey.as<uint64_t>(),
                                                                   h->pc_u_start.as<uint32_t>(), gub, rb, h->sm_ikeys.as<uint64_t>());
    DGE_LAUNCH_CHECK();
    ++h->launches;
    const int kb1 = gub + rb + 3;
    const uint32_t *n_e_ptr = h->sc2.run(h->sm_ikeys.as<uint64_t>(), nullptr, total, kb1, std::min(choose_l1_bits(total), kb1 - 3), nullptr, h->sm_ikeys.as<uint64_t>(),
                                         h->sm_ekey.as<uint64_t>(), h->sm_eval.as<uint32_t>(), h->overflow_flag.as<int>(), st, &h->sc_stats);
    const uint32_t n_e = d2h_scalar<uint32_t>(n_e_ptr, st);
    h->sc2.collect_timing();
    if (d2h_scalar<int>(h->overflow_flag.p, st)) throw std::runtime_error("sub-bucket hash table overflow while indexing UMI-genes");
    tr.mark("merge:  simple: inverted index");

    // ---- per-cell columns the kernels need
    std::vector<uint32_t> ng(n), um(n);
    std::vector<uint64_t> cbs(n);
    for (size_t i = 0; i < n; ++i) { ng[i] = uint32_t(h->real[i].n_genes); um[i] = uint32_t(h->real[i].umis_stat); cbs[i] = h->real[i].cb; }
    h->sm_ngenes.reserve(n * 4); h->sm_umis.reserve(n * 4); h->sm_cb.reserve(n * 8);
    DGE_CUDA(cudaMemcpyAsync(h->sm_ngenes.p, ng.data(), n * 4, cudaMemcpyHostToDevice, st));
    DGE_CUDA(cudaMemcpyAsync(h->sm_umis.p, um.data(), n * 4, cudaMemcpyHostToDevice, st));
    DGE_CUDA(cudaMemcpyAsync(h->sm_cb.p, cbs.data(), n * 8, cudaMemcpyHostToDevice, st));

    // ---- (base, other) pairs inside every run, then their multiplicities
    h->sm_pcnt.reserve((size_t(n_e) + 1) * 4); h->sm_poff.reserve((size_t(n_e) + 1) * 4);
    k_pairs<false><<<grid_for(n_e, 256), 256, 0, st>>>(h->sm_ekey.as<uint64_t>(), n_e, rb, h->sm_ngenes.as<uint32_t>(), h->sm_pcnt.as<uint32_t>(), nullptr, nullptr);
    ++h->launches;
    DGE_CUDA(cudaMemsetAsync(h->sm_pcnt.as<uint32_t>() + n_e, 0, 4, st));
    device_exclusive_scan(h->sm_pcnt.as<uint32_t>(), h->sm_poff.as<uint32_t>(), size_t(n_e) + 1, h->scan_scratch.as<uint32_t>(), st, &h->launches);
    const uint32_t n_pairs = d2h_scalar<uint32_t>(h->sm_poff.as<uint32_t>() + n_e, st);
    tr.mark("merge:  simple: pair count");
    if (n_pairs == 0) return;
    h->sm_pairs.reserve(size_t(n_pairs) * 8); h->sm_pkey.reserve(size_t(n_pairs) * 8); h->sm_pval.reserve(size_t(n_pairs) * 4);
    k_pairs<true><<<grid_for(n_e, 256), 256, 0, st>>>(h->sm_ekey.as<uint64_t>(), n_e, rb, h->sm_ngenes.as<uint32_t>(), nullptr, h->sm_poff.as<uint32_t>(),
                                                      h->sm_pairs.as<uint64_t>());
    DGE_LAUNCH_CHECK();
    ++h->launches;
    const int kb2 = 2 * rb + 3;
    const uint32_t *n_p_ptr = h->sc2.run(h->sm_pairs.as<uint64_t>(), nullptr, n_pairs, kb2, std::min(choose_l1_bits(n_pairs), kb2 - 3), nullptr, h->sm_pairs.as<uint64_t>(),
                                         h->sm_pkey.as<uint64_t>(), h->sm_pval.as<uint32_t>(), h->overflow_flag.as<int>(), st, &h->sc_stats);
    const uint32_t n_p = d2h_scalar<uint32_t>(n_p_ptr, st);
    h->sc2.collect_timing();
    if (d2h_scalar<int>(h->overflow_flag.p, st)) throw std::runtime_error("sub-bucket hash table overflow while counting common UMI-genes");
    tr.mark("merge:  simple: common counts");

    if (poisson)
    {   // ---- PoissonSimpleMergeStrategy::get_merge_target (PoissonSimpleMergeStrategy.cpp:15-42)
        h->pp_flag.reserve(size_t(n_p) * 4); h->pp_prob.reserve(size_t(n_p) * 8); h->pp_best.reserve(n * sizeof(PoissonBest));
        k_pp_admissible<<<grid_for(n_p, 256), 256, 0, st>>>(h->sm_pkey.as<uint64_t>(), n_p, rb, h->sm_cb.as<uint64_t>(), int(h->cfg.cb_len),
                                                            int(h->cfg.max_cb_merge_edit_distance), h->pp_flag.as<uint32_t>());
        ++h->launches;
        poisson_eval_pairs(h, h->sm_pkey.as<uint64_t>(), h->sm_pval.as<uint32_t>(), h->pp_flag.as<uint32_t>(), n_p, rb, h->pp_prob.as<double>());
        tr.mark("merge:  poisson: pair probabilities");
        std::vector<PoissonBest> init(n, PoissonBest{NONE32, 0u, 2.0, 0u, 0u});
        DGE_CUDA(cudaMemcpyAsync(h->pp_best.p, init.data(), n * sizeof(PoissonBest), cudaMemcpyHostToDevice, st));
        // base (!= first neighbour) is never "real" here: the threshold is max_real_cb_merge_prob / #neighbours (PoissonTargetEstimator.cpp:17-22)
        k_pp_best<<<grid_for(n_p, 256), 256, 0, st>>>(h->sm_pkey.as<uint64_t>(), h->pp_flag.as<uint32_t>(), h->pp_prob.as<double>(), n_p, rb,
                                                      h->cfg.max_real_merge_prob, h->pp_best.as<PoissonBest>());
        DGE_LAUNCH_CHECK();
        ++h->launches;
        std::vector<PoissonBest> pb;
        d2h(pb, h->pp_best.p, n, st);
        DGE_CUDA(cudaStreamSynchronize(st));
        SimpleReplay R;
        std::vector<uint64_t> hk;
        std::vector<double> hp;
        std::vector<uint32_t> hf;
        h->n_poisson_replayed = 0;
        for (uint32_t i = 0; i < n; ++i)
        {
            const PoissonBest r = pb[i];
            if (r.n_nb == 0) continue; // no neighbour within the edit distance: the cell keeps itself (:33-34)
            const double thr = h->cfg.max_real_merge_prob / double(r.n_nb);
            if (!r.ambiguous)
            {
                if (!(r.min_prob > thr)) target[i] = long(r.best);
                continue;
            }
            // near-tie: walk the neighbours in the reference's iteration order with the device's probabilities
            if (!R.ready)
            {
                simple_replay_prepare(h, R, n_e);
                d2h(hk, h->sm_pkey.p, n_p, st); d2h(hp, h->pp_prob.p, n_p, st); d2h(hf, h->pp_flag.p, n_p, st);
                DGE_CUDA(cudaStreamSynchronize(st));
            }
            ++h->n_poisson_replayed;
            const uint64_t lo_key = uint64_t(i) << rb;
            const size_t lo = size_t(std::lower_bound(hk.begin(), hk.end(), lo_key) - hk.begin());
            long best = -1;
            double min_prob = 2;
            for (auto const &kv : simple_replay_candidates(h, R, i, rb))
            {
                size_t q = lo;
                while (q < hk.size() && (hk[q] >> rb) == i && uint32_t(hk[q] & ((1ull << rb) - 1)) != kv.first) ++q;
                if (q >= hk.size() || (hk[q] >> rb) != i || !hf[q]) continue; // beyond the edit distance
                if (hp[q] < min_prob) { min_prob = hp[q]; best = long(kv.first); }
            }
            if (best >= 0 && !(min_prob > thr)) target[i] = best;
        }
        tr.mark("merge:  poisson: targets");
        return;
    }

    // ---- best candidate per base
    h->sm_frac.reserve(size_t(n_p) * 8); h->sm_best.reserve(n * sizeof(BaseBest));
    DGE_CUDA(cudaMemsetAsync(h->sm_best.p, 0xFF, n * sizeof(BaseBest), st)); // best = NONE32: no candidate
    k_pair_eval<<<grid_for(n_p, 256), 256, 0, st>>>(h->sm_pkey.as<uint64_t>(), h->sm_pval.as<uint32_t>(), n_p, rb, h->sm_cb.as<uint64_t>(), h->sm_umis.as<uint32_t>(),
                                                    int(h->cfg.cb_len), int(h->cfg.max_cb_merge_edit_distance), h->sm_frac.as<double>());
    k_base_best<<<grid_for(n_p, 256), 256, 0, st>>>(h->sm_pkey.as<uint64_t>(), h->sm_pval.as<uint32_t>(), h->sm_frac.as<double>(), n_p, rb, 2e-5, h->sm_best.as<BaseBest>());
    DGE_LAUNCH_CHECK();
    h->launches += 2;
    const BaseBest *bb = d2h_pinned<BaseBest>(h->pin_best, h->sm_best.p, n, st);
    DGE_CUDA(cudaStreamSynchronize(st));
    SimpleReplay R;
    for (uint32_t i = 0; i < n; ++i)
    {
        const BaseBest r = bb[i];
        if (r.best == NONE32) continue; // no admissible candidate: top fraction stays -1 < min_merge_fraction -> the cell keeps itself
        if (!r.ambiguous)
        {
            const double frac = 0.5 * size_t(r.count) * (1. / size_t(h->real[i].umis_stat) + 1. / size_t(h->real[r.best].umis_stat));
            if (!(frac < h->cfg.min_merge_fraction)) target[i] = long(r.best);
            continue;
        }
        if (!R.ready) simple_replay_prepare(h, R, n_e);
        target[i] = simple_replay(h, R, i, rb);
        ++h->n_simple_replayed;
    }
    tr.mark("merge:  simple: targets");
}

// PoissonRealBarcodesMergeStrategy (Merge/PoissonRealBarcodesMergeStrategy.cpp:20-51): neighbours from the whitelist walk with
// get_max_merge_dist = (min == 0 ? 2 : min + 1), i.e. always beyond the nearest class -- the exact enumerator of whitelist.hpp on the
// host (this strategy is used on small inputs; the reference spends ~2 ms per cell in the same enumeration), intersections and
// probabilities on the device, PoissonTargetEstimator::get_best_merge_target (.cpp:14-44) over the ordered neighbour lists.
void phase1_poisson_real(dge_handle *h, std::vector<long> &target)
{
    cudaStream_t st = h->stream;
    const size_t n = h->real.size();
    target.assign(n, -1);
    h->n_unresolved = 0;
    if (n == 0) return;
    const int rb = std::max(1, ceil_log2_u64(n));
    std::unordered_map<uint64_t, uint32_t> by_cb;
    by_cb.reserve(n * 2);
    for (uint32_t r = 0; r < n; ++r) by_cb.emplace(h->real[r].cb, r);
    std::vector<std::vector<long>> nbs(n);
    std::vector<uint64_t> pkey;
    std::vector<PairJob> jobs;
    const uint32_t empty_pc = h->n_pc;
    for (uint32_t i = 0; i < n; ++i)
    {
        auto lookup = [&](const std::string &sq) -> long {
            uint64_t v;
            if (!pack_seq(sq, v)) return -1;
            auto it = by_cb.find(v);
            return it == by_cb.end() ? -1 : long(it->second);
        };
        auto eligible = [&](long id) {
            return uint32_t(h->real[size_t(id)].n_genes) >= h->cfg.min_genes_before_merge && h->real[size_t(id)].umis_stat >= h->real[i].umis_stat;
        };
        nbs[i] = h->wl.neighbours(cb_string(h, h->real[i].cb), true, lookup, eligible);
        for (long id : nbs[i])
        {
            if (uint32_t(id) == i) continue;
            pkey.push_back((uint64_t(i) << rb) | uint64_t(id));
            const uint32_t a = h->real[i].pc, b = h->real[size_t(id)].pc;
            jobs.push_back(PairJob{a == NONE32 ? empty_pc : a, b == NONE32 ? empty_pc : b});
        }
    }
    std::vector<double> prob(pkey.size(), 2.0);
    if (!pkey.empty())
    {
        run_intersections(h, jobs);
        const uint32_t n_p = uint32_t(pkey.size());
        h->sm_pkey.reserve(size_t(n_p) * 8); h->sm_pval.reserve(size_t(n_p) * 4); h->pp_flag.reserve(size_t(n_p) * 4); h->pp_prob.reserve(size_t(n_p) * 8);
        DGE_CUDA(cudaMemcpyAsync(h->sm_pkey.p, pkey.data(), size_t(n_p) * 8, cudaMemcpyHostToDevice, st));
        DGE_CUDA(cudaMemcpyAsync(h->sm_pval.p, h->d_isect.p, size_t(n_p) * 4, cudaMemcpyDeviceToDevice, st));
        k_fill_u32<<<grid_for(n_p, 256), 256, 0, st>>>(h->pp_flag.as<uint32_t>(), n_p, 1u);
        poisson_eval_pairs(h, h->sm_pkey.as<uint64_t>(), h->sm_pval.as<uint32_t>(), h->pp_flag.as<uint32_t>(), n_p, rb, h->pp_prob.as<double>());
        DGE_CUDA(cudaMemcpyAsync(prob.data(), h->pp_prob.p, size_t(n_p) * 8, cudaMemcpyDeviceToHost, st));
        DGE_CUDA(cudaStreamSynchronize(st));
    }
    size_t q = 0;
    for (uint32_t i = 0; i < n; ++i)
    {
        const std::vector<long> &lst = nbs[i];
        if (lst.empty()) { target[i] = -1; continue; } // RealBarcodesMergeStrategy.cpp:26-27
        const bool base_is_real_cb = uint32_t(lst[0]) == i;
        double max_prob = (base_is_real_cb ? h->cfg.max_merge_prob : h->cfg.max_real_merge_prob) / double(lst.size());
        long best = -1;
        double min_prob = 2;
        for (long id : lst)
        {
            if (uint32_t(id) == i) continue;
            const double pr = prob[q++];
            if (pr < min_prob) { min_prob = pr; best = id; }
        }
        if (min_prob > max_prob) target[i] = base_is_real_cb ? long(i) : -1;
        else target[i] = best;
    }
}

// MergeAllMergeStrategy::get_merge_target for every filtered cell (MergeAllMergeStrategy.h:16-50): all-pairs on the device.
void phase1_all(dge_handle *h, std::vector<long> &target)
{
    cudaStream_t st = h->stream;
    const size_t n = h->real.size();
    target.resize(n);
    for (size_t i = 0; i < n; ++i) target[i] = long(i);
    const std::vector<uint32_t> &f = h->filtered;
    const size_t m = f.size();
    if (m < 2) return;
    if (m >= (size_t(1) << 27)) throw std::runtime_error("merge_type=all: too many filtered cells");
    if (h->cfg.max_cb_merge_edit_distance > 63) throw std::runtime_error("merge_type=all: max_cb_merge_edit_distance must be < 64");
    std::vector<uint64_t> cbs(m);
    std::vector<uint32_t> um(m), tg(m);
    for (size_t k = 0; k < m; ++k) { cbs[k] = h->real[f[k]].cb; um[k] = uint32_t(h->real[f[k]].umis_stat); }
    h->sm_cb.reserve(m * 8); h->sm_umis.reserve(m * 4); h->sm_ngenes.reserve(m * 4);
    DGE_CUDA(cudaMemcpyAsync(h->sm_cb.p, cbs.data(), m * 8, cudaMemcpyHostToDevice, st));
    DGE_CUDA(cudaMemcpyAsync(h->sm_umis.p, um.data(), m * 4, cudaMemcpyHostToDevice, st));
    k_merge_all_targets<<<unsigned(std::min<size_t>(m, 148 * 16)), 256, 0, st>>>(h->sm_cb.as<uint64_t>(), h->sm_umis.as<uint32_t>(), uint32_t(m), int(h->cfg.cb_len),
                                                                                 h->cfg.max_cb_merge_edit_distance, h->sm_ngenes.as<uint32_t>());
    DGE_LAUNCH_CHECK();
    ++h->launches;
    DGE_CUDA(cudaMemcpyAsync(tg.data(), h->sm_ngenes.p, m * 4, cudaMemcpyDeviceToHost, st));
    DGE_CUDA(cudaStreamSynchronize(st));
    for (size_t k = 0; k < m; ++k) target[f[k]] = long(f[tg[k]]);
}

// Phase 2: MergeStrategyBase::merge_inited second loop + reassign (MergeStrategyBase.cpp:29-82), on real-cell indices.
// `reassigned_to` sets are intrusive singly linked lists (child_head/child_next): a cell sits in at most one list.
void phase2(dge_handle *h, const std::vector<long> &target)
{
    const size_t n = h->real.size();
    // the loop walks cells in size order, i.e. randomly in cell-id order: keep what it touches in a compact 16-byte row
    // ... and, in the same walk, the move jobs of apply_merges: merge_cells copies the source's CURRENT content, i.e. its own UMIs plus
    // everything merged into it earlier -- exactly the cells re-pointed at it so far (the child lists of `reassign`).
    struct Row { int32_t umis, reads; uint32_t intergenic, reassign, pc, slot, n_umis, pad; };
    std::vector<Row> row(n);
    std::vector<uint32_t> child_head(n, NONE32), child_tail(n, NONE32), child_next(n, NONE32);
    for (uint32_t i = 0; i < n; ++i)
    {
        const HostCell &c = h->real[i];
        row[i] = Row{c.umis_stat, c.reads_stat, c.n_intergenic, i, c.pc, c.slot, uint32_t(c.n_umis_distinct), 0};
    }
    h->merge_events.clear();
    h->merge_events.reserve(h->filtered.size());
    h->n_merged = h->n_excluded = 0;
    std::vector<MoveJob> &jobs = h->h_moves;
    jobs.clear();
    jobs.reserve(h->filtered.size());
    uint64_t total = 0;
    auto emit = [&](uint32_t o, uint32_t dst) {
        const Row &src = row[o];
        if (src.pc == NONE32 || src.n_umis == 0) return;
        jobs.push_back(MoveJob{src.pc, row[dst].slot, uint32_t(total)});
        total += uint64_t(src.n_umis);
    };
    auto append = [&](uint32_t t, uint32_t c) {
        child_next[c] = NONE32;
        if (child_head[t] == NONE32) child_head[t] = c; else child_next[child_tail[t]] = c;
        child_tail[t] = c;
    };
    const std::vector<uint32_t> &F = h->filtered;
    for (size_t fi = 0; fi < F.size(); ++fi)
    {
        const uint32_t base = F[fi];
        long t = target[base];
        if (t < 0) { h->real[base].excluded = true; ++h->n_excluded; continue; }
        if (uint32_t(t) == base) continue;              // keeps itself (reassign[base] == base: nothing was merged into a merged cell yet)
        t = long(row[size_t(t)].reassign);
        if (uint32_t(t) == base) continue;
        // merge_cells (CellsDataContainer.cpp:90-104): Stats::merge adds every counter (Stats.cpp:29-43)
        Row &src = row[base];
        Row &dst = row[size_t(t)];
        dst.umis += src.umis; dst.reads += src.reads; dst.intergenic += src.intergenic;
        h->merge_events.emplace_back(base, uint32_t(t));
        ++h->n_merged;
        emit(base, uint32_t(t));
        // reassign: base and everything previously re-pointed at base now point at t
        src.reassign = uint32_t(t);
        uint32_t c = child_head[base];
        child_head[base] = child_tail[base] = NONE32;
        append(uint32_t(t), base);
        while (c != NONE32)
        {
            uint32_t nx = child_next[c];
            row[c].reassign = uint32_t(t);
            emit(c, uint32_t(t));
            append(uint32_t(t), c);
            c = nx;
        }
        if (total >= 0xFFFFFFF0ull) throw std::runtime_error("merge volume exceeds 2^32 entries");
    }
    h->moves_total = total;
    h->moves_ready = true;
    for (uint32_t i = 0; i < n; ++i)
    {
        HostCell &c = h->real[i];
        c.umis_stat = row[i].umis; c.reads_stat = row[i].reads; c.n_intergenic = row[i].intergenic;
        c.target = int32_t(row[i].reassign);
        if (row[i].reassign != i) c.merged = true;
    }
}

// Apply the recorded merges to the device lists.
void apply_moved(dge_handle *h, uint64_t total);

void apply_merges(dge_handle *h)
{
    if (h->merge_events.empty()) return;
    cudaStream_t st = h->stream;
    Tracer tr;
    tr.st = st;
    // the move jobs (source cell -> destination slot, with the sequential merge_cells semantics) were produced by phase2
    if (!h->moves_ready) throw std::runtime_error("apply_merges without phase 2");
    std::vector<MoveJob> &jobs = h->h_moves;
    const uint64_t total = h->moves_total;
    h->moves_ready = false;
    if (jobs.empty()) return;
    tr.mark("  apply: host job list");
    h->d_moves.reserve(jobs.size() * sizeof(MoveJob));
    h->mkeys.reserve(total * 8); h->mvals.reserve(total * 4);
    DGE_CUDA(cudaMemcpyAsync(h->d_moves.p, jobs.data(), jobs.size() * sizeof(MoveJob), cudaMemcpyHostToDevice, st));
    k_gather_relabel<<<grid_for(jobs.size(), 1, 148 * 16), 256, 0, st>>>(h->d_moves.as<MoveJob>(), uint32_t(jobs.size()), h->ukey.as<uint64_t>(),
                                                                        h->uval.as<uint32_t>(), h->pc_u_start.as<uint32_t>(), h->kl.gb + h->kl.ub,
                                                                        h->mkeys.as<uint64_t>(), h->mvals.as<uint32_t>());
    DGE_LAUNCH_CHECK();
    ++h->launches;
    apply_moved(h, total);
}

// h->mkeys / h->mvals hold `total` re-labelled (target cell, gene, umi) entries: combine them and fold them into U
// (Gene::merge, Gene.cpp:26-36: existing UMIs get counts added and marks ORed, new UMIs are inserted), then rebuild the tables.
void apply_moved(dge_handle *h, uint64_t total)
{
    cudaStream_t st = h->stream;
    Tracer tr;
    tr.st = st;
    const int gub = h->kl.gb + h->kl.ub;
    h->ekey.reserve(total * 8); h->eval.reserve(total * 4);
    const int l1_bits = std::min(choose_l1_bits(total), h->kl.kb - 3);
    const uint32_t *n_e_ptr = h->sc2.run(h->mkeys.as<uint64_t>(), h->mvals.as<uint32_t>(), total, h->kl.kb, l1_bits, nullptr, h->mkeys.as<uint64_t>(),
                                         h->ekey.as<uint64_t>(), h->eval.as<uint32_t>(), h->overflow_flag.as<int>(), st, &h->sc_stats);
    const uint32_t n_e = d2h_scalar<uint32_t>(n_e_ptr, st);
    h->sc2.collect_timing();
    tr.mark("  apply: gather + combine");
    if (d2h_scalar<int>(h->overflow_flag.p, st)) throw std::runtime_error("sub-bucket hash table overflow while merging cells");
    h->keep.reserve(size_t(n_e + 1) * 4); h->keep_off.reserve(size_t(n_e + 1) * 4);
    k_probe_merge<<<grid_for(n_e, 256), 256, 0, st>>>(h->ekey.as<uint64_t>(), h->eval.as<uint32_t>(), n_e, h->ukey.as<uint64_t>(), h->uval.as<uint32_t>(),
                                                       h->slot_pc.as<uint32_t>(), h->pc_u_start.as<uint32_t>(), gub, h->keep.as<uint32_t>());
    ++h->launches;
    const uint32_t *n_x_ptr = device_exclusive_scan(h->keep.as<uint32_t>(), h->keep_off.as<uint32_t>(), n_e, h->scan_scratch.as<uint32_t>(), st, &h->launches);
    const uint32_t n_x = d2h_scalar<uint32_t>(n_x_ptr, st);
    if (n_x)
    {
        if (uint64_t(h->n_u) + n_x >= 0xFFFFFFF0ull) throw std::runtime_error("too many distinct UMIs after merging");
        h->xkey.reserve(size_t(n_x) * 8); h->xval.reserve(size_t(n_x) * 4);
        k_compact_keep<<<grid_for(n_e, 256), 256, 0, st>>>(h->ekey.as<uint64_t>(), h->eval.as<uint32_t>(), n_e, h->keep.as<uint32_t>(), h->keep_off.as<uint32_t>(),
                                                            h->xkey.as<uint64_t>(), h->xval.as<uint32_t>());
        const size_t n_new = size_t(h->n_u) + n_x;
        h->ukey2.reserve(n_new * 8); h->uval2.reserve(n_new * 4);
        k_merge_rank<<<148 * 8, 256, 0, st>>>(h->ukey.as<uint64_t>(), h->uval.as<uint32_t>(), h->n_u, h->xkey.as<uint64_t>(), h->xval.as<uint32_t>(), n_x,
                                                            h->ukey2.as<uint64_t>(), h->uval2.as<uint32_t>());
        DGE_LAUNCH_CHECK();
        h->launches += 2;
        std::swap(h->ukey.p, h->ukey2.p); std::swap(h->ukey.bytes, h->ukey2.bytes);
        std::swap(h->uval.p, h->uval2.p); std::swap(h->uval.bytes, h->uval2.bytes);
        h->n_u = uint32_t(n_new);
    }
    tr.mark("  apply: probe + merge_rank");
    build_segments(h);
    tr.mark("  apply: segments");
}

// columns = present-cell indices in DEVICE memory
void build_matrix_cols(dge_handle *h, MatrixDev &m, const uint32_t *d_cols, size_t n_cols, bool filtered, const uint32_t *values_override = nullptr);

void build_matrix(dge_handle *h, MatrixDev &m, const std::vector<uint32_t> &col_pcs, bool filtered)
{
    m.cols.reserve(std::max<size_t>(col_pcs.size(), 1) * 4);
    if (!col_pcs.empty()) DGE_CUDA(cudaMemcpyAsync(m.cols.p, col_pcs.data(), col_pcs.size() * 4, cudaMemcpyHostToDevice, h->stream));
    build_matrix_cols(h, m, m.cols.as<uint32_t>(), col_pcs.size(), filtered);
}

void build_matrix_cols(dge_handle *h, MatrixDev &m, const uint32_t *d_cols, size_t n_cols, bool filtered, const uint32_t *values_override)
{
    cudaStream_t st = h->stream;
    m.n_cols = n_cols;
    m.nnz = 0;
    m.built = true;
    m.cols_used = d_cols;
    m.indptr.reserve((m.n_cols + 2) * 4);
    if (m.n_cols == 0) { DGE_CUDA(cudaMemsetAsync(m.indptr.p, 0, 8, st)); return; }
    h->mat_nnz.reserve((m.n_cols + 2) * 4);
    k_matrix_col_nnz<<<grid_for(m.n_cols * 32, 256), 256, 0, st>>>(d_cols, uint32_t(m.n_cols), h->pc_cg_start.as<uint32_t>(),
                                                                   values_override ? values_override : h->cg_req.as<uint32_t>(), filtered ? 1 : 0, h->mat_nnz.as<uint32_t>());
    ++h->launches;
    DGE_CUDA(cudaMemsetAsync(h->mat_nnz.as<uint32_t>() + m.n_cols, 0, 4, st));
    device_exclusive_scan(h->mat_nnz.as<uint32_t>(), m.indptr.as<uint32_t>(), m.n_cols + 1, h->scan_scratch.as<uint32_t>(), st, &h->launches);
    m.nnz = d2h_scalar<uint32_t>(m.indptr.as<uint32_t>() + m.n_cols, st);
    m.gene.reserve(std::max<size_t>(m.nnz, 1) * 4); m.val.reserve(std::max<size_t>(m.nnz, 1) * 4);
    const uint32_t *values;
    int mode;
    if (values_override) { values = values_override; mode = 0; }
    else if (filtered) { values = h->cfg.reads_output ? h->cg_req_reads.as<uint32_t>() : h->cg_req.as<uint32_t>(); mode = 0; }
    else if (h->cfg.reads_output) { values = h->cg_reads.as<uint32_t>(); mode = 2; }
    else { values = nullptr; mode = 1; }
    k_matrix_fill<<<unsigned(m.n_cols), 256, 0, st>>>(d_cols, uint32_t(m.n_cols), m.indptr.as<uint32_t>(), h->pc_cg_start.as<uint32_t>(),
                                                      h->cg_gene.as<uint32_t>(), values, h->cg_start.as<uint32_t>(), mode, m.gene.as<int32_t>(), m.val.as<int32_t>());
    DGE_LAUNCH_CHECK();
    ++h->launches;
}



// ---------------------------------------------------------------------------------------------------------------------
// Exact host replay of MergeUMIsStrategyDirectional::find_targets (MergeUMIsStrategyDirectional.cpp:57-116) for one segment:
// items arrive in U order (ascending UMI value); the reference lists them in UMI-id (first-seen) order and std::sort-s by reads.
// Returns, per item, the index of its root or NONE32.
struct UmiItem { uint32_t umi, reads, first, idx; };

void umi_directional_replay(const dge_handle *h, std::vector<UmiItem> &items, std::vector<uint32_t> &root)
{
    const unsigned max_ed = h->cfg.max_umi_merge_edit_distance;
    const double mult = h->cfg.umi_merge_mult;
    std::sort(items.begin(), items.end(), [](const UmiItem &a, const UmiItem &b) { return a.first < b.first; }); // distinct keys: a total order
    std::sort(items.begin(), items.end(), [](const UmiItem &a, const UmiItem &b) { return a.reads < b.reads; }); // same call as the reference
    const size_t n = items.size();
    std::vector<std::string> seq(n);
    for (size_t i = 0; i < n; ++i) seq[i] = umi_string(h, items[i].umi);
    std::vector<long> tgt(n, -1);
    for (size_t s = 0; s < n; ++s)
    {
        unsigned min_ed = std::numeric_limits<unsigned>::max();
        for (long d = long(n) - 1; d > long(s); --d)
        {
            if (double(items[s].reads) * mult > double(items[size_t(d)].reads)) break;
            const unsigned ed = edit_distance_ref(seq[s].c_str(), seq[size_t(d)].c_str(), true, max_ed);
            if (ed > max_ed) continue;
            if (ed < min_ed)
            {
                tgt[s] = d;
                if (ed <= 1) break;
                min_ed = ed;
            }
        }
    }
    root.assign(n, NONE32);
    for (long i = long(n) - 1; i >= 0; --i)
    {
        if (tgt[size_t(i)] < 0) continue;
        long r = tgt[size_t(i)];
        while (tgt[size_t(r)] >= 0) r = tgt[size_t(r)];
        root[size_t(i)] = uint32_t(r);
    }
}

// A (cell, gene) segment that holds UMIs with N under MergeUMIsStrategyDirectional: the reference's own sequence of operations, literally,
// on the strings (MergeUMIsStrategyDirectional.cpp:57-116: std::sort by reads, find_target with the N rules -- a source with N only stops at
// distance 0 and, without any target, is renamed by fix_n_umi_with_random, i.e. the process-wide rand() --, one hop of path compression
// through the unordered_map; Cell::merge_umis / Gene::merge, Cell.cpp:31-42, Gene.cpp:38-58: the map is walked in ITS order, a target that
// does not exist is created, TOTAL_UMIS_PER_CB drops by one per entry).  Returns the segment's final content as (UMI code, count | mark) and
// the number of applied entries.  `items` come with the first-seen read index of their UMI, which orders the UMI ids of the reference.
void umi_directional_literal(const dge_handle *h, std::vector<UmiItem> &items, const std::unordered_map<std::string, uint32_t> &n_index,
                             std::vector<std::pair<uint32_t, uint32_t>> &final_entries, uint32_t &n_applied,
                             std::vector<std::array<uint32_t, 3>> *applied_targets = nullptr)
{
    auto code_of = [&](const std::string &seq) -> uint32_t { // internal UMI field of a string
        if (seq.find('N') != std::string::npos)
        {
            auto it = n_index.find(seq);
            if (it == n_index.end()) throw std::runtime_error("internal: a UMI with N that is not in the N-UMI list survived the directional merge");
            return (1u << (h->kl.ub - 1)) | it->second;
        }
        uint64_t packed = 0;
        pack_seq(seq, packed);
        return uint32_t(packed);
    };
    struct Wrap { std::string sequence; size_t n_reads; };
    const unsigned max_ed = h->cfg.max_umi_merge_edit_distance;
    const double mult = h->cfg.umi_merge_mult;
    std::sort(items.begin(), items.end(), [](const UmiItem &a, const UmiItem &b) { return a.first < b.first; }); // Gene::umis(): map over UMI ids
    std::map<std::string, std::pair<uint32_t, uint32_t>> content; // sequence -> (reads, mark)
    std::vector<Wrap> umis;
    for (auto const &it : items)
    {
        const std::string seq = umi_string(h, it.umi);
        umis.push_back(Wrap{seq, size_t(it.reads & VAL_COUNT_MASK)});
        content.emplace(seq, std::make_pair(it.reads & VAL_COUNT_MASK, it.reads >> VAL_MARK_SHIFT));
    }
    std::sort(umis.begin(), umis.end(), [](const Wrap &u1, const Wrap &u2) { return u1.n_reads < u2.n_reads; }); // the same call as the reference
    std::unordered_map<std::string, std::string> merge_targets;
    for (size_t src = 0; src < umis.size(); ++src)
    {
        const bool has_ns = umis[src].sequence.find('N') != std::string::npos;
        std::string target;
        unsigned min_ed = std::numeric_limits<unsigned>::max();
        for (long dst = long(umis.size()) - 1; dst > long(src); --dst)
        {
            if (double(umis[src].n_reads) * mult > double(umis[size_t(dst)].n_reads)) break;
            const unsigned ed = edit_distance_ref(umis[src].sequence.c_str(), umis[size_t(dst)].sequence.c_str(), true, max_ed);
            if (ed > max_ed) continue;
            if (ed < min_ed)
            {
                target = umis[size_t(dst)].sequence;
                if ((!has_ns && ed <= 1) || ed == 0) break;
                min_ed = ed;
            }
        }
        if (has_ns && target.empty())
        {   // MergeUMIsStrategyAbstract::fix_n_umi_with_random
            target = umis[src].sequence;
            for (char &c : target) if (c == 'N') c = "ACGT"[size_t(rand()) % 4];
        }
        if (!target.empty()) merge_targets[umis[src].sequence] = target;
    }
    for (long i = long(umis.size()) - 1; i >= 0; --i)
    {
        auto d = merge_targets.find(umis[size_t(i)].sequence);
        if (d == merge_targets.end()) continue;
        d = merge_targets.find(d->second);
        if (d == merge_targets.end()) continue;
        merge_targets[umis[size_t(i)].sequence] = d->second;
    }
    n_applied = 0;
    for (auto const &t : merge_targets)
    {
        if (t.second == t.first) continue;
        auto s = content.find(t.first);
        if (s == content.end()) throw std::runtime_error("Source UMI doesn't belong to the gene: " + t.first); // Gene.cpp:45
        auto ins = content.emplace(t.second, s->second);
        if (!ins.second) { ins.first->second.first += s->second.first; ins.first->second.second |= s->second.second; }
        content.erase(s);
        ++n_applied;
        if (applied_targets) applied_targets->push_back({code_of(t.first), code_of(t.second), ins.second ? 1u : 0u}); // _merge_targets[source_umi] = target_umi
    }
    final_entries.clear();
    for (auto const &c : content)
    {
        const uint32_t code = code_of(c.first);
        if (c.second.first > VAL_COUNT_MASK) throw std::runtime_error("UMI read count beyond the packed value");
        final_entries.emplace_back(code, c.second.first | (c.second.second << VAL_MARK_SHIFT));
    }
}

// MergeUMIsStrategyDirectional::merge over every (real cell, gene) segment.  Returns true when U changed.
bool umi_merge_directional(dge_handle *h)
{
    cudaStream_t st = h->stream;
    h->n_umis_merged = 0; h->n_umi_segments_replayed = 0;
    h->umi_mt.clear();
    const bool save_targets = h->cfg.save_umi_merge_targets != 0;
    if (h->n_cg == 0 || h->n_u == 0) return false;
    const uint32_t n_cg = h->n_cg, n_pc = h->n_pc;
    // real flags per present cell
    std::vector<uint32_t> real_pcs;
    for (auto const &c : h->real) if (c.real && c.pc != NONE32) real_pcs.push_back(c.pc);
    if (real_pcs.empty()) return false;
    h->flags.reserve((size_t(n_pc) + 2) * 4);
    DGE_CUDA(cudaMemsetAsync(h->flags.p, 0, (size_t(n_pc) + 2) * 4, st));
    h->misc.reserve(real_pcs.size() * 4);
    DGE_CUDA(cudaMemcpyAsync(h->misc.p, real_pcs.data(), real_pcs.size() * 4, cudaMemcpyHostToDevice, st));
    k_flag_list<<<grid_for(real_pcs.size(), 256, 1u << 30), 256, 0, st>>>(h->misc.as<uint32_t>(), uint32_t(real_pcs.size()), h->flags.as<uint32_t>());
    ++h->launches;

    const uint32_t list_cap = n_cg; // worst case: every segment deferred
    h->umi_lists.reserve(size_t(list_cap) * 2 * 4);
    h->umi_ctr.reserve(32);
    h->umi_pc_dec.reserve((size_t(n_pc) + 2) * 4);
    DGE_CUDA(cudaMemsetAsync(h->umi_ctr.p, 0, 32, st));
    DGE_CUDA(cudaMemsetAsync(h->umi_pc_dec.p, 0, (size_t(n_pc) + 2) * 4, st));
    UmiDirParams p{h->kl.ub, int(h->cfg.umi_len), h->cfg.max_umi_merge_edit_distance, h->cfg.umi_merge_mult, h->kl.ne ? h->kl.ub - 1 : -1};
    UmiDirOut o{};
    o.big_list = h->umi_lists.as<uint32_t>(); o.big_cap = list_cap;
    o.host_list = h->umi_lists.as<uint32_t>() + list_cap; o.host_cap = list_cap;
    o.big_count = h->umi_ctr.as<uint32_t>(); o.host_count = h->umi_ctr.as<uint32_t>() + 1;
    o.n_merged = reinterpret_cast<unsigned long long *>(h->umi_ctr.as<uint32_t>() + 2);
    o.pc_dec = h->umi_pc_dec.as<uint32_t>();
    if (save_targets)
    {
        h->u_target.reserve((size_t(h->n_u) + 1) * 4);
        DGE_CUDA(cudaMemsetAsync(h->u_target.p, 0xFF, (size_t(h->n_u) + 1) * 4, st));
        o.u_target = h->u_target.as<uint32_t>();
    }
    size_t n_host_pairs = 0;
    k_umi_dir_warp<<<grid_for(div_up(size_t(n_cg), size_t(32)), 8, 148 * 8), 256, 0, st>>>(h->ukey.as<uint64_t>(), h->uval.as<uint32_t>(), h->cg_start.as<uint32_t>(),
                                                                                          h->cg_pc.as<uint32_t>(), n_cg, h->flags.as<uint32_t>(),
                                                                                          h->umi_first.as<uint32_t>(), p, o);
    DGE_LAUNCH_CHECK();
    ++h->launches;
    uint32_t n_big = d2h_scalar<uint32_t>(o.big_count, st);
    if (n_big)
    {
        static bool attr_set[64] = {};
        int dev = 0; cudaGetDevice(&dev);
        const size_t smem = size_t(UMI_BLOCK_CAP) * 16;
        if (!attr_set[dev & 63]) { DGE_CUDA(cudaFuncSetAttribute(k_umi_dir_block, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem))); attr_set[dev & 63] = true; }
        k_umi_dir_block<<<std::min<uint32_t>(n_big, 148 * 4), 256, smem, st>>>(h->ukey.as<uint64_t>(), h->uval.as<uint32_t>(), h->cg_start.as<uint32_t>(),
                                                                                h->cg_pc.as<uint32_t>(), o.big_list, n_big, h->umi_first.as<uint32_t>(), p, o);
        DGE_LAUNCH_CHECK();
        ++h->launches;
    }
    const uint32_t n_host = d2h_scalar<uint32_t>(o.host_count, st);
    std::vector<uint32_t> host_dec_pc, host_dec_n;
    std::vector<uint32_t> dir_kill, dir_new_vals; // N segments: every old entry is dropped, the final content comes back as new keys
    std::vector<uint64_t> dir_new_keys;
    if (n_host)
    {   // exact replay of the segments whose outcome depends on how std::sort ordered equal read counts
        h->n_umi_segments_replayed = n_host;
        h->umi_seg.reserve(size_t(n_host + 1) * 4 * 4);
        uint32_t *seg_start = h->umi_seg.as<uint32_t>(), *seg_n = seg_start + (n_host + 1), *seg_pc = seg_n + (n_host + 1), *seg_off = seg_pc + (n_host + 1);
        k_umi_seg_meta<<<grid_for(n_host, 256, 1u << 30), 256, 0, st>>>(o.host_list, n_host, h->cg_start.as<uint32_t>(), h->cg_pc.as<uint32_t>(), seg_start, seg_n, seg_pc);
        DGE_CUDA(cudaMemsetAsync(seg_n + n_host, 0, 4, st));
        const uint32_t *tot = device_exclusive_scan(seg_n, seg_off, size_t(n_host) + 1, h->scan_scratch.as<uint32_t>(), st, &h->launches);
        (void)tot;
        std::vector<uint32_t> hs_start, hs_off, hs_pc, hs_gene;
        d2h(hs_start, seg_start, n_host, st); d2h(hs_off, seg_off, size_t(n_host) + 1, st); d2h(hs_pc, seg_pc, n_host, st);
        if (h->kl.ne)
        {   // gene of every deferred segment: the N segments are replayed in the reference's traversal order and get new keys
            h->misc.reserve(size_t(n_host + 1) * 4);
            k_gather_u32<<<grid_for(n_host, 256), 256, 0, st>>>(h->cg_gene.as<uint32_t>(), o.host_list, n_host, h->misc.as<uint32_t>());
            ++h->launches;
            d2h(hs_gene, h->misc.as<uint32_t>(), n_host, st);
        }
        DGE_CUDA(cudaStreamSynchronize(st));
        const size_t flat = hs_off[n_host];
        h->umi_flat.reserve(std::max<size_t>(flat, 1) * 3 * 4);
        uint32_t *f_umi = h->umi_flat.as<uint32_t>(), *f_val = f_umi + flat, *f_first = f_val + flat;
        k_umi_seg_gather<<<std::min<uint32_t>(n_host, 148 * 8), 128, 0, st>>>(seg_start, seg_off, n_host, h->ukey.as<uint64_t>(), h->uval.as<uint32_t>(),
                                                                               h->umi_first.as<uint32_t>(), h->kl.ub, f_umi, f_val, f_first);
        DGE_LAUNCH_CHECK();
        h->launches += 2;
        const uint32_t *hf = d2h_pinned<uint32_t>(h->pin_umi, f_umi, flat * 3, st);
        DGE_CUDA(cudaStreamSynchronize(st));
        std::vector<uint2> pairs;
        std::vector<UmiItem> items;
        std::vector<uint32_t> root;
        uint64_t merged_host = 0;
        std::vector<uint32_t> n_segments; // deferred segments that hold a UMI with N (they sort last: look at the last entry)
        for (uint32_t k = 0; k < n_host; ++k)
        {
            const uint32_t off = hs_off[k], n = hs_off[k + 1] - off;
            if (h->kl.ne && n && umi_is_n(h, hf[off + n - 1])) { n_segments.push_back(k); continue; }
            items.resize(n);
            for (uint32_t i = 0; i < n; ++i) items[i] = UmiItem{hf[off + i], hf[flat + off + i] & VAL_COUNT_MASK, hf[2 * flat + off + i], i};
            umi_directional_replay(h, items, root);
            uint32_t m = 0;
            for (uint32_t i = 0; i < n; ++i)
                if (root[i] != NONE32) { pairs.push_back(make_uint2(hs_start[k] + items[i].idx, hs_start[k] + items[root[i]].idx)); ++m; }
            if (m) { host_dec_pc.push_back(hs_pc[k]); host_dec_n.push_back(m); merged_host += m; }
        }
        if (!pairs.empty())
        {
            h->umi_pairs.reserve(pairs.size() * sizeof(uint2));
            DGE_CUDA(cudaMemcpyAsync(h->umi_pairs.p, pairs.data(), pairs.size() * sizeof(uint2), cudaMemcpyHostToDevice, st));
            for (int phase = 0; phase < 2; ++phase)
                k_umi_apply_pairs<<<grid_for(pairs.size(), 256, 1u << 30), 256, 0, st>>>(h->umi_pairs.as<uint2>(), uint32_t(pairs.size()), h->uval.as<uint32_t>(), phase, o.u_target);
            DGE_LAUNCH_CHECK();
            h->launches += 2;
            DGE_CUDA(cudaStreamSynchronize(st));
            n_host_pairs = pairs.size();
        }
        if (!n_segments.empty())
        {   // literal replay, in the order in which the reference walks cells and genes (that order decides who gets which random number)
            std::vector<uint32_t> pc_to_real(size_t(n_pc) + 2, NONE32);
            for (uint32_t i = 0; i < h->real.size(); ++i) if (h->real[i].pc != NONE32) pc_to_real[h->real[i].pc] = i;
            std::vector<uint32_t> gene_rank(h->cfg.n_genes, NONE32);
            for (size_t r = 0; r < h->gene_order.size(); ++r) gene_rank[size_t(h->gene_order[r])] = uint32_t(r);
            std::sort(n_segments.begin(), n_segments.end(), [&](uint32_t a, uint32_t b) {
                const uint32_t ca = pc_to_real[hs_pc[a]], cb = pc_to_real[hs_pc[b]];
                return ca != cb ? ca < cb : gene_rank[hs_gene[a]] < gene_rank[hs_gene[b]];
            });
            std::unordered_map<std::string, uint32_t> n_index;
            for (size_t k = 0; (k + 1) * h->cfg.umi_len <= h->n_umi_strings.size(); ++k) n_index.emplace(h->n_umi_strings.substr(k * h->cfg.umi_len, h->cfg.umi_len), uint32_t(k));
            srand(1); // the reference never seeds rand() on this path (only MergeUMIsStrategySimple's constructor does): the C default
            std::vector<std::pair<uint32_t, uint32_t>> final_entries;
            std::vector<std::array<uint32_t, 3>> applied_targets;
            for (uint32_t k : n_segments)
            {
                const uint32_t off = hs_off[k], n = hs_off[k + 1] - off;
                items.resize(n);
                for (uint32_t i = 0; i < n; ++i) items[i] = UmiItem{hf[off + i], hf[flat + off + i], hf[2 * flat + off + i], i};
                uint32_t applied = 0;
                applied_targets.clear();
                umi_directional_literal(h, items, n_index, final_entries, applied, save_targets ? &applied_targets : nullptr);
                if (!applied) continue;
                const HostCell &cell = h->real[pc_to_real[hs_pc[k]]];
                for (auto const &t : applied_targets)
                    h->umi_mt.push_back(dge_handle::UmiMergeTarget{cell.cb, hs_gene[k], umi_public_code(h, t[0]), umi_public_code(h, t[1]), t[2]});
                for (uint32_t i = 0; i < n; ++i) dir_kill.push_back(hs_start[k] + i); // the segment is rewritten as a whole
                for (auto const &fe : final_entries)
                {
                    dir_new_keys.push_back(((((uint64_t(cell.slot) << h->kl.gb) | hs_gene[k]) << h->kl.ub) | fe.first) << 3);
                    dir_new_vals.push_back(fe.second);
                }
                host_dec_pc.push_back(hs_pc[k]); host_dec_n.push_back(applied); merged_host += applied;
            }
            if (!dir_kill.empty())
            {
                h->misc.reserve(dir_kill.size() * 4);
                DGE_CUDA(cudaMemcpyAsync(h->misc.p, dir_kill.data(), dir_kill.size() * 4, cudaMemcpyHostToDevice, st));
                k_zero_list<<<grid_for(dir_kill.size(), 256), 256, 0, st>>>(h->misc.as<uint32_t>(), uint32_t(dir_kill.size()), h->uval.as<uint32_t>());
                DGE_LAUNCH_CHECK();
                ++h->launches;
                DGE_CUDA(cudaStreamSynchronize(st));
            }
        }
        h->n_umis_merged += merged_host;
    }
    const unsigned long long merged_dev = d2h_scalar<unsigned long long>(o.n_merged, st);
    h->n_umis_merged += merged_dev;
    if (h->n_umis_merged == 0) return false;
    if (save_targets && merged_dev + n_host_pairs)
    {   // the (source, root) decisions taken on the device and in the tie replays, before the sources leave U
        const size_t cap = size_t(merged_dev) + n_host_pairs;
        h->mt_keys.reserve(cap * 8); h->mt_dst.reserve(cap * 4);
        uint32_t *cnt = h->umi_ctr.as<uint32_t>() + 4;
        DGE_CUDA(cudaMemsetAsync(cnt, 0, 4, st));
        k_umi_targets_collect<<<grid_for(h->n_u, 256), 256, 0, st>>>(h->ukey.as<uint64_t>(), h->u_target.as<uint32_t>(), h->n_u, h->kl.ub, h->mt_keys.as<uint64_t>(),
                                                                      h->mt_dst.as<uint32_t>(), uint32_t(cap), cnt);
        DGE_LAUNCH_CHECK();
        ++h->launches;
        const uint32_t got = d2h_scalar<uint32_t>(cnt, st);
        if (got != cap) throw std::runtime_error("internal: UMI merge targets recorded on the device do not add up");
        std::vector<uint64_t> keys;
        std::vector<uint32_t> dsts;
        d2h(keys, h->mt_keys.p, cap, st); d2h(dsts, h->mt_dst.p, cap, st);
        DGE_CUDA(cudaStreamSynchronize(st));
        std::unordered_map<uint32_t, uint64_t> slot_cb;
        for (auto const &c : h->real) if (c.real && c.pc != NONE32) slot_cb.emplace(c.slot, c.cb);
        const int gub = h->kl.gb + h->kl.ub;
        const uint32_t umask = h->kl.ub >= 32 ? 0xFFFFFFFFu : ((1u << h->kl.ub) - 1);
        for (size_t k = 0; k < cap; ++k)
        {
            const uint32_t slot = uint32_t(keys[k] >> gub), gene = uint32_t(keys[k] >> h->kl.ub) & ((1u << h->kl.gb) - 1);
            h->umi_mt.push_back(dge_handle::UmiMergeTarget{slot_cb.at(slot), gene, umi_public_code(h, uint32_t(keys[k]) & umask), umi_public_code(h, dsts[k]), 0u}); // roots exist
        }
    }

    // TOTAL_UMIS_PER_CB decrements (Cell.cpp:39)
    {
        const uint32_t *dec = d2h_pinned<uint32_t>(h->pin_umi, h->umi_pc_dec.p, n_pc, st);
        DGE_CUDA(cudaStreamSynchronize(st));
        std::vector<uint32_t> dec_all(dec, dec + n_pc);
        for (size_t k = 0; k < host_dec_pc.size(); ++k) dec_all[host_dec_pc[k]] += host_dec_n[k];
        for (auto &c : h->real) if (c.real && c.pc != NONE32) c.umis_stat -= int32_t(dec_all[c.pc]);
    }
    // drop the merged entries and rebuild the segment tables (cells and genes stay, so PC / CG indices keep their meaning)
    h->keep.reserve(size_t(h->n_u + 1) * 4); h->keep_off.reserve(size_t(h->n_u + 1) * 4);
    k_flag_live<<<grid_for(h->n_u, 256), 256, 0, st>>>(h->uval.as<uint32_t>(), h->n_u, h->keep.as<uint32_t>());
    const uint32_t *n_live_ptr = device_exclusive_scan(h->keep.as<uint32_t>(), h->keep_off.as<uint32_t>(), h->n_u, h->scan_scratch.as<uint32_t>(), st, &h->launches);
    const uint32_t n_live = d2h_scalar<uint32_t>(n_live_ptr, st);
    h->ukey2.reserve(size_t(n_live + 1) * 8); h->uval2.reserve(size_t(n_live + 1) * 4);
    k_compact_keep<<<grid_for(h->n_u, 256), 256, 0, st>>>(h->ukey.as<uint64_t>(), h->uval.as<uint32_t>(), h->n_u, h->keep.as<uint32_t>(), h->keep_off.as<uint32_t>(),
                                                           h->ukey2.as<uint64_t>(), h->uval2.as<uint32_t>());
    DGE_LAUNCH_CHECK();
    h->launches += 2;
    std::swap(h->ukey.p, h->ukey2.p); std::swap(h->ukey.bytes, h->ukey2.bytes);
    std::swap(h->uval.p, h->uval2.p); std::swap(h->uval.bytes, h->uval2.bytes);
    h->n_u = n_live;
    build_segments(h);
    if (!dir_new_keys.empty())
    {
        build_slot_pc(h);
        const uint64_t total = dir_new_keys.size();
        h->mkeys.reserve(total * 8); h->mvals.reserve(total * 4);
        DGE_CUDA(cudaMemcpyAsync(h->mkeys.p, dir_new_keys.data(), total * 8, cudaMemcpyHostToDevice, st));
        DGE_CUDA(cudaMemcpyAsync(h->mvals.p, dir_new_vals.data(), total * 4, cudaMemcpyHostToDevice, st));
        DGE_CUDA(cudaStreamSynchronize(st));
        apply_moved(h, total);
    }
    return true;
}

// ---------------------------------------------------------------------------------------------------------------------
// merge_and_filter without host cell rows (DummyMergeStrategy / RealBarcodesMergeStrategy, default UMI strategy, single shard):
// phase 1 (p1_device_pass), phase 2 + merge_cells (k_phase2_*: every target is a whitelist barcode that keeps itself, so the
// sequential loop of MergeStrategyBase.cpp:29-51 is order-free), refreshed sizes, final filter in compare_cells order and the two
// matrices -- all from the device-resident CellRow / CellState tables.  Returns false (nothing modified) when a cell needs the
// exact host logic: a far distance class, an order-dependent tie, a merge chain, counters beyond the packed sort key.
// MergeUMIsStrategySimple::merge (Merge/UMIs/MergeUMIsStrategySimple.cpp:21-112): every UMI with N of a real cell goes to the N-free UMI of
// its (cell, gene) at the smallest Hamming distance (N matches anything; ties: more reads, then the smaller UMI id) when that is within
// max_umi_merge_edit_distance, else its N's are replaced by random bases (MergeUMIsStrategyAbstract.cpp:11-23: the process-wide rand(),
// seeded with 42 by the strategy's constructor).  The segments that hold an N-UMI are found on the device; the decisions replay the
// reference's own traversal on the host -- cells by id, genes by StringIndexer id, bad UMIs in the iteration order of an
// unordered_set<string> filled in UMI-id order -- because that order decides which random numbers a UMI gets.  Returns true when U changed.
bool umi_repair_n(dge_handle *h)
{
    if (!h->kl.ne || h->n_cg == 0 || h->n_u == 0) return false;
    cudaStream_t st = h->stream;
    const uint32_t n_cg = h->n_cg, n_pc = h->n_pc;
    std::vector<uint32_t> pc_real(size_t(n_pc) + 2, 0), pc_to_real(size_t(n_pc) + 2, NONE32);
    for (uint32_t i = 0; i < h->real.size(); ++i)
    {
        const HostCell &c = h->real[i];
        if (c.pc == NONE32) continue;
        pc_to_real[c.pc] = i;
        if (c.real) pc_real[c.pc] = 1;
    }
    h->flags.reserve(pc_real.size() * 4);
    DGE_CUDA(cudaMemcpyAsync(h->flags.p, pc_real.data(), pc_real.size() * 4, cudaMemcpyHostToDevice, st));
    h->umi_lists.reserve(size_t(n_cg) * 4 + 64); h->umi_ctr.reserve(32);
    DGE_CUDA(cudaMemsetAsync(h->umi_ctr.p, 0, 32, st));
    uint32_t *list = h->umi_lists.as<uint32_t>();
    k_seg_has_n<<<grid_for(n_cg, 256), 256, 0, st>>>(h->ukey.as<uint64_t>(), h->cg_start.as<uint32_t>(), h->cg_pc.as<uint32_t>(), n_cg, h->flags.as<uint32_t>(),
                                                     h->kl.ub, list, h->umi_ctr.as<uint32_t>());
    DGE_LAUNCH_CHECK();
    ++h->launches;
    const uint32_t n_seg = d2h_scalar<uint32_t>(h->umi_ctr.p, st);
    if (!n_seg) return false;
    // segment tables -> host
    h->umi_seg.reserve(size_t(n_seg + 1) * 5 * 4);
    uint32_t *seg_start = h->umi_seg.as<uint32_t>(), *seg_n = seg_start + (n_seg + 1), *seg_pc = seg_n + (n_seg + 1), *seg_off = seg_pc + (n_seg + 1),
             *seg_gene = seg_off + (n_seg + 1);
    k_umi_seg_meta<<<grid_for(n_seg, 256, 1u << 30), 256, 0, st>>>(list, n_seg, h->cg_start.as<uint32_t>(), h->cg_pc.as<uint32_t>(), seg_start, seg_n, seg_pc);
    k_gather_u32<<<grid_for(n_seg, 256), 256, 0, st>>>(h->cg_gene.as<uint32_t>(), list, n_seg, seg_gene);
    DGE_CUDA(cudaMemsetAsync(seg_n + n_seg, 0, 4, st));
    device_exclusive_scan(seg_n, seg_off, size_t(n_seg) + 1, h->scan_scratch.as<uint32_t>(), st, &h->launches);
    std::vector<uint32_t> hs_start, hs_off, hs_pc, hs_gene;
    d2h(hs_start, seg_start, n_seg, st); d2h(hs_off, seg_off, size_t(n_seg) + 1, st); d2h(hs_pc, seg_pc, n_seg, st); d2h(hs_gene, seg_gene, n_seg, st);
    DGE_CUDA(cudaStreamSynchronize(st));
    const size_t flat = hs_off[n_seg];
    h->umi_flat.reserve(std::max<size_t>(flat, 1) * 3 * 4);
    uint32_t *f_umi = h->umi_flat.as<uint32_t>(), *f_val = f_umi + flat, *f_first = f_val + flat;
    k_umi_seg_gather<<<std::min<uint32_t>(n_seg, 148 * 8), 128, 0, st>>>(seg_start, seg_off, n_seg, h->ukey.as<uint64_t>(), h->uval.as<uint32_t>(),
                                                                          h->umi_first.as<uint32_t>(), h->kl.ub, f_umi, f_val, f_first);
    DGE_LAUNCH_CHECK();
    h->launches += 3;
    const uint32_t *hf = d2h_pinned<uint32_t>(h->pin_umi, f_umi, flat * 3, st);
    DGE_CUDA(cudaStreamSynchronize(st));
    // the reference's traversal order: cells by id (= index in `real`), genes by StringIndexer id (first-seen rank)
    std::vector<uint32_t> gene_rank(h->cfg.n_genes, NONE32);
    for (size_t r = 0; r < h->gene_order.size(); ++r) gene_rank[size_t(h->gene_order[r])] = uint32_t(r);
    std::vector<uint32_t> order(n_seg);
    std::iota(order.begin(), order.end(), 0u);
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
        const uint32_t ca = pc_to_real[hs_pc[a]], cb = pc_to_real[hs_pc[b]];
        return ca != cb ? ca < cb : gene_rank[hs_gene[a]] < gene_rank[hs_gene[b]];
    });
    const unsigned max_ed = h->cfg.max_umi_merge_edit_distance;
    const int gub = h->kl.gb + h->kl.ub;
    std::vector<uint2> pairs;                // (source U index, existing target U index)
    std::vector<uint32_t> kill;              // sources whose target is a new UMI
    std::vector<uint64_t> new_keys;          // [slot | gene | new umi] << 3
    std::vector<uint32_t> new_vals;
    std::vector<uint32_t> dec_real;          // TOTAL_UMIS_PER_CB decrements (Cell.cpp:39), one per repaired UMI
    struct Item { uint32_t umi, val, first, idx; std::string seq; };
    std::vector<Item> items;
    srand(42); // MergeUMIsStrategySimple::MergeUMIsStrategySimple (MergeUMIsStrategySimple.cpp:15-19)
    for (uint32_t k : order)
    {
        const uint32_t off = hs_off[k], n = hs_off[k + 1] - off;
        items.resize(n);
        for (uint32_t i = 0; i < n; ++i)
        {
            items[i] = Item{hf[off + i], hf[flat + off + i], hf[2 * flat + off + i], i, std::string()};
            items[i].seq = umi_string(h, items[i].umi);
        }
        std::sort(items.begin(), items.end(), [](const Item &a, const Item &b) { return a.first < b.first; }); // Gene::umis(): std::map over UMI ids
        std::unordered_set<std::string> bad_umis; // s_hash_t
        for (auto const &it : items) if (it.seq.find('N') != std::string::npos) bad_umis.insert(it.seq);
        const uint32_t ridx = pc_to_real[hs_pc[k]];
        const HostCell &cell = h->real[ridx];
        std::unordered_map<std::string, std::string> seg_targets; // find_targets' result, filled in the same order (its iteration order decides below)
        const size_t mt_first = h->umi_mt.size();
        for (auto const &bad : bad_umis)
        {
            unsigned min_ed = std::numeric_limits<unsigned>::max();
            long best = -1, best_size = 0;
            for (size_t t = 0; t < items.size(); ++t)
            {
                if (bad_umis.find(items[t].seq) != bad_umis.end()) continue;
                const unsigned ed = hamming_distance_ref(items[t].seq, bad);
                const long reads = long(items[t].val & VAL_COUNT_MASK);
                if (ed < min_ed || (ed == min_ed && reads > best_size)) { min_ed = ed; best = long(t); best_size = reads; }
            }
            size_t src = 0;
            while (items[src].seq != bad) ++src;
            const uint32_t src_u = hs_start[k] + items[src].idx;
            if (best < 0 || min_ed > max_ed)
            {   // fix_n_umi_with_random
                std::string fixed(bad);
                for (char &c : fixed) if (c == 'N') c = "ACGT"[size_t(rand()) % 4];
                uint64_t packed = 0;
                pack_seq(fixed, packed);
                kill.push_back(src_u);
                new_keys.push_back(((((uint64_t(cell.slot) << h->kl.gb) | hs_gene[k]) << h->kl.ub) | packed) << 3);
                new_vals.push_back(items[src].val);
                if (h->cfg.save_umi_merge_targets)
                {
                    h->umi_mt.push_back(dge_handle::UmiMergeTarget{cell.cb, hs_gene[k], umi_public_code(h, items[src].umi), uint32_t(packed), 0u});
                    seg_targets[bad] = fixed;
                }
            }
            else
            {
                pairs.push_back(make_uint2(src_u, hs_start[k] + items[size_t(best)].idx));
                if (h->cfg.save_umi_merge_targets)
                {
                    h->umi_mt.push_back(dge_handle::UmiMergeTarget{cell.cb, hs_gene[k], umi_public_code(h, items[src].umi), umi_public_code(h, items[size_t(best)].umi), 0u});
                    seg_targets[bad] = items[size_t(best)].seq;
                }
            }
            dec_real.push_back(ridx);
        }
        if (h->umi_mt.size() > mt_first)
        {   // which source's UMI object BECAME its target (Gene::merge emplaces the target when it does not exist, Gene.cpp:48): Cell::merge_umis
            // walks the map in its own order (Cell.cpp:34), the first source that reaches a missing target creates it
            std::unordered_set<std::string> present;
            for (auto const &it : items) present.insert(it.seq);
            std::unordered_map<uint32_t, size_t> row_of_src;
            for (size_t r = mt_first; r < h->umi_mt.size(); ++r) row_of_src.emplace(h->umi_mt[r].src, r);
            for (auto const &t : seg_targets)
            {
                if (t.first == t.second) continue;
                size_t si = 0;
                while (items[si].seq != t.first) ++si;
                if (present.insert(t.second).second) h->umi_mt[row_of_src.at(umi_public_code(h, items[si].umi))].created = 1u;
                present.erase(t.first);
            }
        }
    }
    (void)gub;
    // apply: counts / marks into existing targets, sources removed, repaired UMIs that are new (or equal an existing one by chance) folded in
    if (!pairs.empty())
    {
        h->umi_pairs.reserve(pairs.size() * sizeof(uint2));
        DGE_CUDA(cudaMemcpyAsync(h->umi_pairs.p, pairs.data(), pairs.size() * sizeof(uint2), cudaMemcpyHostToDevice, st));
        for (int phase = 0; phase < 2; ++phase)
            k_umi_apply_pairs<<<grid_for(pairs.size(), 256, 1u << 30), 256, 0, st>>>(h->umi_pairs.as<uint2>(), uint32_t(pairs.size()), h->uval.as<uint32_t>(), phase);
    }
    if (!kill.empty())
    {
        h->misc.reserve(kill.size() * 4);
        DGE_CUDA(cudaMemcpyAsync(h->misc.p, kill.data(), kill.size() * 4, cudaMemcpyHostToDevice, st));
        k_zero_list<<<grid_for(kill.size(), 256), 256, 0, st>>>(h->misc.as<uint32_t>(), uint32_t(kill.size()), h->uval.as<uint32_t>());
    }
    DGE_LAUNCH_CHECK();
    DGE_CUDA(cudaStreamSynchronize(st));
    for (uint32_t r : dec_real) h->real[r].umis_stat -= 1;
    h->n_umis_merged += dec_real.size();
    h->keep.reserve(size_t(h->n_u + 1) * 4); h->keep_off.reserve(size_t(h->n_u + 1) * 4);
    k_flag_live<<<grid_for(h->n_u, 256), 256, 0, st>>>(h->uval.as<uint32_t>(), h->n_u, h->keep.as<uint32_t>());
    const uint32_t *n_live_ptr = device_exclusive_scan(h->keep.as<uint32_t>(), h->keep_off.as<uint32_t>(), h->n_u, h->scan_scratch.as<uint32_t>(), st, &h->launches);
    const uint32_t n_live = d2h_scalar<uint32_t>(n_live_ptr, st);
    h->ukey2.reserve(size_t(n_live + 1) * 8); h->uval2.reserve(size_t(n_live + 1) * 4);
    k_compact_keep<<<grid_for(h->n_u, 256), 256, 0, st>>>(h->ukey.as<uint64_t>(), h->uval.as<uint32_t>(), h->n_u, h->keep.as<uint32_t>(), h->keep_off.as<uint32_t>(),
                                                           h->ukey2.as<uint64_t>(), h->uval2.as<uint32_t>());
    DGE_LAUNCH_CHECK();
    h->launches += 2;
    std::swap(h->ukey.p, h->ukey2.p); std::swap(h->ukey.bytes, h->ukey2.bytes);
    std::swap(h->uval.p, h->uval2.p); std::swap(h->uval.bytes, h->uval2.bytes);
    h->n_u = n_live;
    build_segments(h);
    if (!new_keys.empty())
    {
        build_slot_pc(h);
        const uint64_t total = new_keys.size();
        h->mkeys.reserve(total * 8); h->mvals.reserve(total * 4);
        DGE_CUDA(cudaMemcpyAsync(h->mkeys.p, new_keys.data(), total * 8, cudaMemcpyHostToDevice, st));
        DGE_CUDA(cudaMemcpyAsync(h->mvals.p, new_vals.data(), total * 4, cudaMemcpyHostToDevice, st));
        DGE_CUDA(cudaStreamSynchronize(st));
        apply_moved(h, total);
    }
    return true;
}

void merge_device_finish(dge_handle *h);

bool merge_device_flow(dge_handle *h)
{
    cudaStream_t st = h->stream;
    Tracer tr; tr.st = st;
    const size_t n = h->n_real_rows;
    const uint32_t n32 = uint32_t(n);
    CellRow *rows = h->rows_dev2.as<CellRow>();
    h->cell_state.reserve(std::max<size_t>(n, 1) * sizeof(CellState));
    h->df_ctr.reserve(sizeof(DevFlowCounters));
    CellState *cs = h->cell_state.as<CellState>();
    DevFlowCounters *ctr = h->df_ctr.as<DevFlowCounters>();
    DGE_CUDA(cudaMemsetAsync(ctr, 0, sizeof(DevFlowCounters), st));
    const unsigned g = grid_for(n, 256);
    k_state_init<<<g, 256, 0, st>>>(rows, n32, cs, h->kl.ne ? ctr : nullptr);
    ++h->launches;
    DevFlowCounters hc{};
    const bool real_merge = h->cfg.merge_type == DGE_MERGE_REAL;
    if (h->kl.ne && !real_merge)
    {   // barcodes containing N among the real cells: their place in compare_cells ties is a string order -> host flow
        DGE_CUDA(cudaMemcpyAsync(&hc, ctr, sizeof(hc), cudaMemcpyDeviceToHost, st));
        DGE_CUDA(cudaStreamSynchronize(st));
        if (hc.n_todo) return false;
    }
    if (real_merge)
    {
        p1_device_pass(h, n);
        h->move_size.reserve((n + 1) * 4); h->move_off.reserve((n + 1) * 4);
        k_p1_finalize<<<g, 256, 0, st>>>(h->d_count.as<int>(), h->p1_target.as<int>(), h->p1_flag.as<uint32_t>(), n32, cs, ctr);
        k_phase2_sizes<<<g, 256, 0, st>>>(rows, cs, n32, h->move_size.as<uint32_t>(), ctr);
        DGE_CUDA(cudaMemsetAsync(h->move_size.as<uint32_t>() + n, 0, 4, st));
        device_exclusive_scan(h->move_size.as<uint32_t>(), h->move_off.as<uint32_t>(), n + 1, h->scan_scratch.as<uint32_t>(), st, &h->launches);
        h->launches += 2;
        uint32_t total = 0;
        DGE_CUDA(cudaMemcpyAsync(&hc, ctr, sizeof(hc), cudaMemcpyDeviceToHost, st));
        DGE_CUDA(cudaMemcpyAsync(&total, h->move_off.as<uint32_t>() + n, 4, cudaMemcpyDeviceToHost, st));
        DGE_CUDA(cudaStreamSynchronize(st));
        tr.mark("merge: phase 1 (device)");
        if (hc.n_todo || hc.n_chain) return false;
        h->d_moves.reserve(std::max<size_t>(n, 1) * sizeof(MoveJob));
        k_phase2_apply<<<g, 256, 0, st>>>(rows, cs, n32, h->move_off.as<uint32_t>(), h->n_pc, h->d_moves.as<MoveJob>(), ctr);
        ++h->launches;
        if (total)
        {
            h->mkeys.reserve(size_t(total) * 8); h->mvals.reserve(size_t(total) * 4);
            k_gather_relabel<<<grid_for(n, 1, 148 * 16), 256, 0, st>>>(h->d_moves.as<MoveJob>(), n32, h->ukey.as<uint64_t>(), h->uval.as<uint32_t>(),
                                                                      h->pc_u_start.as<uint32_t>(), h->kl.gb + h->kl.ub, h->mkeys.as<uint64_t>(), h->mvals.as<uint32_t>());
            DGE_LAUNCH_CHECK();
            ++h->launches;
            apply_moved(h, total);
        }
        tr.mark("merge: phase 2 + apply (device)");
    }
    DGE_CUDA(cudaEventRecord(h->ev[4], st));
    merge_device_finish(h);
    return true;
}

// Second half of the device flow (also the tail of the cross-rank merge, dge_dist_step): refreshed sizes of the merge targets,
// is_real, final filter in compare_cells order, cm / cm_raw.
void merge_device_finish(dge_handle *h)
{
    cudaStream_t st = h->stream;
    Tracer tr; tr.st = st;
    const size_t n = h->n_real_rows;
    const uint32_t n32 = uint32_t(n);
    CellRow *rows = h->rows_dev2.as<CellRow>();
    CellState *cs = h->cell_state.as<CellState>();
    DevFlowCounters *ctr = h->df_ctr.as<DevFlowCounters>();
    DevFlowCounters hc{};
    const unsigned g = grid_for(n, 256);
    // ---- refreshed sizes, is_real, final filter (update_filtered_gene_counts, CellsDataContainer.cpp:250-276)
    h->real_flag.reserve((n + 1) * 4); h->real_off.reserve((n + 1) * 4);
    k_refresh_rows<<<g, 256, 0, st>>>(rows, cs, n32, h->pc_cg_start.as<uint32_t>(), h->pc_u_start.as<uint32_t>(), h->pc_req_genes.as<uint32_t>(),
                                      h->pc_req_umis.as<uint32_t>(), h->cfg.min_genes_before_merge, h->real_flag.as<uint32_t>());
    DGE_CUDA(cudaMemsetAsync(h->real_flag.as<uint32_t>() + n, 0, 4, st));
    device_exclusive_scan(h->real_flag.as<uint32_t>(), h->real_off.as<uint32_t>(), n + 1, h->scan_scratch.as<uint32_t>(), st, &h->launches);
    for (int b = 0; b < 2; ++b) { h->fsort_k[b].reserve(std::max<size_t>(n, 1) * 8); h->fsort_v[b].reserve(std::max<size_t>(n, 1) * 4); }
    // stable sort by barcode, then stable sort by the packed (genes, umis, stat) counters; cells outside the filter sort to the end
    k_rows_cb_keys<<<g, 256, 0, st>>>(rows, n32, h->fsort_k[0].as<uint64_t>(), h->fsort_v[0].as<uint32_t>());
    device_sort_pairs(h, h->fsort_k[0].as<uint64_t>(), h->fsort_k[1].as<uint64_t>(), h->fsort_v[0].as<uint32_t>(), h->fsort_v[1].as<uint32_t>(), n,
                      int(2 * h->cfg.cb_len));
    k_final_filter_keys<<<g, 256, 0, st>>>(rows, cs, h->fsort_v[1].as<uint32_t>(), n32, h->min_after_eff, h->fsort_k[0].as<uint64_t>(), ctr);
    device_sort_pairs(h, h->fsort_k[0].as<uint64_t>(), h->fsort_k[1].as<uint64_t>(), h->fsort_v[1].as<uint32_t>(), h->fsort_v[0].as<uint32_t>(), n, 64);
    h->dev_filtered_buf = 0;
    h->launches += 3;
    uint32_t n_real_final = 0;
    DGE_CUDA(cudaMemcpyAsync(&hc, ctr, sizeof(hc), cudaMemcpyDeviceToHost, st));
    DGE_CUDA(cudaMemcpyAsync(&n_real_final, h->real_off.as<uint32_t>() + n, 4, cudaMemcpyDeviceToHost, st));
    DGE_CUDA(cudaStreamSynchronize(st));
    if (hc.key_overflow) throw std::runtime_error("per-cell counters beyond the packed compare_cells key of the device flow; rerun with DGE_HOST_FLOW=1");
    h->n_merged = hc.n_merged; h->n_excluded = hc.n_excluded; h->n_unresolved = 0;
    h->n_filtered_dev = hc.n_filtered;
    tr.mark("finish: sizes + filter (device)");

    // ---- matrices: cm over the filtered list (last max_cells of it with -C), cm_raw over the real cells in cell-id order
    const uint32_t nf = hc.n_filtered;
    const uint32_t skip = (h->cfg.max_cells > 0 && uint32_t(h->cfg.max_cells) < nf) ? nf - uint32_t(h->cfg.max_cells) : 0u;
    h->cols_f.reserve(std::max<size_t>(nf, 1) * 4); h->cols_r.reserve(std::max<size_t>(n_real_final, 1) * 4);
    if (nf - skip)
        k_cols_from_list<<<grid_for(nf - skip, 256), 256, 0, st>>>(rows, h->fsort_v[0].as<uint32_t>() + skip, nf - skip, h->n_pc, h->cols_f.as<uint32_t>());
    k_cols_from_flags<<<g, 256, 0, st>>>(rows, h->real_flag.as<uint32_t>(), h->real_off.as<uint32_t>(), n32, h->n_pc, h->cols_r.as<uint32_t>());
    h->launches += 2;
    build_matrix_cols(h, h->cm, h->cols_f.as<uint32_t>(), nf - skip, true);
    build_matrix_cols(h, h->cm_raw, h->cols_r.as<uint32_t>(), n_real_final, false);
    h->sum_real = n_real_final;
    h->sum_filtered = nf - skip;
    h->dev_merged = true;
}

void do_merge_and_filter(dge_handle *h)
{
    DGE_CUDA(cudaSetDevice(h->cfg.device));
    cudaStream_t st = h->stream;
    Tracer tr;
    tr.st = st;
    if (!h->dist_done) DGE_CUDA(cudaEventRecord(h->ev[3], st));
    if (!h->slot_pc_built) build_slot_pc(h);

    // ---- device flow: no host cell rows at all (the common configurations); anything it cannot decide exactly -> host flow
    static const bool no_dev_flow = std::getenv("DGE_HOST_FLOW") != nullptr;
    if (h->counters.n_flagged)
    {
        if (h->cfg.sharded) throw std::runtime_error("barcodes / UMIs with N are not supported on sharded (multi-GPU) handles yet");
    }
    const bool dev_flow = !no_dev_flow && h->lazy_rows && !h->cfg.sharded && !h->dist_done && h->n_real_rows > 0 && !h->counters.n_flagged &&
                          (h->cfg.merge_type == DGE_MERGE_NONE || (h->cfg.merge_type == DGE_MERGE_REAL && h->wl_fast)) &&
                          h->cfg.umi_merge_type == DGE_UMI_MERGE_SIMPLE && h->cfg.min_genes_before_merge > 0;
    // sharded run: phase 1/2 ran across ranks in dge_dist_step; with a UMI merge strategy that changes U (-u) the host tail below
    // continues from the merged state, otherwise the device tail finishes the run
    const bool after_dist = h->dist_done && h->cfg.umi_merge_type != DGE_UMI_MERGE_SIMPLE;
    if (h->dist_done) merge_device_finish(h);
    if (after_dist)
    {
        DGE_CUDA(cudaStreamSynchronize(st));
        h->state = 2;           // materialize_host applies the merge outcome (flags, counters, refreshed target sizes)
        materialize_host(h);
        h->state = 1;
    }
    else if (h->dist_done || (dev_flow && merge_device_flow(h)))
    {
        DGE_CUDA(cudaEventRecord(h->ev[5], st));
        DGE_CUDA(cudaStreamSynchronize(st));
        tr.mark("finish: matrices");
        h->state = 2;
        auto el = [&](int a, int b) { float ms = 0; cudaEventElapsedTime(&ms, h->ev[a], h->ev[b]); return ms; };
        h->timings.ms_fill = el(0, 1);
        h->timings.ms_init = el(1, 2);
        h->timings.ms_merge = el(3, 4);
        h->timings.ms_finish = el(4, 5);
        h->timings.ms_total = h->timings.ms_fill + h->timings.ms_init + h->timings.ms_merge + h->timings.ms_finish;
        h->timings.n_kernel_launches = h->launches + h->sc_stats.launches;
        return;
    }
    if (dev_flow) ++h->n_host_fallback;
    materialize_host(h);

    if (h->kl.ne && (h->cfg.merge_type == DGE_MERGE_SIMPLE || h->cfg.merge_type == DGE_MERGE_ALL || h->cfg.merge_type == DGE_MERGE_POISSON_SIMPLE))
    {   // these strategies compare barcodes on the device (2-bit Levenshtein): a REAL cell whose barcode contains N has no packed form
        for (auto const &c : h->real)
            if (c.cb & CB_N_BIT)
                throw std::runtime_error("a real cell's barcode contains N: not supported by the no-whitelist merge strategies yet (" + cb_string(h, c.cb) + ")");
    }

    // ---- CB merge (MergeStrategyAbstract::merge, MergeStrategyAbstract.cpp:13-23)
    if (after_dist) {}
    else if (h->cfg.merge_type == DGE_MERGE_REAL)
    {
        phase1_real(h, h->h_target);
        tr.mark("merge: phase 1");
        phase2(h, h->h_target);
        tr.mark("merge: phase 2");
        apply_merges(h);
        DGE_CUDA(cudaStreamSynchronize(st));
        tr.mark("merge: apply");
    }
    else if (h->cfg.merge_type == DGE_MERGE_SIMPLE)
    {
        if (h->cfg.sharded) throw std::runtime_error("SimpleMergeStrategy is not available on sharded (multi-GPU) handles yet");
        phase1_simple(h, h->h_target);
        tr.mark("merge: phase 1 (simple)");
        phase2(h, h->h_target);
        tr.mark("merge: phase 2");
        apply_merges(h);
        DGE_CUDA(cudaStreamSynchronize(st));
        tr.mark("merge: apply");
    }
    else if (h->cfg.merge_type == DGE_MERGE_ALL)
    {
        if (h->cfg.sharded) throw std::runtime_error("merge_type=all is not available on sharded (multi-GPU) handles yet");
        phase1_all(h, h->h_target);
        tr.mark("merge: phase 1 (all)");
        phase2(h, h->h_target);
        apply_merges(h);
        DGE_CUDA(cudaStreamSynchronize(st));
        tr.mark("merge: apply");
    }
    else if (h->cfg.merge_type == DGE_MERGE_POISSON_SIMPLE)
    {
        if (h->cfg.sharded) throw std::runtime_error("PoissonSimpleMergeStrategy is not available on sharded (multi-GPU) handles yet");
        phase1_simple(h, h->h_target, true);
        tr.mark("merge: phase 1 (poisson simple)");
        phase2(h, h->h_target);
        tr.mark("merge: phase 2");
        apply_merges(h);
        DGE_CUDA(cudaStreamSynchronize(st));
        tr.mark("merge: apply");
    }
    else if (h->cfg.merge_type == DGE_MERGE_POISSON_REAL)
    {
        if (h->cfg.sharded) throw std::runtime_error("PoissonRealBarcodesMergeStrategy is not available on sharded (multi-GPU) handles yet");
        phase1_poisson_real(h, h->h_target);
        tr.mark("merge: phase 1 (poisson real)");
        phase2(h, h->h_target);
        tr.mark("merge: phase 2");
        apply_merges(h);
        DGE_CUDA(cudaStreamSynchronize(st));
        tr.mark("merge: apply");
    }
    else if (h->cfg.merge_type != DGE_MERGE_NONE) throw std::runtime_error("unknown merge_type");
    DGE_CUDA(cudaEventRecord(h->ev[4], st));

    // ---- sizes of merge targets changed; Cell::is_real (Cell.cpp:125-128) is evaluated on the merged content from here on
    auto refresh_rows = [&](const std::vector<uint32_t> &owners) {
        std::vector<uint32_t> pcs;
        for (uint32_t i : owners) pcs.push_back(h->real[i].pc);
        std::vector<CellRow> rows;
        gather_rows(h, pcs, rows);
        for (size_t k = 0; k < rows.size(); ++k)
        {
            HostCell &c = h->real[owners[k]];
            c.n_genes = int32_t(rows[k].n_genes); c.req_genes = int32_t(rows[k].req_genes); c.req_umis = int32_t(rows[k].req_umis);
            c.n_umis_distinct = int32_t(rows[k].n_umis);
        }
    };
    if (!h->merge_events.empty() || !h->dist_targets.empty())
    {
        // only merge TARGETS changed content: every other cell keeps the sizes read at set_initialized
        std::vector<uint32_t> owners;
        std::vector<char> is_target(h->real.size(), 0);
        for (auto const &e : h->merge_events) is_target[e.second] = 1;
        for (uint32_t t : h->dist_targets) is_target[t] = 1;
        for (uint32_t i = 0; i < h->real.size(); ++i)
            if (is_target[i] && h->real[i].pc != NONE32) owners.push_back(i);
        refresh_rows(owners);
    }
    for (auto &c : h->real) c.real = !c.merged && !c.excluded && uint32_t(c.n_genes) >= h->cfg.min_genes_before_merge;

    // ---- UMI merge (CellsDataContainer.cpp:47).  MergeUMIsStrategySimple only touches UMIs containing 'N'
    // (MergeUMIsStrategySimple.cpp:21-59) and 2-bit records cannot carry N, so it has nothing to repair here.
    if (h->cfg.umi_merge_type == DGE_UMI_MERGE_DIRECTIONAL)
    {
        if (umi_merge_directional(h))
        {   // requested sizes of every real cell may have changed (update_cell_sizes, CellsDataContainer.cpp:111-125)
            std::vector<uint32_t> owners;
            for (uint32_t i = 0; i < h->real.size(); ++i)
                if (h->real[i].real && h->real[i].pc != NONE32) owners.push_back(i);
            refresh_rows(owners);
        }
        tr.mark("umi merge: directional");
    }
    else if (h->cfg.umi_merge_type != DGE_UMI_MERGE_SIMPLE) throw std::runtime_error("unknown umi_merge_type");
    else if (h->kl.ne && h->counters.n_flagged)
    {   // MergeUMIsStrategySimple: UMIs with N of the real cells
        h->n_umis_merged = 0;
        h->umi_mt.clear();
        if (umi_repair_n(h))
        {
            std::vector<uint32_t> owners;
            for (uint32_t i = 0; i < h->real.size(); ++i)
                if (h->real[i].real && h->real[i].pc != NONE32) owners.push_back(i);
            refresh_rows(owners);
        }
        tr.mark("umi merge: N repair");
    }

    update_filtered(h, h->min_after_eff, h->cfg.max_cells);
    tr.mark("finish: sizes + filter");

    // ---- matrices
    std::vector<uint32_t> cols;
    for (uint32_t i : h->filtered) cols.push_back(h->real[i].pc);
    // a filtered cell always owns UMIs unless thresholds are 0; cells without UMIs contribute empty columns
    std::vector<uint32_t> raw_cols;
    for (auto const &c : h->real) if (c.real) raw_cols.push_back(c.pc);
    for (auto &pc : cols) if (pc == NONE32) pc = h->n_pc;     // barcodes without UMIs: the empty sentinel cell
    for (auto &pc : raw_cols) if (pc == NONE32) pc = h->n_pc;
    build_matrix(h, h->cm, cols, true);
    build_matrix(h, h->cm_raw, raw_cols, false);
    DGE_CUDA(cudaEventRecord(h->ev[5], st));
    DGE_CUDA(cudaStreamSynchronize(st));
    tr.mark("finish: matrices");
    h->state = 2;
    h->host_stage = 2;
    h->sum_real = raw_cols.size();
    h->sum_filtered = h->filtered.size();

    auto el = [&](int a, int b) { float ms = 0; cudaEventElapsedTime(&ms, h->ev[a], h->ev[b]); return ms; };
    h->timings.ms_fill = el(0, 1);
    h->timings.ms_init = el(1, 2);
    h->timings.ms_merge = el(3, 4);
    h->timings.ms_finish = el(4, 5);
    h->timings.ms_total = h->timings.ms_fill + h->timings.ms_init + h->timings.ms_merge + h->timings.ms_finish;
    // ms_dedup_kernel / n_dedup_launches stay those of the main grouping pass (set in set_initialized): the dominant launch
    h->timings.n_kernel_launches = h->launches + h->sc_stats.launches;
}

// host copies for the query surface -----------------------------------------------------------------------------------
struct AllCells
{
    std::vector<dge_cell_info> info; // first-seen order
    std::vector<uint32_t> pc;        // present-cell index or NONE32
};

AllCells collect_all_cells(dge_handle *h)
{
    cudaStream_t st = h->stream;
    std::vector<CellSlot> tab;
    d2h(tab, h->tab.p, h->table_cap, st);
    std::vector<uint32_t> pc_slot, pc_cg, pc_u, pc_reads, pc_rg, pc_ru;
    d2h(pc_slot, h->pc_slot.p, h->n_pc, st);
    d2h(pc_cg, h->pc_cg_start.p, size_t(h->n_pc) + 1, st);
    d2h(pc_u, h->pc_u_start.p, size_t(h->n_pc) + 1, st);
    d2h(pc_reads, h->pc_reads.p, h->n_pc, st);
    d2h(pc_rg, h->pc_req_genes.p, h->n_pc, st);
    d2h(pc_ru, h->pc_req_umis.p, h->n_pc, st);
    DGE_CUDA(cudaStreamSynchronize(st));
    std::vector<uint32_t> slot_pc(h->table_cap, NONE32);
    for (uint32_t i = 0; i < h->n_pc; ++i) slot_pc[pc_slot[i]] = i;
    std::unordered_map<uint32_t, uint32_t> real_by_slot;
    for (uint32_t i = 0; i < h->real.size(); ++i) real_by_slot.emplace(h->real[i].slot, i);
    std::vector<std::pair<uint32_t, uint32_t>> order; // (first_idx, slot)
    for (size_t s = 0; s < h->table_cap; ++s)
        if (tab[s].cb != EMPTY64) order.emplace_back(tab[s].first_idx, uint32_t(s));
    std::sort(order.begin(), order.end());
    std::unordered_map<uint32_t, uint32_t> pos_of_slot;
    pos_of_slot.reserve(order.size() * 2);
    for (uint32_t p = 0; p < order.size(); ++p) pos_of_slot.emplace(order[p].second, p);
    AllCells res;
    res.info.resize(order.size());
    res.pc.resize(order.size());
    for (uint32_t p = 0; p < order.size(); ++p)
    {
        const uint32_t s = order[p].second;
        dge_cell_info &ci = res.info[p];
        ci.barcode = tab[s].cb; ci.first_read_idx = tab[s].first_idx;
        const uint32_t pc = slot_pc[s];
        res.pc[p] = pc;
        ci.flags = 0; ci.merge_target = int32_t(p);
        if (pc != NONE32)
        {
            ci.n_genes = int32_t(pc_cg[pc + 1] - pc_cg[pc]); ci.umis_stat = int32_t(pc_u[pc + 1] - pc_u[pc]);
            ci.reads_stat = int32_t(pc_reads[pc]); ci.requested_genes_num = int32_t(pc_rg[pc]); ci.requested_umis_num = int32_t(pc_ru[pc]);
        }
        else ci.n_genes = ci.umis_stat = ci.reads_stat = ci.requested_genes_num = ci.requested_umis_num = 0;
        auto it = real_by_slot.find(s);
        if (it != real_by_slot.end())
        {
            const HostCell &c = h->real[it->second];
            ci.flags = (c.real ? DGE_CELL_REAL : 0u) | (c.merged ? DGE_CELL_MERGED : 0u) | (c.excluded ? DGE_CELL_EXCLUDED : 0u);
            ci.umis_stat = c.umis_stat; ci.reads_stat = c.reads_stat;
            ci.merge_target = int32_t(pos_of_slot.at(h->real[size_t(c.target)].slot));
        }
    }
    return res;
}

void fill_info(const dge_handle *h, const HostCell &c, dge_cell_info &ci)
{
    (void)h;
    ci.barcode = c.cb; ci.first_read_idx = c.first_idx;
    ci.flags = (c.real ? DGE_CELL_REAL : 0u) | (c.merged ? DGE_CELL_MERGED : 0u) | (c.excluded ? DGE_CELL_EXCLUDED : 0u);
    ci.n_genes = c.n_genes; ci.umis_stat = c.umis_stat; ci.reads_stat = c.reads_stat;
    ci.requested_genes_num = c.req_genes; ci.requested_umis_num = c.req_umis;
    ci.merge_target = -1;
}

} // namespace

namespace
{
template <class T> void h2d_vec(DevBuf &dst, const std::vector<T> &src, cudaStream_t st)
{
    dst.reserve(std::max<size_t>(src.size(), 1) * sizeof(T));
    if (!src.empty()) DGE_CUDA(cudaMemcpyAsync(dst.p, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice, st));
}

// Exact host enumeration for the few children the device walk cannot settle (no eligible class-1 neighbour anywhere -> farther
// distance classes; duplicated whitelist tokens): RealBarcodesMergeStrategy::get_real_neighbour_cbs against the gathered self cells.
struct DistHostView
{
    std::unordered_map<uint64_t, uint32_t> by_cb;
    void build(const std::vector<SelfRec> &all)
    {
        by_cb.reserve(all.size() * 2);
        for (uint32_t i = 0; i < all.size(); ++i) by_cb.emplace(all[i].cb, i);
    }
};

std::vector<long> dist_exact_neighbours(const dge_handle *h, const DistHostView &v, const CellRow &child)
{
    auto lookup = [&](const std::string &sq) -> long {
        uint64_t packed;
        if (!pack_seq(sq, packed)) return -1;
        auto it = v.by_cb.find(packed);
        return it == v.by_cb.end() ? -1 : long(it->second);
    };
    auto eligible = [&](long gi) { return h->hx_all[size_t(gi)].umis >= child.n_umis; };
    return h->wl.neighbours(cb_string(h, child.cb), false, lookup, eligible);
}

void dist_fetch_host_tables(dge_handle *h)
{
    cudaStream_t st = h->stream;
    if (h->hx_all.size() != h->n_all || h->hx_rows.size() != h->n_real_rows)
    {
        d2h(h->hx_all, h->x_all.p, h->n_all, st);
        d2h(h->hx_rows, h->rows_dev2.p, h->n_real_rows, st);
        DGE_CUDA(cudaStreamSynchronize(st));
    }
}
} // namespace

// =====================================================================================================================
extern "C" {

void dge_config_default(dge_config *cfg)
{
    std::memset(cfg, 0, sizeof(*cfg));
    cfg->abi_version = DGE_ABI_VERSION;
    cfg->device = 0;
    cfg->cb_len = 16; cfg->umi_len = 10; cfg->n_genes = 1;
    cfg->merge_type = DGE_MERGE_NONE;
    cfg->barcodes_type = DGE_BARCODES_INDROP; // MergeStrategyFactory.cpp:45 default "indrop"
    cfg->umi_merge_type = DGE_UMI_MERGE_SIMPLE;
    cfg->min_genes_before_merge = 10; cfg->min_genes_after_merge = 10;
    cfg->max_cb_merge_edit_distance = 2; cfg->max_umi_merge_edit_distance = 1;
    cfg->min_merge_fraction = 0.2; cfg->max_merge_prob = 1e-4; cfg->max_real_merge_prob = 1e-7; cfg->umi_merge_mult = 2;
    cfg->query_mark_mask = 0xCC; // eEBA
    cfg->max_cells = -1;
}

int dge_create(const dge_config *cfg, dge_handle **out)
{
    if (!cfg || !out) return fail(nullptr, DGE_ERR_INVALID, "null argument");
    *out = nullptr;
    if (cfg->abi_version != DGE_ABI_VERSION) return fail(nullptr, DGE_ERR_INVALID, "abi_version mismatch");
    if (cfg->cb_len == 0 || cfg->cb_len > 20) return fail(nullptr, DGE_ERR_INVALID, "cb_len must be in 1..20");
    if (cfg->umi_len == 0 || cfg->umi_len > 12) return fail(nullptr, DGE_ERR_INVALID, "umi_len must be in 1..12");
    if (cfg->n_genes == 0 || cfg->n_genes >= DGE_NO_GENE) return fail(nullptr, DGE_ERR_INVALID, "n_genes must be in 1..2^24-2");
    if (cfg->merge_type > DGE_MERGE_ALL) return fail(nullptr, DGE_ERR_INVALID, "unknown merge_type");
    if ((cfg->query_mark_mask & ~0xFEu) != 0) return fail(nullptr, DGE_ERR_INVALID, "query_mark_mask uses bits 1..7 only");
    std::unique_ptr<dge_handle> h(new dge_handle());
    h->cfg = *cfg;
    if (cfg->barcodes_file) h->barcodes_file = cfg->barcodes_file;
    h->cfg.barcodes_file = nullptr;
    // MergeStrategyAbstract ctor: min_genes_after_merge = max(after, before)  (MergeStrategyAbstract.cpp:8-11)
    h->min_after_eff = std::max(cfg->min_genes_after_merge, cfg->min_genes_before_merge);
    const bool wants_wl = cfg->merge_type == DGE_MERGE_REAL || cfg->merge_type == DGE_MERGE_POISSON_REAL;
    if (wants_wl)
    {
        if (h->barcodes_file.empty()) return fail(nullptr, DGE_ERR_INVALID, "merge_type needs barcodes_file");
        try { h->wl.load(h->barcodes_file, cfg->barcodes_type == DGE_BARCODES_INDROP); }
        catch (std::exception &e) { return fail(nullptr, DGE_ERR_IO, e.what()); }
        h->wl_fast = h->wl.fast_path_ok(cfg->cb_len);
    }
    int n_dev = 0;
    cudaError_t ce = cudaGetDeviceCount(&n_dev);
    if (ce != cudaSuccess || n_dev <= 0)
        return fail(nullptr, DGE_ERR_CUDA, std::string("no CUDA device available: ") + cudaGetErrorString(ce));
    if (cfg->device < 0 || cfg->device >= n_dev) return fail(nullptr, DGE_ERR_INVALID, "device ordinal out of range");
    dge_handle *raw = h.get();
    int rc = guarded(raw, [&] { ensure_device(raw); return int(DGE_OK); });
    if (rc != DGE_OK) { g_create_error = raw->err; return rc; }
    *out = h.release();
    return DGE_OK;
}

void dge_destroy(dge_handle *h)
{
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    delete h;
}

const char *dge_last_error(const dge_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int dge_set_stream(dge_handle *h, void *cuda_stream)
{
    if (!h) return DGE_ERR_INVALID;
    if (h->n_reads) return fail(h, DGE_ERR_STATE, "dge_set_stream must precede the first batch");
    return guarded(h, [&] {
        DGE_CUDA(cudaSetDevice(h->cfg.device));
        DGE_CUDA(cudaStreamSynchronize(h->stream));
        if (h->own_stream) cudaStreamDestroy(h->stream);
        h->stream = static_cast<cudaStream_t>(cuda_stream);
        h->own_stream = false;
        return int(DGE_OK);
    });
}

int dge_set_n_strings(dge_handle *h, int which, const char *strings, size_t n)
{
    if (!h || (n && !strings) || (which != 0 && which != 1)) return fail(h, DGE_ERR_INVALID, "bad argument");
    if (!h->cfg.allow_n) return fail(h, DGE_ERR_STATE, "dge_set_n_strings needs dge_config.allow_n = 1");
    return guarded(h, [&] {
        const size_t len = which == 0 ? h->cfg.umi_len : h->cfg.cb_len;
        std::string all(strings ? strings : "", n * len);
        for (char c : all)
            if (c != 'A' && c != 'C' && c != 'G' && c != 'T' && c != 'N') throw InvalidInput("N-string lists hold A, C, G, T, N only");
        if (which == 0) h->n_umi_strings = all;
        else
        {
            h->n_cb_list.clear();
            for (size_t k = 0; k < n; ++k) h->n_cb_list.push_back(all.substr(k * len, len));
        }
        return int(DGE_OK);
    });
}

// The DGE_FLAG_CB_N list with strings of ANY length: barcodes whose length differs from dge_config.cb_len (variable-length inDrop v1 / v2
// barcodes, InDropBarcodesParser.cpp:31-38) travel like barcodes with N -- as a cell of their own, identified by the list index on the
// device, by their string wherever the reference looks at the string (whitelist walk, compare_cells ties, output).
int dge_set_cb_strings(dge_handle *h, const char *strings, const uint32_t *lengths, size_t n)
{
    if (!h || (n && (!strings || !lengths))) return fail(h, DGE_ERR_INVALID, "bad argument");
    if (!h->cfg.allow_n) return fail(h, DGE_ERR_STATE, "dge_set_cb_strings needs dge_config.allow_n = 1");
    return guarded(h, [&] {
        std::vector<std::string> list;
        size_t off = 0;
        for (size_t k = 0; k < n; ++k)
        {
            std::string s(strings + off, lengths[k]);
            off += lengths[k];
            if (s.empty()) throw InvalidInput("empty barcode in the escaped-barcode list");
            for (char c : s)
                if (c != 'A' && c != 'C' && c != 'G' && c != 'T' && c != 'N') throw InvalidInput("barcode lists hold A, C, G, T, N only");
            list.push_back(std::move(s));
        }
        h->n_cb_list.swap(list);
        return int(DGE_OK);
    });
}

int dge_add_batch_segments_device(dge_handle *h, const dge_record16 *const *segs, const uint64_t *counts, uint32_t n_segs)
{
    if (!h || (n_segs && (!segs || !counts))) return fail(h, DGE_ERR_INVALID, "null argument");
    if (h->state != 0) return fail(h, DGE_ERR_STATE, "Container is already initialized");
    for (uint32_t k = 0; k < n_segs; ++k)
        if (!segs[k] && counts[k]) return fail(h, DGE_ERR_INVALID, "null segment");
    return guarded(h, [&] { ensure_device(h); fill_from_device(h, nullptr, 0, nullptr, nullptr, 0, segs, counts, n_segs); return int(DGE_OK); });
}

int dge_add_batch_device(dge_handle *h, const dge_record16 *recs, size_t n)
{
    if (!h || (!recs && n)) return fail(h, DGE_ERR_INVALID, "null argument");
    if (h->state != 0) return fail(h, DGE_ERR_STATE, "Container is already initialized");
    return guarded(h, [&] { ensure_device(h); fill_from_device(h, recs, n); return int(DGE_OK); });
}

// Per-chromosome counters of one batch (device arrays); runs behind the batch's fill kernel, which has inserted every barcode.
static void chr_stats_from_device(dge_handle *h, const dge_record16 *recs, const unsigned long long *soa_keys, const uint32_t *soa_genes,
                                  const uint8_t *chr, size_t n)
{
    if (n == 0) return;
    if (h->kl.tb > 24) throw CapacityError("per-chromosome statistics need at most 2^24 barcode slots (lower max_barcodes_hint)");
    if (!h->chr_cap)
    {
        int bits = std::min(26, std::max(12, h->kl.tb + 3));
        h->chr_cap = size_t(1) << bits;
        h->chr_tab.reserve(h->chr_cap * sizeof(ChrEntry));
        h->chr_ctr.reserve(sizeof(ChrCounters) + 16);
        DGE_CUDA(cudaMemsetAsync(h->chr_tab.p, 0, h->chr_cap * sizeof(ChrEntry), h->stream));
        DGE_CUDA(cudaMemsetAsync(h->chr_ctr.p, 0, sizeof(ChrCounters), h->stream));
    }
    h->chr_used = true;
    const unsigned grid = unsigned(std::min<size_t>(div_up(n, size_t(256) * 4), size_t(148) * 8));
    if (soa_keys)
        k_chr_stats<true, 4><<<grid, 256, 0, h->stream>>>(nullptr, soa_keys, soa_genes, chr, n, h->tab.as<CellSlot>(), h->kl, h->chr_tab.as<ChrEntry>(),
                                                      uint32_t(h->chr_cap - 1), h->chr_ctr.as<ChrCounters>());
    else
        k_chr_stats<false, 4><<<grid, 256, 0, h->stream>>>(reinterpret_cast<const Rec16 *>(recs), nullptr, nullptr, chr, n, h->tab.as<CellSlot>(), h->kl,
                                                       h->chr_tab.as<ChrEntry>(), uint32_t(h->chr_cap - 1), h->chr_ctr.as<ChrCounters>());
    DGE_LAUNCH_CHECK();
    ++h->launches;
}

// Host batches: double-buffered device staging filled on a COPY stream, consumed by the fill kernel on the main stream, so the H2D
// copy of slice k+1 overlaps the fill kernel of slice k (the copies are truly asynchronous when the host memory is page-locked).
static void add_host_slices(dge_handle *h, const dge_record16 *recs, const unsigned long long *keys, const uint32_t *genes, size_t n, uint64_t first_idx,
                            const uint8_t *chr = nullptr)
{
    ensure_device(h);
    if (!h->copy_stream)
    {
        DGE_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        for (auto &e : h->copied_ev) DGE_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    const bool soa = keys != nullptr;
    const size_t slice = size_t(32) << 20; // records per staging slice
    // the copy stream must not run ahead of work already queued on the main stream that still reads the staging buffers
    for (size_t off = 0; off < n; off += slice)
    {
        const size_t m = std::min(slice, n - off);
        const int turn = h->staging_turn;
        DevBuf &stg = h->staging[turn];
        h->staging_turn ^= 1;
        // the fill kernel that last read this staging buffer must have finished before it is overwritten
        DGE_CUDA(cudaStreamWaitEvent(h->copy_stream, h->staging_ev[turn], 0));
        if (stg.bytes < m * 16 + 64) { DGE_CUDA(cudaEventSynchronize(h->staging_ev[turn])); stg.reserve(m * 16 + 64); }
        uint8_t *dchr = nullptr;
        if (chr)
        {
            DevBuf &cs = h->chr_staging[turn];
            if (cs.bytes < m + 64) { DGE_CUDA(cudaEventSynchronize(h->staging_ev[turn])); cs.reserve(m + 64); }
            dchr = cs.as<uint8_t>();
            DGE_CUDA(cudaMemcpyAsync(dchr, chr + off, m, cudaMemcpyHostToDevice, h->copy_stream));
        }
        if (soa)
        {
            unsigned long long *dk = stg.as<unsigned long long>();
            uint32_t *dg = reinterpret_cast<uint32_t *>(dk + ((m + 1) & ~size_t(1))); // 16-byte aligned (bulk copies)
            DGE_CUDA(cudaMemcpyAsync(dk, keys + off, m * 8, cudaMemcpyHostToDevice, h->copy_stream));
            DGE_CUDA(cudaMemcpyAsync(dg, genes + off, m * 4, cudaMemcpyHostToDevice, h->copy_stream));
            DGE_CUDA(cudaEventRecord(h->copied_ev[turn], h->copy_stream));
            DGE_CUDA(cudaStreamWaitEvent(h->stream, h->copied_ev[turn], 0));
            fill_from_device(h, nullptr, m, dk, dg, uint32_t(first_idx + off));
            if (chr) chr_stats_from_device(h, nullptr, dk, dg, dchr, m);
        }
        else
        {
            DGE_CUDA(cudaMemcpyAsync(stg.p, recs + off, m * sizeof(dge_record16), cudaMemcpyHostToDevice, h->copy_stream));
            DGE_CUDA(cudaEventRecord(h->copied_ev[turn], h->copy_stream));
            DGE_CUDA(cudaStreamWaitEvent(h->stream, h->copied_ev[turn], 0));
            fill_from_device(h, stg.as<dge_record16>(), m);
            if (chr) chr_stats_from_device(h, stg.as<dge_record16>(), nullptr, nullptr, dchr, m);
        }
        DGE_CUDA(cudaEventRecord(h->staging_ev[turn], h->stream));
    }
    // the caller owns the host arrays again when the call returns: wait for the last H2D copy (not for the fill kernels)
    DGE_CUDA(cudaStreamSynchronize(h->copy_stream));
}

int dge_add_batch(dge_handle *h, const dge_record16 *recs, size_t n)
{
    if (!h || (!recs && n)) return fail(h, DGE_ERR_INVALID, "null argument");
    if (h->state != 0) return fail(h, DGE_ERR_STATE, "Container is already initialized");
    return guarded(h, [&] { add_host_slices(h, recs, nullptr, nullptr, n, 0); return int(DGE_OK); });
}

int dge_add_batch_soa(dge_handle *h, const uint64_t *keys, const uint32_t *genes, size_t n, uint64_t first_read_idx)
{
    if (!h || ((!keys || !genes) && n)) return fail(h, DGE_ERR_INVALID, "null argument");
    if (h->state != 0) return fail(h, DGE_ERR_STATE, "Container is already initialized");
    if (first_read_idx + n > 0xFFFFFFFFull) return fail(h, DGE_ERR_INVALID, "read_idx beyond 2^32");
    return guarded(h, [&] { add_host_slices(h, nullptr, reinterpret_cast<const unsigned long long *>(keys), genes, n, first_read_idx); return int(DGE_OK); });
}

int dge_add_batch_chr(dge_handle *h, const dge_record16 *recs, const uint8_t *chr, size_t n)
{
    if (!h || ((!recs || !chr) && n)) return fail(h, DGE_ERR_INVALID, "null argument");
    if (h->state != 0) return fail(h, DGE_ERR_STATE, "Container is already initialized");
    return guarded(h, [&] { add_host_slices(h, recs, nullptr, nullptr, n, 0, chr); return int(DGE_OK); });
}

int dge_add_batch_soa_chr(dge_handle *h, const uint64_t *keys, const uint32_t *genes, const uint8_t *chr, size_t n, uint64_t first_read_idx)
{
    if (!h || ((!keys || !genes || !chr) && n)) return fail(h, DGE_ERR_INVALID, "null argument");
    if (h->state != 0) return fail(h, DGE_ERR_STATE, "Container is already initialized");
    if (first_read_idx + n > 0xFFFFFFFFull) return fail(h, DGE_ERR_INVALID, "read_idx beyond 2^32");
    return guarded(h, [&] { add_host_slices(h, nullptr, reinterpret_cast<const unsigned long long *>(keys), genes, n, first_read_idx, chr); return int(DGE_OK); });
}

int dge_add_batch_chr_device(dge_handle *h, const dge_record16 *recs, const uint8_t *chr, size_t n)
{
    if (!h || ((!recs || !chr) && n)) return fail(h, DGE_ERR_INVALID, "null argument");
    if (h->state != 0) return fail(h, DGE_ERR_STATE, "Container is already initialized");
    return guarded(h, [&] { ensure_device(h); fill_from_device(h, recs, n); chr_stats_from_device(h, recs, nullptr, nullptr, chr, n); return int(DGE_OK); });
}

// Stats::get(CellChrStatType) for the real cells, Stats::merge applied (Stats.cpp:29-63): counts[cell][chr][0 exon | 1 intron | 2 intergenic]
int dge_get_chr_stats(dge_handle *h, int32_t *counts, size_t capacity_cells, size_t *n_cells, uint32_t *n_chr, uint8_t *presented)
{
    if (!h || !n_cells || !n_chr) return fail(h, DGE_ERR_INVALID, "null argument");
    if (h->state != 2) return fail(h, DGE_ERR_STATE, "per-chromosome statistics are read after merge_and_filter");
    return guarded(h, [&] {
        DGE_CUDA(cudaSetDevice(h->cfg.device));
        if (h->cfg.sharded) throw std::runtime_error("per-chromosome statistics are not available on sharded (multi-GPU) handles yet");
        materialize_host(h);
        std::vector<uint32_t> pos(h->real.size(), NONE32);
        size_t nr = 0;
        for (size_t i = 0; i < h->real.size(); ++i) if (h->real[i].real) pos[i] = uint32_t(nr++);
        *n_cells = nr;
        *n_chr = 0;
        if (!h->chr_used) return int(DGE_OK);
        ChrCounters cc;
        DGE_CUDA(cudaMemcpyAsync(&cc, h->chr_ctr.p, sizeof(cc), cudaMemcpyDeviceToHost, h->stream));
        DGE_CUDA(cudaStreamSynchronize(h->stream));
        if (cc.overflow) throw CapacityError("per-chromosome counter table overflow (raise max_barcodes_hint)");
        if (cc.missing_cell) throw std::runtime_error("internal: a chromosome batch named a barcode the fill had not seen");
        *n_chr = cc.max_chr + 1;
        if (!counts && !presented) return int(DGE_OK);
        if (counts && capacity_cells < nr) throw InvalidInput("dge_get_chr_stats: capacity_cells is smaller than the number of real cells");
        // occupied entries -> host (a query outside the per-read path: a few MB)
        h->chr_export.reserve(h->chr_cap * sizeof(ChrEntry));
        unsigned long long *d_n = reinterpret_cast<unsigned long long *>(h->chr_ctr.as<unsigned char>() + sizeof(ChrCounters));
        DGE_CUDA(cudaMemsetAsync(d_n, 0, 8, h->stream));
        k_chr_export<<<unsigned(div_up(h->chr_cap, size_t(256))), 256, 0, h->stream>>>(h->chr_tab.as<ChrEntry>(), h->chr_cap, h->chr_export.as<ChrEntry>(), d_n);
        DGE_LAUNCH_CHECK();
        unsigned long long n_ent = 0;
        DGE_CUDA(cudaMemcpyAsync(&n_ent, d_n, 8, cudaMemcpyDeviceToHost, h->stream));
        DGE_CUDA(cudaStreamSynchronize(h->stream));
        std::vector<ChrEntry> ent;
        ent.resize(size_t(n_ent));
        if (n_ent) DGE_CUDA(cudaMemcpy(ent.data(), h->chr_export.p, size_t(n_ent) * sizeof(ChrEntry), cudaMemcpyDeviceToHost));
        const uint32_t nc = *n_chr;
        if (counts) std::fill(counts, counts + nr * nc * 3, 0);
        if (presented) std::fill(presented, presented + size_t(3) * nc, uint8_t(0));
        std::unordered_map<uint32_t, uint32_t> by_slot;
        by_slot.reserve(h->real.size() * 2);
        for (size_t i = 0; i < h->real.size(); ++i) by_slot.emplace(h->real[i].slot, uint32_t(i));
        for (auto const &e : ent)
        {
            const uint32_t key = e.key - 1u, slot = key >> 8, c = key & 255u;
            if (presented)
            {
                if (e.exon) presented[0 * nc + c] = 1;
                if (e.intron) presented[1 * nc + c] = 1;
                if (e.intergenic) presented[2 * nc + c] = 1;
            }
            if (!counts) continue;
            auto it = by_slot.find(slot);
            if (it == by_slot.end()) continue; // never a real cell: neither reported nor a merge source
            uint32_t f = it->second;
            for (int hop = 0; hop < 64; ++hop)
            {   // merge targets are final (MergeStrategyBase::reassign); follow defensively
                const int32_t t = h->real[f].target;
                if (t < 0 || uint32_t(t) == f || !h->real[f].merged) break;
                f = uint32_t(t);
            }
            if (pos[f] == NONE32) continue; // excluded, or merged into a cell of another kind
            int32_t *row = counts + (size_t(pos[f]) * nc + c) * 3;
            row[0] += int32_t(e.exon); row[1] += int32_t(e.intron); row[2] += int32_t(e.intergenic);
        }
        return int(DGE_OK);
    });
}

int dge_reset(dge_handle *h)
{
    if (!h) return DGE_ERR_INVALID;
    return guarded(h, [&] {
        DGE_CUDA(cudaSetDevice(h->cfg.device));
        if (h->stream) DGE_CUDA(cudaStreamSynchronize(h->stream));
        for (auto &c : h->chunks) h->chunk_pool.push_back(std::move(c));
        h->chunks.clear();
        h->n_chunk_counters = 0; h->n_fill_ev = 0; h->moves_ready = false; h->n_reads = 0; h->n_keys = 0; h->n_u = h->n_cg = h->n_pc = 0;
        h->real.clear(); h->filtered.clear(); h->gene_order.clear(); h->merge_events.clear(); h->umi_mt.clear();
        h->n_merged = h->n_excluded = h->n_unresolved = 0; h->total_cells = 0;
        h->cm.built = h->cm_raw.built = h->cm_marks.built = false;
        h->host_stage = 0; h->lazy_rows = false; h->dev_merged = false; h->n_real_rows = 0; h->n_filtered_dev = 0;
        h->sum_real = h->sum_filtered = h->sum_genes_seen = 0; h->n_host_fallback = 0; h->pp_ready = false; h->n_poisson_replayed = 0;
        h->dist_done = false; h->slot_pc_built = false; h->dist_targets.clear(); h->n_order_ties = 0; h->dist_stage = 0;
        h->timings = dge_timings{}; h->sc_stats = SortCombineStats{}; h->launches = 0;
        h->state = 0;
        if (h->device_ready) reset_fill_state(h);
        return int(DGE_OK);
    });
}

int dge_set_initialized(dge_handle *h)
{
    if (!h) return DGE_ERR_INVALID;
    if (h->state != 0) return fail(h, DGE_ERR_STATE, "Container is already initialized");
    return guarded(h, [&] { do_set_initialized(h); return int(DGE_OK); });
}

int dge_merge_and_filter(dge_handle *h)
{
    if (!h) return DGE_ERR_INVALID;
    if (h->state == 0) return fail(h, DGE_ERR_STATE, "You must initialize container");
    if (h->state == 2) return fail(h, DGE_ERR_STATE, "merge_and_filter was already run");
    return guarded(h, [&] { do_merge_and_filter(h); return int(DGE_OK); });
}

// ---- cross-rank whitelist merge (see include/dropest_b200.h and csrc/distmerge.cuh) ----------------------------------------------
extern "C" int dge_dist_step(dge_handle *h, dge_dist_io *io)
{
    if (!h || !io) return fail(h, DGE_ERR_INVALID, "null argument");
    return guarded(h, [&]() -> int {
        if (h->state != 1) throw std::runtime_error("dge_dist_step runs between dge_set_initialized and dge_merge_and_filter");
        if (!h->cfg.sharded || h->cfg.merge_type != DGE_MERGE_REAL) throw std::runtime_error("dge_dist_step needs sharded = 1 and merge_type = DGE_MERGE_REAL");
        if (!h->wl_fast) throw std::runtime_error("cross-rank merge needs a whitelist with equal-length, N-free parts");
        if (io->world == 0 || io->world > DGE_DIST_MAX_WORLD || io->rank >= io->world) throw InvalidInput("bad world / rank");
        DGE_CUDA(cudaSetDevice(h->cfg.device));
        cudaStream_t st = h->stream;
        const uint32_t world = io->world, me = io->rank;
        const size_t n = h->n_real_rows;
        const uint32_t n32 = uint32_t(n);
        const unsigned g = grid_for(std::max<size_t>(n, 1), 256);
        CellRow *rows = h->rows_dev2.as<CellRow>();
        const int gub = h->kl.gb + h->kl.ub;
        std::memset(io->send_bytes, 0, sizeof(io->send_bytes));
        io->send = nullptr;
        io->stage = uint32_t(h->dist_stage);

        if (h->dist_stage == 0)
        {   // ---- self cells -> all-gather
            if (n && !h->lazy_rows) throw std::runtime_error("cross-rank merge needs the device-resident cell rows (min_genes_before_merge > 0)");
            if (h->counters.n_flagged) throw std::runtime_error("barcodes / UMIs with N are not supported on sharded (multi-GPU) handles yet");
            DGE_CUDA(cudaEventRecord(h->ev[3], st));
            h->dist_world = world; h->dist_rank = me;
            if (!h->wl_uploaded) { upload_whitelist(h); h->wl_uploaded = true; }
            h->cell_state.reserve(std::max<size_t>(n, 1) * sizeof(CellState));
            h->df_ctr.reserve(sizeof(DevFlowCounters));
            DGE_CUDA(cudaMemsetAsync(h->df_ctr.p, 0, sizeof(DevFlowCounters), st));
            h->x_bad.reserve(16);
            DGE_CUDA(cudaMemsetAsync(h->x_bad.p, 0, 16, st));
            h->x_self_flag.reserve((n + 1) * 4); h->x_self_one.reserve((n + 1) * 4); h->x_self_off.reserve((n + 1) * 4);
            h->n_self = 0;
            if (n)
            {
                k_state_init<<<g, 256, 0, st>>>(rows, n32, h->cell_state.as<CellState>(), nullptr);
                k_wl_self<<<unsigned(div_up(n * 32, size_t(256))), 256, 0, st>>>(rows, n32, h->wl_dev, h->x_self_flag.as<uint32_t>());
                k_self_is_one<<<g, 256, 0, st>>>(h->x_self_flag.as<uint32_t>(), n32, h->x_self_one.as<uint32_t>());
                DGE_CUDA(cudaMemsetAsync(h->x_self_one.as<uint32_t>() + n, 0, 4, st));
                device_exclusive_scan(h->x_self_one.as<uint32_t>(), h->x_self_off.as<uint32_t>(), n + 1, h->scan_scratch.as<uint32_t>(), st, &h->launches);
                h->n_self = d2h_scalar<uint32_t>(h->x_self_off.as<uint32_t>() + n, st);
                h->launches += 3;
            }
            h->x_self_send.reserve(std::max<size_t>(h->n_self, 1) * sizeof(SelfRec)); h->x_self_idx.reserve(std::max<size_t>(h->n_self, 1) * 4);
            if (h->n_self)
                k_self_export<<<g, 256, 0, st>>>(rows, h->x_self_flag.as<uint32_t>(), h->x_self_off.as<uint32_t>(), n32, h->x_self_send.as<SelfRec>(),
                                                 h->x_self_idx.as<uint32_t>());
            DGE_LAUNCH_CHECK();
            DGE_CUDA(cudaStreamSynchronize(st));
            io->collective = DGE_DIST_ALLGATHER;
            io->send = h->x_self_send.p;
            io->send_bytes[0] = uint64_t(h->n_self) * sizeof(SelfRec);
            h->dist_stage = 1;
            return DGE_OK;
        }

        if (h->dist_stage == 1)
        {   // ---- gathered self cells -> table; candidates of the local children; pairs packed per owner of the candidate
            h->hx_rank_off.assign(world + 1, 0);
            for (uint32_t r = 0; r < world; ++r)
            {
                if (io->recv_bytes[r] % sizeof(SelfRec)) throw InvalidInput("all-gather piece is not a whole number of summaries");
                h->hx_rank_off[r + 1] = h->hx_rank_off[r] + uint32_t(io->recv_bytes[r] / sizeof(SelfRec));
            }
            h->n_all = h->hx_rank_off[world];
            if (h->hx_rank_off[me + 1] - h->hx_rank_off[me] != h->n_self) throw InvalidInput("all-gather does not contain this rank's own piece");
            h->hx_all.clear(); h->hx_rows.clear();
            h->x_all.reserve(std::max<size_t>(h->n_all, 1) * sizeof(SelfRec));
            if (h->n_all) DGE_CUDA(cudaMemcpyAsync(h->x_all.p, io->recv, size_t(h->n_all) * sizeof(SelfRec), cudaMemcpyDeviceToDevice, st));
            h2d_vec(h->x_rank_off, h->hx_rank_off, st);
            uint32_t cap = 64;
            while (cap < 2 * h->n_all) cap <<= 1;
            h->g_mask = cap - 1;
            h->x_gcb.reserve(size_t(cap) * 8); h->x_ggi.reserve(size_t(cap) * 4);
            DGE_CUDA(cudaMemsetAsync(h->x_gcb.p, 0xFF, size_t(cap) * 8, st));
            if (h->n_all)
                k_g_build<<<grid_for(h->n_all, 256), 256, 0, st>>>(h->x_all.as<SelfRec>(), h->n_all, h->x_gcb.as<unsigned long long>(), h->x_ggi.as<uint32_t>(), h->g_mask);
            h->x_nb_count.reserve((n + 1) * 4); h->x_nb_gi.reserve(std::max<size_t>(n, 1) * WL_K * 4);
            h->x_pair_cnt.reserve((n + 1) * 4); h->x_pair_off.reserve((n + 1) * 4);
            uint32_t n_pairs = 0;
            std::vector<int> nbc;
            if (n)
            {
                k_wl_class01_g<<<unsigned(div_up(n * 32, size_t(256))), 256, 0, st>>>(rows, h->x_self_flag.as<uint32_t>(), n32, h->wl_dev,
                                                                                      h->x_gcb.as<unsigned long long>(), h->x_ggi.as<uint32_t>(), h->g_mask,
                                                                                      h->x_all.as<SelfRec>(), h->x_nb_count.as<int>(), h->x_nb_gi.as<uint32_t>());
                k_p1_counts<<<g, 256, 0, st>>>(h->x_nb_count.as<int>(), n32, h->x_pair_cnt.as<uint32_t>());
                DGE_CUDA(cudaMemsetAsync(h->x_pair_cnt.as<uint32_t>() + n, 0, 4, st));
                device_exclusive_scan(h->x_pair_cnt.as<uint32_t>(), h->x_pair_off.as<uint32_t>(), n + 1, h->scan_scratch.as<uint32_t>(), st, &h->launches);
                DGE_LAUNCH_CHECK();
                h->launches += 3;
                d2h(nbc, h->x_nb_count.p, n, st);
                n_pairs = d2h_scalar<uint32_t>(h->x_pair_off.as<uint32_t>() + n, st);
            }
            // children the device walk could not settle: exact enumeration on the host (rare)
            std::vector<uint32_t> slow;
            for (uint32_t i = 0; i < n; ++i) if (nbc[i] == NB_SLOW) slow.push_back(i);
            std::vector<uint32_t> extra_child, extra_gi;
            h->n_dist_slow = slow.size();
            if (!slow.empty())
            {
                dist_fetch_host_tables(h);
                DistHostView view; view.build(h->hx_all);
                std::vector<uint32_t> pcnt, poff;
                d2h(pcnt, h->x_pair_cnt.p, n + 1, st); d2h(poff, h->x_pair_off.p, n + 1, st);
                DGE_CUDA(cudaStreamSynchronize(st));
                for (uint32_t i : slow)
                {
                    std::vector<long> ids = dist_exact_neighbours(h, view, h->hx_rows[i]);
                    if (!ids.empty() && h->hx_all[size_t(ids[0])].cb == h->hx_rows[i].cb) { nbc[i] = NB_SELF; continue; } // the cell is a whitelist barcode itself
                    poff[i] = n_pairs + uint32_t(extra_child.size());
                    pcnt[i] = uint32_t(ids.size());
                    nbc[i] = 0; // no neighbour within the distance bound: excluded below
                    for (long id : ids) { extra_child.push_back(i); extra_gi.push_back(uint32_t(id)); }
                }
                DGE_CUDA(cudaMemcpyAsync(h->x_pair_cnt.p, pcnt.data(), (n + 1) * 4, cudaMemcpyHostToDevice, st));
                DGE_CUDA(cudaMemcpyAsync(h->x_pair_off.p, poff.data(), (n + 1) * 4, cudaMemcpyHostToDevice, st));
                DGE_CUDA(cudaMemcpyAsync(h->x_nb_count.p, nbc.data(), n * 4, cudaMemcpyHostToDevice, st));
                DGE_CUDA(cudaStreamSynchronize(st));
            }
            const uint32_t n_dev_pairs = n_pairs;
            n_pairs += uint32_t(extra_child.size());
            h->n_pairs = n_pairs;
            const size_t np1 = std::max<size_t>(n_pairs, 1);
            h->x_pair_child.reserve(np1 * 4); h->x_pair_gi.reserve(np1 * 4); h->x_pair_pos.reserve(np1 * 4); h->x_pair_eoff.reserve(np1 * 4);
            h->x_pair_isect.reserve(np1 * 4); h->x_local_jobs.reserve(np1 * sizeof(PairJob)); h->x_local_isect.reserve(np1 * 4);
            if (n_dev_pairs)
                k_dist_pairs<<<g, 256, 0, st>>>(nbc.empty() ? nullptr : h->x_nb_count.as<int>(), h->x_nb_gi.as<uint32_t>(), h->x_pair_off.as<uint32_t>(), n32,
                                                h->x_pair_child.as<uint32_t>(), h->x_pair_gi.as<uint32_t>());
            if (!extra_child.empty())
            {
                DGE_CUDA(cudaMemcpyAsync(h->x_pair_child.as<uint32_t>() + n_dev_pairs, extra_child.data(), extra_child.size() * 4, cudaMemcpyHostToDevice, st));
                DGE_CUDA(cudaMemcpyAsync(h->x_pair_gi.as<uint32_t>() + n_dev_pairs, extra_gi.data(), extra_gi.size() * 4, cudaMemcpyHostToDevice, st));
            }
            // per destination: pairs and entries
            h->x_dest.reserve(size_t(4) * DIST_MAX_WORLD * 4);
            DGE_CUDA(cudaMemsetAsync(h->x_dest.p, 0, size_t(4) * DIST_MAX_WORLD * 4, st));
            uint32_t *dest_pairs = h->x_dest.as<uint32_t>(), *dest_entries = dest_pairs + DIST_MAX_WORLD, *cur_pairs = dest_entries + DIST_MAX_WORLD,
                     *cur_entries = cur_pairs + DIST_MAX_WORLD;
            std::vector<uint32_t> hd(2 * DIST_MAX_WORLD, 0);
            if (n_pairs)
            {
                k_dist_count<<<grid_for(n_pairs, 256), 256, 0, st>>>(h->x_pair_child.as<uint32_t>(), h->x_pair_gi.as<uint32_t>(), n_pairs, rows, h->x_rank_off.as<uint32_t>(),
                                                                    world, me, dest_pairs, dest_entries);
                DGE_CUDA(cudaMemcpyAsync(hd.data(), h->x_dest.p, 2 * DIST_MAX_WORLD * 4, cudaMemcpyDeviceToHost, st));
                DGE_CUDA(cudaStreamSynchronize(st));
            }
            h->hx_lay.assign(world, DistBlobLayout{~0ull, 0, 0});
            uint64_t total_bytes = 0;
            for (uint32_t d = 0; d < world; ++d)
            {
                const uint32_t np = hd[d], ne = hd[DIST_MAX_WORLD + d];
                if (!np) continue;
                h->hx_lay[d] = DistBlobLayout{total_bytes, np, ne};
                io->send_bytes[d] = dist_blob_bytes(np, ne);
                total_bytes += io->send_bytes[d];
            }
            h2d_vec(h->x_lay, h->hx_lay, st);
            h->x_send.reserve(std::max<uint64_t>(total_bytes, 16));
            if (total_bytes) DGE_CUDA(cudaMemsetAsync(h->x_send.p, 0, total_bytes, st));
            if (n_pairs)
            {
                k_dist_pack_heads<<<grid_for(n_pairs, 256), 256, 0, st>>>(h->x_pair_child.as<uint32_t>(), h->x_pair_gi.as<uint32_t>(), n_pairs, rows, h->x_all.as<SelfRec>(),
                                                                         h->x_rank_off.as<uint32_t>(), world, me, h->x_self_idx.as<uint32_t>(), h->n_pc,
                                                                         h->x_lay.as<DistBlobLayout>(), cur_pairs, cur_entries, h->x_send.as<unsigned char>(),
                                                                         h->x_pair_pos.as<uint32_t>(), h->x_pair_eoff.as<uint32_t>(), h->x_local_jobs.as<PairJob>());
                k_dist_blob_headers<<<1, DIST_MAX_WORLD, 0, st>>>(h->x_lay.as<DistBlobLayout>(), world, h->x_send.as<unsigned char>());
                k_dist_pack_lists<<<unsigned(std::min<uint32_t>(n_pairs, 148 * 16)), 128, 0, st>>>(h->x_pair_child.as<uint32_t>(), h->x_pair_gi.as<uint32_t>(), n_pairs, rows,
                                                                                                  h->x_rank_off.as<uint32_t>(), world, me, h->x_lay.as<DistBlobLayout>(),
                                                                                                  h->x_pair_eoff.as<uint32_t>(), h->ukey.as<uint64_t>(), h->uval.as<uint32_t>(),
                                                                                                  h->pc_u_start.as<uint32_t>(), gub, h->x_send.as<unsigned char>());
                k_intersect<<<n_pairs, 128, 0, st>>>(h->x_local_jobs.as<PairJob>(), n_pairs, h->ukey.as<uint64_t>(), h->pc_u_start.as<uint32_t>(), h->pc_slot.as<uint32_t>(),
                                                     gub, h->x_local_isect.as<uint32_t>());
                DGE_LAUNCH_CHECK();
                h->launches += 5;
            }
            DGE_CUDA(cudaStreamSynchronize(st));
            io->collective = DGE_DIST_ALLTOALL;
            io->send = h->x_send.p;
            h->dist_stage = 2;
            return DGE_OK;
        }

        if (h->dist_stage == 2)
        {   // ---- received pairs: |child ∩ candidate| for every one of them -> replies
            h->hx_rl.assign(world, DistBlobLayout{~0ull, 0, 0});
            uint64_t total = 0;
            for (uint32_t sr = 0; sr < world; ++sr) total += io->recv_bytes[sr];
            h->x_kept.reserve(std::max<uint64_t>(total, 16));
            if (total) DGE_CUDA(cudaMemcpyAsync(h->x_kept.p, io->recv, total, cudaMemcpyDeviceToDevice, st));
            std::vector<unsigned long long> hdr(size_t(2) * world, 0);
            uint64_t base = 0;
            for (uint32_t sr = 0; sr < world; ++sr)
            {
                if (io->recv_bytes[sr])
                {
                    if (io->recv_bytes[sr] < 16 || (io->recv_bytes[sr] & 15)) throw InvalidInput("malformed pair blob");
                    DGE_CUDA(cudaMemcpyAsync(&hdr[size_t(2) * sr], h->x_kept.as<unsigned char>() + base, 16, cudaMemcpyDeviceToHost, st));
                }
                base += io->recv_bytes[sr];
            }
            DGE_CUDA(cudaStreamSynchronize(st));
            std::vector<uint32_t> job_base(world + 1, 0), reply_base(world + 1, 0);
            base = 0;
            for (uint32_t sr = 0; sr < world; ++sr)
            {
                const uint32_t np = uint32_t(hdr[size_t(2) * sr]), ne = uint32_t(hdr[size_t(2) * sr + 1]);
                if (io->recv_bytes[sr])
                {
                    if (dist_blob_bytes(np, ne) != io->recv_bytes[sr]) throw InvalidInput("pair blob size does not match its header");
                    h->hx_rl[sr] = DistBlobLayout{base, np, ne};
                }
                job_base[sr + 1] = job_base[sr] + np;
                reply_base[sr + 1] = reply_base[sr] + ((np + 3u) & ~3u);
                io->send_bytes[sr] = uint64_t((np + 3u) & ~3u) * 4;
                base += io->recv_bytes[sr];
            }
            h2d_vec(h->x_rl, h->hx_rl, st); h2d_vec(h->x_job_base, job_base, st); h2d_vec(h->x_reply_base, reply_base, st);
            const uint32_t n_jobs = job_base[world];
            h->x_fjobs.reserve(std::max<size_t>(n_jobs, 1) * sizeof(ForeignJob)); h->x_fisect.reserve(std::max<size_t>(n_jobs, 1) * 4);
            h->x_reply.reserve(std::max<size_t>(reply_base[world], 4) * 4);
            DGE_CUDA(cudaMemsetAsync(h->x_reply.p, 0, std::max<size_t>(reply_base[world], 4) * 4, st));
            if (n_jobs)
            {
                k_dist_recv_jobs<<<grid_for(n_jobs, 256), 256, 0, st>>>(h->x_kept.as<unsigned char>(), h->x_rl.as<DistBlobLayout>(), h->x_job_base.as<uint32_t>(), world, rows,
                                                                       h->x_self_idx.as<uint32_t>(), h->n_self, h->n_pc, h->x_fjobs.as<ForeignJob>(), h->x_bad.as<int>());
                k_intersect_foreign<<<n_jobs, 128, 0, st>>>(h->x_fjobs.as<ForeignJob>(), n_jobs, h->x_kept.as<uint64_t>(), h->ukey.as<uint64_t>(), h->pc_u_start.as<uint32_t>(),
                                                            h->pc_slot.as<uint32_t>(), gub, h->x_fisect.as<uint32_t>());
                k_dist_reply<<<grid_for(n_jobs, 256), 256, 0, st>>>(h->x_fisect.as<uint32_t>(), h->x_job_base.as<uint32_t>(), h->x_reply_base.as<uint32_t>(), world,
                                                                   h->x_reply.as<uint32_t>());
                DGE_LAUNCH_CHECK();
                h->launches += 3;
            }
            DGE_CUDA(cudaStreamSynchronize(st));
            io->collective = DGE_DIST_ALLTOALL;
            io->send = h->x_reply.p;
            h->dist_stage = 3;
            return DGE_OK;
        }

        if (h->dist_stage == 3)
        {   // ---- replies -> best target per child -> commits to the owners of remote targets
            h->hx_reply_base_recv.assign(world + 1, 0);
            for (uint32_t d = 0; d < world; ++d)
            {
                const uint32_t np = h->hx_lay[d].base == ~0ull ? 0u : h->hx_lay[d].n_pairs;
                if (io->recv_bytes[d] != uint64_t((np + 3u) & ~3u) * 4) throw InvalidInput("reply size does not match the pairs sent");
                h->hx_reply_base_recv[d + 1] = h->hx_reply_base_recv[d] + ((np + 3u) & ~3u);
            }
            h2d_vec(h->x_reply_base, h->hx_reply_base_recv, st);
            const uint32_t n_pairs = h->n_pairs;
            h->x_best.reserve((n + 1) * 4); h->x_tie.reserve((n + 1) * 4); h->x_merged_cb.reserve(std::max<size_t>(n, 1) * 8);
            if (n_pairs)
                k_dist_collect<<<grid_for(n_pairs, 256), 256, 0, st>>>(h->x_pair_gi.as<uint32_t>(), h->x_pair_pos.as<uint32_t>(), n_pairs, h->x_rank_off.as<uint32_t>(), world,
                                                                      h->x_local_isect.as<uint32_t>(), static_cast<const uint32_t *>(io->recv), h->x_reply_base.as<uint32_t>(),
                                                                      h->x_pair_isect.as<uint32_t>());
            std::vector<uint32_t> tie;
            if (n)
            {
                k_dist_best2<<<g, 256, 0, st>>>(h->x_pair_off.as<uint32_t>(), h->x_pair_cnt.as<uint32_t>(), h->x_pair_gi.as<uint32_t>(), h->x_pair_isect.as<uint32_t>(), rows,
                                                h->x_all.as<SelfRec>(), n32, h->cfg.min_merge_fraction, h->x_best.as<int>(), h->x_tie.as<uint32_t>());
                DGE_LAUNCH_CHECK();
                d2h(tie, h->x_tie.p, n, st);
                DGE_CUDA(cudaStreamSynchronize(st));
            }
            // order-dependent ties of the best fraction: replay the reference's neighbour order on the host (rare)
            std::vector<uint32_t> tied;
            for (uint32_t i = 0; i < n; ++i) if (tie[i]) tied.push_back(i);
            h->n_dist_ties = tied.size();
            if (!tied.empty())
            {
                dist_fetch_host_tables(h);
                DistHostView view; view.build(h->hx_all);
                std::vector<uint32_t> pcnt, poff, pgi, pis;
                std::vector<int> best;
                d2h(pcnt, h->x_pair_cnt.p, n + 1, st); d2h(poff, h->x_pair_off.p, n + 1, st);
                d2h(pgi, h->x_pair_gi.p, n_pairs, st); d2h(pis, h->x_pair_isect.p, n_pairs, st); d2h(best, h->x_best.p, n, st);
                DGE_CUDA(cudaStreamSynchronize(st));
                for (uint32_t i : tied)
                {
                    const CellRow &c = h->hx_rows[i];
                    std::vector<long> ids = dist_exact_neighbours(h, view, c);
                    double max_frac = 0;
                    long best_gi = ids.empty() ? -1 : ids[0];
                    for (long id : ids)
                    {
                        uint32_t isect = 0;
                        for (uint32_t k = 0; k < pcnt[i]; ++k) if (pgi[poff[i] + k] == uint32_t(id)) isect = pis[poff[i] + k];
                        const double frac = 0.5 * isect * (1. / c.n_umis + 1. / h->hx_all[size_t(id)].umis); // RealBarcodesMergeStrategy.cpp:46-47
                        if (max_frac < frac) { max_frac = frac; best_gi = id; }
                    }
                    int bp = -1;
                    if (best_gi >= 0 && !(max_frac < h->cfg.min_merge_fraction))
                        for (uint32_t k = 0; k < pcnt[i]; ++k) if (pgi[poff[i] + k] == uint32_t(best_gi)) bp = int(poff[i] + k);
                    best[i] = bp;
                }
                DGE_CUDA(cudaMemcpyAsync(h->x_best.p, best.data(), n * 4, cudaMemcpyHostToDevice, st));
                DGE_CUDA(cudaStreamSynchronize(st));
            }
            // commits: capacity per destination = pairs sent to it
            std::vector<DistBlobLayout> clay(world, DistBlobLayout{0, 0, 0});
            uint64_t cbytes = 0;
            for (uint32_t d = 0; d < world; ++d)
            {
                const uint32_t np = h->hx_lay[d].base == ~0ull ? 0u : h->hx_lay[d].n_pairs;
                clay[d] = DistBlobLayout{cbytes, np, 0};
                cbytes += uint64_t(np) * sizeof(DistCommitRec);
            }
            h2d_vec(h->x_clay, clay, st);
            h->x_commit.reserve(std::max<uint64_t>(cbytes, 16)); h->x_commit_send.reserve(std::max<uint64_t>(cbytes, 16));
            h->x_ccur.reserve(DIST_MAX_WORLD * 4);
            DGE_CUDA(cudaMemsetAsync(h->x_ccur.p, 0, DIST_MAX_WORLD * 4, st));
            std::vector<uint32_t> ccur(DIST_MAX_WORLD, 0);
            if (n)
            {
                k_dist_decide<<<g, 256, 0, st>>>(h->x_nb_count.as<int>(), h->x_best.as<int>(), h->x_pair_gi.as<uint32_t>(), h->x_pair_pos.as<uint32_t>(), rows,
                                                 h->x_all.as<SelfRec>(), h->x_rank_off.as<uint32_t>(), world, me, h->x_self_idx.as<uint32_t>(), n32,
                                                 h->cell_state.as<CellState>(), h->x_merged_cb.as<unsigned long long>(), h->x_clay.as<DistBlobLayout>(),
                                                 h->x_ccur.as<uint32_t>(), h->x_commit.as<unsigned char>());
                k_dist_flag_remote<<<g, 256, 0, st>>>(h->cell_state.as<CellState>(), n32, h->df_ctr.as<DevFlowCounters>());
                DGE_LAUNCH_CHECK();
                h->launches += 4;
                DGE_CUDA(cudaMemcpyAsync(ccur.data(), h->x_ccur.p, DIST_MAX_WORLD * 4, cudaMemcpyDeviceToHost, st));
                DGE_CUDA(cudaStreamSynchronize(st));
            }
            uint64_t off = 0;
            for (uint32_t d = 0; d < world; ++d)
            {
                const uint64_t b = uint64_t(ccur[d]) * sizeof(DistCommitRec);
                if (b) DGE_CUDA(cudaMemcpyAsync(h->x_commit_send.as<unsigned char>() + off, h->x_commit.as<unsigned char>() + clay[d].base, b, cudaMemcpyDeviceToDevice, st));
                io->send_bytes[d] = b;
                off += b;
            }
            DGE_CUDA(cudaStreamSynchronize(st));
            io->collective = DGE_DIST_ALLTOALL;
            io->send = h->x_commit_send.p;
            h->dist_stage = 4;
            return DGE_OK;
        }

        if (h->dist_stage == 4)
        {   // ---- received commits: Stats::merge into the targets, foreign lists + local moves folded into U
            std::vector<DistBlobLayout> cl(world, DistBlobLayout{0, 0, 0});
            std::vector<uint32_t> c_base(world + 1, 0);
            uint64_t base = 0;
            for (uint32_t sr = 0; sr < world; ++sr)
            {
                if (io->recv_bytes[sr] % sizeof(DistCommitRec)) throw InvalidInput("malformed commit piece");
                cl[sr] = DistBlobLayout{base, uint32_t(io->recv_bytes[sr] / sizeof(DistCommitRec)), 0};
                c_base[sr + 1] = c_base[sr] + cl[sr].n_pairs;
                base += io->recv_bytes[sr];
            }
            const uint32_t n_commits = c_base[world];
            h2d_vec(h->x_cl, cl, st); h2d_vec(h->x_cbase, c_base, st);
            h->x_fmoves.reserve(std::max<size_t>(n_commits, 1) * sizeof(ForeignMove)); h->x_fsize.reserve((size_t(n_commits) + 1) * 4); h->x_foff.reserve((size_t(n_commits) + 1) * 4);
            CellState *cs = h->cell_state.as<CellState>();
            DevFlowCounters *ctr = h->df_ctr.as<DevFlowCounters>();
            if (n_commits)
                k_dist_apply_commits<<<grid_for(n_commits, 256), 256, 0, st>>>(static_cast<const unsigned char *>(io->recv), h->x_cl.as<DistBlobLayout>(), h->x_cbase.as<uint32_t>(),
                                                                              world, h->x_kept.as<unsigned char>(), h->x_rl.as<DistBlobLayout>(), rows, h->x_self_idx.as<uint32_t>(),
                                                                              cs, h->x_fsize.as<uint32_t>(), h->x_fmoves.as<ForeignMove>(), h->x_bad.as<int>());
            DGE_CUDA(cudaMemsetAsync(h->x_fsize.as<uint32_t>() + n_commits, 0, 4, st));
            device_exclusive_scan(h->x_fsize.as<uint32_t>(), h->x_foff.as<uint32_t>(), size_t(n_commits) + 1, h->scan_scratch.as<uint32_t>(), st, &h->launches);
            h->move_size.reserve((n + 1) * 4); h->move_off.reserve((n + 1) * 4);
            if (n) k_phase2_sizes<<<g, 256, 0, st>>>(rows, cs, n32, h->move_size.as<uint32_t>(), ctr);
            DGE_CUDA(cudaMemsetAsync(h->move_size.as<uint32_t>() + n, 0, 4, st));
            device_exclusive_scan(h->move_size.as<uint32_t>(), h->move_off.as<uint32_t>(), n + 1, h->scan_scratch.as<uint32_t>(), st, &h->launches);
            uint32_t f_total = 0, l_total = 0;
            int bad[4] = {0, 0, 0, 0};
            DevFlowCounters hc{};
            DGE_CUDA(cudaMemcpyAsync(&f_total, h->x_foff.as<uint32_t>() + n_commits, 4, cudaMemcpyDeviceToHost, st));
            DGE_CUDA(cudaMemcpyAsync(&l_total, h->move_off.as<uint32_t>() + n, 4, cudaMemcpyDeviceToHost, st));
            DGE_CUDA(cudaMemcpyAsync(bad, h->x_bad.p, 16, cudaMemcpyDeviceToHost, st));
            DGE_CUDA(cudaMemcpyAsync(&hc, ctr, sizeof(hc), cudaMemcpyDeviceToHost, st));
            DGE_CUDA(cudaStreamSynchronize(st));
            if (bad[0]) throw InvalidInput("inconsistent cross-rank merge messages");
            if (hc.n_chain) throw std::runtime_error("merge chain in the cross-rank whitelist merge");
            const uint64_t total = uint64_t(f_total) + l_total;
            if (total >= 0xFFFFFFF0ull) throw CapacityError("merge volume exceeds 2^32 entries");
            h->d_moves.reserve(std::max<size_t>(n, 1) * sizeof(MoveJob));
            if (n) k_phase2_apply<<<g, 256, 0, st>>>(rows, cs, n32, h->move_off.as<uint32_t>(), h->n_pc, h->d_moves.as<MoveJob>(), ctr);
            if (total)
            {
                h->mkeys.reserve(total * 8); h->mvals.reserve(total * 4);
                if (l_total)
                    k_gather_relabel<<<grid_for(n, 1, 148 * 16), 256, 0, st>>>(h->d_moves.as<MoveJob>(), n32, h->ukey.as<uint64_t>(), h->uval.as<uint32_t>(),
                                                                              h->pc_u_start.as<uint32_t>(), gub, h->mkeys.as<uint64_t>(), h->mvals.as<uint32_t>());
                if (f_total)
                    k_dist_gather_foreign<<<unsigned(std::min<uint32_t>(n_commits, 148 * 16)), 256, 0, st>>>(h->x_fmoves.as<ForeignMove>(), h->x_foff.as<uint32_t>(), n_commits, l_total,
                                                                                                            h->x_kept.as<unsigned char>(), gub, h->mkeys.as<uint64_t>(),
                                                                                                            h->mvals.as<uint32_t>());
                DGE_LAUNCH_CHECK();
                h->launches += 3;
                if (!h->slot_pc_built) build_slot_pc(h);
                apply_moved(h, total);
            }
            DGE_CUDA(cudaEventRecord(h->ev[4], st));
            DGE_CUDA(cudaStreamSynchronize(st));
            h->dist_done = true;
            h->n_unresolved = 0;
            io->collective = DGE_DIST_DONE;
            h->dist_stage = 5;
            return DGE_OK;
        }
        throw std::runtime_error("dge_dist_step called after it reported DGE_DIST_DONE");
    });
}

int dge_umi_first_size(dge_handle *h, size_t *n_entries)
{
    if (!h || !n_entries) return fail(h, DGE_ERR_INVALID, "null argument");
    return guarded(h, [&] {
        ensure_device(h);
        *n_entries = h->track_umi_first ? size_t(1) << h->kl.ub : 0;
        return DGE_OK;
    });
}

int dge_umi_first_export(dge_handle *h, uint32_t *dst_device)
{
    if (!h || !dst_device) return fail(h, DGE_ERR_INVALID, "null argument");
    return guarded(h, [&]() -> int {
        ensure_device(h);
        if (!h->track_umi_first) return fail(h, DGE_ERR_STATE, "the configured strategies do not track UMI first-seen order");
        DGE_CUDA(cudaMemcpyAsync(dst_device, h->umi_first.p, (size_t(1) << h->kl.ub) * 4, cudaMemcpyDeviceToDevice, h->stream));
        DGE_CUDA(cudaStreamSynchronize(h->stream));
        return DGE_OK;
    });
}

int dge_umi_first_import(dge_handle *h, const uint32_t *src_device)
{
    if (!h || !src_device) return fail(h, DGE_ERR_INVALID, "null argument");
    return guarded(h, [&]() -> int {
        ensure_device(h);
        if (!h->track_umi_first) return fail(h, DGE_ERR_STATE, "the configured strategies do not track UMI first-seen order");
        if (h->state == 2) return fail(h, DGE_ERR_STATE, "import the UMI first-seen table before merge_and_filter");
        DGE_CUDA(cudaMemcpyAsync(h->umi_first.p, src_device, (size_t(1) << h->kl.ub) * 4, cudaMemcpyDeviceToDevice, h->stream));
        DGE_CUDA(cudaStreamSynchronize(h->stream));
        return DGE_OK;
    });
}

int dge_get_summary(dge_handle *h, dge_summary *out)
{
    if (!h || !out) return fail(h, DGE_ERR_INVALID, "null argument");
    if (h->state == 0) return fail(h, DGE_ERR_STATE, "You must initialize container");
    std::memset(out, 0, sizeof(*out));
    out->n_reads = h->n_reads;
    out->total_cells_number = h->total_cells;
    out->real_cells_number = h->sum_real;
    out->filtered_cells_number = h->sum_filtered;
    out->n_genes_seen = h->sum_genes_seen;
    out->n_umigs = h->n_u;
    out->intergenic_reads = h->counters.intergenic;
    out->has_exon_reads = h->counters.has_exon;
    out->has_intron_reads = h->counters.has_intron;
    out->has_not_annotated_reads = h->counters.has_not_annotated;
    out->cm_nnz = h->cm.built ? h->cm.nnz : 0;
    out->cm_raw_nnz = h->cm_raw.built ? h->cm_raw.nnz : 0;
    out->n_merged = h->n_merged;
    out->n_excluded = h->n_excluded;
    out->n_unresolved = h->n_unresolved;
    out->n_umis_merged = h->n_umis_merged;
    out->n_umi_segments_replayed = h->n_umi_segments_replayed;
    out->n_cb_merge_replayed = h->n_simple_replayed;
    out->n_host_flow = h->n_host_fallback;
    return DGE_OK;
}

int dge_get_timings(dge_handle *h, dge_timings *out)
{
    if (!h || !out) return fail(h, DGE_ERR_INVALID, "null argument");
    *out = h->timings;
    return DGE_OK;
}

int dge_get_cells(dge_handle *h, int which, dge_cell_info *out, size_t capacity, size_t *n_out)
{
    if (!h || !n_out) return fail(h, DGE_ERR_INVALID, "null argument");
    if (h->state == 0) return fail(h, DGE_ERR_STATE, "You must initialize container");
    return guarded(h, [&] {
        DGE_CUDA(cudaSetDevice(h->cfg.device));
        materialize_host(h);
        if (which == DGE_CELLS_ALL)
        {
            AllCells all = collect_all_cells(h);
            *n_out = all.info.size();
            if (out && capacity >= all.info.size()) std::copy(all.info.begin(), all.info.end(), out);
            return int(DGE_OK);
        }
        std::vector<uint32_t> ids;
        if (which == DGE_CELLS_REAL) { for (uint32_t i = 0; i < h->real.size(); ++i) if (h->real[i].real) ids.push_back(i); }
        else if (which == DGE_CELLS_FILTERED) ids = h->filtered;
        else return fail(h, DGE_ERR_INVALID, "unknown cell class");
        *n_out = ids.size();
        if (out && capacity >= ids.size())
        {
            std::unordered_map<uint32_t, int32_t> pos;
            for (size_t k = 0; k < ids.size(); ++k) pos.emplace(ids[k], int32_t(k));
            for (size_t k = 0; k < ids.size(); ++k)
            {
                fill_info(h, h->real[ids[k]], out[k]);
                auto it = pos.find(uint32_t(h->real[ids[k]].target));
                out[k].merge_target = it == pos.end() ? -1 : it->second;
            }
        }
        return int(DGE_OK);
    });
}

int dge_get_matrix(dge_handle *h, int which, int64_t *indptr, int32_t *gene_ids, int32_t *values, size_t *n_cols, size_t *nnz)
{
    if (!h || !n_cols || !nnz) return fail(h, DGE_ERR_INVALID, "null argument");
    if (h->state != 2) return fail(h, DGE_ERR_STATE, "matrices exist after merge_and_filter");
    if (which != DGE_MATRIX_CM && which != DGE_MATRIX_CM_RAW) return fail(h, DGE_ERR_INVALID, "unknown matrix");
    return guarded(h, [&] {
        DGE_CUDA(cudaSetDevice(h->cfg.device));
        MatrixDev &m = which == DGE_MATRIX_CM ? h->cm : h->cm_raw;
        *n_cols = m.n_cols; *nnz = m.nnz;
        std::vector<uint32_t> ip;
        if (indptr)
        {
            d2h(ip, m.indptr.p, m.n_cols + 1, h->stream);
            DGE_CUDA(cudaStreamSynchronize(h->stream));
            for (size_t i = 0; i <= m.n_cols; ++i) indptr[i] = int64_t(ip[i]);
        }
        if (gene_ids && m.nnz) DGE_CUDA(cudaMemcpyAsync(gene_ids, m.gene.p, m.nnz * 4, cudaMemcpyDeviceToHost, h->stream));
        if (values && m.nnz) DGE_CUDA(cudaMemcpyAsync(values, m.val.p, m.nnz * 4, cudaMemcpyDeviceToHost, h->stream));
        DGE_CUDA(cudaStreamSynchronize(h->stream));
        return int(DGE_OK);
    });
}

int dge_get_matrix_marks(dge_handle *h, uint32_t query_mark_mask, int64_t *indptr, int32_t *gene_ids, int32_t *values, size_t *n_cols, size_t *nnz)
{
    if (!h || !n_cols || !nnz) return fail(h, DGE_ERR_INVALID, "null argument");
    if (h->state != 2) return fail(h, DGE_ERR_STATE, "matrices exist after merge_and_filter");
    if (query_mark_mask & ~0xFEu) return fail(h, DGE_ERR_INVALID, "query_mark_mask: bits 1..7 only");
    return guarded(h, [&] {
        DGE_CUDA(cudaSetDevice(h->cfg.device));
        cudaStream_t st = h->stream;
        MatrixDev &m = h->cm_marks;
        if (!m.built || h->cm_marks_mask != query_mark_mask)
        {
            if (!h->cm.built) throw std::runtime_error("internal: the filtered matrix was not built");
            h->cg_marks.reserve((size_t(h->n_cg) + 2) * 4);
            // one more row than n_cg: the empty sentinel cell of barcodes without UMIs points at [n_cg, n_cg)
            DGE_CUDA(cudaMemsetAsync(h->cg_marks.p, 0, (size_t(h->n_cg) + 2) * 4, st));
            if (h->n_cg)
                k_cg_mark_values<<<grid_for(size_t(h->n_cg) * 8, 256), 256, 0, st>>>(h->uval.as<uint32_t>(), h->cg_start.as<uint32_t>(), h->n_cg, query_mark_mask,
                                                                                   h->cfg.reads_output ? 1 : 0, h->cg_marks.as<uint32_t>());
            DGE_LAUNCH_CHECK();
            ++h->launches;
            build_matrix_cols(h, m, h->cm.cols_used, h->cm.n_cols, true, h->cg_marks.as<uint32_t>());
            h->cm_marks_mask = query_mark_mask;
        }
        *n_cols = m.n_cols; *nnz = m.nnz;
        std::vector<uint32_t> ip;
        if (indptr)
        {
            d2h(ip, m.indptr.p, m.n_cols + 1, st);
            DGE_CUDA(cudaStreamSynchronize(st));
            for (size_t i = 0; i <= m.n_cols; ++i) indptr[i] = int64_t(ip[i]);
        }
        if (gene_ids && m.nnz) DGE_CUDA(cudaMemcpyAsync(gene_ids, m.gene.p, m.nnz * 4, cudaMemcpyDeviceToHost, st));
        if (values && m.nnz) DGE_CUDA(cudaMemcpyAsync(values, m.val.p, m.nnz * 4, cudaMemcpyDeviceToHost, st));
        DGE_CUDA(cudaStreamSynchronize(st));
        return int(DGE_OK);
    });
}

int dge_get_gene_order(dge_handle *h, int32_t *gene_ids, size_t capacity, size_t *n_out)
{
    if (!h || !n_out) return fail(h, DGE_ERR_INVALID, "null argument");
    if (h->state == 0) return fail(h, DGE_ERR_STATE, "You must initialize container");
    return guarded(h, [&] {
        materialize_host(h);
        *n_out = h->gene_order.size();
        if (gene_ids && capacity >= h->gene_order.size()) std::copy(h->gene_order.begin(), h->gene_order.end(), gene_ids);
        return int(DGE_OK);
    });
}

int dge_get_merge_pairs(dge_handle *h, uint64_t *from, uint64_t *to, size_t capacity, size_t *n_out)
{
    if (!h || !n_out) return fail(h, DGE_ERR_INVALID, "null argument");
    if (h->state != 2) return fail(h, DGE_ERR_STATE, "merge targets exist after merge_and_filter");
    { const int rc = guarded(h, [&] { materialize_host(h); return int(DGE_OK); }); if (rc != DGE_OK) return rc; }
    size_t n = 0;
    for (auto const &c : h->real) n += (uint32_t(c.target) != uint32_t(&c - h->real.data())) || c.merged_to_cb != EMPTY64;
    *n_out = n;
    if (from && to && capacity >= n)
    {
        size_t k = 0;
        for (uint32_t i = 0; i < h->real.size(); ++i)
        {
            const HostCell &c = h->real[i];
            if (c.merged_to_cb != EMPTY64) { from[k] = c.cb; to[k] = c.merged_to_cb; ++k; }
            else if (uint32_t(c.target) != i) { from[k] = c.cb; to[k] = h->real[size_t(c.target)].cb; ++k; }
        }
    }
    return DGE_OK;
}

int dge_get_umi_merge_targets(dge_handle *h, uint64_t *cell_barcodes, int32_t *gene_ids, uint32_t *source_umis, uint32_t *target_umis, uint8_t *created,
                              size_t capacity, size_t *n_out)
{
    if (!h || !n_out) return fail(h, DGE_ERR_INVALID, "null argument");
    if (h->state != 2) return fail(h, DGE_ERR_STATE, "UMI merge targets exist after merge_and_filter");
    if (!h->cfg.save_umi_merge_targets) return fail(h, DGE_ERR_STATE, "the handle was created without dge_config.save_umi_merge_targets");
    return guarded(h, [&] {
        *n_out = h->umi_mt.size();
        if (capacity < h->umi_mt.size() || !cell_barcodes || !gene_ids || !source_umis || !target_umis) return int(DGE_OK);
        std::sort(h->umi_mt.begin(), h->umi_mt.end(), [](const dge_handle::UmiMergeTarget &x, const dge_handle::UmiMergeTarget &y) {
            return x.cb != y.cb ? x.cb < y.cb : x.gene != y.gene ? x.gene < y.gene : x.src < y.src;
        });
        for (size_t k = 0; k < h->umi_mt.size(); ++k)
        {
            cell_barcodes[k] = h->umi_mt[k].cb; gene_ids[k] = int32_t(h->umi_mt[k].gene);
            source_umis[k] = h->umi_mt[k].src; target_umis[k] = h->umi_mt[k].dst;
            if (created) created[k] = uint8_t(h->umi_mt[k].created);
        }
        return int(DGE_OK);
    });
}

int dge_get_merge_events(dge_handle *h, uint64_t *from, uint64_t *to, size_t capacity, size_t *n_out)
{
    if (!h || !n_out) return fail(h, DGE_ERR_INVALID, "null argument");
    if (h->state != 2) return fail(h, DGE_ERR_STATE, "merge events exist after merge_and_filter");
    if (h->cfg.sharded) return fail(h, DGE_ERR_STATE, "merge events are not kept on sharded handles");
    return guarded(h, [&] {
        materialize_host(h);
        *n_out = h->merge_events.size();
        if (!from || !to || capacity < h->merge_events.size()) return int(DGE_OK);
        for (size_t k = 0; k < h->merge_events.size(); ++k)
        {
            from[k] = h->real[h->merge_events[k].first].cb;
            to[k] = h->real[h->merge_events[k].second].cb;
        }
        return int(DGE_OK);
    });
}

int dge_get_umigs(dge_handle *h, int which, uint32_t *cell_index, int32_t *gene_ids, uint32_t *umis, uint32_t *read_counts,
                  uint8_t *marks, size_t capacity, size_t *n_out)
{
    if (!h || !n_out) return fail(h, DGE_ERR_INVALID, "null argument");
    if (h->state == 0) return fail(h, DGE_ERR_STATE, "You must initialize container");
    return guarded(h, [&] {
        DGE_CUDA(cudaSetDevice(h->cfg.device));
        materialize_host(h);
        std::vector<uint32_t> pcs;
        if (which == DGE_CELLS_ALL) { AllCells all = collect_all_cells(h); pcs = all.pc; }
        else if (which == DGE_CELLS_REAL) { for (auto const &c : h->real) if (c.real) pcs.push_back(c.pc); }
        else if (which == DGE_CELLS_FILTERED) { for (uint32_t i : h->filtered) pcs.push_back(h->real[i].pc); }
        else return fail(h, DGE_ERR_INVALID, "unknown cell class");
        std::vector<uint32_t> pc_u;
        d2h(pc_u, h->pc_u_start.p, size_t(h->n_pc) + 1, h->stream);
        DGE_CUDA(cudaStreamSynchronize(h->stream));
        size_t total = 0;
        for (uint32_t pc : pcs) if (pc != NONE32) total += pc_u[pc + 1] - pc_u[pc];
        *n_out = total;
        if (capacity < total || total == 0) return int(DGE_OK);
        std::vector<uint64_t> uk;
        std::vector<uint32_t> uv;
        d2h(uk, h->ukey.p, h->n_u, h->stream);
        d2h(uv, h->uval.p, h->n_u, h->stream);
        DGE_CUDA(cudaStreamSynchronize(h->stream));
        size_t k = 0;
        for (size_t ci = 0; ci < pcs.size(); ++ci)
        {
            if (pcs[ci] == NONE32) continue;
            for (uint32_t i = pc_u[pcs[ci]]; i < pc_u[pcs[ci] + 1]; ++i, ++k)
            {
                if (cell_index) cell_index[k] = uint32_t(ci);
                if (gene_ids) gene_ids[k] = int32_t(h->kl.gene_of_ukey(uk[i]));
                if (umis)
                {
                    const uint32_t u = h->kl.umi_of_ukey(uk[i]);
                    umis[k] = umi_is_n(h, u) ? (0x80000000u | (u & ((1u << (h->kl.ub - 1)) - 1))) : u; // DGE_UMI_N_BIT | index into the N-UMI list
                }
                if (read_counts) read_counts[k] = uv[i] & VAL_COUNT_MASK;
                if (marks) marks[k] = uint8_t(uv[i] >> VAL_MARK_SHIFT);
            }
        }
        return int(DGE_OK);
    });
}

int dge_collisions_adjusted_sizes(int device, const double *umi_probabilities, size_t n_umis, size_t max_gene_expression,
                                  uint64_t *adjusted_sizes, uint32_t *exact_rerun)
{
    if (exact_rerun) *exact_rerun = 0;
    if ((!umi_probabilities && n_umis) || (!adjusted_sizes && max_gene_expression)) return fail(nullptr, DGE_ERR_INVALID, "null argument");
    if (max_gene_expression == 0) return DGE_OK;
    return guarded(nullptr, [&] {
        DGE_CUDA(cudaSetDevice(device));
        cudaStream_t st = nullptr;
        DGE_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamDestroy(s); } } guard{st};
        DevBuf p, adj;
        p.reserve(std::max<size_t>(n_umis, 1) * 8); adj.reserve(max_gene_expression * 8);
        DGE_CUDA(cudaMemcpyAsync(p.p, umi_probabilities, n_umis * 8, cudaMemcpyHostToDevice, st));
        collisions_adjusted_device(st, p.as<double>(), n_umis, max_gene_expression, adj.as<unsigned long long>(), exact_rerun);
        DGE_CUDA(cudaMemcpyAsync(adjusted_sizes, adj.p, max_gene_expression * 8, cudaMemcpyDeviceToHost, st));
        DGE_CUDA(cudaStreamSynchronize(st));
        return int(DGE_OK);
    });
}

unsigned dge_edit_distance(const char *s1, const char *s2, int skip_n, unsigned max_ed) { return edit_distance_ref(s1, s2, skip_n != 0, max_ed); }

unsigned dge_hamming_distance(const char *s1, const char *s2, int skip_n)
{
    if (std::strlen(s1) != std::strlen(s2)) return 0xFFFFFFFFu;
    return hamming_distance_ref(s1, s2, skip_n != 0);
}

int dge_whitelist_shape(dge_handle *h, uint32_t *n_parts, uint32_t *part_sizes, uint32_t *part_lengths, size_t capacity)
{
    if (!h || !n_parts) return fail(h, DGE_ERR_INVALID, "null argument");
    *n_parts = uint32_t(h->wl.parts.size());
    for (size_t p = 0; p < h->wl.parts.size() && p < capacity; ++p)
    {
        if (part_sizes) part_sizes[p] = uint32_t(h->wl.parts[p].size());
        if (part_lengths) part_lengths[p] = uint32_t(h->wl.parts[p][0].size());
    }
    return DGE_OK;
}

int dge_whitelist_token(dge_handle *h, uint32_t part, uint32_t index, char *out, size_t capacity)
{
    if (!h || !out) return fail(h, DGE_ERR_INVALID, "null argument");
    if (part >= h->wl.parts.size() || index >= h->wl.parts[part].size()) return fail(h, DGE_ERR_INVALID, "index out of range");
    const std::string &t = h->wl.parts[part][index];
    if (capacity < t.size() + 1) return fail(h, DGE_ERR_INVALID, "buffer too small");
    std::memcpy(out, t.c_str(), t.size() + 1);
    return DGE_OK;
}

} // extern "C"
