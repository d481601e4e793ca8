/* dropest_b200.h -- C ABI of the B200-native dropEst count-matrix hot path.
 *
 * dropEst has no plugin/FFI ABI of its own: the seam is the C++ class Estimation::CellsDataContainer
 * (reference Estimation/CellsDataContainer.h:82-122).  This header is the flat C boundary the host-side C++ facade
 * (dropest_b200/host/CellsDataContainer.h, same class/method names as the reference) binds to; each entry point cites the
 * reference interface it replaces.  Plain pointers and sizes only; no exceptions cross the ABI (status codes +
 * dge_last_error()); one handle = one caller thread at a time, exactly like the reference container.
 *
 * All computation happens in hand-written sm_100a CUDA kernels (dropest_b200/csrc).  There is NO CPU fallback: every call
 * fails with DGE_ERR_CUDA when no CUDA device is usable.
 */
#ifndef DROPEST_B200_H
#define DROPEST_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DGE_ABI_VERSION 2

/* ---- the per-read record --------------------------------------------------------------------------------------------
 * Replaces Estimation::ReadInfo (reference Estimation/ReadInfo.h:9-24) + Tools::ReadParameters (Tools/ReadParameters.h:9-50)
 * on the fill path.  16 bytes, little endian:
 *   key      bits[63:24] cell barcode, 2 bits/base (A=0 C=1 G=2 T=3), first base most significant, right aligned (<= 20 bp)
 *            bits[23:0]  UMI, same coding (<= 12 bp)
 *   gene     bits[23:0]  gene id in [0, n_genes) or DGE_NO_GENE (read has no gene: "intergenic", CellsDataContainer.cpp:73-78)
 *            bits[26:24] UMI::Mark bits of the read (reference Estimation/UMI.h:16-22): 1 not-annotated, 2 exon, 4 intron
 *            bit 27      DGE_FLAG_UMI_N: the UMI contains 'N'; key bits[23:0] hold the index of the UMI string in the caller's N-UMI list
 *            bit 28      DGE_FLAG_CB_N : the barcode contains 'N'; key bits[63:24] hold the index of the barcode string in the N-barcode list
 *                        (both need dge_config.allow_n = 1 and the lists passed with dge_set_n_strings; equal strings = equal indices.
 *                        The reference keeps such reads as barcodes / UMIs of their own and repairs N-UMIs of real cells in
 *                        MergeUMIsStrategySimple, Merge/UMIs/MergeUMIsStrategySimple.cpp:21-112)
 *            bits[31:29] reserved, must be 0
 *   read_idx global 0-based position of the read in the input stream.  First-seen order of barcodes and genes
 *            (cell ids, StringIndexer ids) is derived from it, so records may be passed in ANY order / any batching.
 */
typedef struct dge_record16 {
    uint64_t key;
    uint32_t gene;
    uint32_t read_idx;
} dge_record16;

#define DGE_NO_GENE 0xFFFFFFu
#define DGE_MARK_NOT_ANNOTATED 1u
#define DGE_MARK_EXON 2u
#define DGE_MARK_INTRON 4u
#define DGE_FLAG_UMI_N (1u << 27)
#define DGE_FLAG_CB_N (1u << 28)
/* how the query surface reports them: dge_cell_info.barcode = DGE_CB_N_BIT | index, dge_get_umigs umis[] = DGE_UMI_N_BIT | index */
#define DGE_CB_N_BIT (1ull << 40)
#define DGE_UMI_N_BIT (1u << 31)

/* status codes */
enum {
    DGE_OK = 0,
    DGE_ERR_INVALID = 1,   /* bad argument / bad config                                              */
    DGE_ERR_STATE = 2,     /* call order violated (mirrors the reference's runtime_error throws)      */
    DGE_ERR_CUDA = 3,      /* CUDA runtime failure or no device                                       */
    DGE_ERR_CAPACITY = 4,  /* a device table overflowed (too many distinct barcodes for the key width) */
    DGE_ERR_IO = 5,        /* whitelist file unreadable / malformed                                   */
    DGE_ERR_INTERNAL = 6
};

/* CB merge strategy, selected like MergeStrategyFactory::get_cb_strat (reference Merge/MergeStrategyFactory.cpp:61-103) */
enum {
    DGE_MERGE_NONE = 0,           /* DummyMergeStrategy            (no -m)                         */
    DGE_MERGE_REAL = 1,           /* RealBarcodesMergeStrategy     (-m, barcodes_file set)         */
    DGE_MERGE_SIMPLE = 2,         /* SimpleMergeStrategy           (-m, no barcodes_file)          */
    DGE_MERGE_POISSON_REAL = 3,   /* PoissonRealBarcodesMergeStrategy (-M, barcodes_file set)      */
    DGE_MERGE_POISSON_SIMPLE = 4, /* PoissonSimpleMergeStrategy    (-M, no barcodes_file)          */
    DGE_MERGE_ALL = 5             /* MergeAllMergeStrategy         (merge_type=all)                */
};

enum { DGE_BARCODES_CONST = 0, DGE_BARCODES_INDROP = 1 }; /* MergeStrategyFactory.cpp:113-126 */
enum { DGE_UMI_MERGE_SIMPLE = 0, DGE_UMI_MERGE_DIRECTIONAL = 1 }; /* MergeStrategyFactory.cpp:105-111 */

/* Configuration = constructor arguments of CellsDataContainer (CellsDataContainer.h:82-85) + the keys
 * MergeStrategyFactory reads (MergeStrategyFactory.cpp:23-59), same defaults (dge_config_default). */
typedef struct dge_config {
    uint32_t abi_version;       /* DGE_ABI_VERSION */
    int32_t  device;            /* CUDA device ordinal */
    uint32_t cb_len;            /* barcode length in bases (<= 20) */
    uint32_t umi_len;           /* UMI length in bases (<= 12) */
    uint32_t n_genes;           /* gene ids are in [0, n_genes) (<= 2^24 - 1) */
    uint32_t merge_type;        /* DGE_MERGE_* */
    uint32_t barcodes_type;     /* DGE_BARCODES_* */
    uint32_t umi_merge_type;    /* DGE_UMI_MERGE_* */
    uint32_t min_genes_before_merge; /* default 10 */
    uint32_t min_genes_after_merge;  /* default 10; effective value is max(after, before) (MergeStrategyAbstract.cpp:8-11) */
    uint32_t max_cb_merge_edit_distance;
    uint32_t max_umi_merge_edit_distance; /* default 1 */
    double   min_merge_fraction;     /* default 0.2 */
    double   max_merge_prob;         /* default 1e-4 */
    double   max_real_merge_prob;    /* default 1e-7 */
    double   umi_merge_mult;         /* default 2 */
    uint32_t query_mark_mask;   /* bit m (m in 1..7) set <=> an accumulated UMI mark equal to m matches (UMI.cpp:76-85).
                                   Default "eEBA" = marks {2,3,6,7} = 0xCC (CellsDataContainer.cpp:17, UMI.cpp:123-154) */
    int32_t  max_cells;         /* -C: keep the top N filtered cells; <= 0 keeps all (CellsDataContainer.cpp:268-272) */
    uint32_t reads_output;      /* -R: matrix values are read counts instead of UMI counts (ResultsPrinter.cpp:345) */
    uint32_t sharded;           /* 1 = this handle holds one barcode-hash shard of a multi-GPU run: the whitelist merge runs across
                                   ranks with dge_dist_step (exact) before dge_merge_and_filter; 0 = the handle sees every barcode */
    const char *barcodes_file;  /* whitelist in the reference's own file format (BarcodesParser.cpp:117-144); NULL/"" = none */
    uint64_t max_barcodes_hint; /* upper bound on distinct barcodes, 0 = automatic */
    uint32_t allow_n;           /* 1 = records may carry DGE_FLAG_UMI_N / DGE_FLAG_CB_N.  The grouping key then spends one more bit on the UMI field
                                   (and at least 21): with 12-base UMIs and > 16 k genes the barcode table halves (2^21 slots) */
    uint32_t save_umi_merge_targets; /* 1 = keep, per (cell, gene), which UMI every UMI merged by the UMI merge strategy went to: the
                                   `save_umi_merge_targets` constructor argument (CellsDataContainer.h:83, Gene.cpp:54-57); read back with
                                   dge_get_umi_merge_targets.  Was reserved0 (0 = off): the layout of ABI version 2 is unchanged */
} dge_config;

typedef struct dge_handle dge_handle;

/* Global / per-stage numbers.  Counter names follow the reference getters (CellsDataContainer.h:113-118). */
typedef struct dge_summary {
    uint64_t n_reads;               /* records passed to dge_add_batch* */
    uint64_t total_cells_number;    /* distinct barcodes seen                 (total_cells_number)        */
    uint64_t real_cells_number;     /*                                         (real_cells_number)         */
    uint64_t filtered_cells_number; /* filtered_cells().size()                                            */
    uint64_t n_genes_seen;          /* gene_indexer().values().size()                                     */
    uint64_t n_umigs;               /* distinct (cell, gene, UMI) currently held                          */
    uint64_t intergenic_reads;      /* intergenic_reads_num()        */
    uint64_t has_exon_reads;        /* has_exon_reads_num()          */
    uint64_t has_intron_reads;      /* has_intron_reads_num()        */
    uint64_t has_not_annotated_reads; /* has_not_annotated_reads_num() */
    uint64_t cm_nnz;                /* non-zeros of the filtered matrix `cm`   */
    uint64_t cm_raw_nnz;            /* non-zeros of `cm_raw`                    */
    uint64_t n_merged;              /* cells merged into another cell (MergeStrategyBase.cpp:53) */
    uint64_t n_excluded;            /* cells excluded by the merge      (MergeStrategyBase.cpp:54) */
    uint64_t n_unresolved;          /* sharded handles whose merge_and_filter ran WITHOUT dge_dist_step: cells left unmerged because their
                                       candidates may live on another shard (0 after dge_dist_step) */
    uint64_t n_umis_merged;         /* UMIs merged into another UMI by the UMI merge strategy (MergeUMIsStrategyDirectional.cpp:43) */
    uint64_t n_umi_segments_replayed; /* (cell, gene) segments whose UMI merge was replayed on the host for exact tie order */
    uint64_t n_cb_merge_replayed;   /* SimpleMergeStrategy: base cells whose target was replayed on the host (near-ties of the top fraction) */
    uint64_t n_host_flow;           /* 1 when merge_and_filter left the device-resident flow for the host logic (far distance classes,
                                       order-dependent ties of the best fraction: exactness first), 0 when everything ran on the device */
} dge_summary;

/* Per-cell row returned by dge_get_cells; one per requested cell, in the requested order. */
typedef struct dge_cell_info {
    uint64_t barcode;        /* 2-bit packed, same coding as dge_record16.key >> 24                */
    uint32_t first_read_idx; /* smallest read_idx of the barcode (defines the reference's cell id) */
    uint32_t flags;          /* bit0 is_real, bit1 is_merged, bit2 is_excluded (Cell.cpp:110-128)  */
    int32_t  n_genes;        /* Cell::size()                                                       */
    int32_t  umis_stat;      /* Cell::umis_number()  = Stats TOTAL_UMIS_PER_CB (a counter) */
    int32_t  reads_stat;     /* Stats TOTAL_READS_PER_CB                                           */
    int32_t  requested_genes_num;
    int32_t  requested_umis_num;
    int32_t  merge_target;   /* index INTO THE SAME RETURNED LIST of the cell this one was merged into, or own index */
} dge_cell_info;

#define DGE_CELL_REAL 1u
#define DGE_CELL_MERGED 2u
#define DGE_CELL_EXCLUDED 4u

enum { DGE_CELLS_ALL = 0, DGE_CELLS_REAL = 1, DGE_CELLS_FILTERED = 2 };
enum { DGE_MATRIX_CM = 0, DGE_MATRIX_CM_RAW = 1 };

/* Per-stage device time of the last dge_set_initialized + dge_merge_and_filter, CUDA-event timed on the handle's stream. */
typedef struct dge_timings {
    float ms_fill;        /* records -> sorted distinct (cell,gene,UMI) + per-cell tables (add_record work)  */
    float ms_init;        /* set_initialized: requested sizes, real/filtered cells                            */
    float ms_merge;       /* CB merge phase 1 + 2 + applying merges                                           */
    float ms_finish;      /* UMI merge, final sizes/filter, matrices                                          */
    float ms_total;
    float ms_dedup_kernel; /* sub-bucket sort + dedup (k_sort_dedup size classes + hash tail), summed over its launches */
    uint32_t n_kernel_launches;
    uint32_t n_dedup_launches;
    float ms_fill_kernel; /* k_fill_pipe (the fill kernel), summed over the batches of this run (CUDA events on the launching stream) */
    uint32_t n_fill_launches;
} dge_timings;

/* ---- lifecycle ------------------------------------------------------------------------------------------------------ */

/* Fill `cfg` with the reference defaults (MergeStrategyFactory.cpp:23-59; query marks "eEBA"). */
void dge_config_default(dge_config *cfg);

/* = CellsDataContainer::CellsDataContainer (CellsDataContainer.cpp:20-37) + MergeStrategyFactory strategy selection. */
int dge_create(const dge_config *cfg, dge_handle **out);
void dge_destroy(dge_handle *h);

/* Last error message of this handle (or of a failed dge_create when h == NULL). Never NULL. */
const char *dge_last_error(const dge_handle *h);

/* ---- fill: replaces CellsDataContainer::add_record (CellsDataContainer.cpp:59-88), n reads per call ------------------
 * dge_add_batch        : `recs` is HOST memory; copied to the device asynchronously (pinned staging); no ownership taken.
 * dge_add_batch_device : `recs` is DEVICE memory on cfg.device; referenced, NOT copied -- it must stay valid and unmodified
 *                        until dge_set_initialized returns.
 * dge_add_batch_soa    : HOST memory, structure of arrays: keys[i] = dge_record16.key, genes[i] = dge_record16.gene, and
 *                        read_idx = first_read_idx + i (the stream position is implicit when reads arrive in stream order, which is
 *                        how BamProcessor::save_read calls add_record): 12 bytes per read cross PCIe instead of 16.
 * Host batches are staged in slices on a copy stream, so the copy of one slice overlaps the fill kernel of the previous one; the
 * host arrays may be reused as soon as the call returns (it waits for the last copy, not for the kernels).
 * All fail with DGE_ERR_STATE after dge_set_initialized ("Container is already initialized", CellsDataContainer.cpp:61-62). */
int dge_add_batch(dge_handle *h, const dge_record16 *recs, size_t n);
int dge_add_batch_device(dge_handle *h, const dge_record16 *recs, size_t n);

/* Several device-resident record arrays in ONE fill launch (same result as one dge_add_batch_device per array, in any order).  Arrays may
 * live in a peer GPU's HBM (dge_peer_open): the kernel's bulk-async copies then pull them over NVLink, and local and remote tiles
 * alternate inside every block, so the transfer is hidden behind the per-read work -- the exchange of a sharded run
 * without an all-to-all pass.  Empty segments are allowed; at most 64 non-empty ones; 16-byte aligned. */
int dge_add_batch_segments_device(dge_handle *h, const dge_record16 *const *segs, const uint64_t *counts, uint32_t n_segs);
int dge_add_batch_soa(dge_handle *h, const uint64_t *keys, const uint32_t *genes, size_t n, uint64_t first_read_idx);

/* The same batches with the chromosome of every read as a 1-byte side array (ids < 256, assigned by the caller in first-seen order like
 * Stats::get_index, reference Estimation/Stats.cpp:79-88).  Feeds the per-(cell, chromosome) counters of Stats
 * (EXON / INTRON / INTERGENIC_READS_PER_CHR_PER_CELL: CellsDataContainer.cpp:73-78, 309-327, Stats.cpp:23-28); batches added without a
 * chromosome array simply do not count there.  chr is HOST memory for the first two, DEVICE memory for the third. */
int dge_add_batch_chr(dge_handle *h, const dge_record16 *recs, const uint8_t *chr, size_t n);
int dge_add_batch_soa_chr(dge_handle *h, const uint64_t *keys, const uint32_t *genes, const uint8_t *chr, size_t n, uint64_t first_read_idx);
int dge_add_batch_chr_device(dge_handle *h, const dge_record16 *recs, const uint8_t *chr, size_t n);

/* = CellsDataContainer::set_initialized (CellsDataContainer.cpp:163-175): runs the whole per-read grouping on the device.
 * DGE_ERR_STATE when called twice (":165-166"). */
int dge_set_initialized(dge_handle *h);

/* = CellsDataContainer::merge_and_filter (CellsDataContainer.cpp:39-57). DGE_ERR_STATE before dge_set_initialized
 * ("You must initialize container", ":41-42"). */
int dge_merge_and_filter(dge_handle *h);

/* Forget all reads and results but keep the configuration and every device workspace, so that the next run of the
 * same size allocates nothing (the reference equivalent is constructing a fresh CellsDataContainer). */
int dge_reset(dge_handle *h);

/* The strings behind DGE_FLAG_UMI_N (which = 0: umi_len characters each) or DGE_FLAG_CB_N (which = 1: cb_len characters each) indices, concatenated
 * without separators; replaces the previous list.  Needed by the host-side exact paths only (N repair of MergeUMIsStrategySimple, whitelist
 * walk of barcodes containing N): call any time before dge_merge_and_filter. */
int dge_set_n_strings(dge_handle *h, int which, const char *strings, size_t n);

/* The DGE_FLAG_CB_N list with strings of ANY length (concatenated, lengths[k] characters each; replaces the list): barcodes whose length
 * differs from dge_config.cb_len -- variable-length inDrop v1 / v2 barcodes, InDropBarcodesParser.cpp:31-38 -- travel like barcodes with
 * N: a cell of their own, identified by the list index on the device and by the string wherever the reference looks at the string
 * (whitelist walk, compare_cells ties, output).  Same restrictions as barcodes with N (merge none / whitelist merge, one GPU). */
int dge_set_cb_strings(dge_handle *h, const char *strings, const uint32_t *lengths, size_t n);

/* Optional: run on this CUDA stream (a cudaStream_t passed as void*); default is a stream owned by the handle. */
int dge_set_stream(dge_handle *h, void *cuda_stream);

/* ---- query surface (CellsDataContainer.h:99-121, ResultsPrinter.cpp:334-396) ----------------------------------------- */
int dge_get_summary(dge_handle *h, dge_summary *out);
int dge_get_timings(dge_handle *h, dge_timings *out);

/* Cells of one class, in the reference's order: ALL and REAL by cell id (= first-seen order), FILTERED in
 * filtered_cells() order (ascending compare_cells, CellsDataContainer.cpp:329-344).  `out` has room for `capacity` rows;
 * *n_out receives the number available (call with capacity 0 to size). */
int dge_get_cells(dge_handle *h, int which, dge_cell_info *out, size_t capacity, size_t *n_out);

/* Count matrix in compressed-column form, columns = cells (ResultsPrinter.cpp:334-396, 433-442):
 *   DGE_MATRIX_CM     columns = filtered_cells() in order, values = UMIs (or reads with reads_output) whose mark matches the query
 *   DGE_MATRIX_CM_RAW columns = real cells by cell id, values = all UMIs (or reads)
 * indptr has n_cols+1 entries; `gene_ids`/`values` have nnz entries, gene ids ascending within a column and expressed as the
 * caller's gene ids.  Any output pointer may be NULL; *n_cols / *nnz are always written. */
int dge_get_matrix(dge_handle *h, int which, int64_t *indptr, int32_t *gene_ids, int32_t *values,
                   size_t *n_cols, size_t *nnz);

/* The filtered count matrix for ANOTHER set of query marks than dge_config.query_mark_mask, over the same columns as DGE_MATRIX_CM
 * (filtered_cells() in order): ResultsPrinter::get_count_matrix_filtered(container, query_marks), ResultsPrinter.cpp:334-361, which `-V`
 * (save_intron_exon_matrices, :455-474) calls with "e" (mask 0x04), "i" (0x10) and "BA" (0xC0).  Values are UMIs -- reads with
 * reads_output -- whose accumulated mark is one of the mask's; zero entries are dropped.  Built on the device on request; the last mask is
 * kept, so the size query and the fetch cost one build.  Same output convention as dge_get_matrix. */
int dge_get_matrix_marks(dge_handle *h, uint32_t query_mark_mask, int64_t *indptr, int32_t *gene_ids, int32_t *values,
                         size_t *n_cols, size_t *nnz);

/* Per-chromosome read counters of the real cells (CellsDataContainer::get_stat_by_real_cells(CellChrStatType, ...), reference
 * CellsDataContainer.cpp:292-307 + Stats::get, Stats.cpp:50-63), after merge_and_filter, with Stats::merge applied (the counters of a merged
 * cell belong to its merge target, Stats.cpp:36-42).  counts[(cell * n_chr + chr) * 3 + t], t = 0 exon, 1 intron, 2 intergenic; cells in
 * DGE_CELLS_REAL order; n_chr = largest chromosome id seen + 1.  presented[t * n_chr + chr] = 1 when ANY cell counted chromosome chr for
 * statistic t (Stats::presented_chromosomes; the reference lists a cell in a statistic's table only if one of its counters is non-zero).
 * Size query: counts = presented = NULL.  Not available on sharded handles yet. */
int dge_get_chr_stats(dge_handle *h, int32_t *counts, size_t capacity_cells, size_t *n_cells, uint32_t *n_chr, uint8_t *presented);

/* Gene ids in first-seen order (= gene_indexer().values(), StringIndexer.cpp:10-18). */
int dge_get_gene_order(dge_handle *h, int32_t *gene_ids, size_t capacity, size_t *n_out);

/* (barcode_from, barcode_to) for every cell whose merge target is not itself, ordered by source cell id
 * (= ResultsPrinter::get_merge_targets, ResultsPrinter.cpp:316-332). */
int dge_get_merge_pairs(dge_handle *h, uint64_t *from, uint64_t *to, size_t capacity, size_t *n_out);

/* Every distinct (cell, gene, UMI) currently held, for the cells of class `which` (order: cell order of dge_get_cells,
 * then gene id, then UMI value): = walking Cell::genes() / Gene::umis() (Cell.h:22, Gene.h:19).  cell_index indexes the
 * list dge_get_cells(which) returns. Any output pointer may be NULL. */
int dge_get_umigs(dge_handle *h, int which, uint32_t *cell_index, int32_t *gene_ids, uint32_t *umis,
                  uint32_t *read_counts, uint8_t *marks, size_t capacity, size_t *n_out);

/* Gene::merge_targets() (Gene.h:41, filled by Gene::merge(source_umi, target_umi), Gene.cpp:38-58) of every (cell, gene): one row per UMI
 * that MergeUMIsStrategySimple (N repair) or MergeUMIsStrategyDirectional moved into another UMI -- the cell that owns the gene after the
 * barcode merge, the gene, the source UMI and the UMI it went to (codes as in dge_get_umigs; a target created by the N repair is a new,
 * N-free UMI).  Sorted by (barcode code, gene, source).  created[k] (optional, may be NULL) = 1 when the target did not exist in the gene
 * and the source's UMI object BECAME it (Gene.cpp:48: emplace(target, source object) -- the target then carries the source's base-quality sums,
 * which is all that distinguishes the two cases).  Consumers in the reference: the filtered-BAM writer, FilteringBamProcessor::write_alignment
 * (BamProcessing/FilteringBamProcessor.cpp:73-88), and through UMI::mean_quality the `reads_per_umi_per_cell` table.  Needs
 * dge_config.save_umi_merge_targets. */
int dge_get_umi_merge_targets(dge_handle *h, uint64_t *cell_barcodes, int32_t *gene_ids, uint32_t *source_umis, uint32_t *target_umis,
                              uint8_t *created, size_t capacity, size_t *n_out);

/* The merge_cells(source, target) calls of the barcode merge in the order they were applied (MergeStrategyBase::merge_inited,
 * MergeStrategyBase.cpp:29-51 -> merge_force, :84-89): `to` is the cell the source's content went into AT THAT TIME, which with merge chains
 * is not the final target dge_get_merge_pairs reports.  Needed where the order shows: Gene::merge keeps the target's own UMI object and only
 * copies a source's when the target has none (Gene.cpp:26-36), so per-UMI base-quality sums (UMI::mean_quality) belong to the first holder.
 * Single-GPU handles only. */
int dge_get_merge_events(dge_handle *h, uint64_t *from, uint64_t *to, size_t capacity, size_t *n_out);

/* ---- helpers that mirror reference utilities on the path (host-side, exact restatements used by the facade and tests) -- */

/* Tools::CollisionsAdjuster (reference Tools/CollisionsAdjuster.cpp:12-49; used by PoissonTargetEstimator.cpp:99-100):
 * adjusted_sizes[s-1] = CollisionsAdjuster::estimate_adjusted_gene_expression(s) for s = 1..max_gene_expression after
 * init(umi_probabilities).  Host pointers; the recurrence runs on `device` (FP64, parallel over the UMI space).  *exact_rerun
 * (optional) receives the first step at which the parallel sum came too close to a rounding boundary, in which case the whole
 * table was recomputed with the reference's sequential summation order (0 = the fast pass was provably sufficient). */
int dge_collisions_adjusted_sizes(int device, const double *umi_probabilities, size_t n_umis, size_t max_gene_expression,
                                  uint64_t *adjusted_sizes, uint32_t *exact_rerun);

/* Tools::edit_distance (Tools/UtilFunctions.cpp:32-65), literal behaviour including the banded quirks. */
unsigned dge_edit_distance(const char *s1, const char *s2, int skip_n, unsigned max_ed);
/* Tools::hamming_distance (Tools/UtilFunctions.cpp:67-82); returns UINT32_MAX when lengths differ. */
unsigned dge_hamming_distance(const char *s1, const char *s2, int skip_n);

/* Whitelist loaded by the handle (BarcodesParser::init, BarcodesParser.cpp:88-100): number of parts, then per part the
 * token count and token length; tokens are returned reverse-complemented exactly as the reference stores them. */
int dge_whitelist_shape(dge_handle *h, uint32_t *n_parts, uint32_t *part_sizes, uint32_t *part_lengths, size_t capacity);
int dge_whitelist_token(dge_handle *h, uint32_t part, uint32_t index, char *out, size_t capacity);

/* ---- synthetic read streams (bench + tests) -------------------------------------------------------------
 * Counter-based generator: record i depends only on (seed, i) and the tables, so any slice can be produced anywhere
 * (host mirror: dropest_b200/synth.py).  `out_device` is DEVICE memory for `count` records, filled with reads
 * [first, first+count). */
typedef struct dge_synth_params {
    uint64_t seed;
    uint64_t n_reads_total;       /* sizes the per-(cell,gene) UMI pools */
    uint32_t n_cells, n_genes, cb_len, umi_len;
    const uint64_t *cell_cdf;     /* [n_cells] HOST: inclusive cumulative weights scaled to 2^64 (last = UINT64_MAX) */
    const uint64_t *cell_barcode; /* [n_cells] HOST: 2-bit packed true barcodes                                     */
    const uint64_t *cell_reads;   /* [n_cells] HOST: expected reads per cell                                        */
    const uint64_t *gene_cdf;     /* [n_genes] HOST: inclusive cumulative weights scaled to 2^64                    */
    const uint32_t *gene_weight;  /* [n_genes] HOST: weight scaled to 2^32                                          */
    uint32_t cb_error_ppm;        /* reads per million with one substituted barcode base */
    uint32_t intergenic_ppm;
    uint32_t intron_ppm, not_annotated_ppm; /* remaining reads are exonic */
    uint32_t reads_per_umi;       /* pool divisor (>= 1) */
    uint32_t reserved;
} dge_synth_params;

int dge_synth_generate_device(int device, const dge_synth_params *p, uint64_t first, uint64_t count,
                              dge_record16 *out_device, void *cuda_stream);

/* Route records to `n_ranks` owners by barcode hash (the multi-GPU partition step before the all-to-all).
 * in/out are DEVICE memory, out has room for n records; counts[n_ranks] (HOST) receives the per-rank segment sizes;
 * segment r starts at sum(counts[0..r)). */
int dge_route_by_barcode_device(int device, const dge_record16 *in, size_t n, uint32_t n_ranks, dge_record16 *out,
                                uint64_t *counts, void *cuda_stream);

/* Routing for a PIPELINED exchange (route kernel -> all-to-all -> fill, slice by slice).
 * Step 1, dge_route_count_slices_device: `in` is cut into n_slices slices of slice_len records (a multiple of 2048; the last may be
 * shorter); one pass counts the destinations of every slice: counts[s * n_ranks + r] (HOST) and, in `cursors_device` (n_slices * 64
 * uint64, DEVICE, caller-owned), the exclusive prefix of every slice's segments.  The only host synchronisation of the exchange.
 * Step 2, dge_route_scatter_slice_device, once per slice and asynchronous: the slice grouped by destination rank into out_slice,
 * segment r at the prefix of step 1 (slice_cursors_device = cursors_device + 64 * slice, consumed by the launch).  The caller queues the
 * slice's all-to-all behind it and goes on; received slices are filled with dge_add_batch_device while later ones still travel. */
int dge_route_count_slices_device(int device, const dge_record16 *in, size_t n, uint32_t n_ranks, size_t slice_len, uint32_t n_slices,
                                  uint64_t *counts, uint64_t *cursors_device, void *cuda_stream);
int dge_route_scatter_slice_device(int device, const dge_record16 *in_slice, size_t n_slice, uint32_t n_ranks, uint64_t *slice_cursors_device,
                                   dge_record16 *out_slice, void *cuda_stream);

/* Routing in ONE pass (no counting pass), for the peer-memory exchange: destination d owns the window out[d * seg_capacity, (d + 1) *
 * seg_capacity) of this rank's buffer (pulled by d's fill kernel later).  push16 > 0: that many of every 16 tiles store their records for a
 * remote destination d straight into d's HBM at push_base_device[d] (DEVICE array of n_ranks pointers obtained with dge_peer_open; room for
 * push_capacity records each) -- NVLink is idle while the scatter streams through local HBM, so part of the exchange rides along.
 * state_device (2 * n_ranks + 1 uint64, DEVICE, zeroed by the call): sizes of the local windows, records pushed per destination, overflow flag.
 * When a window overflowed the routing is repeated with the exact two-pass scheme above: results never depend on the capacities.  Asynchronous. */
int dge_route_scatter_bounded_device(int device, const dge_record16 *in, size_t n, uint32_t n_ranks, uint32_t my_rank, size_t seg_capacity,
                                     uint64_t *state_device, dge_record16 *out, uint32_t push16, size_t push_capacity,
                                     dge_record16 *const *push_base_device, void *cuda_stream);

/* Peer memory for the exchange (one process per GPU of one NVLink/NVSwitch node).  Instead of scatter -> all-to-all -> fill, the routed
 * records stay in the SOURCE rank's HBM (a dge_peer_alloc buffer, exported as a 64-byte CUDA IPC handle that the caller sends to the
 * peers by any transport) and every owner's fill kernel pulls its segment out of it over NVLink: dge_add_batch_device accepts pointers
 * returned by dge_peer_open.  The caller orders the steps (a barrier before a source overwrites the buffer, one after its scatter). */
int dge_peer_alloc(int device, size_t bytes, void **ptr, unsigned char handle[64]);
int dge_peer_free(int device, void *ptr);
int dge_peer_open(int device, const unsigned char handle[64], void **ptr);
int dge_peer_close(int device, void *ptr);

/* ---- cross-rank whitelist merge for sharded runs ---------------------------------------------------
 * With reads sharded by barcode hash a cell and its merge candidates usually live on different ranks.  The merge is exact
 * (RealBarcodesMergeStrategy.cpp:22-114 incl. the fall-through to farther distance classes and the neighbour order on ties) and every
 * rank only works on ITS OWN cells.  The library is a state machine, the caller owns the transport (NCCL in bench.py /
 * dropest_b200/dist.py, plain copies in the single-GPU multi-handle tests): call dge_dist_step after dge_set_initialized; while it
 * asks for a collective (io->collective != DGE_DIST_DONE) perform it on the bytes it points to and call again with the received bytes:
 *   step 1  -> ALLGATHER  16-byte summaries of the real cells that are whitelist barcodes (the only possible merge targets)
 *   step 2  -> ALLTOALL   per owner of a candidate: the (child, candidate) pairs with the child's (gene|umi, value) list
 *   step 3  -> ALLTOALL   one u32 per received pair: |child ∩ candidate| (MergeStrategyBase.cpp:100-147)
 *   step 4  -> ALLTOALL   one 16-byte commit per child merged into a cell of another rank (its Stats counters, Stats.cpp:29-43)
 *   step 5  -> DONE       commits applied; continue with dge_merge_and_filter (UMI merge, final sizes, filter, matrices)
 * ALLGATHER: every rank contributes send_bytes[0] bytes; recv = the pieces of rank 0..world-1 concatenated, recv_bytes[r] their sizes.
 * ALLTOALL : piece d (send_bytes[d] bytes, pieces concatenated in rank order starting at `send`) goes to rank d; recv likewise by source.
 * All pointers are DEVICE memory on cfg.device; `send` stays valid until the next call, `recv` must stay valid during the call only. */
#define DGE_DIST_MAX_WORLD 64
enum { DGE_DIST_DONE = 0, DGE_DIST_ALLGATHER = 1, DGE_DIST_ALLTOALL = 2 };
typedef struct dge_dist_io {
    uint32_t world, rank;                          /* set by the caller before the first step */
    uint32_t collective;                           /* out: DGE_DIST_* */
    uint32_t stage;                                /* out: number of the step just executed (diagnostics) */
    const void *send;                              /* out */
    uint64_t send_bytes[DGE_DIST_MAX_WORLD];       /* out */
    const void *recv;                              /* in: result of the collective requested by the previous step */
    uint64_t recv_bytes[DGE_DIST_MAX_WORLD];       /* in */
} dge_dist_io;

int dge_dist_step(dge_handle *h, dge_dist_io *io);

/* Sharded runs with a strategy that depends on the UMI indexer's first-seen order (directional UMI merge): every rank tracks
 * min(read_idx) per packed UMI over ITS reads; the reference's StringIndexer is global, so the tables have to be min-reduced
 * across ranks before dge_merge_and_filter.  dge_umi_first_size gives the table length (4^umi_len entries, 0 when the
 * configured strategies do not need it); export/import copy it to / from caller-owned DEVICE memory (u32 per entry,
 * 0xFFFFFFFF = UMI not seen) around the caller's all-reduce(min).  Reference: StringIndexer.cpp:10-18 via Gene.cpp:17-24. */
int dge_umi_first_size(dge_handle *h, size_t *n_entries);
int dge_umi_first_export(dge_handle *h, uint32_t *dst_device);
int dge_umi_first_import(dge_handle *h, const uint32_t *src_device);

#ifdef __cplusplus
}
#endif
#endif /* DROPEST_B200_H */
