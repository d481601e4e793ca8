/* TEST INFRASTRUCTURE ONLY -- never shipped, never linked into the product.  No-op stand-ins for the C-ABI entries the host mirror calls
 * while reads are added, so that the HOST side of the BAM ingest (loader / producer / hand-over threads, counters, error propagation;
 * tests/cpp/test_ingest_pipeline_host.cpp) can run in the CPU test suite, also under ThreadSanitizer.  Nothing is computed: a program linked
 * against this cannot produce a result, and the real library keeps failing loudly without a GPU. */
#include <string.h>
static int dummy_handle;
static unsigned long long n_reads_seen;
int dge_create(const void *cfg, void **out) { (void)cfg; *out = &dummy_handle; return 0; }
extern const unsigned long stub_dge_config_size; /* stub_device_sizes.c: sizeof(dge_config) from the real header */
void dge_config_default(void *cfg) { memset(cfg, 0, stub_dge_config_size); }
const char *dge_last_error(const void *h) { (void)h; return "stub device: nothing is computed"; }
/* what reaches the device, read by read in arrival order: (key, gene word, read index, chromosome id) folded into one digest, so that two
 * host paths can be compared on everything the device would see (packing, gene ids, marks, stream positions, chromosome ids) */
static unsigned long long digest = 0xCBF29CE484222325ull;
static void fold(const void *p, unsigned long n) { const unsigned char *b = (const unsigned char *)p; for (unsigned long i = 0; i < n; ++i) { digest ^= b[i]; digest *= 0x100000001B3ull; } }
static void fold_read(unsigned long long key, unsigned gene, unsigned idx, unsigned char chr) { fold(&key, 8); fold(&gene, 4); fold(&idx, 4); fold(&chr, 1); }
struct rec16 { unsigned long long key; unsigned gene, read_idx; };
int dge_add_batch_chr(void *h, const void *recs, const void *chr, unsigned long n)
{
	(void)h;
	const struct rec16 *r = (const struct rec16 *)recs;
	const unsigned char *c = (const unsigned char *)chr;
	for (unsigned long i = 0; i < n; ++i) fold_read(r[i].key, r[i].gene, r[i].read_idx, c ? c[i] : 0);
	n_reads_seen += n;
	return 0;
}
int dge_add_batch(void *h, const void *recs, unsigned long n) { return dge_add_batch_chr(h, recs, 0, n); }
int dge_add_batch_soa_chr(void *h, const void *k, const void *g, const void *c, unsigned long n, unsigned long long first)
{
	(void)h;
	const unsigned long long *keys = (const unsigned long long *)k;
	const unsigned *genes = (const unsigned *)g;
	const unsigned char *chr = (const unsigned char *)c;
	for (unsigned long i = 0; i < n; ++i) fold_read(keys[i], genes[i], (unsigned)(first + i), chr ? chr[i] : 0);
	n_reads_seen += n;
	return 0;
}
int dge_add_batch_soa(void *h, const void *k, const void *g, unsigned long n, unsigned long long first) { return dge_add_batch_soa_chr(h, k, g, 0, n, first); }
unsigned long long stub_digest(void) { return digest; }
unsigned long long stub_reads_seen(void) { return n_reads_seen; }
#define STUB(name) int name(void) { return 1; } /* DGE_ERR_*: anything past the fill is refused */
STUB(dge_collisions_adjusted_sizes) STUB(dge_edit_distance) STUB(dge_get_cells) STUB(dge_get_chr_stats) STUB(dge_get_matrix) STUB(dge_get_matrix_marks)
STUB(dge_get_merge_events) STUB(dge_get_summary) STUB(dge_get_umi_merge_targets) STUB(dge_get_umigs) STUB(dge_hamming_distance) STUB(dge_merge_and_filter)
int dge_set_cb_strings(void) { return 0; }
int dge_set_n_strings(void) { return 0; }
int dge_set_initialized(void) { return 0; } /* set_initialized flushes the last batch: the test counts the reads that arrived */
int dge_destroy(void *h) { (void)h; return 0; }
