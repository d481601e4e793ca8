// BamIngest.h -- the caller of the count-matrix path: BAM alignments -> ReadInfo -> CellsDataContainer::add_record.
// Replaces, for the tag-driven modes of `dropest`, the BamTools-based loop of the reference:
//   Estimation/BamProcessing/BamController.cpp:70-172 (parse_bam_file / process_alignment), FilledBamParamsParser.cpp:12-40 (-f: barcode and
//   UMI from tags, base-quality threshold), ReadParamsParser.cpp:21-90,179-197 (read-name codec, gene tag, read-type tag), BamTags.cpp:7-25.
// Own BGZF / BAM reader (zlib is the only dependency; the blocks are inflated by FastInflate.h, zlib checks the CRC and judges what that
// decoder refuses): the compressed blocks of a chunk are inflated by a pool of threads straight into their place of one contiguous buffer
// (block sizes are known from the headers and trailers without inflating) -- in the background, while the previous chunk is framed and
// parsed --, records are handed to the container in stream order: the order is what defines cell / gene ids downstream.  The gene lookup in
// an annotation (-g, row f3) lives in GeneAnnotation.h; this file calls it for the first and last aligned base of a read.
#pragma once
#include "Estimation.h"
#include "GeneAnnotation.h"

#include <cstdint>
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <future>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

namespace Estimation
{
namespace BamProcessing
{
	struct BamTags // BamTags.cpp:7-25: names and defaults of the tags that are read
	{
		std::string cell_barcode = "CB", umi = "UB", gene = "GX", cell_barcode_quality = "CQ", umi_quality = "UQ";
		std::string read_type, intronic_read_value, intergenic_read_value; // BamTags.Type.*: empty = every read with a gene is exonic
		// written only (BamOutput.h): the raw barcode / UMI tags and the read-type values of the output BAMs (BamTags.cpp:11-21)
		std::string cell_barcode_raw = "CR", umi_raw = "UR", exonic_read_value;
		std::string intronic_read_value_out() const { return intronic_read_value.empty() ? "INTRONIC" : intronic_read_value; }
		std::string intergenic_read_value_out() const { return intergenic_read_value.empty() ? "INTERGENIC" : intergenic_read_value; }
		std::string exonic_read_value_out() const { return exonic_read_value.empty() ? "EXONIC" : exonic_read_value; }
	};

	// One alignment, as a view into the reader's buffer (valid until the next call of BamReader::next)
	struct BamAlignment
	{
		int32_t ref_id = -1, position = -1;
		int32_t end_position = -1; // position + reference bases consumed by the CIGAR (M, D, N, =, X): half-open end, like BamTools' GetEndPosition()
		uint16_t flag = 0;
		std::string name;
		const uint8_t *tag_data = nullptr;
		size_t tag_bytes = 0;

		bool is_mapped() const { return !(flag & 0x4); }
		bool is_primary_alignment() const { return !(flag & 0x100); }
		// type of a tag ('Z', 'A', 'i', ...) or 0 when absent; throws std::runtime_error on a malformed tag block
		char tag_type(const std::string &tag) const;
		// string tags only (Z, H; A gives its one character): false when the tag is absent or of another type
		bool get_string_tag(const std::string &tag, std::string &value) const;
		// ONE walk over the tag block for several tags: found[k] = {type, value pointer, value bytes} of tags[k] (type 0 = absent)
		struct TagValue { char type = 0; const uint8_t *value = nullptr; size_t bytes = 0; bool string(std::string &out) const; };
		void find_tags(const std::string *const *tags, size_t n_tags, TagValue *found) const;
	};

	// the tags of a record's tag block, each as the span [name(2) type(1) value...]; throws std::runtime_error on a malformed block
	struct TagSpan { const uint8_t *begin; size_t bytes; };
	void list_tags(const uint8_t *tag_data, size_t tag_bytes, std::vector<TagSpan> &out);

	class BamReader
	{
	public:
		explicit BamReader(const std::string &file_name, unsigned threads = 0);
		~BamReader();
		BamReader(const BamReader &) = delete;
		BamReader &operator=(const BamReader &) = delete;
		const std::vector<std::string> &reference_names() const { return _refs; }
		const std::vector<uint32_t> &reference_lengths() const { return _ref_lengths; }
		const std::string &header_text() const { return _header_text; }
		bool next(BamAlignment &alignment); // false at the end of the file
		// Every complete record that is already inflated (at least one; up to `max_records`), as views into the reader's buffer that stay
		// valid until the next call of next / next_batch.  Empty at the end of the file.  `name` is left empty (use `name_data`).
		// `raw` = the whole record after its 4-byte block_size field (what a writer copies), raw_bytes = block_size
		struct RecordView { BamAlignment al; const char *name_data = nullptr; size_t name_len = 0; const uint8_t *raw = nullptr; size_t raw_bytes = 0; };
		void next_batch(std::vector<RecordView> &out, size_t max_records);

	private:
		std::string _file_name;
		std::FILE *_f = nullptr;
		unsigned _threads;
		// byte buffers that grow without zero-filling (a vector's resize would touch every new byte once more)
		struct Bytes
		{
			std::unique_ptr<uint8_t[]> p;
			size_t n = 0, cap = 0;
			uint8_t *data() { return p.get(); }
			const uint8_t *data() const { return p.get(); }
			size_t size() const { return n; }
			void grow_to(size_t want) // keeps [0, n)
			{
				if (want <= cap) return;
				size_t c = std::max(want, cap + cap / 2);
				std::unique_ptr<uint8_t[]> q(new uint8_t[c]);
				if (n) std::memcpy(q.get(), p.get(), n);
				p = std::move(q); cap = c;
			}
			void drop_front(size_t k) { if (k) { std::memmove(p.get(), p.get() + k, n - k); n -= k; } }
		};
		// what the background loader hands over: inflated bytes at [begin, bytes.n) of a buffer whose first `begin` bytes are free
		struct Chunk
		{
			Bytes bytes;
			size_t begin = 0;
			bool end_of_file = false; // nothing left (bytes holds no data)
		};
		Bytes _comp;                  // loader: compressed bytes not yet inflated (whole blocks + a partial one at the end)
		bool _eof = false;            // loader: the file has been read to its end
		Bytes _data;                  // inflated bytes; [_pos, _data.n) not yet consumed
		size_t _pos = 0;              // read position in _data
		Bytes _spare;                 // the buffer the next load may reuse
		bool _finished = false;       // the loader reported the end of the file
		std::future<Chunk> _ahead;    // the load in flight (at most one)
		size_t _chunk_bytes = size_t(4) << 20; // compressed bytes read per load (the inflated chunk stays of the order of the last-level cache)
		size_t _headroom = size_t(1) << 20;    // free bytes in front of a chunk for the record cut by the previous chunk's end
		std::vector<std::string> _refs;
		std::vector<uint32_t> _ref_lengths;
		std::string _header_text;

		bool fill(size_t need); // makes at least `need` bytes available at _pos; false at a clean end of file
		Chunk load_chunk(Bytes buffer);
		void start_loading();
		void read_header();
		bool view_at(size_t pos, RecordView &v, size_t &next_pos) const; // the record at _data[pos], false when it is not complete yet
	};

	// -r: barcode, UMI and UMI quality of every read from separate (gzipped) text files written by droptag, one row per read:
	// "name barcode UMI barcode_quality UMI_quality" (ReadMapParamsParser.cpp:50-109, ReadParameters::parse_from_string, ReadParameters.cpp:59-81).
	// Rows that cannot be parsed and repeated names are skipped like the reference does; a read is handed out ONCE (get_read_params erases
	// it, :22-48), a second alignment of the same name cannot be parsed.
	class ReadParamsMap
	{
	public:
		ReadParamsMap(const std::string &read_param_filenames, int min_barcode_quality); // file names separated by blanks / tabs
		struct Entry { uint32_t barcode, umi, umi_quality; bool pass_quality; bool taken; };
		Entry *find(const std::string &read_name) { auto it = _reads.find(read_name); return it == _reads.end() ? nullptr : &it->second; }
		Tools::ReadParameters parameters(const Entry &e) const; // barcode quality is not kept (ReadParametersEfficient.cpp:15-23)
		size_t size() const { return _reads.size(); }

	private:
		std::unordered_map<std::string, Entry> _reads;
		StringIndexer _barcodes, _umis, _umi_qualities;
	};

	// -r bookkeeping of one read that is settled by whoever sees the reads in stream order: the row it found, and whether the annotation did
	// not know its chromosome (which the reference only finds out after the row was taken and the quality checked, BamController.cpp:139-166)
	struct MapLookup { ReadParamsMap::Entry *entry = nullptr; bool chr_not_found = false; };

	struct IngestParams
	{
		bool filled_bam = true;               // -f: barcode / UMI from tags; false: from the read name "prefix!CB#UMI" (ReadParameters::parse_encoded_id)
		// not -f and non-empty: -r, barcode / UMI looked up by read name in these files (BamController::get_parser, BamController.cpp:118-129).
		// Loaded by parse_bam_files / for_each_alignment when `read_params` is not set yet.
		std::string read_param_filenames;
		std::shared_ptr<ReadParamsMap> read_params;
		BamTags tags;
		bool gene_in_chromosome_name = false; // pseudo-aligner output: the reference name is the gene
		int min_barcode_quality = 0;          // -f only: reads with a barcode / UMI base below this Phred quality are dropped (0 = off)
		unsigned threads = 0;                 // BGZF inflate / deflate threads (0 = hardware concurrency)
		std::string output_dir;               // where the tagged / filtered BAMs go; empty = the working directory, like the reference
		// -g: gene and mark from an annotation (GTF / BED, optionally .gz) instead of the gene tag: the positions of the first and the last
		// aligned base are looked up (ReadParamsParser::get_gene_from_reference, ReadParamsParser.cpp:92-150).  Loaded by parse_bam_files /
		// for_each_read when `genes` is not set yet.
		std::string genes_filename;
		std::shared_ptr<const Tools::GeneAnnotation::RefGenesContainer> genes;
	};

	struct IngestStats // the counters BamProcessorAbstract keeps (BamProcessorAbstract.cpp)
	{
		size_t total_reads = 0, cant_parse = 0, low_quality = 0, skipped_unmapped_or_secondary = 0;
	};

	// BamController::parse_bam_files + process_alignment for the tag / read-name modes: every primary mapped alignment of every file, in
	// order, becomes one add_record call.
	// With print_result_bams (-b) every accepted read is also written to "<bam name>.tagged.bam" (BamProcessor, see BamOutput.h).
	void parse_bam_files(const std::vector<std::string> &bam_files, const IngestParams &params, CellsDataContainer &container, IngestStats &stats,
	                     bool print_result_bams = false);

	// The same loop with the ReadInfo handed to a callback instead of a container (tests, other consumers)
	template <class F> void for_each_read(const std::vector<std::string> &bam_files, const IngestParams &params, IngestStats &stats, F &&sink);

	// one alignment -> ReadInfo; false = counted in `stats` and skipped (process_alignment, BamController.cpp:131-172)
	// In -r mode `map` (when given) receives the row the parameters came from; its quality flag, its "taken" state and an unknown chromosome
	// are left to the caller, which sees the reads in stream order.
	bool read_info_from_alignment(const BamAlignment &alignment, const std::string &chr_name, const IngestParams &params, IngestStats &stats,
	                              Tools::ReadParameters &read_params, std::string &gene, UMI::Mark &mark, MapLookup *map = nullptr);

	// One parsed alignment of a batch (filled by several threads, consumed in order)
	struct ParsedRead
	{
		enum Status : uint8_t { OK, SKIPPED, NO_CHROMOSOME, CANT_PARSE, LOW_QUALITY } status = SKIPPED;
		MapLookup map; // -r: the row this read took its parameters from (claimed in stream order by the consumer)
		int32_t ref_id = -1;
		Tools::ReadParameters params;
		std::string gene;
		UMI::Mark mark;
	};

	// parses every record of the batch with `threads` threads (they share nothing); refs = the file's reference sequence names
	void parse_batch(const std::vector<BamReader::RecordView> &records, const std::vector<std::string> &refs, const IngestParams &params,
	                 std::vector<ParsedRead> &out, unsigned threads);

	// The loop of BamController::process_bam_files / parse_bam_file (BamController.cpp:49-116): on_open(file name, reader) once per file, then
	// sink(read info, record view or nullptr) for every read that process_alignment accepts.  With keep_records the sink also gets the raw
	// record (for the BAM writers, BamOutput.h); the next batch is then produced only after the current one was consumed, because the views
	// point into the reader's buffer.  Without it the next batch is read, inflated and parsed while the current one is handed over.
	template <class O, class F> void for_each_alignment(const std::vector<std::string> &bam_files, const IngestParams &params_in, IngestStats &stats,
	                                                    bool keep_records, O &&on_open, F &&sink)
	{
		IngestParams params(params_in);
		if (!params.genes && !params.genes_filename.empty())
			params.genes = std::make_shared<const Tools::GeneAnnotation::RefGenesContainer>(params.genes_filename);
		if (!params.filled_bam && !params.read_params && !params.read_param_filenames.empty())
			params.read_params = std::make_shared<ReadParamsMap>(params.read_param_filenames, params.min_barcode_quality);
		std::vector<BamReader::RecordView> views;
		std::vector<ParsedRead> parsed, parsed_next;
		for (auto const &file : bam_files)
		{
			BamReader reader(file, params.threads);
			on_open(file, reader);
			const auto &refs = reader.reference_names();
			auto produce = [&]() -> bool {
				reader.next_batch(views, size_t(1) << 15); // a few similar batches per inflated chunk: parsing one overlaps with consuming the previous one
				if (views.empty()) return false;
				parse_batch(views, refs, params, parsed_next, params.threads);
				return true;
			};
			bool have = produce();
			while (have)
			{
				parsed.swap(parsed_next);
				std::future<bool> next;
				if (!keep_records) next = std::async(std::launch::async, produce);
				try
				{
					for (size_t k = 0; k < parsed.size(); ++k) // stream order: it defines cell / gene / chromosome ids downstream
					{
						ParsedRead &r = parsed[k];
						switch (r.status)
						{
						case ParsedRead::SKIPPED: ++stats.skipped_unmapped_or_secondary; break;
						case ParsedRead::NO_CHROMOSOME: ++stats.cant_parse; break; // unknown chromosome: not counted as a read (BamController.cpp:93-105)
						case ParsedRead::CANT_PARSE: ++stats.total_reads; ++stats.cant_parse; break;
						case ParsedRead::LOW_QUALITY: ++stats.total_reads; ++stats.low_quality; break;
						case ParsedRead::OK:
							++stats.total_reads;
							if (r.map.entry)
							{   // -r: a row serves one alignment; the next one of that name is "can't find read name" (ReadMapParamsParser.cpp:27-42)
								if (r.map.entry->taken) { ++stats.cant_parse; break; }
								r.map.entry->taken = true;
								if (!r.map.entry->pass_quality) { ++stats.low_quality; break; }
								if (r.map.chr_not_found) { ++stats.cant_parse; break; }
							}
							sink(ReadInfo(std::move(r.params), std::move(r.gene), refs[size_t(r.ref_id)], r.mark), keep_records ? &views[k] : nullptr);
							break;
						}
					}
				}
				catch (...) { if (next.valid()) next.wait(); throw; }
				have = keep_records ? produce() : next.get();
			}
		}
	}

	// The bulk form of a parsed batch for CellsDataContainer::add_records (-f mode without an annotation): no std::string per read, barcode
	// and UMI 2-bit packed by the parsing threads; the strings a read may still need are copied into per-thread arenas that live as long
	// as the batch.
	struct PackedBatch
	{
		std::vector<uint8_t> status;            // ParsedRead::Status per record
		std::vector<PackedRead> reads;          // per record; meaningful where status == OK
		std::vector<std::vector<char>> arenas;  // what the views of `reads` point into
	};
	bool packed_path_applies(const IngestParams &params); // -f, gene from the gene tag or the chromosome name
	void parse_batch_packed(const std::vector<BamReader::RecordView> &records, const std::vector<std::string> &refs, const IngestParams &params,
	                        PackedBatch &out, unsigned threads);

	template <class F> void for_each_read(const std::vector<std::string> &bam_files, const IngestParams &params, IngestStats &stats, F &&sink)
	{
		for_each_alignment(bam_files, params, stats, false, [](const std::string &, const BamReader &) {},
		                   [&](const ReadInfo &ri, const BamReader::RecordView *) { sink(ri); });
	}
}
}
