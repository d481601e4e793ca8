// BamOutput.h -- the BAM writers of dropEst: `-b` "<name>.tagged.bam" (every accepted read with its barcode / UMI /
// gene tags) and `-F` "<name>.filtered.bam" (the reads of the cells that survive merge_and_filter, tagged with the CORRECTED barcode and UMI).
// Replaces, without BamTools:
//   Estimation/BamProcessing/BamProcessorAbstract.cpp:31-114 (update_bam, save_alignment: which tags are edited, in which order),
//   Estimation/BamProcessing/BamProcessor.cpp:58-72 (`-b`), Estimation/BamProcessing/FilteringBamProcessor.cpp:14-111 (`-F`: the barcode map built from
//   merge_targets() / filtered_cells(), the per-read lookup through cell_id_by_cb, Cell::genes, Gene::merge_targets, Gene::has, the counters),
//   BamController::write_filtered_bam_files (BamController.cpp:38-47).
// The BGZF / BAM writer is our own (SAM/BAM specification 4.1-4.2; zlib is the only dependency): records are copied from the reader's buffer
// with their tag block edited like BamTools' EditTag does (an existing tag is removed, the new value is appended), blocks are deflated by a
// pool of threads.
#pragma once
#include "BamIngest.h"

#include <unordered_map>

namespace Estimation
{
namespace BamProcessing
{
	// Block-compressed output stream: 0xff00 bytes of payload per BGZF block, the 28-byte end-of-file marker on close
	class BgzfWriter
	{
	public:
		BgzfWriter(const std::string &file_name, unsigned threads = 0, int level = -1);
		~BgzfWriter();
		BgzfWriter(const BgzfWriter &) = delete;
		BgzfWriter &operator=(const BgzfWriter &) = delete;
		void write(const void *data, size_t n);
		void close(); // flushes, writes the end-of-file block; throws on an I/O error (the destructor swallows it)

	private:
		std::string _file_name;
		std::FILE *_f = nullptr;
		unsigned _threads;
		int _level;
		std::vector<uint8_t> _pending; // payload not yet compressed
		void flush(bool all);
	};

	class BamWriter
	{
	public:
		// BamWriter::Open (header text and reference dictionary copied from the input BAM, BamProcessorAbstract.cpp:43)
		BamWriter(const std::string &file_name, const std::string &header_text, const std::vector<std::string> &ref_names,
		          const std::vector<uint32_t> &ref_lengths, unsigned threads = 0);
		// One string-tag edit: BamAlignment::EditTag(tag, "Z", value).  Tags whose name is not two characters are ignored, like BamTools does.
		struct TagEdit { std::string tag, value; };
		// the record `raw` (RecordView::raw / raw_bytes) with the edits applied in order
		void save_alignment(const uint8_t *raw, size_t raw_bytes, const std::vector<TagEdit> &edits);
		void close() { _out.close(); }
		size_t written() const { return _written; }

	private:
		BgzfWriter _out;
		std::vector<uint8_t> _rec;
		size_t _written = 0;
	};

	// The tag edits of BamProcessorAbstract::save_alignment (BamProcessorAbstract.cpp:65-114) for one read
	void tag_edits(const BamTags &tags, const ReadInfo &read_info_raw, const std::string &cell_barcode_corrected, const std::string &umi_corrected,
	               std::vector<BamWriter::TagEdit> &edits);

	// BamProcessorAbstract::update_bam's file name: get_result_bam_name without its directory (the reference writes into the working directory)
	std::string result_bam_name(const std::string &bam_name, const std::string &suffix, const std::string &output_dir);

	// FilteringBamProcessor: decides, per read, whether it is written and with which corrected barcode / UMI
	class FilteringBamProcessor
	{
	public:
		explicit FilteringBamProcessor(const CellsDataContainer &container); // FilteringBamProcessor.cpp:14-40
		// write_alignment (:62-96) up to the save: false = not written (no gene, cell filtered out, gene / UMI not found)
		bool corrected_tags(const ReadInfo &read_info, std::string &cell_barcode, std::string &umi);
		size_t merge_cbs_size() const { return _merge_cbs.size(); }
		size_t written_reads() const { return _written_reads; }
		size_t wrong_genes() const { return _wrong_genes; }
		size_t wrong_umis() const { return _wrong_umis; }

	private:
		const CellsDataContainer &_container;
		std::unordered_map<std::string, std::string> _merge_cbs;
		size_t _written_reads = 0, _wrong_genes = 0, _wrong_umis = 0;
	};

	struct FilteredBamStats
	{
		IngestStats reads;
		size_t written_reads = 0, wrong_genes = 0, wrong_umis = 0;
		std::string file_name; // the BAM that was written (empty when there was no input file)
	};

	// BamController::write_filtered_bam_files (BamController.cpp:38-47): a second pass over the input BAMs after merge_and_filter.  Like the
	// reference, ONE output file is opened -- named after the first input, with its header (FilteringBamProcessor::update_bam, :103-110).
	// The container must have been built with save_umi_merge_targets = true for the UMI corrections of the UMI merge strategy to show.
	void write_filtered_bam_files(const std::vector<std::string> &bam_files, const IngestParams &params, const CellsDataContainer &container,
	                              FilteredBamStats &stats);
}
}
