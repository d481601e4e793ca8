// BamOutput.cpp -- see BamOutput.h.  BGZF: SAM/BAM specification 4.1 (gzip members with a "BC" extra subfield, <= 64 KiB each, an empty
// member as the end-of-file marker); BAM header and records: 4.2.  Nothing here is taken from BamTools.
#include "BamOutput.h"

#include <cstring>
#include <stdexcept>
#include <thread>

#include <zlib.h>

namespace Estimation
{
namespace BamProcessing
{
	namespace
	{
		constexpr size_t BGZF_PAYLOAD = 0xff00; // payload bytes per block: even incompressible data stays below the 64 KiB block limit
		constexpr size_t BGZF_BATCH = 64;       // blocks compressed per flush

		inline void put16(uint8_t *p, uint32_t v) { p[0] = uint8_t(v); p[1] = uint8_t(v >> 8); }
		inline void put32(uint8_t *p, uint32_t v) { p[0] = uint8_t(v); p[1] = uint8_t(v >> 8); p[2] = uint8_t(v >> 16); p[3] = uint8_t(v >> 24); }
		inline uint16_t get16(const uint8_t *p) { return uint16_t(p[0] | (p[1] << 8)); }
		inline uint32_t get32(const uint8_t *p) { return uint32_t(p[0]) | (uint32_t(p[1]) << 8) | (uint32_t(p[2]) << 16) | (uint32_t(p[3]) << 24); }

		// one BGZF block holding data[0, n), n <= BGZF_PAYLOAD
		void deflate_block(const uint8_t *data, size_t n, int level, std::vector<uint8_t> &out)
		{
			out.resize(18 + 65536 + 8);
			static const uint8_t header[16] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 'B', 'C', 2, 0};
			std::memcpy(out.data(), header, 16);
			z_stream zs;
			std::memset(&zs, 0, sizeof(zs));
			if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) throw std::runtime_error("zlib: deflateInit2 failed");
			zs.next_in = const_cast<Bytef *>(data);
			zs.avail_in = uInt(n);
			zs.next_out = out.data() + 18;
			zs.avail_out = uInt(65536 - 18 - 8);
			const int rc = deflate(&zs, Z_FINISH);
			const size_t clen = zs.total_out;
			deflateEnd(&zs);
			if (rc != Z_STREAM_END) throw std::runtime_error("zlib: a BGZF block did not fit 64 KiB");
			const size_t total = 18 + clen + 8;
			put16(out.data() + 16, uint32_t(total - 1));
			put32(out.data() + 18 + clen, uint32_t(crc32(crc32(0L, Z_NULL, 0), data, uInt(n))));
			put32(out.data() + 18 + clen + 4, uint32_t(n));
			out.resize(total);
		}
	}

	BgzfWriter::BgzfWriter(const std::string &file_name, unsigned threads, int level)
		: _file_name(file_name), _f(std::fopen(file_name.c_str(), "wb")), _threads(threads ? threads : std::max(1u, std::thread::hardware_concurrency()))
		, _level(level)
	{
		if (!_f) throw std::runtime_error("Could not open BAM file to write: " + file_name); // BamProcessorAbstract.cpp:43-44
	}

	BgzfWriter::~BgzfWriter()
	{
		try { close(); }
		catch (...) {}
	}

	void BgzfWriter::write(const void *data, size_t n)
	{
		if (!_f) throw std::runtime_error("write to a closed BAM file: " + _file_name);
		const uint8_t *p = static_cast<const uint8_t *>(data);
		_pending.insert(_pending.end(), p, p + n);
		if (_pending.size() >= BGZF_PAYLOAD * BGZF_BATCH) flush(false);
	}

	void BgzfWriter::flush(bool all)
	{
		const size_t n_blocks = all ? (_pending.size() + BGZF_PAYLOAD - 1) / BGZF_PAYLOAD : _pending.size() / BGZF_PAYLOAD;
		if (!n_blocks) return;
		std::vector<std::vector<uint8_t>> out(n_blocks);
		const unsigned nt = unsigned(std::min<size_t>(_threads, (n_blocks + 3) / 4));
		std::vector<std::string> errors(std::max(1u, nt));
		auto work = [&](unsigned t) {
			try
			{
				for (size_t b = t; b < n_blocks; b += std::max(1u, nt))
					deflate_block(_pending.data() + b * BGZF_PAYLOAD, std::min(BGZF_PAYLOAD, _pending.size() - b * BGZF_PAYLOAD), _level, out[b]);
			}
			catch (std::exception &e) { errors[t] = e.what(); }
		};
		if (nt <= 1) work(0);
		else
		{
			std::vector<std::thread> pool;
			for (unsigned t = 0; t < nt; ++t) pool.emplace_back(work, t);
			for (auto &th : pool) th.join();
		}
		for (auto const &e : errors) if (!e.empty()) throw std::runtime_error(e);
		for (auto const &o : out)
			if (std::fwrite(o.data(), 1, o.size(), _f) != o.size()) throw std::runtime_error("write error on " + _file_name);
		const size_t done = std::min(_pending.size(), n_blocks * BGZF_PAYLOAD);
		_pending.erase(_pending.begin(), _pending.begin() + long(done));
	}

	void BgzfWriter::close()
	{
		if (!_f) return;
		std::FILE *f = _f;
		try
		{
			flush(true);
			static const uint8_t eof_marker[28] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 'B', 'C', 2, 0, 27, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
			if (std::fwrite(eof_marker, 1, 28, f) != 28) throw std::runtime_error("write error on " + _file_name);
		}
		catch (...) { _f = nullptr; std::fclose(f); throw; }
		_f = nullptr;
		if (std::fclose(f) != 0) throw std::runtime_error("write error on " + _file_name);
	}

	BamWriter::BamWriter(const std::string &file_name, const std::string &header_text, const std::vector<std::string> &ref_names,
	                     const std::vector<uint32_t> &ref_lengths, unsigned threads)
		: _out(file_name, threads)
	{
		if (ref_names.size() != ref_lengths.size()) throw std::runtime_error("BAM header: reference names and lengths differ in number");
		std::vector<uint8_t> h(12 + header_text.size());
		std::memcpy(h.data(), "BAM\1", 4);
		put32(h.data() + 4, uint32_t(header_text.size()));
		std::memcpy(h.data() + 8, header_text.data(), header_text.size());
		put32(h.data() + 8 + header_text.size(), uint32_t(ref_names.size()));
		_out.write(h.data(), h.size());
		for (size_t r = 0; r < ref_names.size(); ++r)
		{
			std::vector<uint8_t> e(4 + ref_names[r].size() + 1 + 4);
			put32(e.data(), uint32_t(ref_names[r].size() + 1));
			std::memcpy(e.data() + 4, ref_names[r].c_str(), ref_names[r].size() + 1);
			put32(e.data() + 4 + ref_names[r].size() + 1, ref_lengths[r]);
			_out.write(e.data(), e.size());
		}
	}

	void BamWriter::save_alignment(const uint8_t *raw, size_t raw_bytes, const std::vector<TagEdit> &edits)
	{
		if (raw_bytes < 32) throw std::runtime_error("malformed alignment record");
		const size_t l_read_name = raw[8], n_cigar = get16(raw + 12), l_seq = get32(raw + 16);
		const size_t tag_off = 32 + l_read_name + n_cigar * 4 + (l_seq + 1) / 2 + l_seq;
		if (tag_off > raw_bytes) throw std::runtime_error("malformed alignment record");
		// the tag block after the edits: EditTag = RemoveTag + AddTag (the new value goes to the end of the block)
		struct Piece { char a, b; const uint8_t *p; size_t n; const TagEdit *edit; };
		std::vector<TagSpan> spans;
		list_tags(raw + tag_off, raw_bytes - tag_off, spans);
		std::vector<Piece> pieces;
		pieces.reserve(spans.size() + edits.size());
		for (auto const &s : spans) pieces.push_back(Piece{char(s.begin[0]), char(s.begin[1]), s.begin, s.bytes, nullptr});
		for (auto const &e : edits)
		{
			if (e.tag.size() != 2) continue;
			for (size_t k = 0; k < pieces.size();)
				if (pieces[k].a == e.tag[0] && pieces[k].b == e.tag[1]) pieces.erase(pieces.begin() + long(k)); else ++k;
			pieces.push_back(Piece{e.tag[0], e.tag[1], nullptr, 3 + e.value.size() + 1, &e});
		}
		size_t total = tag_off;
		for (auto const &pc : pieces) total += pc.n;
		_rec.resize(4 + total);
		put32(_rec.data(), uint32_t(total));
		std::memcpy(_rec.data() + 4, raw, tag_off);
		uint8_t *w = _rec.data() + 4 + tag_off;
		for (auto const &pc : pieces)
		{
			if (pc.edit)
			{
				w[0] = uint8_t(pc.a); w[1] = uint8_t(pc.b); w[2] = 'Z';
				std::memcpy(w + 3, pc.edit->value.c_str(), pc.edit->value.size() + 1);
			}
			else std::memcpy(w, pc.p, pc.n);
			w += pc.n;
		}
		_out.write(_rec.data(), _rec.size());
		++_written;
	}

	void tag_edits(const BamTags &tags, const ReadInfo &read_info_raw, const std::string &cell_barcode_corrected, const std::string &umi_corrected,
	               std::vector<BamWriter::TagEdit> &edits)
	{
		edits.clear();
		auto const &raw_params = read_info_raw.params;
		if (!read_info_raw.gene.empty()) edits.push_back({tags.gene, read_info_raw.gene});
		edits.push_back({tags.cell_barcode_raw, raw_params.cell_barcode()});
		edits.push_back({tags.umi_raw, raw_params.umi()});
		if (!raw_params.cell_barcode_quality().empty()) edits.push_back({tags.cell_barcode_quality, raw_params.cell_barcode_quality()});
		if (!raw_params.umi_quality().empty()) edits.push_back({tags.umi_quality, raw_params.umi_quality()});
		// read type: only the three pure marks get a value (BamProcessorAbstract.cpp:88-100)
		if (read_info_raw.umi_mark == UMI::Mark::HAS_EXONS) edits.push_back({tags.read_type, tags.exonic_read_value_out()});
		else if (read_info_raw.umi_mark == UMI::Mark::HAS_INTRONS) edits.push_back({tags.read_type, tags.intronic_read_value_out()});
		else if (read_info_raw.umi_mark == UMI::Mark::HAS_NOT_ANNOTATED) edits.push_back({tags.read_type, tags.intergenic_read_value_out()});
		if (!cell_barcode_corrected.empty()) edits.push_back({tags.cell_barcode, cell_barcode_corrected});
		if (!umi_corrected.empty()) edits.push_back({tags.umi, umi_corrected});
	}

	std::string result_bam_name(const std::string &bam_name, const std::string &suffix, const std::string &output_dir)
	{
		std::string name = bam_name.substr(0, bam_name.find_last_of('.')) + suffix; // get_result_bam_name
		const std::string::size_type path_end = name.find_last_of("\\/");            // update_bam, BamProcessorAbstract.cpp:35-39
		if (path_end != std::string::npos) name = name.substr(path_end + 1);
		if (output_dir.empty()) return name;
		return output_dir.back() == '/' ? output_dir + name : output_dir + '/' + name;
	}

	FilteringBamProcessor::FilteringBamProcessor(const CellsDataContainer &container)
		: _container(container)
	{
		auto const &merge_targets = container.merge_targets();
		std::vector<bool> good_cells_mask(merge_targets.size(), false);
		for (size_t id : container.filtered_cells()) good_cells_mask[id] = true;
		for (size_t base_cell_id = 0; base_cell_id < merge_targets.size(); ++base_cell_id)
		{
			const size_t target_cell = merge_targets[base_cell_id];
			if (!good_cells_mask[target_cell]) continue;
			_merge_cbs[container.cell(base_cell_id).barcode()] = container.cell(target_cell).barcode();
		}
	}

	bool FilteringBamProcessor::corrected_tags(const ReadInfo &read_info, std::string &cell_barcode, std::string &umi)
	{
		if (read_info.gene.empty()) return false;
		auto cb_iter = _merge_cbs.find(read_info.params.cell_barcode());
		if (cb_iter == _merge_cbs.end()) return false; // the barcode did not pass the size threshold
		auto const &genes = _container.cell(_container.cell_id_by_cb(cb_iter->second)).genes();
		Cell::genes_t::const_iterator gene_iter = genes.end();
		try { gene_iter = genes.find(_container.gene_indexer().get_index(read_info.gene)); }
		catch (std::out_of_range &) {} // a gene the first pass never saw (the reference would stop with out_of_range here)
		if (gene_iter == genes.end()) { ++_wrong_genes; return false; }
		auto const &targets = gene_iter->second.merge_targets();
		auto umi_target_it = targets.find(read_info.params.umi());
		if (umi_target_it != targets.end()) umi = umi_target_it->second;
		else if (gene_iter->second.has(read_info.params.umi())) umi = read_info.params.umi();
		else { ++_wrong_umis; return false; }
		cell_barcode = cb_iter->second;
		++_written_reads;
		return true;
	}

	void write_filtered_bam_files(const std::vector<std::string> &bam_files, const IngestParams &params, const CellsDataContainer &container,
	                              FilteredBamStats &stats)
	{
		FilteringBamProcessor processor(container);
		std::unique_ptr<BamWriter> writer;
		std::vector<BamWriter::TagEdit> edits;
		std::string cb, umi;
		for_each_alignment(bam_files, params, stats.reads, true,
			[&](const std::string &file, const BamReader &reader) {
				if (writer) return; // FilteringBamProcessor::update_bam: only the first input opens the output
				stats.file_name = result_bam_name(file, ".filtered.bam", params.output_dir);
				writer.reset(new BamWriter(stats.file_name, reader.header_text(), reader.reference_names(), reader.reference_lengths(), params.threads));
			},
			[&](const ReadInfo &ri, const BamReader::RecordView *view) {
				if (!processor.corrected_tags(ri, cb, umi)) return;
				tag_edits(params.tags, ri, cb, umi, edits);
				writer->save_alignment(view->raw, view->raw_bytes, edits);
			});
		if (writer) writer->close();
		stats.written_reads = processor.written_reads();
		stats.wrong_genes = processor.wrong_genes();
		stats.wrong_umis = processor.wrong_umis();
	}
}
}
