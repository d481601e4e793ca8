// BamIngest.cpp -- see BamIngest.h.  BGZF: RFC 1952 members with a "BC" extra subfield holding the block size (SAM/BAM specification 4.1);
// BAM records: specification 4.2.  Nothing here is taken from BamTools.
#include "BamIngest.h"
#include "BamOutput.h"
#include "FastInflate.h"

#include <algorithm>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <exception>
#include <mutex>
#include <stdexcept>
#include <thread>

#include <zlib.h>

namespace Estimation
{
namespace BamProcessing
{
	namespace
	{
		inline uint16_t le16(const uint8_t *p) { return uint16_t(p[0] | (p[1] << 8)); }
		inline uint32_t le32(const uint8_t *p) { return uint32_t(p[0]) | (uint32_t(p[1]) << 8) | (uint32_t(p[2]) << 16) | (uint32_t(p[3]) << 24); }

		struct Block { size_t in_off, in_len, out_off, out_len; };

		// Length of the BGZF block starting at p (n bytes available), 0 when the header is not complete yet
		size_t bgzf_block_size(const uint8_t *p, size_t n, const std::string &fname)
		{
			if (n < 18) return 0;
			if (p[0] != 31 || p[1] != 139 || p[2] != 8 || !(p[3] & 4)) throw std::runtime_error("not a BGZF block in " + fname);
			const size_t xlen = le16(p + 10);
			if (n < 12 + xlen) return 0;
			size_t off = 12;
			while (off + 4 <= 12 + xlen)
			{
				const size_t slen = le16(p + off + 2);
				if (p[off] == 'B' && p[off + 1] == 'C' && slen == 2) return size_t(le16(p + off + 4)) + 1;
				off += 4 + slen;
			}
			throw std::runtime_error("BGZF block without a BC subfield in " + fname);
		}

		void inflate_block(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len, const std::string &fname)
		{
			const size_t xlen = le16(in + 10), hdr = 12 + xlen;
			if (in_len < hdr + 8) throw std::runtime_error("truncated BGZF block in " + fname);
			if (out_len == 0) return;
			const uint32_t want_crc = le32(in + in_len - 8);
			auto crc_ok = [&] { return uint32_t(crc32(crc32(0L, Z_NULL, 0), out, uInt(out_len))) == want_crc; };
			// our own decoder first (FastInflate.h); whatever it does not accept -- or gets wrong by the CRC -- is zlib's to judge
			static const bool use_zlib_only = std::getenv("DGE_BAM_ZLIB_INFLATE") != nullptr;
			if (!use_zlib_only && FastInflate::inflate_raw(in + hdr, in_len - hdr - 8, out, out_len) && crc_ok()) return;
			z_stream zs;
			std::memset(&zs, 0, sizeof(zs));
			if (inflateInit2(&zs, -15) != Z_OK) throw std::runtime_error("zlib: inflateInit2 failed");
			zs.next_in = const_cast<Bytef *>(in + hdr);
			zs.avail_in = uInt(in_len - hdr - 8);
			zs.next_out = out;
			zs.avail_out = uInt(out_len);
			const int rc = inflate(&zs, Z_FINISH);
			const bool ok = rc == Z_STREAM_END && zs.avail_out == 0;
			inflateEnd(&zs);
			if (!ok) throw std::runtime_error("corrupt BGZF block in " + fname);
			if (!crc_ok()) throw std::runtime_error("BGZF CRC mismatch in " + fname);
		}
	}

	BamReader::BamReader(const std::string &file_name, unsigned threads)
		: _file_name(file_name), _f(std::fopen(file_name.c_str(), "rb")), _threads(threads ? threads : std::max(1u, std::thread::hardware_concurrency()))
	{
		if (!_f) throw std::runtime_error("Can't open BAM file: " + file_name); // BamController.cpp:77-78
		// test hooks: small chunks / no headroom put records across every chunk boundary
		if (const char *e = std::getenv("DGE_BAM_CHUNK_BYTES")) _chunk_bytes = std::max<size_t>(1, std::strtoull(e, nullptr, 10));
		if (const char *e = std::getenv("DGE_BAM_HEADROOM_BYTES")) _headroom = std::strtoull(e, nullptr, 10);
		read_header();
	}

	BamReader::~BamReader()
	{
		if (_ahead.valid()) { try { _ahead.get(); } catch (...) {} } // the loader reads from _f
		if (_f) std::fclose(_f);
	}

	// One step of the background loader: reads compressed bytes until at least one whole BGZF block is there, inflates every whole block
	// (thread pool, straight into place) behind `headroom` free bytes of the buffer it was given.  Only the loader touches _f, _comp, _eof.
	BamReader::Chunk BamReader::load_chunk(Bytes buffer)
	{
		Chunk c;
		c.bytes = std::move(buffer);
		c.bytes.n = 0;
		while (true)
		{
			// more compressed bytes (a few MB at a time: the inflated chunk stays of the order of the last-level cache)
			const size_t chunk = _chunk_bytes;
			if (!_eof)
			{
				_comp.grow_to(_comp.n + chunk + 16);
				const size_t got = std::fread(_comp.data() + _comp.n, 1, chunk, _f);
				_comp.n += got;
				std::memset(_comp.data() + _comp.n, 0, 16); // FastInflate may look (not depend) on up to 15 bytes behind a block
				if (got < chunk) _eof = true;
			}
			// the complete blocks in _comp
			std::vector<Block> blocks;
			size_t off = 0, out_total = 0;
			while (off < _comp.size())
			{
				const size_t bs = bgzf_block_size(_comp.data() + off, _comp.size() - off, _file_name);
				if (bs == 0 || off + bs > _comp.size()) break;
				const size_t isize = le32(_comp.data() + off + bs - 4);
				if (isize > 65536) throw std::runtime_error("BGZF block larger than 64 KiB in " + _file_name);
				blocks.push_back(Block{off, bs, out_total, isize});
				out_total += isize;
				off += bs;
			}
			if (blocks.empty())
			{
				if (_eof)
				{
					if (off < _comp.size()) throw std::runtime_error("truncated BGZF block at the end of " + _file_name);
					c.end_of_file = true;
					return c;
				}
				continue;
			}
			const size_t base = _headroom;
			c.bytes.grow_to(base + out_total);
			c.bytes.n = base + out_total;
			c.begin = base;
			const unsigned nt = unsigned(std::min<size_t>(_threads, (blocks.size() + 15) / 16));
			std::vector<std::string> errors(std::max(1u, nt));
			auto work = [&](unsigned t, size_t first, size_t last) { // contiguous runs of blocks per thread
				try
				{
					for (size_t b = first; b < last; ++b)
						inflate_block(_comp.data() + blocks[b].in_off, blocks[b].in_len, c.bytes.data() + base + blocks[b].out_off, blocks[b].out_len, _file_name);
				}
				catch (std::exception &e) { errors[t] = e.what(); }
			};
			if (nt <= 1) work(0, 0, blocks.size());
			else
			{
				std::vector<std::thread> pool;
				const size_t per = (blocks.size() + nt - 1) / nt;
				for (unsigned t = 0; t < nt; ++t) pool.emplace_back(work, t, std::min(blocks.size(), size_t(t) * per), std::min(blocks.size(), size_t(t + 1) * per));
				for (auto &th : pool) th.join();
			}
			for (auto const &e : errors) if (!e.empty()) throw std::runtime_error(e);
			_comp.drop_front(off);
			return c;
		}
	}

	void BamReader::start_loading()
	{
		_ahead = std::async(std::launch::async, [this](Bytes buffer) { return load_chunk(std::move(buffer)); }, std::move(_spare));
		_spare = Bytes();
	}

	// The chunk after the current one is read and inflated in the background while the caller frames and parses the current one; the
	// bytes of a record cut by the chunk boundary are copied in front of the new chunk (into its headroom).
	bool BamReader::fill(size_t need)
	{
		while (_data.size() - _pos < need)
		{
			if (_finished) return false;
			if (!_ahead.valid()) start_loading();
			Chunk c = _ahead.get(); // rethrows what the loader threw
			if (c.end_of_file)
			{
				_finished = true;
				return false;
			}
			const size_t tail = _data.size() - _pos;
			if (tail > c.begin)
			{   // more left over than the headroom takes (a record of megabytes): a buffer of the exact size
				Bytes whole;
				const size_t len = c.bytes.n - c.begin;
				whole.grow_to(tail + len);
				std::memcpy(whole.data() + tail, c.bytes.data() + c.begin, len);
				whole.n = tail + len;
				c.bytes = std::move(whole);
				c.begin = tail;
			}
			if (tail) std::memcpy(c.bytes.data() + c.begin - tail, _data.data() + _pos, tail);
			_pos = c.begin - tail;
			_spare = std::move(_data);
			_data = std::move(c.bytes);
			start_loading();
		}
		return true;
	}

	void BamReader::read_header()
	{
		if (!fill(12) || std::memcmp(_data.data() + _pos, "BAM\1", 4) != 0) throw std::runtime_error("not a BAM file: " + _file_name);
		const size_t l_text = le32(_data.data() + _pos + 4);
		_pos += 8;
		if (!fill(l_text + 4)) throw std::runtime_error("truncated BAM header in " + _file_name);
		_header_text.assign(reinterpret_cast<const char *>(_data.data() + _pos), l_text);
		_pos += l_text;
		const size_t n_ref = le32(_data.data() + _pos);
		_pos += 4;
		for (size_t r = 0; r < n_ref; ++r)
		{
			if (!fill(4)) throw std::runtime_error("truncated BAM header in " + _file_name);
			const size_t l_name = le32(_data.data() + _pos);
			_pos += 4;
			if (l_name == 0 || !fill(l_name + 4)) throw std::runtime_error("truncated BAM header in " + _file_name);
			_refs.emplace_back(reinterpret_cast<const char *>(_data.data() + _pos), l_name - 1);
			_ref_lengths.push_back(le32(_data.data() + _pos + l_name));
			_pos += l_name + 4;
		}
	}

	bool BamReader::view_at(size_t pos, RecordView &v, size_t &next_pos) const
	{
		if (_data.size() - pos < 4) return false;
		const size_t block_size = le32(_data.data() + pos);
		if (block_size < 32) throw std::runtime_error("malformed alignment record in " + _file_name);
		if (_data.size() - pos < 4 + block_size) return false;
		const uint8_t *p = _data.data() + pos + 4, *end = p + block_size;
		next_pos = pos + 4 + block_size;
		v.raw = p;
		v.raw_bytes = block_size;
		v.al.ref_id = int32_t(le32(p));
		v.al.position = int32_t(le32(p + 4));
		const size_t l_read_name = p[8];
		const size_t n_cigar = le16(p + 12);
		v.al.flag = le16(p + 14);
		const size_t l_seq = le32(p + 16);
		const uint8_t *q = p + 32;
		const size_t fixed = l_read_name + n_cigar * 4 + (l_seq + 1) / 2 + l_seq;
		if (size_t(end - q) < fixed || l_read_name == 0) throw std::runtime_error("malformed alignment record in " + _file_name);
		v.name_data = reinterpret_cast<const char *>(q);
		v.name_len = l_read_name - 1;
		{
			int32_t ref_len = 0;
			const uint8_t *cg = q + l_read_name;
			for (size_t k = 0; k < n_cigar; ++k)
			{
				const uint32_t op = le32(cg + 4 * k);
				switch (op & 0xF) { case 0: case 2: case 3: case 7: case 8: ref_len += int32_t(op >> 4); break; default: break; } // M D N = X
			}
			v.al.end_position = v.al.position + ref_len;
		}
		v.al.tag_data = q + fixed;
		v.al.tag_bytes = size_t(end - v.al.tag_data);
		return true;
	}

	bool BamReader::next(BamAlignment &al)
	{
		if (!fill(4)) return false;
		const size_t block_size = le32(_data.data() + _pos);
		if (block_size < 32) throw std::runtime_error("malformed alignment record in " + _file_name);
		if (!fill(4 + block_size)) throw std::runtime_error("truncated alignment record in " + _file_name);
		RecordView v;
		size_t np = 0;
		view_at(_pos, v, np);
		_pos = np;
		al = v.al;
		al.name.assign(v.name_data, v.name_len);
		return true;
	}

	void BamReader::next_batch(std::vector<RecordView> &out, size_t max_records)
	{
		out.clear();
		if (!fill(4)) return;
		const size_t block_size = le32(_data.data() + _pos);
		if (block_size < 32) throw std::runtime_error("malformed alignment record in " + _file_name);
		if (!fill(4 + block_size)) throw std::runtime_error("truncated alignment record in " + _file_name);
		RecordView v;
		size_t np = 0;
		const uint8_t *const base = _data.data(), *const limit = base + _data.size();
		while (out.size() < max_records && view_at(_pos, v, np)) // no fill() in here: the views stay valid
		{
			// the bytes were written by the inflating threads; where the next record starts is known only after this one's length was read,
			// so the walk is a chain of cache misses unless the lines a kilobyte ahead are already on their way
			for (const uint8_t *q = base + ((_pos + 1024) & ~size_t(63)); q < base + np + 1024 && q < limit; q += 64) __builtin_prefetch(q);
			out.push_back(v);
			_pos = np;
		}
	}

	namespace
	{
		// walks the tag block once; calls on_tag(two-character name, type, value, value bytes) for every tag
		template <class F> void walk_tags(const uint8_t *p, size_t n, F &&on_tag)
		{
			const uint8_t *end = p + n;
			auto bad = [] { throw std::runtime_error("malformed tag block in a BAM record"); };
			while (p < end)
			{
				if (end - p < 3) bad();
				const char t = char(p[2]);
				const uint8_t *v = p + 3;
				size_t len = 0;
				switch (t)
				{
				case 'A': case 'c': case 'C': len = 1; break;
				case 's': case 'S': len = 2; break;
				case 'i': case 'I': case 'f': len = 4; break;
				case 'Z': case 'H':
				{
					const void *z = std::memchr(v, 0, size_t(end - v));
					if (!z) bad();
					len = size_t(static_cast<const uint8_t *>(z) - v) + 1;
					break;
				}
				case 'B':
				{
					if (end - v < 5) bad();
					const char st = char(v[0]);
					const size_t cnt = le32(v + 1);
					const size_t es = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : (st == 'i' || st == 'I' || st == 'f') ? 4 : 0;
					if (!es) bad();
					len = 5 + cnt * es;
					break;
				}
				default: bad();
				}
				if (size_t(end - v) < len) bad();
				on_tag(p, t, v, len);
				p = v + len;
			}
		}

		const uint8_t *find_tag(const uint8_t *p, size_t n, const std::string &tag, char *type, size_t *value_bytes)
		{
			if (tag.size() != 2) return nullptr;
			const uint8_t *hit = nullptr;
			walk_tags(p, n, [&](const uint8_t *name, char t, const uint8_t *v, size_t len) {
				if (!hit && char(name[0]) == tag[0] && char(name[1]) == tag[1]) { hit = v; *type = t; *value_bytes = len; }
			});
			return hit;
		}
	}

	void list_tags(const uint8_t *tag_data, size_t tag_bytes, std::vector<TagSpan> &out)
	{
		out.clear();
		walk_tags(tag_data, tag_bytes, [&](const uint8_t *name, char, const uint8_t *v, size_t len) { out.push_back(TagSpan{name, size_t(v - name) + len}); });
	}

	bool BamAlignment::TagValue::string(std::string &out) const
	{
		if (type == 'Z' || type == 'H') { out.assign(reinterpret_cast<const char *>(value), bytes - 1); return true; }
		if (type == 'A') { out.assign(1, char(value[0])); return true; }
		return false;
	}

	void BamAlignment::find_tags(const std::string *const *tags, size_t n_tags, TagValue *found) const
	{
		for (size_t k = 0; k < n_tags; ++k) found[k] = TagValue();
		walk_tags(tag_data, tag_bytes, [&](const uint8_t *name, char t, const uint8_t *v, size_t len) {
			for (size_t k = 0; k < n_tags; ++k)
				if (!found[k].type && tags[k]->size() == 2 && char(name[0]) == (*tags[k])[0] && char(name[1]) == (*tags[k])[1])
				{
					found[k].type = t; found[k].value = v; found[k].bytes = len;
				}
		});
	}

	char BamAlignment::tag_type(const std::string &tag) const
	{
		char t = 0;
		size_t n = 0;
		return find_tag(tag_data, tag_bytes, tag, &t, &n) ? t : char(0);
	}

	bool BamAlignment::get_string_tag(const std::string &tag, std::string &value) const
	{
		char t = 0;
		size_t n = 0;
		const uint8_t *v = find_tag(tag_data, tag_bytes, tag, &t, &n);
		if (!v) return false;
		if (t == 'Z' || t == 'H') { value.assign(reinterpret_cast<const char *>(v), n - 1); return true; }
		if (t == 'A') { value.assign(1, char(v[0])); return true; }
		return false;
	}

	namespace
	{
		using Tools::GeneAnnotation::RefGenesContainer;

		void add_type(UMI::Mark &mark, Tools::GeneAnnotation::RecordType type) // UMI::Mark::add(GtfRecord::RecordType), UMI.cpp:87-100
		{
			if (type == Tools::GeneAnnotation::EXON) mark.add(UMI::Mark::HAS_EXONS);
			else if (type == Tools::GeneAnnotation::INTRON) mark.add(UMI::Mark::HAS_INTRONS);
			else throw std::runtime_error("Unexpected GtfRecord type: " + std::to_string(int(type)));
		}

		// ReadParamsParser::find_exon (.cpp:152-172): the one gene whose exon is hit; false when exons of two genes are
		bool find_exon(const RefGenesContainer::query_results_t &results, RefGenesContainer::QueryResult &exon)
		{
			for (auto const &r : results)
			{
				if (r.type != Tools::GeneAnnotation::EXON) continue;
				if (exon.gene_name.empty()) { exon = r; continue; }
				if (exon.gene_name != r.gene_name) return false;
			}
			return true;
		}

		// ReadParamsParser::get_gene_from_reference (.cpp:92-150): the annotation at the first and at the last aligned base decides
		UMI::Mark gene_from_reference(const RefGenesContainer &genes, const std::string &chr_name, const BamAlignment &al, std::string &gene)
		{
			UMI::Mark mark;
			const auto pos = RefGenesContainer::pos_t(al.position);
			const int end_position = al.end_position;
			auto set1 = genes.get_gene_info(chr_name, pos, pos + 1);
			auto set2 = genes.get_gene_info(chr_name, RefGenesContainer::pos_t(end_position - 1), RefGenesContainer::pos_t(end_position));
			if (set1.empty() && set2.empty()) return mark;
			if (set1.size() == 1 && set2.size() == 1)
			{
				if (set1.begin()->gene_name == set2.begin()->gene_name)
				{
					add_type(mark, set1.begin()->type);
					add_type(mark, set2.begin()->type);
					gene = set1.begin()->gene_name;
				}
				return mark;
			}
			if (set1.size() <= 1 && set2.size() <= 1)
			{   // one end in a gene, the other outside every gene
				auto const &hit = set1.empty() ? *set2.begin() : *set1.begin();
				gene = hit.gene_name;
				add_type(mark, hit.type);
				mark.add(UMI::Mark::HAS_NOT_ANNOTATED);
				return mark;
			}
			if (set1.empty() || set2.empty()) return mark;
			RefGenesContainer::QueryResult exon1, exon2;
			if (!find_exon(set1, exon1) || !find_exon(set2, exon2)) return mark;
			if (!exon1.gene_name.empty() && !exon2.gene_name.empty())
			{
				if (exon1.gene_name != exon2.gene_name) return mark;
				gene = exon1.gene_name;
				add_type(mark, exon1.type);
				add_type(mark, exon2.type);
			}
			return mark;
		}
	}

	ReadParamsMap::ReadParamsMap(const std::string &read_param_filenames, int min_barcode_quality)
	{
		const int min_phred = min_barcode_quality + 33; // ReadParameters::quality_to_phred
		size_t start = 0;
		while (start <= read_param_filenames.size())
		{
			size_t end = read_param_filenames.find_first_of(" \t", start);
			if (end == std::string::npos) end = read_param_filenames.size();
			std::string name = read_param_filenames.substr(start, end - start);
			start = end + 1;
			if (name.empty()) continue;
			if (name[0] == '~')
				if (const char *home = std::getenv("HOME")) name = std::string(home) + name.substr(1); // Tools::expand_tilde_in_path
			gzFile f = gzopen(name.c_str(), "rb");
			if (!f) throw std::runtime_error("Can't open file with read parameters'" + name + "'");
			std::string row;
			char buf[1 << 16];
			auto take_row = [&]() {
				if (row.empty()) return;
				// parse_from_string: four blanks split off name, barcode, UMI, barcode quality; the rest is the UMI quality
				size_t pos[4], p0 = 0;
				bool ok = true;
				for (int k = 0; k < 4 && ok; ++k)
				{
					pos[k] = row.find(' ', p0);
					if (pos[k] == std::string::npos) ok = false; else p0 = pos[k] + 1;
				}
				if (!ok) return; // "can't parse read parameters from string": logged and skipped
				std::string read_name = row.substr(0, pos[0]);
				if (!read_name.empty() && read_name[0] == '@') read_name.erase(0, 1);
				const std::string cb = row.substr(pos[0] + 1, pos[1] - pos[0] - 1), umi = row.substr(pos[1] + 1, pos[2] - pos[1] - 1),
				                  cbq = row.substr(pos[2] + 1, pos[3] - pos[2] - 1), umiq = row.substr(pos[3] + 1);
				if (cb.empty() || umi.empty()) return; // "Wrong read parameters"
				bool pass = true;
				if (min_phred > 33)
				{
					for (char c : cbq) if (c < min_phred) pass = false;
					for (char c : umiq) if (c < min_phred) pass = false;
				}
				// the indexers grow even when the name turns out to be a repeat, as the reference's do; only the first row of a name counts
				const Entry e{uint32_t(_barcodes.add(cb)), uint32_t(_umis.add(umi)), uint32_t(_umi_qualities.add(umiq)), pass, false};
				_reads.emplace(read_name, e);
			};
			while (gzgets(f, buf, int(sizeof(buf))))
			{
				const size_t n = std::strlen(buf);
				if (n && buf[n - 1] == '\n') { row.append(buf, n - 1); take_row(); row.clear(); }
				else row.append(buf, n);
			}
			take_row();
			gzclose(f);
		}
	}

	Tools::ReadParameters ReadParamsMap::parameters(const Entry &e) const
	{
		return Tools::ReadParameters(_barcodes.get_value(e.barcode), _umis.get_value(e.umi), "", _umi_qualities.get_value(e.umi_quality));
	}

	bool read_info_from_alignment(const BamAlignment &al, const std::string &chr_name, const IngestParams &params, IngestStats &stats,
	                              Tools::ReadParameters &read_params, std::string &gene, UMI::Mark &mark, MapLookup *map)
	{
		if (map) *map = MapLookup();
		// every tag this read can need, found in ONE walk over its tag block
		const std::string *wanted[6] = {&params.tags.cell_barcode, &params.tags.umi, &params.tags.cell_barcode_quality, &params.tags.umi_quality,
		                                &params.tags.gene, &params.tags.read_type};
		BamAlignment::TagValue tv[6];
		al.find_tags(wanted, 6, tv);
		// ---- barcode + UMI: FilledBamParamsParser::get_read_params (.cpp:12-40) / ReadParamsParser::get_read_params (.cpp:21-34)
		bool pass_quality = true;
		try
		{
			if (params.filled_bam)
			{
				std::string cb, umi, cbq, umiq;
				if (!tv[0].string(cb) || !tv[1].string(umi)) { ++stats.cant_parse; return false; }
				tv[2].string(cbq);
				tv[3].string(umiq);
				// ReadParameters::check_quality (Tools/ReadParameters.cpp:122-141): Phred+33 characters, every barcode and UMI base
				const int min_phred = params.min_barcode_quality + 33;
				if (min_phred > 33)
				{
					for (char c : cbq) if (c < min_phred) pass_quality = false;
					for (char c : umiq) if (c < min_phred) pass_quality = false;
				}
				read_params = Tools::ReadParameters(std::move(cb), std::move(umi), std::move(cbq), std::move(umiq));
			}
			else if (params.read_params)
			{   // ReadMapParamsParser::get_read_params (.cpp:22-48)
				ReadParamsMap::Entry *e = params.read_params->find(!al.name.empty() && al.name[0] == '@' ? al.name.substr(1) : al.name);
				if (!e) { ++stats.cant_parse; return false; }
				read_params = params.read_params->parameters(*e);
				if (map) map->entry = e;
				else
				{   // a caller that walks the reads itself, in order
					if (e->taken) { ++stats.cant_parse; return false; }
					e->taken = true;
					pass_quality = e->pass_quality;
				}
			}
			else read_params = Tools::ReadParameters::parse_encoded_id(al.name);
		}
		catch (std::runtime_error &)
		{
			++stats.cant_parse; // empty barcode / UMI, or a read name without "!CB#UMI"
			return false;
		}
		if (!pass_quality) { ++stats.low_quality; return false; }

		// ---- gene + mark: ReadParamsParser::get_gene / parse_read_type (.cpp:36-90)
		mark = UMI::Mark();
		gene.clear();
		if (params.gene_in_chromosome_name)
		{
			gene = chr_name;
			if (!chr_name.empty()) mark.add(UMI::Mark::HAS_EXONS);
			return true;
		}
		if (params.genes && !params.genes->is_empty())
		{
			try { mark = gene_from_reference(*params.genes, chr_name, al, gene); }
			catch (Tools::GeneAnnotation::RefGenesContainer::ChrNotFoundException &)
			{   // BamController.cpp:158-166
				if (map && map->entry) { map->chr_not_found = true; return true; }
				++stats.cant_parse;
				return false;
			}
			return true;
		}
		if (!tv[4].string(gene))
		{
			gene.clear();
			mark.add(UMI::Mark::HAS_NOT_ANNOTATED);
			return true;
		}
		std::string read_type;
		bool have_type = false;
		if (!params.tags.read_type.empty())
		{
			const char t = tv[5].type;
			if (t && t != 'Z' && t != 'A') throw std::runtime_error(std::string("Expected string tag, but got ") + t); // get_bam_tag, .cpp:179-197
			have_type = tv[5].string(read_type);
		}
		if (!have_type) mark.add(UMI::Mark::HAS_EXONS);
		else if (read_type == params.tags.intronic_read_value) mark.add(UMI::Mark::HAS_INTRONS);
		else if (!params.tags.intergenic_read_value.empty() && read_type == params.tags.intergenic_read_value) mark.add(UMI::Mark::HAS_NOT_ANNOTATED);
		else mark.add(UMI::Mark::HAS_EXONS);
		return true;
	}

	void parse_batch(const std::vector<BamReader::RecordView> &records, const std::vector<std::string> &refs, const IngestParams &params,
	                 std::vector<ParsedRead> &out, unsigned threads)
	{
		const size_t n_refs = refs.size();
		out.clear();
		out.resize(records.size());
		const unsigned nt = unsigned(std::min<size_t>(threads ? threads : std::max(1u, std::thread::hardware_concurrency()), (records.size() + 4095) / 4096));
		std::vector<std::string> errors(std::max(1u, nt));
		auto work = [&](unsigned t, size_t first, size_t last) {
			try
			{
				BamAlignment al;
				for (size_t k = first; k < last; ++k)
				{
					ParsedRead &r = out[k];
					al = records[k].al;
					r.ref_id = al.ref_id;
					if (!al.is_mapped() || !al.is_primary_alignment()) { r.status = ParsedRead::SKIPPED; continue; }
					if (al.ref_id < 0 || size_t(al.ref_id) >= n_refs) { r.status = ParsedRead::NO_CHROMOSOME; continue; }
					if (!params.filled_bam) al.name.assign(records[k].name_data, records[k].name_len);
					IngestStats st;
					if (read_info_from_alignment(al, refs[size_t(al.ref_id)], params, st, r.params, r.gene, r.mark, &r.map)) r.status = ParsedRead::OK;
					else r.status = st.low_quality ? ParsedRead::LOW_QUALITY : ParsedRead::CANT_PARSE;
				}
			}
			catch (std::exception &e) { errors[t] = e.what(); }
		};
		if (nt <= 1) work(0, 0, records.size());
		else
		{
			std::vector<std::thread> pool;
			const size_t per = (records.size() + nt - 1) / nt;
			for (unsigned t = 0; t < nt; ++t) pool.emplace_back(work, t, std::min(records.size(), size_t(t) * per), std::min(records.size(), size_t(t + 1) * per));
			for (auto &th : pool) th.join();
		}
		for (auto const &e : errors) if (!e.empty()) throw std::runtime_error(e);
	}

	bool packed_path_applies(const IngestParams &params)
	{
		// every mode but -r (the rows of the read-parameter files are claimed in stream order): barcode / UMI from the tags (-f) or from the
		// read name, gene from the gene tag, the chromosome name or the annotation (-g)
		return params.filled_bam || (!params.read_params && params.read_param_filenames.empty());
	}

	namespace
	{
		// 2-bit code per base (A C G T), 0xFF for every other character
		struct BaseCodes
		{
			uint8_t code[256];
			BaseCodes() { std::memset(code, 0xFF, sizeof(code)); code[uint8_t('A')] = 0; code[uint8_t('C')] = 1; code[uint8_t('G')] = 2; code[uint8_t('T')] = 3; }
		};
		const BaseCodes base_codes;

		inline bool pack_bases(const char *p, size_t n, uint64_t &out)
		{
			if (n > 32) return false;
			uint64_t v = 0;
			unsigned bad = 0;
			for (size_t i = 0; i < n; ++i) // no branch per base: random bases defeat the predictor
			{
				const unsigned b = base_codes.code[uint8_t(p[i])];
				bad |= b;
				v = (v << 2) | (b & 3u);
			}
			if (bad & 0x80u) return false;
			out = v;
			return true;
		}

		// TagValue::string as a view
		inline bool tag_view(const BamAlignment::TagValue &t, const char *&p, size_t &n)
		{
			if (t.type == 'Z' || t.type == 'H') { p = reinterpret_cast<const char *>(t.value); n = t.bytes - 1; return true; }
			if (t.type == 'A') { p = reinterpret_cast<const char *>(t.value); n = 1; return true; }
			p = nullptr; n = 0;
			return false;
		}
	}

	// read_info_from_alignment for every mode but -r, producing a PackedRead: same decisions, same counters, same exception.
	// params.genes must be loaded already when an annotation is used (parse_bam_files does that).
	void parse_batch_packed(const std::vector<BamReader::RecordView> &records, const std::vector<std::string> &refs, const IngestParams &params,
	                        PackedBatch &out, unsigned threads)
	{
		const size_t n_refs = refs.size(), n = records.size();
		out.status.assign(n, uint8_t(ParsedRead::SKIPPED));
		out.reads.resize(n);
		const unsigned nt = std::max(1u, unsigned(std::min<size_t>(threads ? threads : std::max(1u, std::thread::hardware_concurrency()), (n + 4095) / 4096)));
		if (out.arenas.size() < nt) out.arenas.resize(nt);
		std::vector<std::string> errors(nt);
		const int min_phred = params.min_barcode_quality + 33;
		const std::string *wanted[6] = {&params.tags.cell_barcode, &params.tags.umi, &params.tags.cell_barcode_quality, &params.tags.umi_quality,
		                                &params.tags.gene, &params.tags.read_type};
		const size_t max_gene_name = params.genes ? params.genes->max_gene_name_length() : 0; // arena room per read for a name from the annotation
		// the five strings kept per read are values of different tags, together shorter than the tag block -- unless one tag name was
		// configured for several of them, in which case the same value is kept several times
		size_t tag_copies = 1;
		for (int a = 0; a < 5; ++a)
			for (int b = a + 1; b < 5; ++b)
				if (*wanted[a] == *wanted[b]) tag_copies = 5;
		auto work = [&](unsigned t, size_t first, size_t last) {
			try
			{
				// every copied string is part of a tag block: an arena of their total size never moves.  The write position is a local of the
				// thread (the headers of out.arenas[] share cache lines: bumping them per string makes the threads fight over those lines)
				std::vector<char> &arena = out.arenas[t];
				size_t cap = 0;
				for (size_t k = first; k < last; ++k) cap += tag_copies * records[k].al.tag_bytes + records[k].name_len + max_gene_name;
				if (arena.size() < cap + 1) arena.resize(cap + 1);
				char *top = arena.data();
				auto keep = [&](const char *p, size_t len) -> const char * {
					char *at = top;
					if (len) std::memcpy(at, p, len);
					top += len;
					return at;
				};
				BamAlignment::TagValue tv[6];
				std::string from_reference;
				for (size_t k = first; k < last; ++k)
				{
					const BamAlignment &al = records[k].al;
					if (k + 8 < last)
					{   // the tag blocks were written by the inflating threads: fetch the one needed a few records from now
						const uint8_t *ahead = records[k + 8].al.tag_data;
						__builtin_prefetch(ahead); __builtin_prefetch(ahead + 64);
					}
					if (!al.is_mapped() || !al.is_primary_alignment()) continue;
					if (al.ref_id < 0 || size_t(al.ref_id) >= n_refs) { out.status[k] = ParsedRead::NO_CHROMOSOME; continue; }
					al.find_tags(wanted, 6, tv);
					const char *cb, *umi, *cbq = nullptr, *umiq = nullptr, *gene;
					size_t cb_n, umi_n, cbq_n = 0, umiq_n = 0, gene_n;
					if (params.filled_bam)
					{
						if (!tag_view(tv[0], cb, cb_n) || !tag_view(tv[1], umi, umi_n) || cb_n == 0 || umi_n == 0) { out.status[k] = ParsedRead::CANT_PARSE; continue; }
						tag_view(tv[2], cbq, cbq_n);
						tag_view(tv[3], umiq, umiq_n);
					}
					else
					{   // ReadParameters::parse_encoded_id on the name, as views: "...!<barcode>#<UMI>", the last '#', the last '!' before it
						const char *const begin = records[k].name_data, *const end = begin + records[k].name_len;
						const char *hash = nullptr, *bang = nullptr;
						for (const char *c = end; c != begin && !hash;) if (*--c == '#') hash = c;
						if (hash) for (const char *c = hash + 1; c != begin && !bang;) if (*--c == '!') bang = c;
						if (!hash || !bang || bang + 1 == hash || hash + 1 == end) { out.status[k] = ParsedRead::CANT_PARSE; continue; }
						cb = bang + 1; cb_n = size_t(hash - cb);
						umi = hash + 1; umi_n = size_t(end - umi);
					}
					bool pass_quality = true;
					if (min_phred > 33)
					{
						for (size_t i = 0; i < cbq_n; ++i) if (cbq[i] < min_phred) pass_quality = false;
						for (size_t i = 0; i < umiq_n; ++i) if (umiq[i] < min_phred) pass_quality = false;
					}
					if (!pass_quality) { out.status[k] = ParsedRead::LOW_QUALITY; continue; }
					unsigned mark = 0;
					if (params.gene_in_chromosome_name)
					{
						const std::string &chr = refs[size_t(al.ref_id)];
						gene = chr.data(); gene_n = chr.size();
						if (gene_n) mark = UMI::Mark::HAS_EXONS;
					}
					else if (params.genes && !params.genes->is_empty())
					{   // -g: the annotation decides (ReadParamsParser::get_gene_from_reference)
						from_reference.clear();
						try { mark = gene_from_reference(*params.genes, refs[size_t(al.ref_id)], al, from_reference).bits(); }
						catch (Tools::GeneAnnotation::RefGenesContainer::ChrNotFoundException &) { out.status[k] = ParsedRead::CANT_PARSE; continue; }
						gene = from_reference.data(); gene_n = from_reference.size();
					}
					else if (!tag_view(tv[4], gene, gene_n)) mark = UMI::Mark::HAS_NOT_ANNOTATED;
					else
					{
						const char *rt = nullptr;
						size_t rt_n = 0;
						bool have_type = false;
						if (!params.tags.read_type.empty())
						{
							const char t5 = tv[5].type;
							if (t5 && t5 != 'Z' && t5 != 'A') throw std::runtime_error(std::string("Expected string tag, but got ") + t5); // get_bam_tag
							have_type = tag_view(tv[5], rt, rt_n);
						}
						auto equals = [&](const std::string &v) { return v.size() == rt_n && std::memcmp(v.data(), rt, rt_n) == 0; };
						if (!have_type) mark = UMI::Mark::HAS_EXONS;
						else if (equals(params.tags.intronic_read_value)) mark = UMI::Mark::HAS_INTRONS;
						else if (!params.tags.intergenic_read_value.empty() && equals(params.tags.intergenic_read_value)) mark = UMI::Mark::HAS_NOT_ANNOTATED;
						else mark = UMI::Mark::HAS_EXONS;
					}
					if (cb_n > 0xFFFF || umi_n > 0xFFFF || gene_n > 0xFFFF || cbq_n > 0xFFFF || umiq_n > 0xFFFF) { out.status[k] = ParsedRead::CANT_PARSE; continue; }
					PackedRead &r = out.reads[k];
					r.packable = 0;
					uint64_t v = 0;
					if (pack_bases(cb, cb_n, v)) { r.cb_packed = v; r.packable |= 1; }
					if (umi_n <= 16 && pack_bases(umi, umi_n, v)) { r.umi_packed = uint32_t(v); r.packable |= 2; }
					r.cb = keep(cb, cb_n); r.cb_len = uint16_t(cb_n);
					r.umi = keep(umi, umi_n); r.umi_len = uint16_t(umi_n);
					r.cb_quality = keep(cbq, cbq_n); r.cb_quality_len = uint16_t(cbq_n);
					r.umi_quality = keep(umiq, umiq_n); r.umi_quality_len = uint16_t(umiq_n);
					r.gene = params.gene_in_chromosome_name ? gene : keep(gene, gene_n); r.gene_len = uint16_t(gene_n);
					r.gene_hash = StringIndexer::hash_of(gene, gene_n);
					r.mark_bits = uint8_t(mark);
					r.chromosome = al.ref_id;
					out.status[k] = ParsedRead::OK;
				}
			}
			catch (std::exception &e) { errors[t] = e.what(); }
		};
		if (nt <= 1) work(0, 0, n);
		else
		{
			std::vector<std::thread> pool;
			const size_t per = (n + nt - 1) / nt;
			for (unsigned t = 0; t < nt; ++t) pool.emplace_back(work, t, std::min(n, size_t(t) * per), std::min(n, size_t(t + 1) * per));
			for (auto &th : pool) th.join();
		}
		for (auto const &e : errors) if (!e.empty()) throw std::runtime_error(e);
	}

	namespace
	{
		// parse_bam_files over PackedBatch: the next batch is read, inflated and parsed while the current one goes to the container
		void parse_bam_files_packed(const std::vector<std::string> &bam_files, const IngestParams &params, CellsDataContainer &container, IngestStats &stats)
		{
			// Three stages run side by side: the reader's loader inflates chunk k + 1, a producer thread frames and parses the records of
			// chunk k batch by batch into a small ring, this thread hands the batches to the container in stream order.
			constexpr size_t RING = 4;
			struct Slot { PackedBatch batch; bool full = false; };
			for (auto const &file : bam_files)
			{
				BamReader reader(file, params.threads);
				const auto &refs = reader.reference_names();
				Slot ring[RING];
				std::mutex m;
				std::condition_variable cv;
				bool done = false, stop = false;
				std::exception_ptr failure;
				std::thread producer([&] {
					std::vector<BamReader::RecordView> views;
					try
					{
						for (size_t head = 0;; ++head)
						{
							Slot &slot = ring[head % RING];
							{
								std::unique_lock<std::mutex> lock(m);
								cv.wait(lock, [&] { return !slot.full || stop; });
								if (stop) break;
							}
							reader.next_batch(views, size_t(1) << 15); // a few batches per inflated chunk
							if (views.empty()) break;
							parse_batch_packed(views, refs, params, slot.batch, params.threads);
							{
								std::lock_guard<std::mutex> lock(m);
								slot.full = true;
							}
							cv.notify_all();
						}
					}
					catch (...) { failure = std::current_exception(); }
					{
						std::lock_guard<std::mutex> lock(m);
						done = true;
					}
					cv.notify_all();
				});
				try
				{
					for (size_t tail = 0;; ++tail)
					{
						Slot &slot = ring[tail % RING];
						{
							std::unique_lock<std::mutex> lock(m);
							cv.wait(lock, [&] { return slot.full || done; });
							if (!slot.full) break; // done, and every batch before it was taken
						}
						PackedBatch &cur = slot.batch;
						// the accepted reads, in stream order (it defines cell / gene / chromosome ids): compacted in place, so a batch
						// without a rejected record is handed over as it is
						size_t n_ok = 0;
						for (size_t k = 0; k < cur.status.size(); ++k)
						{
							switch (ParsedRead::Status(cur.status[k]))
							{
							case ParsedRead::SKIPPED: ++stats.skipped_unmapped_or_secondary; break;
							case ParsedRead::NO_CHROMOSOME: ++stats.cant_parse; break;
							case ParsedRead::CANT_PARSE: ++stats.total_reads; ++stats.cant_parse; break;
							case ParsedRead::LOW_QUALITY: ++stats.total_reads; ++stats.low_quality; break;
							case ParsedRead::OK:
								++stats.total_reads;
								if (n_ok != k) cur.reads[n_ok] = cur.reads[k];
								++n_ok;
								break;
							}
						}
						container.add_records(cur.reads.data(), n_ok, refs);
						{
							std::lock_guard<std::mutex> lock(m);
							slot.full = false;
						}
						cv.notify_all();
					}
				}
				catch (...)
				{
					{
						std::lock_guard<std::mutex> lock(m);
						stop = true;
					}
					cv.notify_all();
					producer.join();
					throw;
				}
				producer.join();
				if (failure) std::rethrow_exception(failure); // after the batches parsed before it went to the container, like the one-read path
			}
		}
	}

	void parse_bam_files(const std::vector<std::string> &bam_files, const IngestParams &params, CellsDataContainer &container, IngestStats &stats,
	                     bool print_result_bams)
	{
		if (!params.tags.read_type.empty() && params.tags.intronic_read_value.empty()) // BamTags.cpp:22-23
			throw std::runtime_error("You have to specify tag values to be able to parse info about read types");
		if (!print_result_bams && packed_path_applies(params) && !std::getenv("DGE_BAM_ONE_BY_ONE"))
		{
			IngestParams loaded(params);
			if (!loaded.genes && !loaded.genes_filename.empty())
				loaded.genes = std::make_shared<const Tools::GeneAnnotation::RefGenesContainer>(loaded.genes_filename);
			parse_bam_files_packed(bam_files, loaded, container, stats);
			return;
		}
		if (!print_result_bams)
		{
			for_each_read(bam_files, params, stats, [&](const ReadInfo &ri) { container.add_record(ri); });
			return;
		}
		// -b: BamProcessor::update_bam opens "<name>.tagged.bam" per input, write_alignment saves every accepted read before save_read
		// (BamProcessor.cpp:50-72, BamController.cpp:169-171)
		std::unique_ptr<BamWriter> writer;
		std::vector<BamWriter::TagEdit> edits;
		for_each_alignment(bam_files, params, stats, true,
			[&](const std::string &file, const BamReader &reader) {
				if (writer) writer->close();
				writer.reset(new BamWriter(result_bam_name(file, ".tagged.bam", params.output_dir), reader.header_text(), reader.reference_names(),
				                           reader.reference_lengths(), params.threads));
			},
			[&](const ReadInfo &ri, const BamReader::RecordView *view) {
				tag_edits(params.tags, ri, std::string(), std::string(), edits);
				writer->save_alignment(view->raw, view->raw_bytes, edits);
				container.add_record(ri);
			});
		if (writer) writer->close();
	}
}
}
