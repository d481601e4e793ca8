// oracle/common/dge_io.h -- TEST INFRASTRUCTURE ONLY.
// File formats shared by the oracle drivers (oracle/ref_driver, oracle/port) and read from Python
// (dropest_b200/oracle_io.py).  Nothing in the product path includes this header.
//
//   DGER0001  packed read stream  : header + optional gene-name blob + dge_record16[n]
//   DGER0002  the same + two string lists (UMIs / barcodes containing N) between the gene names and the records; a record whose
//             gene word has bit 27 (bit 28) set carries an index into the N-UMI (N-barcode) list instead of a packed sequence
//             either format: header n_chr > 0 => one chromosome id (uint8) per record follows the records (chromosome name "chr<id>")
//   DGEO0001  oracle output       : directory of named typed arrays
//
// The 16-byte record is the same layout as include/dropest_b200.h:dge_record16 (restated here so the
// oracle does not depend on product headers):
//   key      bits[63:24] cell barcode, 2 bit/base A=0 C=1 G=2 T=3, first base most significant, right-aligned
//            bits[23:0]  UMI, same coding
//   gene     bits[23:0] gene id (0xFFFFFF = no gene / intergenic), bits[26:24] UMI::Mark bits, bit 27 / 28 N escapes (DGER0002)
//   read_idx global 0-based position of the read in the stream
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace dge_io
{
	struct Record16 { uint64_t key; uint32_t gene; uint32_t read_idx; };
	static const uint32_t NO_GENE = 0xFFFFFFu;

	inline std::string unpack_seq(uint64_t v, unsigned len)
	{
		std::string s(len, 'A');
		for (unsigned i = 0; i < len; ++i)
			s[len - 1 - i] = "ACGT"[(v >> (2 * i)) & 3];
		return s;
	}

	inline bool pack_seq(const std::string &s, uint64_t &out)
	{
		out = 0;
		for (char c : s)
		{
			unsigned b;
			switch (c) { case 'A': b = 0; break; case 'C': b = 1; break; case 'G': b = 2; break; case 'T': b = 3; break; default: return false; }
			out = (out << 2) | b;
		}
		return true;
	}

	struct ReadStream
	{
		uint32_t cb_len = 0, umi_len = 0, n_genes = 0, n_chr = 0;
		std::vector<std::string> gene_names; // empty => "g<id>"
		std::vector<std::string> n_umis, n_cbs; // DGER0002: sequences containing N, referenced by index
		std::vector<Record16> recs;
		std::vector<uint8_t> chr; // empty, or one chromosome id per record

		std::string chr_name(size_t i) const { return chr.empty() ? std::string() : "chr" + std::to_string(unsigned(chr[i])); }

		std::string cb_of(const Record16 &r) const
		{
			return (r.gene & (1u << 28)) ? n_cbs.at(size_t(r.key >> 24)) : unpack_seq(r.key >> 24, cb_len);
		}
		std::string umi_of(const Record16 &r) const
		{
			return (r.gene & (1u << 27)) ? n_umis.at(size_t(r.key & 0xFFFFFFu)) : unpack_seq(r.key & 0xFFFFFFu, umi_len);
		}

		std::string gene_name(uint32_t id) const
		{
			if (id == NO_GENE) return "";
			if (id < gene_names.size()) return gene_names[id];
			return "g" + std::to_string(id);
		}
	};

	inline ReadStream read_packed(const std::string &fname)
	{
		std::ifstream f(fname, std::ios::binary);
		if (!f) throw std::runtime_error("can't open " + fname);
		char magic[8];
		f.read(magic, 8);
		const bool v2 = std::memcmp(magic, "DGER0002", 8) == 0;
		if (!v2 && std::memcmp(magic, "DGER0001", 8) != 0) throw std::runtime_error("bad magic in " + fname);
		uint64_t n = 0, names_bytes = 0;
		ReadStream s;
		f.read((char *)&n, 8);
		f.read((char *)&s.cb_len, 4); f.read((char *)&s.umi_len, 4);
		f.read((char *)&s.n_genes, 4); f.read((char *)&s.n_chr, 4);
		f.read((char *)&names_bytes, 8);
		if (names_bytes)
		{
			std::string blob(names_bytes, '\0');
			f.read(&blob[0], names_bytes);
			size_t start = 0;
			while (start < blob.size())
			{
				size_t end = blob.find('\n', start);
				if (end == std::string::npos) end = blob.size();
				s.gene_names.push_back(blob.substr(start, end - start));
				start = end + 1;
			}
		}
		if (v2)
		{
			for (std::vector<std::string> *lst : {&s.n_umis, &s.n_cbs})
			{
				uint64_t bytes = 0;
				f.read((char *)&bytes, 8);
				std::string blob(bytes, '\0');
				if (bytes) f.read(&blob[0], bytes);
				size_t start = 0;
				while (start < blob.size())
				{
					size_t end = blob.find('\n', start);
					if (end == std::string::npos) end = blob.size();
					lst->push_back(blob.substr(start, end - start));
					start = end + 1;
				}
			}
		}
		s.recs.resize(n);
		f.read((char *)s.recs.data(), n * sizeof(Record16));
		if (s.n_chr)
		{
			s.chr.resize(n);
			f.read((char *)s.chr.data(), n);
		}
		if (!f) throw std::runtime_error("short read in " + fname);
		return s;
	}

	// ---- output: directory of named arrays --------------------------------------------------------
	enum DType : uint32_t { U8 = 0, I32 = 1, U32 = 2, I64 = 3, U64 = 4, F64 = 5 };

	class Writer
	{
		struct Section { std::string name; uint32_t dtype; std::vector<char> data; uint64_t count; };
		std::vector<Section> _sections;

		static size_t dsize(uint32_t t) { static const size_t s[] = {1, 4, 4, 8, 8, 8}; return s[t]; }

	public:
		template <class T> void add(const std::string &name, DType t, const std::vector<T> &v)
		{
			if (sizeof(T) != dsize(t)) throw std::runtime_error("dtype size mismatch for " + name);
			Section s; s.name = name; s.dtype = t; s.count = v.size();
			s.data.resize(v.size() * sizeof(T));
			if (!v.empty()) std::memcpy(s.data.data(), v.data(), s.data.size());
			_sections.push_back(std::move(s));
		}

		void add_blob(const std::string &name, const std::string &blob)
		{
			std::vector<uint8_t> v(blob.begin(), blob.end());
			add(name, U8, v);
		}

		void add_strings(const std::string &name, const std::vector<std::string> &strs)
		{
			std::string blob;
			for (auto const &s : strs) { blob += s; blob += '\n'; }
			add_blob(name, blob);
		}

		void add_scalar_i64(const std::string &name, int64_t x) { add(name, I64, std::vector<int64_t>(1, x)); }
		void add_scalar_f64(const std::string &name, double x) { add(name, F64, std::vector<double>(1, x)); }

		void write(const std::string &fname) const
		{
			std::ofstream f(fname, std::ios::binary);
			if (!f) throw std::runtime_error("can't write " + fname);
			f.write("DGEO0001", 8);
			uint64_t n = _sections.size();
			f.write((const char *)&n, 8);
			uint64_t offset = 16 + n * 56;
			for (auto const &s : _sections)
			{
				char name[32] = {0};
				std::strncpy(name, s.name.c_str(), 31);
				uint32_t pad = 0;
				uint64_t off = offset;
				f.write(name, 32); f.write((const char *)&s.dtype, 4); f.write((const char *)&pad, 4);
				f.write((const char *)&s.count, 8); f.write((const char *)&off, 8);
				offset += s.data.size();
			}
			for (auto const &s : _sections) f.write(s.data.data(), s.data.size());
		}
	};
}
