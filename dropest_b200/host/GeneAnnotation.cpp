// GeneAnnotation.cpp -- see GeneAnnotation.h
#include "GeneAnnotation.h"

#include <algorithm>
#include <cstdlib>
#include <iostream>
#include <sstream>

#include <zlib.h>

namespace Tools
{
namespace GeneAnnotation
{
	namespace
	{
		std::vector<std::string> split_ws(const std::string &line) // RefGenesContainer::split: any run of whitespace separates
		{
			std::istringstream in(line);
			std::vector<std::string> cols;
			std::string c;
			while (in >> c) cols.push_back(c);
			return cols;
		}

		// lines of a plain or gzip-compressed text file (zlib reads both), without the line terminator, like std::getline
		template <class F> void for_each_line(const std::string &fname, F &&f)
		{
			gzFile in = gzopen(fname.c_str(), "rb");
			if (!in) throw std::runtime_error("Can't open GTF file: '" + fname + "'");
			std::string line;
			char buf[1 << 16];
			try
			{
				while (true)
				{
					char *got = gzgets(in, buf, sizeof(buf));
					if (!got)
					{
						if (!line.empty()) f(line);
						break;
					}
					line += buf;
					if (!line.empty() && line.back() == '\n')
					{
						line.pop_back();
						f(line);
						line.clear();
					}
				}
			}
			catch (...) { gzclose(in); throw; }
			gzclose(in);
		}
	}

	RefGenesContainer::RefGenesContainer(const std::string &genes_filename) : _is_empty(false)
	{
		const std::string wrong_format = "Wrong genes file format: '" + genes_filename + "'";
		if (genes_filename.length() < 3) throw std::runtime_error(wrong_format);
		_file_format = genes_filename.substr(genes_filename.length() - 3);
		if (_file_format == ".gz")
		{
			if (genes_filename.length() < 6) throw std::runtime_error(wrong_format);
			_file_format = genes_filename.substr(genes_filename.length() - 6, 3);
		}
		if (_file_format != "bed" && _file_format != "gtf") throw std::runtime_error(wrong_format);

		for_each_line(genes_filename, [&](const std::string &line) {
			Record rec;
			try
			{
				rec = _file_format == "gtf" ? parse_gtf_line(line) : parse_bed_line(line);
			}
			catch (std::runtime_error &err) // unparsable lines are reported and skipped (RefGenesContainer.cpp:73-77); other exceptions end the load
			{
				std::cerr << err.what() << "\n";
				return;
			}
			if (rec.valid()) save(rec);
		});

		for (auto const &chr : _transcript_span)
		{
			auto &index = _transcripts.emplace(chr.first, IntervalIndex<std::string>(true)).first->second;
			for (auto const &tr : chr.second)
			{
				index.add(tr.second.first, tr.second.second, tr.first);
				_exons.at(chr.first).at(tr.first).seal(); // throws when an exon and an intron of one transcript overlap
			}
			index.seal();
		}
	}

	void RefGenesContainer::save(const Record &rec)
	{
		auto &span = _transcript_span[rec.chr].emplace(rec.transcript_id(), std::make_pair(rec.start, rec.end)).first->second;
		span.first = std::min(span.first, rec.start);
		span.second = std::max(span.second, rec.end);
		_exons[rec.chr].emplace(rec.transcript_id(), IntervalIndex<RecordType>(false)).first->second.add(rec.start, rec.end, rec.type);
		auto gene = _gene_of_transcript.emplace(rec.transcript_id(), rec.gene_name());
		_max_gene_name = std::max(_max_gene_name, rec.gene_name().size());
		if (!gene.second && gene.first->second != rec.gene_name())
			throw std::runtime_error("Different gene names (" + rec.gene_name() + ", " + gene.first->second + ") for the same transcript (" +
			                         rec.transcript_id() + ")");
	}

	RefGenesContainer::Record RefGenesContainer::parse_gtf_line(const std::string &line)
	{
		Record none;
		if (line.at(0) == '#') return none; // an empty line throws std::out_of_range here, as in the reference
		std::vector<std::string> col(split_ws(line));
		if (col.size() < 9) throw std::runtime_error("Can't parse record: \n" + line);
		if (col[0] == "." || col[3] == "." || col[4] == "." || col.size() == 9) return none;
		RecordType type;
		if (col[2] == "exon") type = EXON;
		else if (col[2] == "intron") { type = INTRON; _use_introns_from_gtf = true; }
		else return none;
		std::string id, name, transcript;
		for (size_t k = 8; k + 1 < col.size(); ++k)
		{
			const std::string &key = col[k], &value = col[k + 1];
			// "value"; -> value (the same substr call as the reference, corner cases included)
			if (key == "gene_id") id = value.substr(1, value.length() - 3);
			if (key == "gene_name") name = value.substr(1, value.length() - 3);
			if (key == "transcript_id") transcript = value.substr(1, value.length() - 3);
		}
		if (transcript.empty()) _gtf_has_transcripts = false;
		if (id.empty())
		{
			if (name.empty()) throw std::runtime_error("GTF record doesn't contain either gene name or id:\n" + line);
			id = name;
		}
		Record r;
		r.chr = col[0]; r.gene_id = id; r.gene_name_raw = name == id ? std::string() : name; r.transcript_raw = transcript;
		r.start = strtoul(col[3].c_str(), nullptr, 10) - 1; // GTF is 1-based, inclusive
		r.end = strtoul(col[4].c_str(), nullptr, 10);
		r.type = type;
		return r;
	}

	RefGenesContainer::Record RefGenesContainer::parse_bed_line(const std::string &line)
	{
		Record none;
		const size_t first = line.find_first_not_of("\t ");
		if (first == std::string::npos || line[first] == '#') return none;
		std::vector<std::string> col(split_ws(line));
		if (col.size() < 4) throw std::runtime_error("Bed record is too short:\n" + line);
		Record r;
		r.chr = col[0]; r.gene_id = col[3];
		r.start = strtoul(col[1].c_str(), nullptr, 10);
		r.end = strtoul(col[2].c_str(), nullptr, 10);
		r.type = EXON;
		return r;
	}

	RefGenesContainer::query_results_t RefGenesContainer::get_gene_info(const std::string &chr_name, pos_t start_pos, pos_t end_pos) const
	{
		if (end_pos < start_pos) return query_results_t();
		auto chr = _transcripts.find(chr_name);
		if (chr == _transcripts.end()) throw ChrNotFoundException(chr_name);
		query_results_t results;
		if (end_pos == start_pos + 1)
		{   // one nucleotide (what the BAM ingest asks, twice per read): the covering label sets are read in place, same iteration order
			const std::set<std::string> *transcripts = chr->second.query_point(start_pos);
			if (!transcripts) return results;
			auto const &exons_of = _exons.at(chr_name);
			for (const std::string &transcript : *transcripts)
			{
				const std::set<RecordType> *types = exons_of.at(transcript).query_point(start_pos);
				const std::string &gene = _gene_of_transcript.at(transcript);
				if (!types) { if (!_use_introns_from_gtf) results.emplace(gene, INTRON); continue; }
				for (RecordType t : *types) results.emplace(gene, t);
			}
			return results;
		}
		for (const std::string &transcript : chr->second.query(start_pos, end_pos))
		{
			auto types = _exons.at(chr_name).at(transcript).query(start_pos, end_pos);
			const std::string &gene = _gene_of_transcript.at(transcript);
			if (types.empty() && !_use_introns_from_gtf) { results.emplace(gene, INTRON); continue; } // inside the transcript, outside its exons
			for (RecordType t : types) results.emplace(gene, t);
		}
		return results;
	}
}
}
