// GeneAnnotation.h -- gene / exon / intron lookup from a GTF or BED annotation: what `dropest -g` uses to turn an
// alignment position into a gene name and a UMI::Mark.  Mirrors the observable behaviour of the reference's
//   Tools/GeneAnnotation/RefGenesContainer.{h,cpp} (parsing rules :118-189, 222-237; transcripts and their exons :100-116, 84-97; queries
//   :191-220), IntervalsContainer.h (same-label intervals merged on insertion :152-188, sweep into homogeneous segments :101-133, queries
//   :214-236), GtfRecord.cpp, Interval.cpp
// including its quirks (asymmetric "touching" test of Interval::is_intercept, one-nucleotide queries, std::out_of_range on an empty GTF
// line), with its own data structures.  Pinned against the compiled reference (oracle/_ref/ref_gtf) in tests/test_gene_annotation.py.
#pragma once
#include <algorithm>
#include <cstddef>
#include <list>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace Tools
{
namespace GeneAnnotation
{
	enum RecordType { NONE = 0, INTRON = 1, EXON = 2 }; // GtfRecord::RecordType (GtfRecord.h:23-28)

	// Labelled half-open intervals -> "which labels cover [start, end)".  Same-label intervals are merged as they are added; seal()
	// sweeps all interval ends once and keeps the maximal segments over which the set of covering labels is constant.
	template <class Label> class IntervalIndex
	{
		struct Span { size_t start, end; };
		struct Segment { size_t start, end; std::set<Label> labels; };
		bool _allow_overlap, _sealed = false;
		std::map<Label, std::list<Span>> _spans;
		std::vector<Segment> _segments;

		// Interval::is_intercept (Interval.cpp:15-18): asymmetric -- `a` starting exactly where `b` ends counts, the converse does not
		static bool touches(const Span &a, const Span &b) { return a.start <= b.end && a.end > b.start; }
		static void hull(Span &a, const Span &b) { a.start = std::min(a.start, b.start); a.end = std::max(a.end, b.end); }

	public:
		explicit IntervalIndex(bool allow_overlap = true) : _allow_overlap(allow_overlap) {}

		void add(size_t start, size_t end, const Label &label)
		{
			if (_sealed) throw std::runtime_error("IntervalsContainer is already initialized");
			Span q{start, end};
			auto &spans = _spans[label];
			auto it = spans.begin();
			while (it != spans.end() && !touches(q, *it))
			{
				if (it->start > q.end) { spans.insert(it, q); return; }
				++it;
			}
			if (it == spans.end()) { spans.push_back(q); return; }
			auto last = std::next(it);
			while (last != spans.end() && touches(q, *last)) { hull(q, *last); ++last; }
			hull(*it, q);
			spans.erase(std::next(it), last);
		}

		void seal()
		{
			_sealed = true;
			struct Event { size_t pos; bool open; const Label *label; };
			std::vector<Event> events; // opening and closing event of every span, spans in (label, list) order: ties keep this order
			for (auto const &kv : _spans)
				for (auto const &s : kv.second) { events.push_back(Event{s.start, true, &kv.first}); events.push_back(Event{s.end, false, &kv.first}); }
			std::stable_sort(events.begin(), events.end(), [](const Event &a, const Event &b) { return a.pos < b.pos; });
			_segments.clear();
			size_t from = 0;
			std::set<Label> covering;
			for (auto const &e : events)
			{
				if (!covering.empty() && e.pos - from >= 1)
				{
					if (!_allow_overlap && covering.size() > 1)
						throw std::runtime_error("Intervals intersection at (" + std::to_string(from) + ", " + std::to_string(e.pos) + ")");
					_segments.push_back(Segment{from, e.pos, covering});
				}
				if (e.open) covering.insert(*e.label); else covering.erase(*e.label);
				from = e.pos;
			}
			_spans.clear();
		}

		std::set<Label> query(size_t start, size_t end) const
		{
			if (!_sealed) throw std::runtime_error("Interval must be initialized");
			auto it = std::lower_bound(_segments.begin(), _segments.end(), start, [](const Segment &s, size_t pos) { return s.end <= pos; });
			std::set<Label> out;
			for (; it != _segments.end() && it->start < end; ++it) out.insert(it->labels.begin(), it->labels.end());
			return out;
		}

		// query(pos, pos + 1) without the copy: the segments are disjoint, so at most one holds pos (nullptr: none)
		const std::set<Label> *query_point(size_t pos) const
		{
			if (!_sealed) throw std::runtime_error("Interval must be initialized");
			auto it = std::lower_bound(_segments.begin(), _segments.end(), pos, [](const Segment &s, size_t p) { return s.end <= p; });
			return it != _segments.end() && it->start <= pos ? &it->labels : nullptr;
		}

		size_t n_segments() const { return _segments.size(); }
	};

	class RefGenesContainer
	{
	public:
		using pos_t = unsigned long;

		class ChrNotFoundException : public std::runtime_error
		{
		public:
			const std::string chr_name;
			explicit ChrNotFoundException(const std::string &chr_name) : std::runtime_error("Can't find chromosome " + chr_name), chr_name(chr_name) {}
		};

		struct QueryResult
		{
			std::string gene_name;
			RecordType type;
			explicit QueryResult(const std::string &gene_name = "", RecordType type = NONE) : gene_name(gene_name), type(type) {}
			bool operator<(const QueryResult &other) const { return type == other.type ? gene_name < other.gene_name : type < other.type; }
		};
		using query_results_t = std::set<QueryResult>;

		RefGenesContainer() = default;                                  // empty: no annotation given
		explicit RefGenesContainer(const std::string &genes_filename);  // .gtf / .bed, optionally .gz
		// genes covering [start_pos, end_pos) (0-based) on chr_name with the kind of region; throws ChrNotFoundException
		query_results_t get_gene_info(const std::string &chr_name, pos_t start_pos, pos_t end_pos) const;
		bool is_empty() const { return _is_empty; }
		bool has_introns() const { return _gtf_has_transcripts || _use_introns_from_gtf; }
		size_t max_gene_name_length() const { return _max_gene_name; } // of every name a query can return

	private:
		struct Record // GtfRecord
		{
			std::string chr, gene_id, gene_name_raw, transcript_raw;
			size_t start = 0, end = 0;
			RecordType type = NONE;
			bool valid() const { return !gene_id.empty(); }
			const std::string &gene_name() const { return gene_name_raw.empty() ? gene_id : gene_name_raw; }
			const std::string &transcript_id() const { return transcript_raw.empty() ? gene_id : transcript_raw; }
		};
		bool _is_empty = true, _use_introns_from_gtf = false, _gtf_has_transcripts = true;
		size_t _max_gene_name = 0;
		std::string _file_format;
		std::unordered_map<std::string, IntervalIndex<std::string>> _transcripts;                                   // chr -> transcripts
		std::unordered_map<std::string, std::unordered_map<std::string, IntervalIndex<RecordType>>> _exons;          // chr -> transcript -> exon / intron spans
		std::unordered_map<std::string, std::unordered_map<std::string, std::pair<size_t, size_t>>> _transcript_span; // chr -> transcript -> [start, end)
		std::unordered_map<std::string, std::string> _gene_of_transcript;

		Record parse_gtf_line(const std::string &line);
		static Record parse_bed_line(const std::string &line);
		void save(const Record &record);
	};
}
}
