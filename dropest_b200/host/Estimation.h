// Estimation.h -- host-side C++ mirror of the reference's container surface for the count-matrix hot path.
//
// Same namespaces, class names, method names, argument meaning and exception behaviour as the reference
// (Estimation/CellsDataContainer.h:82-122, Cell.h, Gene.h, UMI.h, Stats.h, StringIndexer.h, ReadInfo.h,
// Merge/MergeStrategyFactory.h:55-58), so that dropest.cpp:239-254 and ResultsPrinter-style consumers compile against it
// unchanged.  All grouping / merging work is done by the CUDA library behind include/dropest_b200.h; this layer only packs
// reads into 16-byte records, batches them, and materialises query results lazily.  No CPU implementation of the path lives here.
#pragma once

#include "../../include/dropest_b200.h"

#include <map>
#include <memory>
#include <stdexcept>
#include <cstring>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace Tools
{
	// Tools::ReadParameters (Tools/ReadParameters.h:9-50): only what the container reads.
	class ReadParameters
	{
		std::string _cell_barcode, _umi, _cell_barcode_quality, _umi_quality;

	public:
		ReadParameters() = default; // the reference's empty parameters (ReadParameters.cpp:34-41)
		ReadParameters(const std::string &cell_barcode, const std::string &umi, const std::string &cell_barcode_quality,
		               const std::string &umi_quality)
			: _cell_barcode(cell_barcode), _umi(umi), _cell_barcode_quality(cell_barcode_quality), _umi_quality(umi_quality)
		{
			if (cell_barcode.empty() || umi.empty())
				throw std::runtime_error("Wrong read parameters: '" + cell_barcode + "' '" + umi + "'");
		}
		ReadParameters(std::string &&cell_barcode, std::string &&umi, std::string &&cell_barcode_quality, std::string &&umi_quality)
			: _cell_barcode(std::move(cell_barcode)), _umi(std::move(umi)), _cell_barcode_quality(std::move(cell_barcode_quality)), _umi_quality(std::move(umi_quality))
		{
			if (_cell_barcode.empty() || _umi.empty())
				throw std::runtime_error("Wrong read parameters: '" + _cell_barcode + "' '" + _umi + "'");
		}
		// "prefix!CB#UMI" read-name codec (Tools/ReadParameters.cpp:42-56)
		static ReadParameters parse_encoded_id(const std::string &encoded_id);
		std::string encoded_id(const std::string &id_prefix) const { return id_prefix + '!' + _cell_barcode + '#' + _umi; }
		const std::string &cell_barcode() const { return _cell_barcode; }
		const std::string &umi() const { return _umi; }
		const std::string &cell_barcode_quality() const { return _cell_barcode_quality; }
		const std::string &umi_quality() const { return _umi_quality; }
	};

	unsigned edit_distance(const char *s1, const char *s2, bool skip_n = true, unsigned max_ed = 10000); // UtilFunctions.cpp:32-65
	unsigned hamming_distance(const std::string &s1, const std::string &s2, bool skip_n = true);         // UtilFunctions.cpp:67-82

	// Tools::CollisionsAdjuster (Tools/CollisionsAdjuster.h:8-33): same interface; the table of adjusted sizes is computed on the
	// device (dge_collisions_adjusted_sizes) and extended on demand like update_adjusted_sizes does.
	class CollisionsAdjuster
	{
	public:
		using size_vec_t = std::vector<size_t>;
		using probs_vec_t = std::vector<double>;

	private:
		size_vec_t _adjusted_sizes;
		probs_vec_t _umi_probabilities;
		int _device;
		void update_adjusted_sizes(size_t max_gene_expression);

	public:
		explicit CollisionsAdjuster(int device = 0) : _device(device) {}
		void init(const probs_vec_t &umi_probabilities, size_t max_gene_expression = 0);
		size_t estimate_adjusted_gene_expression(size_t expression);
	};
}

namespace Estimation
{
	class StringIndexer // StringIndexer.h: strings <-> ids in first-seen order
	{
	public:
		using index_t = size_t;
		using values_t = std::vector<std::string>;

	private:
		values_t _values;
		// open addressing over _values (index + 1, 0 = empty): a lookup by (pointer, length) builds no std::string, which is what the per-read
		// paths need (gene names of BAM records)
		std::vector<uint32_t> _slots;

		static uint64_t hash(const char *p, size_t n)
		{
			uint64_t h = 0xCBF29CE484222325ull;
			for (size_t i = 0; i < n; ++i) { h ^= uint8_t(p[i]); h *= 0x100000001B3ull; }
			return h ^ (h >> 29);
		}
		// slot of the string, or the empty slot where it would go
		size_t probe(const char *p, size_t n) const { return probe(p, n, hash(p, n)); }
		size_t probe(const char *p, size_t n, uint64_t h) const
		{
			const size_t mask = _slots.size() - 1;
			for (size_t i = size_t(h) & mask;; i = (i + 1) & mask)
			{
				const uint32_t v = _slots[i];
				if (!v) return i;
				const std::string &s = _values[v - 1];
				if (s.size() == n && std::memcmp(s.data(), p, n) == 0) return i;
			}
		}
		void grow()
		{
			std::vector<uint32_t> fresh(std::max<size_t>(64, _slots.size() * 2), 0u);
			_slots.swap(fresh);
			for (size_t k = 0; k < _values.size(); ++k) _slots[probe(_values[k].data(), _values[k].size())] = uint32_t(k + 1);
		}

	public:
		const values_t &values() const { return _values; }
		const std::string &get_value(index_t index) const { return _values.at(index); }
		index_t get_index(const std::string &value) const
		{
			if (!_slots.empty()) { const uint32_t v = _slots[probe(value.data(), value.size())]; if (v) return v - 1; }
			throw std::out_of_range("StringIndexer::get_index: unknown value '" + value + "'"); // the reference's unordered_map::at
		}
		index_t add(const char *p, size_t n) { return add(p, n, hash(p, n)); }
		// the same with the hash computed by the caller (the BAM parsing threads hash the gene names; the thread that owns the indexer only probes)
		static uint64_t hash_of(const char *p, size_t n) { return hash(p, n); }
		void prefetch(uint64_t h) const { if (!_slots.empty()) __builtin_prefetch(&_slots[size_t(h) & (_slots.size() - 1)]); }
		index_t add(const char *p, size_t n, uint64_t h)
		{
			if ((_values.size() + 1) * 2 > _slots.size()) grow();
			const size_t i = probe(p, n, h);
			if (_slots[i]) return _slots[i] - 1;
			if (_values.size() >= 0xFFFFFFFEull) throw std::runtime_error("StringIndexer: too many values");
			_values.emplace_back(p, n);
			_slots[i] = uint32_t(_values.size());
			return _values.size() - 1;
		}
		index_t add(const std::string &value) { return add(value.data(), value.size()); }
	};

	class UMI // UMI.h
	{
	public:
		class Mark
		{
		public:
			enum MarkType { NONE = 0, HAS_NOT_ANNOTATED = 1, HAS_EXONS = 2, HAS_INTRONS = 4 };
			using query_t = std::vector<Mark>;

		private:
			char _mark;

		public:
			static const std::string DEFAULT_CODE; // "eEBA" (CellsDataContainer.cpp:17)
			explicit Mark(MarkType type = NONE) : _mark(char(type)) {}
			void add(const Mark &mark) { _mark |= mark._mark; }
			void add(MarkType type) { _mark |= char(type); }
			bool check(MarkType type) const { return _mark & type; }
			bool match(const std::vector<Mark> &match_levels) const // exact equality, UMI.cpp:76-85
			{
				for (auto const &m : match_levels)
					if (_mark == m._mark) return true;
				return false;
			}
			bool operator==(const MarkType &other) const { return _mark == other; }
			bool operator==(const Mark &other) const { return _mark == other._mark; }
			int bits() const { return _mark; }
			static Mark get_by_code(char code);                       // UMI.cpp:123-154
			static std::vector<Mark> get_by_code(const std::string &code);
		};

	private:
		size_t _read_count;
		Mark _mark;
		std::vector<unsigned> _sum_quality; // UMI.h:51; filled only by a container built with save_umi_qualities
		friend class CellsDataContainer;

	public:
		explicit UMI(size_t read_count = 0, Mark mark = Mark()) : _read_count(read_count), _mark(mark) {}
		size_t read_count() const { return _read_count; }
		const Mark &mark() const { return _mark; }
		// UMI.cpp:46-55, literally: the Phred offset is subtracted ONCE from the per-base sum of quality characters and the difference is
		// divided by the read count in unsigned integer arithmetic.  The sum covers the reads added to THIS object (UMI::add_read): UMI::merge
		// (:15-19) adds read counts and marks but not qualities, so after barcode / UMI merges it is the first holder's sum over the merged count.
		std::vector<double> mean_quality() const
		{
			std::vector<double> res(_sum_quality.size());
			for (size_t i = 0; i < res.size(); ++i) res[i] = double(size_t(unsigned(_sum_quality[i] - unsigned(33))) / _read_count);
			return res;
		}
	};

	class ReadInfo // ReadInfo.h:9-24
	{
	public:
		const Tools::ReadParameters params;
		const std::string gene;
		const std::string chromosome_name;
		const UMI::Mark umi_mark;
		ReadInfo(const Tools::ReadParameters &params, const std::string &gene, const std::string &chromosome_name, const UMI::Mark &umi_mark)
			: params(params), gene(gene), chromosome_name(chromosome_name), umi_mark(umi_mark) {}
		ReadInfo(Tools::ReadParameters &&params, std::string &&gene, const std::string &chromosome_name, const UMI::Mark &umi_mark) // no string copies (BAM ingest)
			: params(std::move(params)), gene(std::move(gene)), chromosome_name(chromosome_name), umi_mark(umi_mark) {}
	};

	// One accepted read as the BAM ingest hands it over in bulk (CellsDataContainer::add_records): views into the ingest's buffers instead of
	// the five std::strings of a ReadInfo, and the barcode / UMI already 2-bit packed when they consist of A, C, G, T only.
	struct PackedRead
	{
		const char *cb, *umi, *gene, *cb_quality, *umi_quality;
		uint16_t cb_len, umi_len, gene_len, cb_quality_len, umi_quality_len;
		uint8_t mark_bits;     // UMI::Mark bits of the read
		uint8_t packable;      // bit 0: `cb_packed` is valid, bit 1: `umi_packed` is valid
		uint64_t cb_packed;
		uint32_t umi_packed;
		int32_t chromosome;    // index into the chromosome-name list passed with the batch
		uint64_t gene_hash;    // StringIndexer::hash_of(gene, gene_len)
	};

	class Stats // Stats.h (per-cell counters; the per-chromosome tables are read from the device by CellsDataContainer::get_stat_by_real_cells)
	{
	public:
		enum CellStatType { TOTAL_READS_PER_CB, TOTAL_UMIS_PER_CB, CELL_STAT_SIZE };
		enum CellChrStatType { EXON_READS_PER_CHR_PER_CELL = 0, INTRON_READS_PER_CHR_PER_CELL, INTERGENIC_READS_PER_CHR_PER_CELL, CHROMOSOME_STAT_SIZE };
		using stat_t = int;

	private:
		int _stat_data[CELL_STAT_SIZE];

	public:
		Stats() { _stat_data[0] = _stat_data[1] = 0; }
		Stats(int reads, int umis) { _stat_data[TOTAL_READS_PER_CB] = reads; _stat_data[TOTAL_UMIS_PER_CB] = umis; }
		stat_t get(CellStatType type) const { return _stat_data[type]; }
	};

	class Gene // Gene.h
	{
	public:
		using umis_t = std::map<StringIndexer::index_t, UMI>;
		using s_s_hash_t = std::unordered_map<std::string, std::string>;

	private:
		umis_t _umis;
		s_s_hash_t _merge_targets; // Gene.h:27, filled by Gene::merge(source_umi, target_umi) when the container keeps them (Gene.cpp:54-57)
		const StringIndexer *_umi_indexer;
		friend class CellsDataContainer;

	public:
		explicit Gene(const StringIndexer *umi_indexer = nullptr) : _umi_indexer(umi_indexer) {}
		const UMI &at(const std::string &umi) const { return _umis.at(_umi_indexer->get_index(umi)); }
		const umis_t &umis() const { return _umis; }
		size_t size() const { return _umis.size(); }
		bool has(const std::string &umi) const;
		// the UMIs the UMI merge strategy moved away from this gene and where they went (empty unless save_umi_merge_targets; Gene.cpp:125-128)
		const s_s_hash_t &merge_targets() const { return _merge_targets; }
		size_t number_of_requested_umis(const UMI::Mark::query_t &query, bool return_reads) const; // Gene.cpp:60-79
		size_t number_of_umis(bool return_reads) const;                                             // Gene.cpp:81-93
		using s_ul_hash_t = std::unordered_map<std::string, size_t>;
		s_ul_hash_t requested_reads_per_umi(const UMI::Mark::query_t &query) const;                 // Gene.cpp:95-107
	};

	class Cell // Cell.h
	{
	public:
		using genes_t = std::map<StringIndexer::index_t, Gene>;
		using s_ul_hash_t = std::unordered_map<std::string, size_t>;

	private:
		std::string _barcode;
		bool _is_real = false, _is_merged = false, _is_excluded = false;
		size_t _requested_genes_num = 0, _requested_umis_num = 0, _n_genes = 0;
		genes_t _genes;
		Stats _stats;
		const StringIndexer *_gene_indexer = nullptr;
		friend class CellsDataContainer;

	public:
		bool is_merged() const { return _is_merged; }
		bool is_excluded() const { return _is_excluded; }
		bool is_real() const { return _is_real; }                       // Cell.cpp:125-128
		std::string barcode() const { return _barcode; }
		const char *barcode_c() const { return _barcode.c_str(); }
		size_t umis_number() const { return size_t(_stats.get(Stats::TOTAL_UMIS_PER_CB)); } // Cell.cpp:105-108
		size_t requested_genes_num() const { return _requested_genes_num; }
		size_t requested_umis_num() const { return _requested_umis_num; }
		const Stats &stats() const { return _stats; }
		const genes_t &genes() const { return _genes; }
		size_t size() const { return _n_genes; }
		const Gene &at(const std::string &gene) const { return _genes.at(_gene_indexer->get_index(gene)); }
		s_ul_hash_t requested_umis_per_gene(const UMI::Mark::query_t &query_marks, bool return_reads) const; // Cell.cpp:54-68
		using ss_ul_hash_t = std::unordered_map<std::string, s_ul_hash_t>;
		ss_ul_hash_t requested_reads_per_umi_per_gene(const UMI::Mark::query_t &query_marks) const;          // Cell.cpp:70-83
	};

	namespace Merge
	{
		// Strategy objects are descriptors here: they carry exactly the constructor arguments of the reference classes and are
		// translated into dge_config; the strategy code itself runs on the device (dropest_b200/csrc/merge.cuh + engine.cu).
		class MergeStrategyAbstract
		{
			size_t _min_genes_before_merge, _min_genes_after_merge;

		public:
			MergeStrategyAbstract(size_t min_genes_before_merge, size_t min_genes_after_merge)
				: _min_genes_before_merge(min_genes_before_merge)
				, _min_genes_after_merge(std::max(min_genes_after_merge, min_genes_before_merge)) {} // MergeStrategyAbstract.cpp:8-11
			virtual ~MergeStrategyAbstract() {}
			virtual std::string merge_type() const = 0;
			virtual void configure(dge_config &cfg) const = 0;
			size_t min_genes_before_merge() const { return _min_genes_before_merge; }
			size_t min_genes_after_merge() const { return _min_genes_after_merge; }
		};

		class DummyMergeStrategy : public MergeStrategyAbstract
		{
		public:
			DummyMergeStrategy(size_t before, size_t after) : MergeStrategyAbstract(before, after) {}
			std::string merge_type() const override { return "No"; }
			void configure(dge_config &cfg) const override { cfg.merge_type = DGE_MERGE_NONE; }
		};

		namespace BarcodesParsing
		{
			// file name + flavour; loading/reverse-complementing happens in the library (whitelist.hpp)
			class BarcodesParser
			{
			public:
				std::string filename;
				bool indrop;
				BarcodesParser(const std::string &barcodes_filename, bool indrop) : filename(barcodes_filename), indrop(indrop) {}
				virtual ~BarcodesParser() {}
			};
			class InDropBarcodesParser : public BarcodesParser
			{
			public:
				explicit InDropBarcodesParser(const std::string &f) : BarcodesParser(f, true) {}
			};
			class ConstLengthBarcodesParser : public BarcodesParser
			{
			public:
				explicit ConstLengthBarcodesParser(const std::string &f) : BarcodesParser(f, false) {}
			};
		}

		class RealBarcodesMergeStrategy : public MergeStrategyAbstract
		{
		public:
			using barcodes_parser_ptr = std::shared_ptr<BarcodesParsing::BarcodesParser>;

		private:
			barcodes_parser_ptr _parser;
			unsigned _max_merge_edit_distance;
			double _min_merge_fraction;

		public:
			RealBarcodesMergeStrategy(const barcodes_parser_ptr &barcodes_parser, size_t min_genes_before_merge, size_t min_genes_after_merge,
			                          unsigned max_merge_edit_distance, double min_merge_fraction)
				: MergeStrategyAbstract(min_genes_before_merge, min_genes_after_merge), _parser(barcodes_parser)
				, _max_merge_edit_distance(max_merge_edit_distance), _min_merge_fraction(min_merge_fraction) {}
			std::string merge_type() const override { return "Real CBs"; }
			void configure(dge_config &cfg) const override;
		};

		// SimpleMergeStrategy (SimpleMergeStrategy.h:12-38): no whitelist, candidates = cells sharing UMI-genes
		class SimpleMergeStrategy : public MergeStrategyAbstract
		{
			unsigned _max_merge_edit_distance;
			double _min_merge_fraction;

		public:
			SimpleMergeStrategy(size_t min_genes_before_merge, size_t min_genes_after_merge, unsigned max_merge_edit_distance, double min_merge_fraction)
				: MergeStrategyAbstract(min_genes_before_merge, min_genes_after_merge)
				, _max_merge_edit_distance(max_merge_edit_distance), _min_merge_fraction(min_merge_fraction) {}
			std::string merge_type() const override { return "Simple"; }
			void configure(dge_config &cfg) const override
			{
				cfg.merge_type = DGE_MERGE_SIMPLE;
				cfg.max_cb_merge_edit_distance = _max_merge_edit_distance;
				cfg.min_merge_fraction = _min_merge_fraction;
			}
		};

		// MergeAllMergeStrategy (MergeAllMergeStrategy.h:13-61): nearest larger cell within the edit distance
		class MergeAllMergeStrategy : public MergeStrategyAbstract
		{
			unsigned _max_merge_edit_distance;

		public:
			MergeAllMergeStrategy(size_t min_genes_before_merge, size_t min_genes_after_merge, unsigned max_merge_edit_distance)
				: MergeStrategyAbstract(min_genes_before_merge, min_genes_after_merge), _max_merge_edit_distance(max_merge_edit_distance) {}
			std::string merge_type() const override { return "Merge all"; }
			void configure(dge_config &cfg) const override
			{
				cfg.merge_type = DGE_MERGE_ALL;
				cfg.max_cb_merge_edit_distance = _max_merge_edit_distance;
				cfg.min_merge_fraction = 0;
			}
		};

		// PoissonTargetEstimator (PoissonTargetEstimator.h:20-66): here the two thresholds; the estimator itself (UMI distribution,
		// CollisionsAdjuster, per-pair lambda and Poisson tail) runs on the device (dropest_b200/csrc/poisson.cuh)
		class PoissonTargetEstimator
		{
		public:
			const double max_merge_prob, max_real_cb_merge_prob;
			PoissonTargetEstimator(double max_merge_prob, double max_real_cb_merge_prob)
				: max_merge_prob(max_merge_prob), max_real_cb_merge_prob(max_real_cb_merge_prob) {}
		};

		// PoissonSimpleMergeStrategy (PoissonSimpleMergeStrategy.h): -M without a whitelist
		class PoissonSimpleMergeStrategy : public MergeStrategyAbstract
		{
			PoissonTargetEstimator _target_estimator;
			unsigned _max_merge_edit_distance;

		public:
			PoissonSimpleMergeStrategy(const PoissonTargetEstimator &target_estimator, unsigned min_genes_before_merge, unsigned min_genes_after_merge,
			                           unsigned max_merge_edit_distance)
				: MergeStrategyAbstract(min_genes_before_merge, min_genes_after_merge), _target_estimator(target_estimator)
				, _max_merge_edit_distance(max_merge_edit_distance) {}
			std::string merge_type() const override { return "Poisson Simple"; }
			void configure(dge_config &cfg) const override
			{
				cfg.merge_type = DGE_MERGE_POISSON_SIMPLE;
				cfg.max_cb_merge_edit_distance = _max_merge_edit_distance;
				cfg.min_merge_fraction = 0;
				cfg.max_merge_prob = _target_estimator.max_merge_prob;
				cfg.max_real_merge_prob = _target_estimator.max_real_cb_merge_prob;
			}
		};

		// PoissonRealBarcodesMergeStrategy (PoissonRealBarcodesMergeStrategy.h): -M with a whitelist
		class PoissonRealBarcodesMergeStrategy : public MergeStrategyAbstract
		{
			PoissonTargetEstimator _target_estimator;
			RealBarcodesMergeStrategy::barcodes_parser_ptr _parser;
			unsigned _max_merge_edit_distance;

		public:
			PoissonRealBarcodesMergeStrategy(const PoissonTargetEstimator &target_estimator, const RealBarcodesMergeStrategy::barcodes_parser_ptr &barcodes_parser,
			                                 size_t min_genes_before_merge, size_t min_genes_after_merge, unsigned max_merge_edit_distance)
				: MergeStrategyAbstract(min_genes_before_merge, min_genes_after_merge), _target_estimator(target_estimator), _parser(barcodes_parser)
				, _max_merge_edit_distance(max_merge_edit_distance) {}
			std::string merge_type() const override { return "Poisson Real CBs"; }
			void configure(dge_config &cfg) const override
			{
				cfg.merge_type = DGE_MERGE_POISSON_REAL;
				cfg.barcodes_type = _parser->indrop ? DGE_BARCODES_INDROP : DGE_BARCODES_CONST;
				cfg.barcodes_file = _parser->filename.c_str();
				cfg.max_cb_merge_edit_distance = _max_merge_edit_distance;
				cfg.min_merge_fraction = 0;
				cfg.max_merge_prob = _target_estimator.max_merge_prob;
				cfg.max_real_merge_prob = _target_estimator.max_real_cb_merge_prob;
			}
		};

		namespace UMIs
		{
			class MergeUMIsStrategyAbstract
			{
			public:
				virtual ~MergeUMIsStrategyAbstract() {}
				virtual void configure(dge_config &cfg) const = 0;
			};
			class MergeUMIsStrategySimple : public MergeUMIsStrategyAbstract
			{
				unsigned _max_merge_distance;

			public:
				explicit MergeUMIsStrategySimple(unsigned max_merge_distance) : _max_merge_distance(max_merge_distance) {}
				void configure(dge_config &cfg) const override
				{
					cfg.umi_merge_type = DGE_UMI_MERGE_SIMPLE;
					cfg.max_umi_merge_edit_distance = _max_merge_distance;
				}
			};
			// MergeUMIsStrategyDirectional (MergeUMIsStrategyDirectional.h:13-44): the container-wide merge runs on the device;
			// find_targets is the reference's public per-segment entry point, kept as a host function for callers and tests.
			class MergeUMIsStrategyDirectional : public MergeUMIsStrategyAbstract
			{
			public:
				struct UmiWrap
				{
					std::string sequence;
					size_t n_reads;
					UmiWrap(const std::string &sequence, size_t n_reads) : sequence(sequence), n_reads(n_reads) {}
				};
				using umi_vec_t = std::vector<UmiWrap>;
				using merge_targets_t = std::unordered_map<std::string, std::string>;

				explicit MergeUMIsStrategyDirectional(double mult = 2, unsigned max_edit_distance = 1) : _mult(mult), _max_edit_distance(max_edit_distance) {}
				merge_targets_t find_targets(umi_vec_t &umis) const;
				void configure(dge_config &cfg) const override
				{
					cfg.umi_merge_type = DGE_UMI_MERGE_DIRECTIONAL;
					cfg.max_umi_merge_edit_distance = _max_edit_distance;
					cfg.umi_merge_mult = _mult;
				}

			private:
				double _mult;
				unsigned _max_edit_distance;
			};
		}

		// MergeStrategyFactory (MergeStrategyFactory.cpp:23-126) with the XML values passed in directly (no boost::property_tree here)
		class MergeStrategyFactory
		{
		public:
			std::string merge_type = "none", barcodes_type = "indrop", barcodes_filename;
			size_t min_genes_before_merge = 10, min_genes_after_merge = 10;
			unsigned max_merge_edit_distance = 2, max_umi_merge_edit_distance = 1;
			double min_merge_fraction = 0.2, max_merge_prob = 1e-4, max_real_cb_merge_prob = 1e-7, umi_merge_mult = 2;

			MergeStrategyFactory() = default;
			// MergeStrategyFactory(const ptree &config, const std::string &config_file_name, int min_genes_after_merge), MergeStrategyFactory.cpp:23-59:
			// the <Estimation> block of a dropEst XML configuration (configs/*.xml).  Same keys, defaults and errors: max_cb_merge_edit_distance
			// is mandatory, barcodes_file is resolved against the configuration file's directory (and ~/) and must exist, a positive
			// min_genes_after_merge argument (the -G option) overrides the file.
			static MergeStrategyFactory from_xml(const std::string &config_file_name, int min_genes_after_merge = -1);

			std::shared_ptr<MergeStrategyAbstract> get_cb_strat(bool merge_tags, bool use_poisson) const;
			std::shared_ptr<BarcodesParsing::BarcodesParser> get_barcodes_parser() const; // MergeStrategyFactory.cpp:113-126
			std::shared_ptr<UMIs::MergeUMIsStrategyAbstract> get_umi(bool advanced) const;
		};
	}

	class CellsDataContainer // CellsDataContainer.h:33-122
	{
	public:
		using s_ul_hash_t = std::unordered_map<std::string, size_t>;
		using s_i_hash_t = std::unordered_map<std::string, int>;
		using ids_t = std::vector<size_t>;
		using names_t = std::vector<std::string>;

	private:
		std::shared_ptr<Merge::MergeStrategyAbstract> _merge_strategy;
		std::shared_ptr<Merge::UMIs::MergeUMIsStrategyAbstract> _umi_merge_strategy;
		const int _max_cells_num;
		const UMI::Mark::query_t _query_marks;
		bool _is_initialized = false, _is_merged = false;
		int _device;
		bool _reads_output;
		bool _save_umi_merge_targets;

		dge_handle *_h = nullptr;
		unsigned _cb_len = 0, _umi_len = 0;
		std::vector<uint64_t> _batch_keys;  // pending records as two arrays (dge_add_batch_soa), flushed in blocks of _batch_capacity
		std::vector<uint32_t> _batch_genes;
		std::vector<uint32_t> _batch_idx;   // stream positions of the pending records (used when skipped reads left gaps)
		std::vector<uint8_t> _batch_chr;    // chromosome id of the pending records (Stats' per-chromosome tables)
		// Stats' process-wide statics (Stats.cpp:5-7, 23-28, 79-88): chromosome ids in first-seen order -- assigned when a read is COUNTED for a
		// chromosome, not when it is merely seen -- and, per statistic, the ids it has counted (iterated in this container's order, :65-73)
		StringIndexer _chromosome_indexer;
		std::unordered_set<size_t> _presented_chromosomes[Stats::CHROMOSOME_STAT_SIZE];
		bool _chr_overflow = false;         // more than 256 chromosome names: the per-chromosome tables are dropped (1-byte side array)
		bool _batch_gaps = false;
		// add_records: per chromosome-list entry, the id Stats assigned (or -1) and which statistics already list it
		std::vector<int32_t> _bulk_chr_id;
		std::vector<uint8_t> _bulk_chr_presented;
		std::vector<std::string> _bulk_chr_names;
		StringIndexer _n_umis, _n_cbs;      // UMIs / barcodes containing N, passed to the device as indices (DGE_FLAG_UMI_N / DGE_FLAG_CB_N)
		bool _n_dirty = false, _allow_n = false, _allow_n_cb = false;
		uint64_t _skipped_n_reads = 0, _skipped_length_reads = 0;
		void upload_n_strings();
		uint64_t _batch_first = 0;          // stream position of the first pending record
		size_t _batch_capacity;
		uint64_t _n_records = 0;
		std::vector<ReadInfo> _deferred; // reads buffered until the first flush fixes cb/umi lengths and the gene-id space

		StringIndexer _gene_indexer; // built on the host at ingest: ids are first-seen ranks like the reference's
		mutable StringIndexer _umi_indexer;

		// lazily materialised query state
		mutable bool _cells_loaded = false, _genes_loaded = false;
		mutable std::vector<Cell> _cells;
		mutable std::unordered_map<std::string, size_t> _cell_ids_by_cb;
		mutable ids_t _filtered_cells, _merge_targets;
		mutable std::vector<uint64_t> _cell_codes; // barcode codes of the cells (dge_cell_info.barcode), cell-id order
		mutable dge_summary _summary;
		// per-base sums of the UMI quality characters of every (barcode, gene, UMI) as add_record saw it (UMI::add_read, UMI.cpp:21-34); host side
		struct QualityTable;
		std::unique_ptr<QualityTable> _qualities;
		void load_qualities() const;
		struct PairHash { size_t operator()(const std::pair<size_t, uint64_t> &p) const { return std::hash<uint64_t>()(p.second * 0x9E3779B97F4A7C15ull ^ p.first); } };
		mutable std::unordered_map<std::pair<size_t, uint64_t>, uint32_t, PairHash> _created; // (cell id, gene << 32 | UMI) created by the UMI merge -> source UMI
		mutable std::vector<uint32_t> _loaded_cell, _loaded_umi;                            // the rows of the last dge_get_umigs (load_genes)
		mutable std::vector<int32_t> _loaded_gene;

		void ensure_handle();
		void flush();
		void check(int rc) const;
		void load_cells() const;
		void load_genes() const;

	public:
		// `n_genes_hint`: capacity of the gene-id space given to the device (genes are still indexed first-seen on the host)
		CellsDataContainer(const std::shared_ptr<Merge::MergeStrategyAbstract> &merge_strategy,
		                   const std::shared_ptr<Merge::UMIs::MergeUMIsStrategyAbstract> &umi_merge_strategy,
		                   const std::vector<UMI::Mark> &gene_match_levels, bool save_umi_merge_targets = false, int max_cells_num = -1,
		                   int device = 0, size_t n_genes_hint = 1u << 17, bool reads_output = false, bool save_umi_qualities = false);
		~CellsDataContainer();
		CellsDataContainer(const CellsDataContainer &) = delete;
		CellsDataContainer &operator=(const CellsDataContainer &) = delete;

		void add_record(const ReadInfo &read_info);   // throws std::runtime_error("Container is already initialized") after set_initialized
		// n x add_record in order, for reads given as views: the common read (nominal lengths, no N, no base-quality bookkeeping) is packed
		// without building a string; every other read goes through add_record itself.  chromosome_names[read.chromosome] is its chromosome.
		void add_records(const PackedRead *reads, size_t n, const std::vector<std::string> &chromosome_names);
		void set_initialized();                        // throws if called twice
		void merge_and_filter();                       // throws std::runtime_error("You must initialize container")

		size_t total_cells_number() const;
		size_t cell_id_by_cb(const std::string &barcode) const; // throws std::out_of_range
		const ids_t &filtered_cells() const;
		const ids_t &merge_targets() const;
		const UMI::Mark::query_t &gene_match_level() const { return _query_marks; }
		s_i_hash_t get_stat_by_real_cells(Stats::CellStatType type) const;
		// CellsDataContainer.cpp:292-307: real cells that counted anything for `stat` (cell-id order), the chromosomes `stat` has seen, and one
		// count per (listed cell, listed chromosome), cell-major
		using counts_t = std::vector<int>;
		void get_stat_by_real_cells(Stats::CellChrStatType stat, names_t &cell_barcodes, names_t &chromosome_names, counts_t &counts) const;
		bool chromosome_stats_available() const { return !_chr_overflow; }
		bool umi_qualities_saved() const { return bool(_qualities); } // built with save_umi_qualities: UMI::mean_quality is meaningful
		s_ul_hash_t umi_distribution() const; // CellsDataContainer.cpp:182-197: occurrences of every UMI string over the (gene, UMI) entries of the filtered cells
		const Cell &cell(size_t index) const;
		size_t intergenic_reads_num() const;
		size_t has_exon_reads_num() const;
		size_t has_intron_reads_num() const;
		size_t has_not_annotated_reads_num() const;
		size_t real_cells_number() const;
		std::string merge_type() const { return _merge_strategy->merge_type(); }
		const StringIndexer &gene_indexer() const { return _gene_indexer; }
		const StringIndexer &umi_indexer() const; // UMI strings of the held UMIs (ids in (cell, gene, umi) order, not first-seen)

		// direct access for high-volume consumers (ResultsPrinter): matrices straight from the device, no per-cell maps
		dge_handle *handle() const { return _h; }
		bool reads_output() const { return _reads_output; }
		unsigned cb_length() const { return _cb_len; }
		uint64_t skipped_length_reads() const { return _skipped_length_reads; } // reads of another barcode / UMI length than the first read's
		uint64_t skipped_n_reads() const { return _skipped_n_reads; } // reads dropped because the key had no room for the N flag (see ensure_handle)
		std::string barcode_string(uint64_t packed) const; // 2-bit unpacked, or the N-string behind DGE_CB_N_BIT | index
		std::string umi_string(uint32_t packed) const;
	};

	// ResultsPrinter (ResultsPrinter.cpp:23-91,334-453): count matrices -> MatrixMarket + cells/genes tsv and an R-readable .rds
	class ResultsPrinter
	{
		bool write_matrix, reads_output, validation_stats, umi_correction_info;

	public:
		// ResultsPrinter.cpp:16-21.  umi_correction_info (dropest passes !umi_merge, dropest.cpp:311) adds `reads_per_umi_per_cell` to the .rds and
		// needs a container built with save_umi_qualities; validation_stats (-S, MergeProbabilityValidator) is not mirrored: save_results throws.
		ResultsPrinter(bool write_matrix, bool reads_output, bool validation_stats = false, bool umi_correction_info = false)
			: write_matrix(write_matrix), reads_output(reads_output), validation_stats(validation_stats), umi_correction_info(umi_correction_info) {}
		// get_reads_per_umi_per_cell (ResultsPrinter.cpp:261-314): filtered cells, requested UMIs.  Entry k belongs to cells[cell_indexes[k]] and
		// genes[gene_indexes[k]] and lists, per UMI, its read count and UMI::mean_quality.  Cells, genes and entries come in the reference's
		// order; the UMIs inside an entry are an R named list whose order follows an unordered_map over UMI-indexer ids and is not reproduced.
		struct ReadsPerUmiPerCell
		{
			std::vector<std::string> cells, genes;
			std::vector<unsigned> cell_indexes, gene_indexes;
			struct Entry { std::vector<std::string> umis; std::vector<unsigned> reads; std::vector<std::vector<double>> mean_quality; };
			std::vector<Entry> reads_per_umi;
		};
		ReadsPerUmiPerCell get_reads_per_umi_per_cell(const CellsDataContainer &container) const;
		void save_results(const CellsDataContainer &container, const std::string &filename) const;

		struct SparseMatrix // dgCMatrix layout: column-compressed, rows ascending inside a column
		{
			std::vector<int32_t> i, p;
			std::vector<double> x;
			std::vector<std::string> row_names, col_names;
		};
		SparseMatrix get_count_matrix(const CellsDataContainer &container, bool filtered) const; // :334-396 incl. the first-met row order
		// the filtered matrix for another set of query marks (:334-361), and the three matrices of -V written to "<base>.matrices.rds" (:455-474)
		SparseMatrix get_count_matrix_filtered(const CellsDataContainer &container, const UMI::Mark::query_t &query_marks) const;
		void save_intron_exon_matrices(const CellsDataContainer &container, const std::string &filename) const;
		static void save_mtx(const SparseMatrix &m, const std::string &filename_base);            // :81-91
		void save_rds(const CellsDataContainer &container, const SparseMatrix &cm, const SparseMatrix &cm_raw, const std::string &filename_base) const;

	private:
		SparseMatrix assemble(const CellsDataContainer &container, bool filtered, const std::vector<int64_t> &indptr, const std::vector<int32_t> &genes,
		                      const std::vector<int32_t> &vals) const;
	};
}
