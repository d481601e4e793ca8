// Estimation.cpp -- implementation of the host-side mirror declared in Estimation.h (see there for the reference citations).
#include "Estimation.h"

#include <limits>

#include <algorithm>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <sstream>

#include <zlib.h>

namespace Tools
{
	// Read names written by droptag end in "!<cell barcode>#<UMI>" (Tools/ReadParameters.cpp:42-56): the last '#' starts the UMI, the last
	// '!' before it starts the barcode.  Same two failure messages as the reference, since callers print them.
	ReadParameters ReadParameters::parse_encoded_id(const std::string &encoded_id)
	{
		const char *const begin = encoded_id.data();
		const char *hash = nullptr, *bang = nullptr;
		for (const char *c = begin + encoded_id.size(); c != begin && !hash;) if (*--c == '#') hash = c;
		if (!hash) throw std::runtime_error("ERROR: unable to parse out UMI in: " + encoded_id);
		for (const char *c = hash + 1; c != begin && !bang;) if (*--c == '!') bang = c;
		if (!bang) throw std::runtime_error("ERROR: unable to parse out cell barcode in: " + encoded_id);
		return ReadParameters(std::string(bang + 1, hash), std::string(hash + 1, begin + encoded_id.size()), "", "");
	}

	unsigned edit_distance(const char *s1, const char *s2, bool skip_n, unsigned max_ed) { return dge_edit_distance(s1, s2, skip_n ? 1 : 0, max_ed); }

	unsigned hamming_distance(const std::string &s1, const std::string &s2, bool skip_n)
	{
		if (s1.size() != s2.size()) throw std::runtime_error("Strings should have equal length");
		return dge_hamming_distance(s1.c_str(), s2.c_str(), skip_n ? 1 : 0);
	}

	void CollisionsAdjuster::init(const probs_vec_t &umi_probabilities, size_t max_gene_expression)
	{
		_umi_probabilities = umi_probabilities;
		_adjusted_sizes.clear();
		update_adjusted_sizes(max_gene_expression);
	}

	void CollisionsAdjuster::update_adjusted_sizes(size_t max_gene_expression)
	{
		if (max_gene_expression <= _adjusted_sizes.size()) return;
		// the recurrence has no closed form to resume from on the host side: recompute the table up to the new size on the device
		std::vector<uint64_t> table(max_gene_expression);
		if (dge_collisions_adjusted_sizes(_device, _umi_probabilities.data(), _umi_probabilities.size(), max_gene_expression, table.data(), nullptr) != DGE_OK)
			throw std::runtime_error(std::string("dropest_b200: ") + dge_last_error(nullptr));
		_adjusted_sizes.assign(table.begin(), table.end());
	}

	size_t CollisionsAdjuster::estimate_adjusted_gene_expression(size_t expression)
	{
		if (expression > _adjusted_sizes.size()) update_adjusted_sizes(expression);
		return _adjusted_sizes.at(expression - 1);
	}
}

namespace Estimation
{
	const std::string UMI::Mark::DEFAULT_CODE = "eEBA";

	UMI::Mark UMI::Mark::get_by_code(char code)
	{
		Mark mark;
		switch (code)
		{
			case 'e': mark.add(HAS_EXONS); return mark;
			case 'i': mark.add(HAS_INTRONS); return mark;
			case 'E': mark.add(HAS_EXONS); mark.add(HAS_NOT_ANNOTATED); return mark;
			case 'I': mark.add(HAS_INTRONS); mark.add(HAS_NOT_ANNOTATED); return mark;
			case 'B': mark.add(HAS_EXONS); mark.add(HAS_INTRONS); return mark;
			case 'A': mark.add(HAS_EXONS); mark.add(HAS_INTRONS); mark.add(HAS_NOT_ANNOTATED); return mark;
			default: throw std::runtime_error(std::string("Unexpected gene match levels: ") + code);
		}
	}

	std::vector<UMI::Mark> UMI::Mark::get_by_code(const std::string &code)
	{
		std::vector<Mark> levels;
		for (char c : code) levels.push_back(get_by_code(c));
		return levels;
	}

	bool Gene::has(const std::string &umi) const
	{
		try { return _umis.find(_umi_indexer->get_index(umi)) != _umis.end(); }
		catch (std::out_of_range &) { return false; }
	}

	size_t Gene::number_of_requested_umis(const UMI::Mark::query_t &query, bool return_reads) const
	{
		size_t n = 0;
		for (auto const &u : _umis)
			if (u.second.mark().match(query)) n += return_reads ? u.second.read_count() : 1;
		return n;
	}

	size_t Gene::number_of_umis(bool return_reads) const
	{
		if (!return_reads) return _umis.size();
		size_t n = 0;
		for (auto const &u : _umis) n += u.second.read_count();
		return n;
	}

	Gene::s_ul_hash_t Gene::requested_reads_per_umi(const UMI::Mark::query_t &query) const
	{
		s_ul_hash_t res;
		for (auto const &u : _umis)
			if (u.second.mark().match(query)) res.emplace(_umi_indexer->get_value(u.first), u.second.read_count());
		return res;
	}

	Cell::ss_ul_hash_t Cell::requested_reads_per_umi_per_gene(const UMI::Mark::query_t &query_marks) const
	{
		ss_ul_hash_t res;
		for (auto const &g : _genes)
		{
			Gene::s_ul_hash_t r(g.second.requested_reads_per_umi(query_marks));
			if (!r.empty()) res.emplace(_gene_indexer->get_value(g.first), r);
		}
		return res;
	}

	Cell::s_ul_hash_t Cell::requested_umis_per_gene(const UMI::Mark::query_t &query_marks, bool return_reads) const
	{
		s_ul_hash_t res;
		for (auto const &g : _genes)
		{
			size_t n = g.second.number_of_requested_umis(query_marks, return_reads);
			if (n) res.emplace(_gene_indexer->get_value(g.first), n);
		}
		return res;
	}

	namespace Merge
	{
		void RealBarcodesMergeStrategy::configure(dge_config &cfg) const
		{
			cfg.merge_type = DGE_MERGE_REAL;
			cfg.barcodes_type = _parser->indrop ? DGE_BARCODES_INDROP : DGE_BARCODES_CONST;
			cfg.barcodes_file = _parser->filename.c_str();
			cfg.max_cb_merge_edit_distance = _max_merge_edit_distance;
			cfg.min_merge_fraction = _min_merge_fraction;
		}

		namespace
		{
			// Just enough XML for dropEst's configuration files: nested elements with text, comments, an optional declaration; attributes are
			// skipped.  Yields "config.Estimation.Merge.barcodes_file" -> text (first occurrence wins, like ptree::get).
			std::map<std::string, std::string> xml_leaves(const std::string &text, const std::string &fname)
			{
				std::map<std::string, std::string> out;
				std::vector<std::string> path;
				std::string value;
				size_t i = 0;
				auto fail = [&](const std::string &what) { throw std::runtime_error(fname + ": " + what); };
				while (i < text.size())
				{
					if (text[i] != '<') { value += text[i++]; continue; }
					if (text.compare(i, 4, "<!--") == 0)
					{
						const size_t e = text.find("-->", i + 4);
						if (e == std::string::npos) fail("unterminated comment");
						i = e + 3;
						continue;
					}
					if (text.compare(i, 2, "<?") == 0)
					{
						const size_t e = text.find("?>", i + 2);
						if (e == std::string::npos) fail("unterminated declaration");
						i = e + 2;
						continue;
					}
					const size_t e = text.find('>', i);
					if (e == std::string::npos) fail("unterminated tag");
					std::string tag = text.substr(i + 1, e - i - 1);
					i = e + 1;
					if (!tag.empty() && tag[0] == '/')
					{
						if (path.empty() || path.back() != tag.substr(1)) fail("mismatched closing tag </" + tag.substr(1) + ">");
						std::string key;
						for (auto const &p : path) key += (key.empty() ? "" : ".") + p;
						const size_t a = value.find_first_not_of(" \t\r\n"), b = value.find_last_not_of(" \t\r\n");
						out.emplace(key, a == std::string::npos ? std::string() : value.substr(a, b - a + 1));
						path.pop_back();
						value.clear();
						continue;
					}
					const bool self_closing = !tag.empty() && tag.back() == '/';
					if (self_closing) tag.pop_back();
					const size_t sp = tag.find_first_of(" \t\r\n");
					if (sp != std::string::npos) tag = tag.substr(0, sp);
					if (tag.empty()) fail("empty tag");
					value.clear();
					if (!self_closing) path.push_back(tag);
				}
				if (!path.empty()) fail("unclosed element <" + path.back() + ">");
				return out;
			}
		}

		MergeStrategyFactory MergeStrategyFactory::from_xml(const std::string &config_file_name, int min_genes_after_merge_arg)
		{
			std::ifstream f(config_file_name);
			if (!f) throw std::runtime_error("Can't open config file: '" + config_file_name + "'");
			std::stringstream ss;
			ss << f.rdbuf();
			const auto leaves = xml_leaves(ss.str(), config_file_name);
			auto get = [&](const std::string &block, const std::string &key) -> const std::string * {
				auto it = leaves.find("config.Estimation." + block + "." + key);
				return it == leaves.end() ? nullptr : &it->second;
			};
			auto number = [&](const std::string &key, const std::string &text) {
				size_t used = 0;
				double v = 0;
				try { v = std::stod(text, &used); } catch (std::exception &) { used = 0; }
				if (used != text.size() || text.empty()) throw std::runtime_error("conversion of data to type failed for '" + key + "': '" + text + "'");
				return v;
			};
			MergeStrategyFactory fac;
			if (auto v = get("Merge", "merge_type")) fac.merge_type = *v;
			if (auto v = get("Merge", "min_genes_before_merge")) fac.min_genes_before_merge = size_t(number("min_genes_before_merge", *v));
			if (min_genes_after_merge_arg > 0) fac.min_genes_after_merge = unsigned(min_genes_after_merge_arg);
			else if (auto v = get("Merge", "min_genes_after_merge")) fac.min_genes_after_merge = size_t(number("min_genes_after_merge", *v));
			auto ed = get("Merge", "max_cb_merge_edit_distance"); // no default in the reference: ptree::get throws
			if (!ed) throw std::runtime_error("No such node (max_cb_merge_edit_distance)");
			fac.max_merge_edit_distance = unsigned(number("max_cb_merge_edit_distance", *ed));
			if (auto v = get("Merge", "min_merge_fraction")) fac.min_merge_fraction = number("min_merge_fraction", *v);
			if (auto v = get("Merge", "barcodes_type")) fac.barcodes_type = *v;
			if (auto v = get("Merge", "barcodes_file")) fac.barcodes_filename = *v;
			// Tools::ltrim, expand_tilde_in_path, expand_relative_path (UtilFunctions.cpp:117-149)
			std::string &bf = fac.barcodes_filename;
			if (!bf.empty())
			{
				const size_t a = bf.find_first_not_of(" \t");
				bf = a == std::string::npos ? std::string() : bf.substr(a);
			}
			if (bf.size() >= 2 && bf.compare(0, 2, "~/") == 0 && std::getenv("HOME")) bf = std::getenv("HOME") + bf.substr(1);
			if (!bf.empty() && bf[0] != '/')
			{
				const size_t slash = config_file_name.find_last_of('/');
				if (slash != std::string::npos) bf = config_file_name.substr(0, slash) + "/" + bf;
			}
			if (!bf.empty() && !std::ifstream(bf)) throw std::runtime_error("Can't open file with barcodes: '" + bf + "'");
			if (auto v = get("PreciseMerge", "max_merge_prob")) fac.max_merge_prob = number("max_merge_prob", *v);
			if (auto v = get("PreciseMerge", "max_real_merge_prob")) fac.max_real_cb_merge_prob = number("max_real_merge_prob", *v);
			if (auto v = get("Merge", "max_umi_merge_edit_distance")) fac.max_umi_merge_edit_distance = unsigned(number("max_umi_merge_edit_distance", *v));
			if (auto v = get("Merge", "umi_merge_multiplier")) fac.umi_merge_mult = number("umi_merge_multiplier", *v);
			return fac;
		}

		std::shared_ptr<BarcodesParsing::BarcodesParser> MergeStrategyFactory::get_barcodes_parser() const
		{
			if (barcodes_type == "indrop") return std::make_shared<BarcodesParsing::InDropBarcodesParser>(barcodes_filename);
			if (barcodes_type == "const") return std::make_shared<BarcodesParsing::ConstLengthBarcodesParser>(barcodes_filename);
			throw std::runtime_error("Unexpected barcodes type: " + barcodes_type);
		}

		// MergeStrategyFactory::get_cb_strat / get_cb_poisson_strat (MergeStrategyFactory.cpp:61-103)
		std::shared_ptr<MergeStrategyAbstract> MergeStrategyFactory::get_cb_strat(bool merge_tags, bool use_poisson) const
		{
			if (!merge_tags) return std::make_shared<DummyMergeStrategy>(min_genes_before_merge, min_genes_after_merge);
			if (use_poisson)
			{
				PoissonTargetEstimator target_estimator(max_merge_prob, max_real_cb_merge_prob);
				if (barcodes_filename.empty())
					return std::make_shared<PoissonSimpleMergeStrategy>(target_estimator, unsigned(min_genes_before_merge), unsigned(min_genes_after_merge),
					                                                    max_merge_edit_distance);
				return std::make_shared<PoissonRealBarcodesMergeStrategy>(target_estimator, get_barcodes_parser(), min_genes_before_merge,
				                                                          min_genes_after_merge, max_merge_edit_distance);
			}
			if (merge_type == "all") return std::make_shared<MergeAllMergeStrategy>(min_genes_before_merge, min_genes_after_merge, max_merge_edit_distance);
			if (barcodes_filename.empty())
				return std::make_shared<SimpleMergeStrategy>(min_genes_before_merge, min_genes_after_merge, max_merge_edit_distance, min_merge_fraction);
			return std::make_shared<RealBarcodesMergeStrategy>(get_barcodes_parser(), min_genes_before_merge, min_genes_after_merge, max_merge_edit_distance,
			                                                   min_merge_fraction);
		}

		std::shared_ptr<UMIs::MergeUMIsStrategyAbstract> MergeStrategyFactory::get_umi(bool advanced) const
		{
			if (advanced) return std::make_shared<UMIs::MergeUMIsStrategyDirectional>(umi_merge_mult, max_umi_merge_edit_distance);
			return std::make_shared<UMIs::MergeUMIsStrategySimple>(max_umi_merge_edit_distance);
		}

		// MergeUMIsStrategyDirectional::find_targets / find_target (MergeUMIsStrategyDirectional.cpp:57-116) for N-free UMIs:
		// sort by reads, scan candidates from the largest, keep the first one at the smallest distance, follow chains to the root.
		UMIs::MergeUMIsStrategyDirectional::merge_targets_t UMIs::MergeUMIsStrategyDirectional::find_targets(umi_vec_t &umis) const
		{
			std::sort(umis.begin(), umis.end(), [](const UmiWrap &a, const UmiWrap &b) { return a.n_reads < b.n_reads; });
			const size_t n = umis.size();
			std::vector<long> tgt(n, -1);
			for (size_t s = 0; s < n; ++s)
			{
				if (umis[s].sequence.find('N') != std::string::npos)
					throw std::runtime_error("UMIs containing N are not supported by the packed record yet");
				unsigned best_ed = std::numeric_limits<unsigned>::max();
				for (long d = long(n) - 1; d > long(s); --d)
				{
					if (umis[s].n_reads * _mult > umis[size_t(d)].n_reads) break;
					const unsigned ed = Tools::edit_distance(umis[s].sequence.c_str(), umis[size_t(d)].sequence.c_str(), true, _max_edit_distance);
					if (ed > _max_edit_distance || ed >= best_ed) continue;
					tgt[s] = d;
					if (ed <= 1) break;
					best_ed = ed;
				}
			}
			merge_targets_t res;
			for (size_t s = 0; s < n; ++s)
			{
				if (tgt[s] < 0) continue;
				long r = tgt[s];
				while (tgt[size_t(r)] >= 0) r = tgt[size_t(r)];
				res[umis[s].sequence] = umis[size_t(r)].sequence;
			}
			return res;
		}
	}

	// ---- CellsDataContainer ---------------------------------------------------------------------------------------------
	static bool pack2bit(const std::string &s, uint64_t &out)
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

	static std::string unpack2bit(uint64_t v, unsigned len)
	{
		std::string s(len, 'A');
		for (unsigned i = 0; i < len; ++i) s[len - 1 - i] = "ACGT"[(v >> (2 * i)) & 3];
		return s;
	}

	// (barcode code, gene id, UMI code) -> per-base sums of the quality characters of the reads add_record saw for it.  Open addressing, grows by
	// doubling; the sums live in one flat array (`len` entries from `off`).
	struct CellsDataContainer::QualityTable
	{
		struct Slot { uint64_t cb, gu; uint32_t off; uint16_t len; uint16_t used; };
		std::vector<Slot> slots;
		std::vector<unsigned> sums;
		size_t n = 0;
		QualityTable() : slots(size_t(1) << 16, Slot{0, 0, 0, 0, 0}) {}
		static size_t hash(uint64_t cb, uint64_t gu)
		{
			uint64_t x = cb * 0x9E3779B97F4A7C15ull ^ (gu + 0x7F4A7C15ull) * 0xC2B2AE3D27D4EB4Full;
			x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
			return size_t(x);
		}
		Slot *find(uint64_t cb, uint64_t gu)
		{
			const size_t mask = slots.size() - 1;
			for (size_t i = hash(cb, gu) & mask;; i = (i + 1) & mask)
			{
				Slot &s = slots[i];
				if (!s.used) return nullptr;
				if (s.cb == cb && s.gu == gu) return &s;
			}
		}
		void grow()
		{
			std::vector<Slot> old(slots.size() * 2, Slot{0, 0, 0, 0, 0});
			old.swap(slots);
			const size_t mask = slots.size() - 1;
			for (auto const &s : old)
			{
				if (!s.used) continue;
				size_t i = hash(s.cb, s.gu) & mask;
				while (slots[i].used) i = (i + 1) & mask;
				slots[i] = s;
			}
		}
		// UMI::add_read (UMI.cpp:21-34): the first read of a UMI fixes the length of its quality vector
		void add(uint64_t cb, uint64_t gu, const std::string &quality)
		{
			if ((n + 1) * 10 > slots.size() * 6) grow();
			const size_t mask = slots.size() - 1;
			size_t i = hash(cb, gu) & mask;
			while (slots[i].used && !(slots[i].cb == cb && slots[i].gu == gu)) i = (i + 1) & mask;
			Slot &s = slots[i];
			if (!s.used)
			{
				if (quality.size() > 0xFFFF || sums.size() + quality.size() > 0xFFFFFFFFull) throw std::runtime_error("UMI quality table is full");
				s = Slot{cb, gu, uint32_t(sums.size()), uint16_t(quality.size()), 1};
				sums.resize(sums.size() + quality.size(), 0u);
				++n;
			}
			if (quality.size() != s.len)
				throw std::runtime_error("Wrong quality length: " + std::to_string(quality.size()) + ", expected: " + std::to_string(s.len));
			unsigned *p = sums.data() + s.off;
			for (size_t k = 0; k < quality.size(); ++k) p[k] += unsigned(quality[k]); // a char, sign-extended like the reference's `+=`
		}
	};

	CellsDataContainer::CellsDataContainer(const std::shared_ptr<Merge::MergeStrategyAbstract> &merge_strategy,
	                                       const std::shared_ptr<Merge::UMIs::MergeUMIsStrategyAbstract> &umi_merge_strategy,
	                                       const std::vector<UMI::Mark> &gene_match_levels, bool save_umi_merge_targets, int max_cells_num, int device,
	                                       size_t n_genes_hint, bool reads_output, bool save_umi_qualities)
		: _merge_strategy(merge_strategy), _umi_merge_strategy(umi_merge_strategy), _max_cells_num(max_cells_num)
		, _query_marks(gene_match_levels), _device(device), _reads_output(reads_output), _save_umi_merge_targets(save_umi_merge_targets)
		, _batch_capacity(n_genes_hint)
	{
		// _batch_capacity temporarily carries the gene-space hint until the handle exists
		std::memset(&_summary, 0, sizeof(_summary));
		if (save_umi_qualities)
		{   // where a UMI object ends up after the merges decides whose quality sums it shows: the UMI merge targets are needed too
			_qualities.reset(new QualityTable());
			_save_umi_merge_targets = true;
		}
	}

	CellsDataContainer::~CellsDataContainer() { if (_h) dge_destroy(_h); }

	void CellsDataContainer::check(int rc) const
	{
		if (rc == DGE_OK) return;
		std::string msg = dge_last_error(_h);
		if (rc == DGE_ERR_STATE) throw std::runtime_error(msg); // same messages as the reference's throws
		throw std::runtime_error("dropest_b200: " + msg);
	}

	void CellsDataContainer::ensure_handle()
	{
		if (_h) return;
		dge_config cfg;
		dge_config_default(&cfg);
		cfg.device = _device;
		cfg.cb_len = _cb_len ? _cb_len : 16;
		cfg.umi_len = _umi_len ? _umi_len : 10;
		cfg.n_genes = uint32_t(_batch_capacity);
		cfg.min_genes_before_merge = uint32_t(_merge_strategy->min_genes_before_merge());
		cfg.min_genes_after_merge = uint32_t(_merge_strategy->min_genes_after_merge());
		cfg.max_cells = _max_cells_num;
		cfg.reads_output = _reads_output ? 1 : 0;
		cfg.save_umi_merge_targets = _save_umi_merge_targets ? 1 : 0;
		cfg.query_mark_mask = 0;
		for (auto const &m : _query_marks) cfg.query_mark_mask |= 1u << m.bits();
		_merge_strategy->configure(cfg);
		_umi_merge_strategy->configure(cfg);
		// Barcodes / UMIs containing N travel as indices into _n_cbs / _n_umis (DGE_FLAG_CB_N / DGE_FLAG_UMI_N).  That costs the grouping key
		// one more UMI bit; when it would leave fewer than 2^22 barcode slots (12-base UMIs with a gene space > 2^14) N reads are skipped
		// and counted instead (skipped_n_reads()) -- pass a tighter n_genes_hint to keep them.
		{
			unsigned gb = 1;
			while ((size_t(1) << gb) < size_t(cfg.n_genes)) ++gb;
			const unsigned ub_n = std::max(2 * cfg.umi_len, 20u) + 1;
			_allow_n = 61 >= gb + ub_n + 22;
		}
		// What the library refuses (an explicit error at merge_and_filter, never a silent difference): a REAL cell whose barcode is escaped (N,
		// or another length) with the strategies that compare barcodes on the device.  Rather than losing the whole run, reads with such
		// barcodes are skipped and counted here, with a warning, under the no-whitelist / Poisson strategies.
		_allow_n_cb = _allow_n && (cfg.merge_type == DGE_MERGE_NONE || cfg.merge_type == DGE_MERGE_REAL);
		cfg.allow_n = _allow_n ? 1 : 0;
		int rc = dge_create(&cfg, &_h);
		if (rc != DGE_OK) throw std::runtime_error(std::string("dropest_b200: ") + dge_last_error(nullptr));
		_batch_capacity = size_t(1) << 20;
		_batch_keys.reserve(_batch_capacity); _batch_genes.reserve(_batch_capacity);
	}

	void CellsDataContainer::upload_n_strings()
	{
		if (!_n_dirty || !_h) return;
		std::string blob;
		for (auto const &v : _n_umis.values()) blob += v;
		check(dge_set_n_strings(_h, 0, blob.c_str(), _n_umis.values().size()));
		blob.clear();
		std::vector<uint32_t> lengths;
		for (auto const &v : _n_cbs.values()) { blob += v; lengths.push_back(uint32_t(v.size())); }
		check(dge_set_cb_strings(_h, blob.c_str(), lengths.data(), lengths.size())); // any length: barcodes with N and variable-length barcodes
		_n_dirty = false;
	}

	std::string CellsDataContainer::barcode_string(uint64_t packed) const
	{
		return (packed & DGE_CB_N_BIT) ? _n_cbs.get_value(size_t(packed & (DGE_CB_N_BIT - 1))) : unpack2bit(packed, _cb_len);
	}

	std::string CellsDataContainer::umi_string(uint32_t packed) const
	{
		return (packed & DGE_UMI_N_BIT) ? _n_umis.get_value(size_t(packed & ~DGE_UMI_N_BIT)) : unpack2bit(packed, _umi_len);
	}

	void CellsDataContainer::flush()
	{
		if (_batch_keys.empty()) return;
		if (!_batch_gaps)
		{   // add_record is called in stream order, so the read index is implicit: 12 bytes per read go to the device
			if (_chr_overflow) check(dge_add_batch_soa(_h, _batch_keys.data(), _batch_genes.data(), _batch_keys.size(), _batch_first));
			else check(dge_add_batch_soa_chr(_h, _batch_keys.data(), _batch_genes.data(), _batch_chr.data(), _batch_keys.size(), _batch_first));
		}
		else
		{   // skipped reads left gaps in the numbering: 16-byte records with explicit stream positions
			std::vector<dge_record16> recs(_batch_keys.size());
			for (size_t i = 0; i < recs.size(); ++i) recs[i] = dge_record16{_batch_keys[i], _batch_genes[i], _batch_idx[i]};
			if (_chr_overflow) check(dge_add_batch(_h, recs.data(), recs.size()));
			else check(dge_add_batch_chr(_h, recs.data(), _batch_chr.data(), recs.size()));
		}
		_batch_first = _n_records;
		_batch_gaps = false;
		_batch_keys.clear(); _batch_genes.clear(); _batch_idx.clear(); _batch_chr.clear();
	}

	void CellsDataContainer::add_record(const ReadInfo &read_info)
	{
		if (_is_initialized) throw std::runtime_error("Container is already initialized");
		const std::string &cb = read_info.params.cell_barcode(), &umi = read_info.params.umi();
		if (!_h)
		{
			_cb_len = unsigned(cb.size()); _umi_len = unsigned(umi.size());
			if (_cb_len > 20 || _umi_len > 12) throw std::runtime_error("barcode/UMI too long for the packed record (20/12 bp)");
			ensure_handle();
		}
		// The packed record has one barcode / UMI length per run (the first read's).  A barcode of another length (variable-length inDrop v1 / v2
		// barcodes) travels like a barcode with N: through the escaped-barcode list, as a cell of its own (dge_set_cb_strings).  A UMI of another
		// length -- or such a barcode under a strategy that cannot take escaped barcodes -- is counted and skipped, not fatal; the skipped read
		// keeps its stream position like a skipped N read.
		const bool odd_cb = cb.size() != _cb_len;
		if (umi.size() != _umi_len || (odd_cb && (!_allow_n_cb || cb.empty() || cb.size() > 64)))
		{
			if (_skipped_length_reads++ == 0)
				std::cerr << "dropest_b200: reads whose UMI length differs from the first read's (" << _umi_len << "), or whose barcode length does (" << _cb_len
				          << ") under a strategy that cannot take such barcodes, are skipped; skipped_length_reads() reports how many\n";
			++_n_records;
			_batch_gaps = true;
			return;
		}
		// N-free sequences are 2-bit packed; a sequence with N is passed as its index in the container's N-string list (the reference keeps
		// such reads: the barcode is a cell of its own, the UMI is repaired by MergeUMIsStrategySimple after the barcode merge)
		uint64_t cbv, umiv;
		uint32_t flags = 0;
		if ((!_allow_n && umi.find('N') != std::string::npos) || (!_allow_n_cb && cb.find('N') != std::string::npos))
		{
			if (_skipped_n_reads++ == 0)
				std::cerr << "dropest_b200: reads whose barcode contains N (no-whitelist / Poisson barcode merge) or whose barcode / UMI contains N (no room "
				             "for the N flag in the grouping key) are skipped; skipped_n_reads() reports how many\n";
			++_n_records;   // the skipped read keeps its position in the stream: the pending batch now has a gap (flush sends explicit read indices)
			_batch_gaps = true;
			return;
		}
		if (odd_cb || !pack2bit(cb, cbv))
		{
			if (cb.find_first_not_of("ACGTN") != std::string::npos) throw std::runtime_error("unexpected character in the cell barcode: " + cb);
			cbv = _n_cbs.add(cb); flags |= DGE_FLAG_CB_N; _n_dirty = true;
		}
		if (!pack2bit(umi, umiv))
		{
			if (umi.find_first_not_of("ACGTN") != std::string::npos) throw std::runtime_error("unexpected character in the UMI: " + umi);
			umiv = _n_umis.add(umi); flags |= DGE_FLAG_UMI_N; _n_dirty = true;
			if (umiv >> (std::max(2 * _umi_len, 20u))) throw std::runtime_error("too many distinct UMIs containing N");
		}
		uint32_t gene = DGE_NO_GENE;
		if (!read_info.gene.empty()) gene = uint32_t(_gene_indexer.add(read_info.gene));
		// Stats::inc(CellChrStatType, chromosome) of add_record / update_cell_stats (CellsDataContainer.cpp:73-78, 313-322): the id is
		// assigned and the statistic's chromosome set grows only when this read is counted
		uint8_t chr_id = 0;
		if (!_chr_overflow)
		{
			const unsigned m = read_info.umi_mark.bits();
			const bool inter = read_info.gene.empty();
			if (inter || (m & 6u))
			{
				const size_t id = _chromosome_indexer.add(read_info.chromosome_name);
				if (id > 255)
				{
					_chr_overflow = true;
					std::cerr << "dropest_b200: more than 256 chromosome names; reads_per_chr_per_cells is not produced\n";
				}
				else
				{
					chr_id = uint8_t(id);
					if (inter) _presented_chromosomes[Stats::INTERGENIC_READS_PER_CHR_PER_CELL].insert(id);
					else
					{
						if (m & 2u) _presented_chromosomes[Stats::EXON_READS_PER_CHR_PER_CELL].insert(id);
						if (m & 4u) _presented_chromosomes[Stats::INTRON_READS_PER_CHR_PER_CELL].insert(id);
					}
				}
			}
		}
		if (_qualities && gene != DGE_NO_GENE)
			_qualities->add((flags & DGE_FLAG_CB_N) ? (DGE_CB_N_BIT | cbv) : cbv, (uint64_t(gene) << 32) | ((flags & DGE_FLAG_UMI_N) ? (DGE_UMI_N_BIT | uint32_t(umiv)) : uint32_t(umiv)),
			                read_info.params.umi_quality());
		_batch_chr.push_back(chr_id);
		_batch_keys.push_back((cbv << 24) | umiv);
		_batch_genes.push_back(gene | (uint32_t(read_info.umi_mark.bits()) << 24) | flags);
		_batch_idx.push_back(uint32_t(_n_records));
		++_n_records;
		if (_batch_keys.size() >= _batch_capacity) flush();
	}

	void CellsDataContainer::add_records(const PackedRead *reads, size_t n, const std::vector<std::string> &chromosome_names)
	{
		if (_is_initialized) throw std::runtime_error("Container is already initialized");
		if (_bulk_chr_names != chromosome_names)
		{   // another file: its chromosome list maps afresh (ids come from the names, so this only resets the cache)
			_bulk_chr_names = chromosome_names;
			_bulk_chr_id.assign(chromosome_names.size(), -1);
			_bulk_chr_presented.assign(chromosome_names.size(), 0);
		}
		for (size_t k = 0; k < n; ++k)
		{
			const PackedRead &r = reads[k];
			if (k + 8 < n)
			{   // the gene name lies in a parsing thread's arena, its slot in the indexer's table: both are fetched a few reads ahead
				const PackedRead &ahead = reads[k + 8];
				__builtin_prefetch(ahead.gene);
				if (ahead.gene_len) _gene_indexer.prefetch(ahead.gene_hash);
			}
			const bool common = _h && !_qualities && r.packable == 3 && r.cb_len == _cb_len && r.umi_len == _umi_len && r.chromosome >= 0 &&
			                    size_t(r.chromosome) < chromosome_names.size();
			if (!common)
			{   // the first read (it fixes the lengths and creates the handle), N, other lengths, quality bookkeeping: the one-read path
				add_record(ReadInfo(Tools::ReadParameters(std::string(r.cb, r.cb_len), std::string(r.umi, r.umi_len), std::string(r.cb_quality, r.cb_quality_len),
				                                          std::string(r.umi_quality, r.umi_quality_len)),
				                    std::string(r.gene, r.gene_len), chromosome_names.at(size_t(r.chromosome)), UMI::Mark(UMI::Mark::MarkType(r.mark_bits))));
				continue;
			}
			// ---- exactly what add_record does for such a read
			uint32_t gene = DGE_NO_GENE;
			if (r.gene_len) gene = uint32_t(_gene_indexer.add(r.gene, r.gene_len, r.gene_hash));
			uint8_t chr_id = 0;
			if (!_chr_overflow)
			{
				const unsigned m = r.mark_bits;
				const bool inter = r.gene_len == 0;
				if (inter || (m & 6u))
				{
					int32_t &cached = _bulk_chr_id[size_t(r.chromosome)];
					if (cached < 0) cached = int32_t(_chromosome_indexer.add(chromosome_names[size_t(r.chromosome)]));
					const size_t id = size_t(cached);
					if (id > 255)
					{
						_chr_overflow = true;
						std::cerr << "dropest_b200: more than 256 chromosome names; reads_per_chr_per_cells is not produced\n";
					}
					else
					{
						chr_id = uint8_t(id);
						const uint8_t want = inter ? 1u : uint8_t(((m & 2u) ? 2u : 0u) | ((m & 4u) ? 4u : 0u));
						uint8_t &have = _bulk_chr_presented[size_t(r.chromosome)];
						if ((have & want) != want)
						{
							if (want & 1u) _presented_chromosomes[Stats::INTERGENIC_READS_PER_CHR_PER_CELL].insert(id);
							if (want & 2u) _presented_chromosomes[Stats::EXON_READS_PER_CHR_PER_CELL].insert(id);
							if (want & 4u) _presented_chromosomes[Stats::INTRON_READS_PER_CHR_PER_CELL].insert(id);
							have |= want;
						}
					}
				}
			}
			_batch_chr.push_back(chr_id);
			_batch_keys.push_back((r.cb_packed << 24) | r.umi_packed);
			_batch_genes.push_back(gene | (uint32_t(r.mark_bits) << 24));
			_batch_idx.push_back(uint32_t(_n_records));
			++_n_records;
			if (_batch_keys.size() >= _batch_capacity) flush();
		}
	}

	void CellsDataContainer::set_initialized()
	{
		if (_is_initialized) throw std::runtime_error("Container is already initialized");
		ensure_handle();
		flush();
		upload_n_strings();
		check(dge_set_initialized(_h));
		_is_initialized = true;
		_cells_loaded = _genes_loaded = false;
	}

	void CellsDataContainer::merge_and_filter()
	{
		if (!_is_initialized) throw std::runtime_error("You must initialize container");
		check(dge_merge_and_filter(_h));
		_is_merged = true;
		_cells_loaded = _genes_loaded = false;
	}

	void CellsDataContainer::load_cells() const
	{
		if (_cells_loaded) return;
		if (!_is_initialized) throw std::runtime_error("You must initialize container");
		size_t n = 0;
		check(dge_get_cells(_h, DGE_CELLS_ALL, nullptr, 0, &n));
		std::vector<dge_cell_info> info(n);
		if (n) check(dge_get_cells(_h, DGE_CELLS_ALL, info.data(), n, &n));
		_cells.assign(n, Cell());
		_cell_ids_by_cb.clear();
		_merge_targets.resize(n);
		_cell_codes.resize(n);
		for (size_t i = 0; i < n; ++i)
		{
			Cell &c = _cells[i];
			c._barcode = barcode_string(info[i].barcode);
			_cell_codes[i] = info[i].barcode;
			c._is_real = info[i].flags & DGE_CELL_REAL; c._is_merged = info[i].flags & DGE_CELL_MERGED; c._is_excluded = info[i].flags & DGE_CELL_EXCLUDED;
			c._n_genes = size_t(info[i].n_genes);
			c._requested_genes_num = size_t(info[i].requested_genes_num); c._requested_umis_num = size_t(info[i].requested_umis_num);
			c._stats = Stats(info[i].reads_stat, info[i].umis_stat);
			c._gene_indexer = &_gene_indexer;
			_cell_ids_by_cb.emplace(c._barcode, i);
			_merge_targets[i] = size_t(info[i].merge_target);
		}
		size_t nf = 0;
		check(dge_get_cells(_h, DGE_CELLS_FILTERED, nullptr, 0, &nf));
		std::vector<dge_cell_info> finfo(nf);
		if (nf) check(dge_get_cells(_h, DGE_CELLS_FILTERED, finfo.data(), nf, &nf));
		_filtered_cells.resize(nf);
		for (size_t k = 0; k < nf; ++k) _filtered_cells[k] = _cell_ids_by_cb.at(barcode_string(finfo[k].barcode));
		check(dge_get_summary(_h, &_summary));
		_cells_loaded = true;
		_genes_loaded = false;
	}

	void CellsDataContainer::load_genes() const
	{
		load_cells();
		if (_genes_loaded) return;
		size_t n = 0;
		check(dge_get_umigs(_h, DGE_CELLS_ALL, nullptr, nullptr, nullptr, nullptr, nullptr, 0, &n));
		std::vector<uint32_t> cell(n), umi(n), reads(n);
		std::vector<int32_t> gene(n);
		std::vector<uint8_t> mark(n);
		if (n) check(dge_get_umigs(_h, DGE_CELLS_ALL, cell.data(), gene.data(), umi.data(), reads.data(), mark.data(), n, &n));
		for (auto &c : _cells) c._genes.clear();
		for (size_t k = 0; k < n; ++k)
		{
			Cell &c = _cells.at(cell[k]);
			auto git = c._genes.emplace(size_t(gene[k]), Gene(&_umi_indexer)).first;
			UMI::Mark m;
			if (mark[k] & 1) m.add(UMI::Mark::HAS_NOT_ANNOTATED);
			if (mark[k] & 2) m.add(UMI::Mark::HAS_EXONS);
			if (mark[k] & 4) m.add(UMI::Mark::HAS_INTRONS);
			git->second._umis.emplace(_umi_indexer.add(umi_string(umi[k])), UMI(reads[k], m));
		}
		if (_save_umi_merge_targets && _is_merged)
		{   // Gene::_merge_targets: the source UMIs are gone from the gene, their strings still belong to the UMI indexer (Gene.cpp:38-58)
			size_t nt = 0;
			check(dge_get_umi_merge_targets(_h, nullptr, nullptr, nullptr, nullptr, nullptr, 0, &nt));
			std::vector<uint64_t> t_cb(nt);
			std::vector<int32_t> t_gene(nt);
			std::vector<uint32_t> t_src(nt), t_dst(nt);
			std::vector<uint8_t> t_created(nt);
			if (nt) check(dge_get_umi_merge_targets(_h, t_cb.data(), t_gene.data(), t_src.data(), t_dst.data(), t_created.data(), nt, &nt));
			for (size_t k = 0; k < nt; ++k)
			{
				Cell &c = _cells.at(_cell_ids_by_cb.at(barcode_string(t_cb[k])));
				auto git = c._genes.emplace(size_t(t_gene[k]), Gene(&_umi_indexer)).first;
				const std::string src(umi_string(t_src[k])), dst(umi_string(t_dst[k]));
				_umi_indexer.add(src);
				_umi_indexer.add(dst);
				git->second._merge_targets[src] = dst;
			}
			if (_qualities)
			{
				_created.clear();
				for (size_t k = 0; k < nt; ++k)
					if (t_created[k]) _created[std::make_pair(_cell_ids_by_cb.at(barcode_string(t_cb[k])), (uint64_t(uint32_t(t_gene[k])) << 32) | t_dst[k])] = t_src[k];
			}
		}
		if (_qualities)
		{
			_loaded_cell = std::move(cell); _loaded_gene = std::move(gene); _loaded_umi = std::move(umi);
			load_qualities();
		}
		_genes_loaded = true;
	}

	// UMI::_sum_quality of every held UMI.  A UMI object keeps the quality sums of the reads that were ADDED to it; merges move or drop objects:
	//   * merge_cells -> Cell::merge -> Gene::merge(const Gene &) (Gene.cpp:26-36): the target keeps its own object, and takes a copy of the
	//     source's only when it has none.  So a merged cell's UMI shows the sums of the first holder along the merge_cells calls, in the order
	//     they were applied (dge_get_merge_events), the target itself first;
	//   * Gene::merge(source_umi, target_umi) (Gene.cpp:38-58) of the UMI merge strategies: the source's object becomes the target when the
	//     target does not exist (`created`), otherwise it is dropped.
	void CellsDataContainer::load_qualities() const
	{
		QualityTable &qt = *_qualities;
		// merge_cells calls into every cell, in time order
		std::unordered_map<size_t, std::vector<std::pair<size_t, size_t>>> incoming; // target cell id -> (time, source cell id)
		if (_is_merged)
		{
			size_t ne = 0;
			check(dge_get_merge_events(_h, nullptr, nullptr, 0, &ne));
			std::vector<uint64_t> from(ne), to(ne);
			if (ne) check(dge_get_merge_events(_h, from.data(), to.data(), ne, &ne));
			for (size_t t = 0; t < ne; ++t)
				incoming[_cell_ids_by_cb.at(barcode_string(to[t]))].emplace_back(t, _cell_ids_by_cb.at(barcode_string(from[t])));
		}
		// the object (cell, gene, umi) held just before time `limit` of the barcode merge: its own, else the first one merged in
		std::function<const QualityTable::Slot *(size_t, uint64_t, size_t)> held = [&](size_t cell_id, uint64_t gu, size_t limit) -> const QualityTable::Slot * {
			if (const QualityTable::Slot *own = qt.find(_cell_codes[cell_id], gu)) return own;
			auto in = incoming.find(cell_id);
			if (in == incoming.end()) return nullptr;
			for (auto const &e : in->second)
			{
				if (e.first >= limit) break;
				if (const QualityTable::Slot *s = held(e.second, gu, e.first)) return s;
			}
			return nullptr;
		};
		const size_t forever = ~size_t(0);
		for (size_t k = 0; k < _loaded_cell.size(); ++k)
		{
			const size_t cell_id = _loaded_cell[k];
			uint64_t gu = (uint64_t(uint32_t(_loaded_gene[k])) << 32) | _loaded_umi[k];
			for (int hop = 0; hop < 8; ++hop)
			{   // an object that the UMI merge moved under another name
				auto c = _created.find(std::make_pair(cell_id, gu));
				if (c == _created.end()) break;
				gu = (gu & 0xFFFFFFFF00000000ull) | c->second;
			}
			const QualityTable::Slot *s = held(cell_id, gu, forever);
			if (!s) throw std::runtime_error("internal: no base-quality record for a held UMI");
			UMI &u = _cells[cell_id]._genes.at(size_t(_loaded_gene[k]))._umis.at(_umi_indexer.get_index(umi_string(_loaded_umi[k])));
			u._sum_quality.assign(qt.sums.begin() + s->off, qt.sums.begin() + s->off + s->len);
		}
	}

	size_t CellsDataContainer::total_cells_number() const { load_cells(); return _cells.size(); }
	size_t CellsDataContainer::cell_id_by_cb(const std::string &barcode) const { load_cells(); return _cell_ids_by_cb.at(barcode); }
	const CellsDataContainer::ids_t &CellsDataContainer::filtered_cells() const { load_cells(); return _filtered_cells; }
	const CellsDataContainer::ids_t &CellsDataContainer::merge_targets() const { load_cells(); return _merge_targets; }
	const Cell &CellsDataContainer::cell(size_t index) const { load_genes(); return _cells.at(index); }
	size_t CellsDataContainer::intergenic_reads_num() const { load_cells(); return size_t(_summary.intergenic_reads); }
	size_t CellsDataContainer::has_exon_reads_num() const { load_cells(); return size_t(_summary.has_exon_reads); }
	size_t CellsDataContainer::has_intron_reads_num() const { load_cells(); return size_t(_summary.has_intron_reads); }
	size_t CellsDataContainer::has_not_annotated_reads_num() const { load_cells(); return size_t(_summary.has_not_annotated_reads); }
	size_t CellsDataContainer::real_cells_number() const { load_cells(); return size_t(_summary.real_cells_number); }
	const StringIndexer &CellsDataContainer::umi_indexer() const { load_genes(); return _umi_indexer; }

	CellsDataContainer::s_i_hash_t CellsDataContainer::get_stat_by_real_cells(Stats::CellStatType type) const
	{
		load_cells();
		s_i_hash_t res;
		for (auto const &c : _cells)
			if (c.is_real()) res[c.barcode()] = c.stats().get(type);
		return res;
	}

	CellsDataContainer::s_ul_hash_t CellsDataContainer::umi_distribution() const
	{
		load_genes();
		s_ul_hash_t umi_dist;
		for (size_t cell_id : _filtered_cells)
			for (auto const &gene : _cells[cell_id].genes())
				for (auto const &umi : gene.second.umis()) umi_dist[_umi_indexer.get_value(umi.first)]++;
		return umi_dist;
	}

	void CellsDataContainer::get_stat_by_real_cells(Stats::CellChrStatType stat, names_t &cell_barcodes, names_t &chromosome_names, counts_t &counts) const
	{
		if (_chr_overflow) throw std::runtime_error("per-chromosome statistics were dropped: more than 256 chromosome names");
		load_cells();
		size_t n_cells = 0;
		uint32_t n_chr = 0;
		check(dge_get_chr_stats(_h, nullptr, 0, &n_cells, &n_chr, nullptr));
		std::vector<int32_t> dense(n_cells * n_chr * 3 + 1);
		if (n_chr) check(dge_get_chr_stats(_h, dense.data(), n_cells, &n_cells, &n_chr, nullptr));
		const auto &presented = _presented_chromosomes[stat];
		size_t k = 0; // index among the real cells (DGE_CELLS_REAL order = cell-id order)
		for (auto const &c : _cells)
		{
			if (!c.is_real()) continue;
			const int32_t *row = dense.data() + k * n_chr * 3;
			++k;
			bool any = false;
			for (uint32_t ch = 0; ch < n_chr && !any; ++ch) any = row[ch * 3 + stat] != 0;
			if (!any) continue; // Stats::get returns false for a cell without an entry for this statistic (Stats.cpp:50-54)
			for (size_t id : presented) counts.push_back(id < n_chr ? row[id * 3 + stat] : 0);
			cell_barcodes.push_back(c.barcode());
		}
		for (size_t id : presented) chromosome_names.push_back(_chromosome_indexer.get_value(id)); // Stats::presented_chromosomes, :65-73
	}

	// ---- ResultsPrinter ---------------------------------------------------------------------------------------------------
	ResultsPrinter::SparseMatrix ResultsPrinter::get_count_matrix(const CellsDataContainer &container, bool filtered) const
	{
		dge_handle *h = container.handle();
		size_t n_cols = 0, nnz = 0;
		const int which = filtered ? DGE_MATRIX_CM : DGE_MATRIX_CM_RAW;
		if (dge_get_matrix(h, which, nullptr, nullptr, nullptr, &n_cols, &nnz) != DGE_OK) throw std::runtime_error(dge_last_error(h));
		std::vector<int64_t> indptr(n_cols + 1);
		std::vector<int32_t> genes(nnz), vals(nnz);
		if (dge_get_matrix(h, which, indptr.data(), genes.data(), vals.data(), &n_cols, &nnz) != DGE_OK) throw std::runtime_error(dge_last_error(h));
		return assemble(container, filtered, indptr, genes, vals);
	}

	ResultsPrinter::SparseMatrix ResultsPrinter::get_count_matrix_filtered(const CellsDataContainer &container, const UMI::Mark::query_t &query_marks) const
	{
		dge_handle *h = container.handle();
		uint32_t mask = 0;
		for (auto const &m : query_marks) mask |= 1u << m.bits();
		size_t n_cols = 0, nnz = 0;
		if (dge_get_matrix_marks(h, mask, nullptr, nullptr, nullptr, &n_cols, &nnz) != DGE_OK) throw std::runtime_error(dge_last_error(h));
		std::vector<int64_t> indptr(n_cols + 1);
		std::vector<int32_t> genes(nnz), vals(nnz);
		if (dge_get_matrix_marks(h, mask, indptr.data(), genes.data(), vals.data(), &n_cols, &nnz) != DGE_OK) throw std::runtime_error(dge_last_error(h));
		return assemble(container, true, indptr, genes, vals);
	}

	// device CSC (columns = cells in the reference's order, gene ids ascending) -> the reference's matrix: its row numbering and names
	ResultsPrinter::SparseMatrix ResultsPrinter::assemble(const CellsDataContainer &container, bool filtered, const std::vector<int64_t> &indptr,
	                                                      const std::vector<int32_t> &genes, const std::vector<int32_t> &vals) const
	{
		SparseMatrix m;
		dge_handle *h = container.handle();
		const size_t n_cols = indptr.size() - 1, nnz = genes.size();
		size_t n_cells = 0;
		const int cls = filtered ? DGE_CELLS_FILTERED : DGE_CELLS_REAL;
		dge_get_cells(h, cls, nullptr, 0, &n_cells);
		std::vector<dge_cell_info> info(n_cells);
		if (n_cells) dge_get_cells(h, cls, info.data(), n_cells, &n_cells);
		for (auto const &ci : info) m.col_names.push_back(container.barcode_string(ci.barcode));

		// Row numbering = first time a gene name is met while walking columns (ResultsPrinter.cpp:345-356 / :379-388).  For the
		// filtered matrix the reference walks, per column, a std::unordered_map<std::string,size_t> built in gene-index order
		// (Cell.cpp:54-68): reproduce that walk with the same container so row order matches under the same libstdc++.
		const StringIndexer &gi = container.gene_indexer();
		std::unordered_map<int32_t, int32_t> row_of_gene;
		std::vector<int32_t> row(nnz);
		for (size_t c = 0; c < n_cols; ++c)
		{
			if (filtered)
			{
				std::unordered_map<std::string, size_t> per_gene;
				for (int64_t k = indptr[c]; k < indptr[c + 1]; ++k) per_gene.emplace(gi.get_value(size_t(genes[size_t(k)])), size_t(genes[size_t(k)]));
				for (auto const &pg : per_gene)
					if (row_of_gene.emplace(int32_t(pg.second), int32_t(row_of_gene.size())).second) m.row_names.push_back(pg.first);
			}
			else
				for (int64_t k = indptr[c]; k < indptr[c + 1]; ++k)
					if (row_of_gene.emplace(genes[size_t(k)], int32_t(row_of_gene.size())).second) m.row_names.push_back(gi.get_value(size_t(genes[size_t(k)])));
			for (int64_t k = indptr[c]; k < indptr[c + 1]; ++k) row[size_t(k)] = row_of_gene.at(genes[size_t(k)]);
		}
		// Eigen::SparseMatrix::setFromTriplets (ResultsPrinter.cpp:436-437): column-major, row indices ascending inside a column
		m.p.resize(n_cols + 1);
		m.i.resize(nnz); m.x.resize(nnz);
		std::vector<std::pair<int32_t, int32_t>> colbuf;
		for (size_t c = 0; c < n_cols; ++c)
		{
			m.p[c] = int32_t(indptr[c]);
			colbuf.clear();
			for (int64_t k = indptr[c]; k < indptr[c + 1]; ++k) colbuf.emplace_back(row[size_t(k)], vals[size_t(k)]);
			std::sort(colbuf.begin(), colbuf.end());
			for (size_t k = 0; k < colbuf.size(); ++k) { m.i[size_t(indptr[c]) + k] = colbuf[k].first; m.x[size_t(indptr[c]) + k] = double(colbuf[k].second); }
		}
		m.p[n_cols] = int32_t(nnz);
		return m;
	}

	void ResultsPrinter::save_mtx(const SparseMatrix &m, const std::string &filename_base)
	{
		// Matrix::writeMM(d$cm, base.mtx) + write.table(colnames / rownames)  (ResultsPrinter.cpp:81-91).  writeMM hands a dgCMatrix to
		// CHOLMOD's MatrixMarket writer (third party, not in the reference tree: Matrix >= 1.2 / SuiteSparse cholmod_write_sparse), which
		// declares a real matrix whose entries are all integer-valued -- as every count matrix is -- as "integer" and prints the
		// entries without a fraction, 1-based, column by column.  Parity of this file is unpinned (no reference test reads it back).
		std::ofstream f(filename_base + ".mtx");
		f << "%%MatrixMarket matrix coordinate integer general\n";
		f << m.row_names.size() << " " << m.col_names.size() << " " << m.i.size() << "\n";
		for (size_t c = 0; c + 1 < m.p.size(); ++c)
			for (int32_t k = m.p[c]; k < m.p[c + 1]; ++k) f << (m.i[size_t(k)] + 1) << " " << (c + 1) << " " << int64_t(m.x[size_t(k)]) << "\n";
		std::ofstream fc(filename_base + ".cells.tsv");
		for (auto const &n : m.col_names) fc << n << "\n";
		std::ofstream fg(filename_base + ".genes.tsv");
		for (auto const &n : m.row_names) fg << n << "\n";
	}

	// ---- minimal writer of R's serialization format (XDR, version 2), gzip-compressed like saveRDS(compress=TRUE) --------------
	namespace
	{
		class RdsWriter
		{
			gzFile _f;
			std::unordered_map<std::string, int> _sym; // symbol reference table (REFSXP indices start at 1)

			void raw(const void *p, size_t n) { if (gzwrite(_f, p, unsigned(n)) != int(n)) throw std::runtime_error("rds: write failed"); }

		public:
			explicit RdsWriter(const std::string &fname) : _f(gzopen(fname.c_str(), "wb"))
			{
				if (!_f) throw std::runtime_error("can't write " + fname);
				raw("X\n", 2);
				i32(2); i32(0x00030600); i32(0x00020300); // format 2, written by R 3.6.0, readable from 2.3.0
			}
			~RdsWriter() { if (_f) gzclose(_f); }
			void i32(int32_t v) { unsigned char b[4] = {(unsigned char)(v >> 24), (unsigned char)(v >> 16), (unsigned char)(v >> 8), (unsigned char)v}; raw(b, 4); }
			void f64(double d) { uint64_t u; std::memcpy(&u, &d, 8); unsigned char b[8]; for (int k = 0; k < 8; ++k) b[k] = (unsigned char)(u >> (56 - 8 * k)); raw(b, 8); }
			void flags(int type, bool obj = false, bool attr = false, bool tag = false, int levels = 0)
			{
				i32(type | (obj ? 1 << 8 : 0) | (attr ? 1 << 9 : 0) | (tag ? 1 << 10 : 0) | (levels << 12));
			}
			void charsxp(const std::string &s) { flags(9, false, false, false, 64 /* ASCII */); i32(int32_t(s.size())); raw(s.data(), s.size()); }
			void symbol(const std::string &s)
			{
				auto it = _sym.find(s);
				if (it != _sym.end()) { i32((it->second << 8) | 255); return; } // REFSXP
				_sym.emplace(s, int(_sym.size()) + 1);
				flags(1);
				charsxp(s);
			}
			void nil() { i32(254); }
			void tagged(const std::string &name) { flags(2, false, false, true); symbol(name); } // pairlist cell header; value follows
			void strvec_body(const std::vector<std::string> &v) { i32(int32_t(v.size())); for (auto const &s : v) charsxp(s); }
			void strvec(const std::vector<std::string> &v) { flags(16); strvec_body(v); }
			void intvec(const std::vector<int32_t> &v, const std::vector<std::string> *names = nullptr)
			{
				flags(13, false, names != nullptr);
				i32(int32_t(v.size()));
				for (int32_t x : v) i32(x);
				if (names) { tagged("names"); strvec(*names); nil(); }
			}
			void realvec(const std::vector<double> &v, const std::vector<std::string> *names = nullptr)
			{
				flags(14, false, names != nullptr);
				i32(int32_t(v.size()));
				for (double x : v) f64(x);
				if (names) { tagged("names"); strvec(*names); nil(); }
			}
			// integer matrix, column-major, with dimnames (what Rcpp's IntegerMatrix + rownames/colnames wraps to)
			void intmatrix(const std::vector<int32_t> &col_major, const std::vector<std::string> &row_names, const std::vector<std::string> &col_names)
			{
				flags(13, false, true);
				i32(int32_t(col_major.size()));
				for (int32_t x : col_major) i32(x);
				tagged("dim"); intvec({int32_t(row_names.size()), int32_t(col_names.size())});
				tagged("dimnames"); flags(19); i32(2); strvec(row_names); strvec(col_names);
				nil();
			}
			void named_strvec(const std::vector<std::string> &v, const std::vector<std::string> &names)
			{
				flags(16, false, true); strvec_body(v);
				tagged("names"); strvec(names); nil();
			}
			void list_header(size_t n) { flags(19, false, true); i32(int32_t(n)); }
			void plain_list_header(size_t n) { flags(19); i32(int32_t(n)); } // a list without attributes (List::create without names)
			void list_names(const std::vector<std::string> &names) { tagged("names"); strvec(names); nil(); }
			void dgcmatrix(const ResultsPrinter::SparseMatrix &m)
			{
				// S4 object of class "dgCMatrix" (package Matrix): slots i, p, Dim, Dimnames, x, factors
				flags(25, true, true);
				tagged("i"); intvec(m.i);
				tagged("p"); intvec(m.p);
				tagged("Dim"); intvec({int32_t(m.row_names.size()), int32_t(m.col_names.size())});
				tagged("Dimnames"); flags(19); i32(2); strvec(m.row_names); strvec(m.col_names);
				tagged("x"); realvec(m.x);
				tagged("factors"); flags(19); i32(0);
				tagged("class"); flags(16, false, true); strvec_body({"dgCMatrix"}); tagged("package"); strvec({"Matrix"}); nil();
				nil();
			}
		};
	}

	ResultsPrinter::ReadsPerUmiPerCell ResultsPrinter::get_reads_per_umi_per_cell(const CellsDataContainer &container) const
	{
		if (!container.umi_qualities_saved())
			throw std::runtime_error("dropest_b200: reads_per_umi_per_cell needs a CellsDataContainer built with save_umi_qualities = true");
		ReadsPerUmiPerCell res;
		StringIndexer cell_indexer, gene_indexer;
		for (size_t container_cell_id : container.filtered_cells())
		{
			const Cell &cur_cell = container.cell(container_cell_id);
			const unsigned cell_id = unsigned(cell_indexer.add(cur_cell.barcode()));
			for (auto const &gene_rpus : cur_cell.requested_reads_per_umi_per_gene(container.gene_match_level()))
			{
				const unsigned gene_id = unsigned(gene_indexer.add(gene_rpus.first));
				ReadsPerUmiPerCell::Entry e;
				for (auto const &umi_reads : gene_rpus.second)
				{
					e.umis.push_back(umi_reads.first);
					e.reads.push_back(unsigned(umi_reads.second));
					e.mean_quality.push_back(cur_cell.at(gene_rpus.first).at(umi_reads.first).mean_quality());
				}
				res.reads_per_umi.push_back(std::move(e));
				res.cell_indexes.push_back(cell_id);
				res.gene_indexes.push_back(gene_id);
			}
		}
		res.cells = cell_indexer.values();
		res.genes = gene_indexer.values();
		return res;
	}

	void ResultsPrinter::save_rds(const CellsDataContainer &container, const SparseMatrix &cm, const SparseMatrix &cm_raw,
	                              const std::string &filename_base) const
	{
		// d <- list(cm, cm_raw, reads_per_chr_per_cells, mean_reads_per_umi, saturation_info, merge_targets, aligned_reads_per_cell,
		// aligned_umis_per_cell, requested_umis_per_cb, requested_reads_per_cb); saveRDS(d, base.rds) -- field names, order and meaning:
		// ResultsPrinter.cpp:36-57, docs/dropest.rst:178-194.  reads_per_umi_per_cell (only with -u info, needs per-UMI base qualities) is not produced.
		std::vector<std::string> mt_from, mt_to, real_names, sat_cbs, sat_umis;
		std::vector<int32_t> reads, umis, req_umis, req_reads, sat_reads;
		std::vector<double> mean_rpu;
		const auto &targets = container.merge_targets();
		const auto &query = container.gene_match_level();
		for (size_t i = 0; i < container.total_cells_number(); ++i)
		{
			const Cell &c = container.cell(i);
			if (targets[i] != i) { mt_from.push_back(c.barcode()); mt_to.push_back(container.cell(targets[i]).barcode()); }
			if (!c.is_real()) continue;
			real_names.push_back(c.barcode());
			reads.push_back(c.stats().get(Stats::TOTAL_READS_PER_CB));
			umis.push_back(c.stats().get(Stats::TOTAL_UMIS_PER_CB));
			req_umis.push_back(int32_t(c.requested_umis_num()));
			size_t rr = 0;                                            // get_requested_umis_per_cb(container, true), :398-431
			for (auto const &g : c.requested_umis_per_gene(query, true)) rr += g.second;
			req_reads.push_back(int32_t(rr));
			size_t n_umis = 0;                                        // get_mean_reads_per_umi, :227-258 (0 / 0 = NaN for a cell without UMIs, as in R)
			double n_reads = 0.0;
			for (auto const &g : c.genes())
			{
				for (auto const &u : g.second.umis()) n_reads += double(u.second.read_count());
				n_umis += g.second.size();
			}
			mean_rpu.push_back(n_reads / double(n_umis));
			for (auto const &g : c.requested_reads_per_umi_per_gene(query)) // get_saturation_analysis_info, :113-138
				for (auto const &u : g.second)
				{
					sat_cbs.push_back(c.barcode()); sat_umis.push_back(u.first); sat_reads.push_back(int32_t(u.second));
				}
		}
		const bool chr_tables = container.chromosome_stats_available();
		ReadsPerUmiPerCell rpupc;
		if (umi_correction_info) rpupc = get_reads_per_umi_per_cell(container);
		RdsWriter w(filename_base + ".rds");
		w.list_header((chr_tables ? 10 : 9) + (umi_correction_info ? 1 : 0));
		w.dgcmatrix(cm);
		w.dgcmatrix(cm_raw);
		if (chr_tables)
		{   // get_reads_per_chr_per_cell_info, :140-166: list(Exon, Intron, Intergenic) of cells x chromosomes integer matrices
			w.list_header(3);
			for (auto stat : {Stats::EXON_READS_PER_CHR_PER_CELL, Stats::INTRON_READS_PER_CHR_PER_CELL, Stats::INTERGENIC_READS_PER_CHR_PER_CELL})
			{
				CellsDataContainer::names_t cells, chrs;
				CellsDataContainer::counts_t counts;
				container.get_stat_by_real_cells(stat, cells, chrs, counts);
				std::vector<int32_t> col_major(counts.size());
				for (size_t r = 0; r < cells.size(); ++r)
					for (size_t ch = 0; ch < chrs.size(); ++ch) col_major[ch * cells.size() + r] = counts[r * chrs.size() + ch]; // create_matrix transposes, :102-111
				w.intmatrix(col_major, cells, chrs);
			}
			w.list_names({"Exon", "Intron", "Intergenic"});
		}
		w.realvec(mean_rpu, &real_names);
		w.list_header(3); w.intvec(sat_reads); w.strvec(sat_cbs); w.strvec(sat_umis); w.list_names({"reads", "cbs", "umis"});
		w.named_strvec(mt_to, mt_from);
		w.intvec(reads, &real_names);
		w.intvec(umis, &real_names);
		w.intvec(req_umis, &real_names);
		w.intvec(req_reads, &real_names);
		if (umi_correction_info)
		{   // d$reads_per_umi_per_cell <- list(cells, genes, cell_indexes, gene_indexes, reads_per_umi) (ResultsPrinter.cpp:59-64, 308-313); Rcpp wraps
			// unsigned values as numeric
			w.list_header(5);
			w.strvec(rpupc.cells);
			w.strvec(rpupc.genes);
			w.realvec(std::vector<double>(rpupc.cell_indexes.begin(), rpupc.cell_indexes.end()));
			w.realvec(std::vector<double>(rpupc.gene_indexes.begin(), rpupc.gene_indexes.end()));
			w.plain_list_header(rpupc.reads_per_umi.size());
			for (auto const &e : rpupc.reads_per_umi)
			{
				w.list_header(e.umis.size());
				for (size_t k = 0; k < e.umis.size(); ++k)
				{
					w.plain_list_header(2);
					w.realvec({double(e.reads[k])});
					w.realvec(e.mean_quality[k]);
				}
				w.list_names(e.umis);
			}
			w.list_names({"cells", "genes", "cell_indexes", "gene_indexes", "reads_per_umi"});
		}
		std::vector<std::string> fields = {"cm", "cm_raw"};
		if (chr_tables) fields.push_back("reads_per_chr_per_cells");
		for (const char *f : {"mean_reads_per_umi", "saturation_info", "merge_targets", "aligned_reads_per_cell", "aligned_umis_per_cell",
		                      "requested_umis_per_cb", "requested_reads_per_cb"}) fields.push_back(f);
		if (umi_correction_info) fields.push_back("reads_per_umi_per_cell");
		w.list_names(fields);
	}

	void ResultsPrinter::save_results(const CellsDataContainer &container, const std::string &filename) const
	{
		std::string base = filename;
		auto pos = filename.find_last_of('.');
		if (pos != std::string::npos && filename.substr(pos + 1) == "rds") base = filename.substr(0, pos); // extract_filename_base, :93-100
		if (validation_stats) throw std::runtime_error("dropest_b200: the merge validation statistics (-S, MergeProbabilityValidator) are not available");
		SparseMatrix cm = get_count_matrix(container, true), cm_raw = get_count_matrix(container, false);
		save_rds(container, cm, cm_raw, base);
		if (write_matrix) save_mtx(cm, base);
	}

	void ResultsPrinter::save_intron_exon_matrices(const CellsDataContainer &container, const std::string &filename) const
	{
		// -V: matrices <- list(exon, intron, spanning); saveRDS(matrices, base.matrices.rds)  (ResultsPrinter.cpp:455-474)
		std::string base = filename;
		auto pos = filename.find_last_of('.');
		if (pos != std::string::npos && filename.substr(pos + 1) == "rds") base = filename.substr(0, pos);
		const SparseMatrix exon = get_count_matrix_filtered(container, UMI::Mark::get_by_code("e"));
		const SparseMatrix intron = get_count_matrix_filtered(container, UMI::Mark::get_by_code("i"));
		const SparseMatrix spanning = get_count_matrix_filtered(container, UMI::Mark::get_by_code("BA"));
		RdsWriter w(base + ".matrices.rds");
		w.list_header(3);
		w.dgcmatrix(exon); w.dgcmatrix(intron); w.dgcmatrix(spanning);
		w.list_names({"exon", "intron", "spanning"});
	}
}
