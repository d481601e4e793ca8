// oracle/port/oracle_port.cpp -- TEST INFRASTRUCTURE ONLY: our CPU restatement of the dropEst count-matrix hot path.
//
// Never linked into, called by, or shipped with the product.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs execute it (as the checker / the baseline).
// PINNED: tests/test_oracle.py checks this program against (a) the literal expectations of the reference's own unit tests
// (Tests/TestEstimation.cpp, Tests/TestTools.cpp:47-54), (b) tests/golden/ref_pins.json = outputs of the compiled, unmodified
// reference (oracle/_ref), and (c) oracle/_ref/dropest_ref on seeded random streams whenever that binary is present.
//
// Same command line and DGEO0001 output as oracle/ref_driver/ref_driver.cpp.  Every block cites the reference lines it restates.
// Containers whose iteration order leaks into results in the reference (std::unordered_map / std::unordered_set / std::sort)
// are used here with the same key types and insertion sequences so that the leak is reproduced under the same libstdc++.
#include "../common/dge_io.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <functional>
#include <iostream>
#include <map>
#include <numeric>
#include <sstream>
#include <unordered_map>
#include <unordered_set>

namespace port
{
typedef std::vector<std::string> strs_t;

// ---- Tools/UtilFunctions.cpp:32-82 -----------------------------------------------------------------------------------
static unsigned edit_distance(const std::string &s1, const std::string &s2, bool skip_n = true, unsigned max_ed = 10000)
{
	const int n1 = int(s1.size()), n2 = int(s2.size());
	std::vector<int> column(n1 + 1);
	std::iota(column.begin(), column.end(), 0);
	for (int x = 1; x <= n2; ++x)
	{
		const int lo = std::max(0, x - int(max_ed)), hi = std::min(n1, x + int(max_ed));
		int lastdiag = column[lo];
		column[lo] = x;
		int best = x;
		for (int y = lo + 1; y <= hi; ++y)
		{
			const int olddiag = column[y];
			const bool match = s1[y - 1] == s2[x - 1] || (skip_n && (s1[y - 1] == 'N' || s2[x - 1] == 'N'));
			const int cand = std::min({column[y] + 1, column[y - 1] + 1, lastdiag + (match ? 0 : 1)});
			best = std::min(best, cand + std::abs(y - x));
			column[y] = cand;
			lastdiag = olddiag;
		}
		if (best > int(max_ed)) return unsigned(best);
	}
	return unsigned(column[n1]);
}

static unsigned hamming_distance(const std::string &a, const std::string &b, bool skip_n = true)
{
	if (a.size() != b.size()) throw std::runtime_error("Strings should have equal length");
	unsigned d = 0;
	for (size_t i = 0; i < a.size(); ++i)
		d += (a[i] != b[i] && !(skip_n && (a[i] == 'N' || b[i] == 'N'))) ? 1u : 0u;
	return d;
}

// ---- container state: Cell.h / Gene.h / UMI.h / Stats.h ----------------------------------------------------------------
struct Umi { size_t reads = 0; int mark = 0; };
typedef std::map<size_t, Umi> umis_t;          // Gene::umis_t, keyed by umi index (first-seen order)
typedef std::map<size_t, umis_t> genes_t;      // Cell::genes_t, keyed by gene index (first-seen order)

struct Indexer // StringIndexer.cpp:5-28
{
	strs_t values;
	std::unordered_map<std::string, size_t> index;
	size_t add(const std::string &v)
	{
		auto it = index.emplace(v, index.size());
		if (it.second) values.push_back(v);
		return it.first->second;
	}
};

struct Cell
{
	std::string barcode;
	genes_t genes;
	int total_reads = 0, total_umis = 0; // Stats TOTAL_READS_PER_CB / TOTAL_UMIS_PER_CB: counters, not set sizes
	bool merged = false, excluded = false;
	size_t requested_genes = 0, requested_umis = 0;
	std::unordered_map<size_t, int> chr_stat[3]; // Stats::_chromosome_stat_data: exon, intron, intergenic reads per chromosome id
};

struct Params
{
	std::string merge = "none", barcodes, barcodes_type = "const", umi_merge = "simple", marks = "eEBA";
	size_t min_genes_before = 10, min_genes_after = 10;
	unsigned max_cb_ed = 2, max_umi_ed = 1;
	double min_frac = 0.2, max_merge_prob = 1e-4, max_real_merge_prob = 1e-7, umi_mult = 2;
	int max_cells = -1;
	bool reads_output = false;
};

struct Container
{
	Params p;
	std::vector<Cell> cells;
	std::unordered_map<std::string, size_t> cell_by_cb;
	Indexer gene_idx, umi_idx;
	std::vector<int> query; // accumulated mark values that match (UMI::Mark::get_by_code, UMI.cpp:123-154)
	std::vector<size_t> filtered, merge_targets;
	size_t intergenic = 0, has_exon = 0, has_intron = 0, has_na = 0, n_real = 0;
	bool initialized = false;
	// Stats' process-wide statics (Stats.cpp:5-7): chromosome ids in first-seen order, the ids each statistic has seen
	std::unordered_map<std::string, size_t> chr_ids;
	strs_t chr_names;
	std::unordered_set<size_t> presented[3];

	void chr_inc(Cell &cell, int stat, const std::string &chr) // Stats::inc(CellChrStatType, subtype), Stats.cpp:23-28 + get_index :79-88
	{
		auto it = chr_ids.emplace(chr, chr_ids.size());
		if (it.second) chr_names.push_back(chr);
		presented[stat].insert(it.first->second);
		cell.chr_stat[stat][it.first->second]++;
	}

	size_t min_after() const { return std::max(p.min_genes_after, p.min_genes_before); } // MergeStrategyAbstract.cpp:8-11

	bool is_real(const Cell &c) const { return !c.excluded && !c.merged && c.genes.size() >= p.min_genes_before; } // Cell.cpp:125-128

	bool mark_matches(int m) const // UMI.cpp:76-85: exact equality with one of the query marks
	{
		return std::find(query.begin(), query.end(), m) != query.end();
	}

	// CellsDataContainer::add_record, CellsDataContainer.cpp:59-88 (+ :356-364, :309-327)
	void add_record(const std::string &cb, const std::string &umi, const std::string &gene, int mark, const std::string &chr = std::string())
	{
		if (initialized) throw std::runtime_error("Container is already initialized");
		auto res = cell_by_cb.emplace(cb, cell_by_cb.size());
		if (res.second) { cells.emplace_back(); cells.back().barcode = cb; }
		Cell &cell = cells[res.first->second];
		if (gene.empty()) { chr_inc(cell, 2, chr); ++intergenic; return; }
		const size_t g = gene_idx.add(gene);
		const size_t u = umi_idx.add(umi);
		umis_t &umis = cell.genes[g];
		auto ins = umis.emplace(u, Umi());
		ins.first->second.reads++;
		ins.first->second.mark |= mark;
		if (ins.second) cell.total_umis++;
		cell.total_reads++;
		if (mark & 2) { chr_inc(cell, 0, chr); ++has_exon; }
		if (mark & 4) { chr_inc(cell, 1, chr); ++has_intron; }
		if (mark & 1) ++has_na;
	}

	size_t requested_in_gene(const umis_t &umis, bool reads) const // Gene.cpp:60-79
	{
		size_t n = 0;
		for (auto const &u : umis)
			if (mark_matches(u.second.mark)) n += reads ? u.second.reads : 1;
		return n;
	}

	// CellsDataContainer::update_cell_sizes / update_filtered_gene_counts / compare_cells, CellsDataContainer.cpp:111-125,250-276,329-344
	void update_cell_sizes(size_t threshold, int cell_threshold)
	{
		n_real = 0;
		for (auto &c : cells)
		{
			c.requested_genes = c.requested_umis = 0;
			for (auto const &g : c.genes)
			{
				size_t k = requested_in_gene(g.second, false);
				if (k) { c.requested_umis += k; c.requested_genes++; }
			}
			if (is_real(c)) ++n_real;
		}
		filtered.clear();
		for (size_t i = 0; i < cells.size(); ++i)
			if (is_real(cells[i]) && cells[i].requested_genes >= threshold) filtered.push_back(i);
		std::sort(filtered.begin(), filtered.end(), [this](size_t a, size_t b) {
			const Cell &x = cells[a], &y = cells[b];
			if (x.requested_genes != y.requested_genes) return x.requested_genes < y.requested_genes;
			if (x.requested_umis != y.requested_umis) return x.requested_umis < y.requested_umis;
			if (x.total_umis != y.total_umis) return size_t(x.total_umis) < size_t(y.total_umis);
			return x.barcode < y.barcode;
		});
		if (cell_threshold > 0 && size_t(cell_threshold) < filtered.size())
			filtered.erase(filtered.begin(), filtered.end() - cell_threshold);
	}

	void set_initialized() // CellsDataContainer.cpp:163-175
	{
		if (initialized) throw std::runtime_error("Container is already initialized");
		update_cell_sizes(0, -1);
		initialized = true;
	}

	// MergeStrategyBase::get_umigs_intersect_size, MergeStrategyBase.cpp:100-147
	static size_t intersect(const Cell &a, const Cell &b)
	{
		size_t n = 0;
		auto ga = a.genes.begin(), gb = b.genes.begin();
		while (ga != a.genes.end() && gb != b.genes.end())
		{
			if (ga->first < gb->first) { ++ga; continue; }
			if (gb->first < ga->first) { ++gb; continue; }
			auto ua = ga->second.begin(), ub = gb->second.begin();
			while (ua != ga->second.end() && ub != gb->second.end())
			{
				if (ua->first < ub->first) ++ua;
				else if (ub->first < ua->first) ++ub;
				else { ++n; ++ua; ++ub; }
			}
			++ga; ++gb;
		}
		return n;
	}

	// CellsDataContainer::merge_cells, CellsDataContainer.cpp:90-104 (Gene::merge Gene.cpp:26-36, UMI::merge UMI.cpp:15-19, Stats::merge Stats.cpp:29-43)
	void merge_cells(size_t src_id, size_t dst_id)
	{
		Cell &src = cells.at(src_id), &dst = cells.at(dst_id);
		for (auto const &g : src.genes)
		{
			umis_t &t = dst.genes[g.first];
			for (auto const &u : g.second)
			{
				auto ins = t.insert(u);
				if (!ins.second) { ins.first->second.reads += u.second.reads; ins.first->second.mark |= u.second.mark; }
			}
		}
		dst.total_reads += src.total_reads;
		dst.total_umis += src.total_umis;
		for (int t = 0; t < 3; ++t)
			for (auto const &kv : src.chr_stat[t]) dst.chr_stat[t][kv.first] += kv.second;
		src.merged = true;
	}
};

// ---- whitelist: BarcodesParser.cpp / ConstLengthBarcodesParser.cpp / InDropBarcodesParser.cpp ---------------------------------
struct Whitelist
{
	std::vector<strs_t> parts;
	bool indrop = false;

	static std::string revcomp(const std::string &s) // Tools::ReverseComplement, UtilFunctions.cpp:97-115
	{
		std::string r(s.rbegin(), s.rend());
		for (char &c : r) c = c == 'A' ? 'T' : c == 'T' ? 'A' : c == 'G' ? 'C' : c == 'C' ? 'G' : c;
		return r;
	}

	void load(const std::string &fname, bool indrop_type) // BarcodesParser.cpp:117-144, ConstLength...cpp:50-69, InDrop...cpp:15-29
	{
		indrop = indrop_type;
		std::ifstream f(fname);
		if (f.fail()) throw std::runtime_error("Can't open file with barcodes: '" + fname + "'");
		std::string line;
		while (std::getline(f, line))
		{
			std::istringstream in(line);
			strs_t toks;
			std::string t;
			while (in >> t)
			{
				if (!indrop && !toks.empty() && toks[0].size() != t.size())
					throw std::runtime_error("All barcodes in one line must have the same length");
				toks.push_back(revcomp(t));
			}
			if (toks.empty()) throw std::runtime_error("File with barcodes (" + fname + ") has wrong format");
			parts.push_back(toks);
			if (indrop && parts.size() == 2) break;
		}
		if (parts.empty() || (indrop && parts.size() != 2)) throw std::runtime_error("File with barcodes (" + fname + ") has wrong format");
	}

	strs_t split(const std::string &cb) const // ConstLength...cpp:34-48, InDrop...cpp:31-38
	{
		strs_t res;
		if (indrop)
		{
			const size_t l2 = parts[1][0].size();
			res.push_back(cb.substr(0, cb.size() - l2));
			res.push_back(cb.substr(cb.size() - l2));
			return res;
		}
		size_t total = 0;
		for (auto const &p : parts) total += p[0].size();
		if (cb.size() != total) throw std::runtime_error("Barcode '" + cb + "' has wrong length (" + std::to_string(total) + " expected)");
		size_t pos = 0;
		for (auto const &p : parts) { res.push_back(cb.substr(pos, p[0].size())); pos += p[0].size(); }
		return res;
	}

	struct Combo { std::vector<size_t> inds; unsigned dist; };

	// BarcodesParser::get_distances_to_barcode + push_remaining_dists (BarcodesParser.cpp:21-74) followed by the sort of
	// RealBarcodesMergeStrategy.cpp:74-76
	std::vector<Combo> sorted_combos(const std::string &cb) const
	{
		struct IdxVal { size_t index; long value; };
		strs_t cb_parts = split(cb);
		std::vector<std::vector<IdxVal>> d(parts.size());
		for (size_t p = 0; p < parts.size(); ++p)
		{
			for (size_t t = 0; t < parts[p].size(); ++t) d[p].push_back(IdxVal{t, long(edit_distance(cb_parts[p], parts[p][t]))});
			std::sort(d[p].begin(), d[p].end(), [](const IdxVal &a, const IdxVal &b) { return a.value < b.value; });
		}
		std::vector<Combo> out;
		std::vector<size_t> cur(parts.size(), 0);
		std::function<void(size_t, unsigned)> walk = [&](size_t p, unsigned acc) {
			if (p == parts.size()) { out.push_back(Combo{cur, acc}); return; }
			for (auto const &iv : d[p])
			{
				if (acc + unsigned(iv.value) > 5) return; // MAX_REAL_MERGE_EDIT_DISTANCE, BarcodesParser.h:57
				cur[p] = iv.index;
				walk(p + 1, acc + unsigned(iv.value));
			}
		};
		walk(0, 0);
		std::sort(out.begin(), out.end(), [](const Combo &a, const Combo &b) { return a.dist < b.dist; });
		return out;
	}

	std::string barcode(const Combo &c) const
	{
		std::string s;
		for (size_t p = 0; p < parts.size(); ++p) s += parts[p].at(c.inds[p]);
		return s;
	}
};

// RealBarcodesMergeStrategy::get_real_neighbour_cbs, RealBarcodesMergeStrategy.cpp:63-109
static std::vector<size_t> real_neighbours(const Container &c, const Whitelist &wl, size_t base, bool poisson)
{
	std::vector<size_t> res;
	auto combos = wl.sorted_combos(c.cells[base].barcode);
	if (combos.empty()) return res;
	const unsigned min_d = combos.front().dist;
	unsigned max_dist = poisson ? (min_d == 0 ? 2 : min_d + 1) : min_d; // PoissonRealBarcodesMergeStrategy.cpp:20-23
	for (auto const &cmb : combos)
	{
		if (cmb.dist > max_dist && !res.empty()) break;
		auto it = c.cell_by_cb.find(wl.barcode(cmb));
		if (it != c.cell_by_cb.end() && c.cells[it->second].genes.size() >= c.p.min_genes_before &&
		    c.cells[it->second].total_umis >= c.cells[base].total_umis)
			res.push_back(it->second);
		max_dist = std::max(max_dist, cmb.dist);
	}
	return res;
}

// RealBarcodesMergeStrategy::get_merge_target / get_best_merge_target, RealBarcodesMergeStrategy.cpp:22-61
static long target_real(const Container &c, const Whitelist &wl, size_t base)
{
	std::vector<size_t> nb = real_neighbours(c, wl, base, false);
	if (nb.empty()) return -1;
	if (nb[0] == base) return long(base);
	double best_frac = 0;
	size_t best = nb[0];
	for (size_t n : nb)
	{
		const size_t isz = Container::intersect(c.cells[base], c.cells[n]);
		const double frac = 0.5 * isz * (1. / size_t(c.cells[base].total_umis) + 1. / size_t(c.cells[n].total_umis));
		if (best_frac < frac) { best_frac = frac; best = n; }
	}
	if (best_frac < c.p.min_frac) return -1;
	return long(best);
}

// SimpleMergeStrategy, SimpleMergeStrategy.cpp:16-103.  The inverted index is keyed by (umi, gene) like the reference; the
// per-base candidate counts live in std::unordered_map<size_t,size_t> exactly as in the reference because its iteration
// order decides ties (SimpleMergeStrategy.cpp:54-80).
struct PairHashFixed
{
	size_t operator()(const std::pair<size_t, size_t> &p) const { return std::hash<size_t>()(p.first) * 1000003u ^ std::hash<size_t>()(p.second); }
};
typedef std::unordered_map<std::pair<size_t, size_t>, std::unordered_set<size_t>, PairHashFixed> umig_index_t;

static umig_index_t build_umig_index(const Container &c)
{
	umig_index_t idx;
	for (size_t cell_id : c.filtered)
		for (auto const &g : c.cells[cell_id].genes)
			for (auto const &u : g.second) idx[std::make_pair(u.first, g.first)].emplace(cell_id);
	return idx;
}

static std::unordered_map<size_t, size_t> common_umigs(const Container &c, const umig_index_t &idx, size_t base)
{
	std::unordered_map<size_t, size_t> res;
	for (auto const &g : c.cells[base].genes)
		for (auto const &u : g.second)
			for (size_t other : idx.at(std::make_pair(u.first, g.first)))
			{
				if (other == base) continue;
				if (c.cells[other].genes.size() >= c.cells[base].genes.size()) res[other]++;
			}
	return res;
}

static long target_simple(const Container &c, const umig_index_t &idx, size_t base)
{
	auto cand = common_umigs(c, idx, base);
	long top = -1, top_genes = -1;
	double top_frac = -1;
	for (auto const &kv : cand)
	{
		const double frac = 0.5 * kv.second * (1. / size_t(c.cells[base].total_umis) + 1. / size_t(c.cells[kv.first].total_umis));
		if (frac - top_frac > 0.00001 || (std::abs(frac - top_frac) < 0.00001 && long(c.cells[kv.first].genes.size()) > top_genes))
		{
			const int ed = int(edit_distance(c.cells[base].barcode, c.cells[kv.first].barcode));
			if (ed >= int(c.p.max_cb_ed)) continue;
			top = long(kv.first); top_frac = frac; top_genes = long(c.cells[kv.first].genes.size());
		}
	}
	if (top_frac < c.p.min_frac) return long(base);
	return top;
}

// MergeAllMergeStrategy::get_merge_target, MergeAllMergeStrategy.h:16-50
static long target_all(const Container &c, size_t base)
{
	int min_ed = std::numeric_limits<int>::max(), max_umis = 0;
	size_t target = std::numeric_limits<size_t>::max();
	for (size_t cell : c.filtered)
	{
		const size_t n = size_t(c.cells[cell].total_umis);
		if (n <= size_t(c.cells[base].total_umis)) continue;
		const int ed = int(edit_distance(c.cells[base].barcode, c.cells[cell].barcode, false, c.p.max_cb_ed));
		if (ed > int(c.p.max_cb_ed)) continue;
		if (min_ed > ed) { min_ed = ed; max_umis = int(n); target = cell; }
		else if (min_ed == ed && max_umis < int(n)) { max_umis = int(n); target = cell; }
	}
	return target != std::numeric_limits<size_t>::max() ? long(target) : long(base);
}

// ---- the -M strategies: PoissonTargetEstimator.cpp:14-119, PoissonRealBarcodesMergeStrategy.cpp:20-51, PoissonSimpleMergeStrategy.cpp:15-42,
// CellsDataContainer::umi_distribution (CellsDataContainer.cpp:182-197), Tools::CollisionsAdjuster (CollisionsAdjuster.cpp:12-49),
// Tools::fpow (UtilFunctions.cpp:13-30).  ppois is R's (absent here): the same restatement as oracle/shim/RInside.h, so that port and
// compiled reference agree bit for bit; it feeds a threshold compare only (third-party arithmetic, unpinned by the reference).
static double fpow(double base, long exp)
{
	if (exp == 1) return base;
	double result = 1;
	while (exp)
	{
		if (exp & 1) result *= base;
		exp >>= 1;
		base *= base;
	}
	return result;
}

namespace ppois_restated
{
	static double log_pmf(long k, double lambda) { return -lambda + k * std::log(lambda) - std::lgamma(double(k) + 1.0); }
	static double lower(long x, double lambda)
	{
		if (x < 0) return 0;
		if (lambda <= 0) return 1;
		double s = 0;
		for (long k = x; k >= 0; --k)
		{
			double t = std::exp(log_pmf(k, lambda));
			s += t;
			if (double(k) < lambda && t < s * 1e-17) break;
		}
		return s > 1 ? 1 : s;
	}
	static double upper(long x, double lambda) // P[X > x]
	{
		if (x < 0) return 1;
		if (lambda <= 0) return 0;
		if (double(x + 1) < lambda) return 1 - lower(x, lambda);
		double s = 0;
		for (long k = x + 1;; ++k)
		{
			double t = std::exp(log_pmf(k, lambda));
			s += t;
			if (t <= s * 1e-17 || k > x + 100000) break;
		}
		return s > 1 ? 1 : s;
	}
}

struct PoissonEstimator
{
	double max_merge_prob, max_real_cb_merge_prob;
	std::vector<double> umi_dist, neg_prod;       // _umi_distribution; CollisionsAdjuster::_umi_probabilities_neg_prod
	std::vector<size_t> adjusted;                 // CollisionsAdjuster::_adjusted_sizes
	double sum_collisions = 0;
	size_t last_total = 0;
	std::map<std::pair<size_t, size_t>, double> memo; // _estimated_gene_intersections (a cache: its container type cannot change a value)

	void init(const Container &c) // PoissonTargetEstimator::init over CellsDataContainer::umi_distribution()
	{
		std::unordered_map<std::string, size_t> dist; // s_ul_hash_t: its iteration order fixes the order of the probability vector
		for (size_t cell_id : c.filtered)
			for (auto const &g : c.cells[cell_id].genes)
				for (auto const &u : g.second) dist[c.umi_idx.values[u.first]]++;
		double sum = 0;
		for (auto const &it : dist) sum += it.second;
		for (auto const &it : dist) umi_dist.push_back(it.second / sum);
		neg_prod.assign(umi_dist.size(), 1);
	}

	size_t adjust(size_t expression) // CollisionsAdjuster::estimate_adjusted_gene_expression + update_adjusted_sizes
	{
		for (size_t s = adjusted.size() + 1; s <= expression; ++s)
		{
			const size_t total = s + size_t(sum_collisions);
			double new_umi_prob = 0;
			for (size_t i = 0; i < umi_dist.size(); ++i)
			{
				neg_prod[i] *= fpow(1 - umi_dist[i], long(total - last_total));
				new_umi_prob += umi_dist[i] * (1 - neg_prod[i]);
			}
			last_total = total;
			const double collision_num = 1.0 / (1.0 - new_umi_prob) - 1.0;
			sum_collisions += collision_num;
			adjusted.push_back(size_t(std::lround(s + sum_collisions)));
		}
		return adjusted.at(expression - 1);
	}

	double genes_intersection(size_t g1, size_t g2) // estimate_genes_intersection_size, PoissonTargetEstimator.cpp:92-119
	{
		if (g1 > g2) std::swap(g1, g2);
		g1 = adjust(g1);
		g2 = adjust(g2);
		auto key = std::make_pair(g1, g2);
		auto it = memo.find(key);
		if (it != memo.end()) return it->second;
		const size_t d = g2 - g1;
		double est = 0;
		for (size_t i = 0; i < umi_dist.size(); ++i)
		{
			const double min_prob = fpow(1 - umi_dist[i], long(g1));
			const double max_prob = min_prob * fpow(1 - umi_dist[i], long(d));
			est += (1 - min_prob) * (1 - max_prob);
		}
		memo.emplace(key, est);
		return est;
	}

	double merge_probability(const Container &c, size_t a, size_t b) // estimate_intersection_prob, :66-90
	{
		const Cell &c1 = c.cells[a], &c2 = c.cells[b];
		const size_t isz = Container::intersect(c1, c2);
		if (isz == 0) return 1;
		double expected = 0;
		for (auto const &g1 : c1.genes)
		{
			auto g2 = c2.genes.find(g1.first);
			if (g2 == c2.genes.end()) continue;
			expected += genes_intersection(g1.second.size(), g2->second.size());
		}
		return ppois_restated::upper(long(isz) - 1, expected);
	}

	long best_target(const Container &c, size_t base, const std::vector<size_t> &nb) // get_best_merge_target, :14-44
	{
		const bool base_is_real = base == nb.at(0);
		double thr = base_is_real ? max_merge_prob : max_real_cb_merge_prob;
		thr /= nb.size();
		long best = -1;
		double min_prob = 2;
		for (size_t n : nb)
		{
			if (n == base) continue;
			const double prob = merge_probability(c, base, n);
			if (prob < min_prob) { min_prob = prob; best = long(n); }
		}
		if (min_prob > thr) return base_is_real ? long(base) : -1;
		return best;
	}
};

// PoissonRealBarcodesMergeStrategy: RealBarcodesMergeStrategy::get_merge_target (:22-29) with the Poisson best-target rule
static long target_poisson_real(const Container &c, const Whitelist &wl, PoissonEstimator &est, size_t base)
{
	std::vector<size_t> nb = real_neighbours(c, wl, base, true);
	if (nb.empty()) return -1;
	return est.best_target(c, base, nb);
}

// PoissonSimpleMergeStrategy::get_merge_target, PoissonSimpleMergeStrategy.cpp:15-42
static long target_poisson_simple(const Container &c, const umig_index_t &idx, PoissonEstimator &est, size_t base)
{
	auto cand = common_umigs(c, idx, base);
	std::vector<size_t> nb;
	for (auto const &kv : cand)
	{
		if (edit_distance(c.cells[base].barcode, c.cells[kv.first].barcode) > c.p.max_cb_ed) continue;
		nb.push_back(kv.first);
	}
	if (nb.empty()) return long(base);
	const long t = est.best_target(c, base, nb);
	return t != -1 ? t : long(base);
}

// MergeStrategyBase::merge_inited + reassign, MergeStrategyBase.cpp:11-82
static void merge_cells_stage(Container &c, const Whitelist &wl)
{
	const size_t n = c.cells.size();
	c.merge_targets.resize(n);
	std::iota(c.merge_targets.begin(), c.merge_targets.end(), 0);
	if (c.p.merge == "none") return; // DummyMergeStrategy.h:12-17
	umig_index_t idx;
	if (c.p.merge == "simple" || c.p.merge == "poisson_simple") idx = build_umig_index(c);
	PoissonEstimator est{c.p.max_merge_prob, c.p.max_real_merge_prob};
	if (c.p.merge == "poisson_simple" || c.p.merge == "poisson_real") est.init(c);
	std::vector<long> targets(c.filtered.size());
	for (size_t k = 0; k < c.filtered.size(); ++k)
	{
		if (c.p.merge == "real") targets[k] = target_real(c, wl, c.filtered[k]);
		else if (c.p.merge == "simple") targets[k] = target_simple(c, idx, c.filtered[k]);
		else if (c.p.merge == "all") targets[k] = target_all(c, c.filtered[k]);
		else if (c.p.merge == "poisson_real") targets[k] = target_poisson_real(c, wl, est, c.filtered[k]);
		else if (c.p.merge == "poisson_simple") targets[k] = target_poisson_simple(c, idx, est, c.filtered[k]);
		else throw std::runtime_error("merge type not restated in oracle/port: " + c.p.merge);
	}
	std::unordered_map<size_t, std::unordered_set<size_t>> moved_to;
	for (size_t k = 0; k < c.filtered.size(); ++k)
	{
		const size_t base = c.filtered[k];
		long t = targets[k];
		if (t < 0) { c.cells.at(base).excluded = true; continue; }
		if (size_t(t) != c.merge_targets.at(size_t(t))) t = long(c.merge_targets[size_t(t)]);
		if (size_t(t) == base) continue;
		c.merge_cells(base, size_t(t));
		c.merge_targets[base] = size_t(t);
		moved_to[size_t(t)].insert(base);
		auto it = moved_to.find(base);
		if (it == moved_to.end()) continue;
		std::vector<size_t> kids(it->second.begin(), it->second.end());
		for (size_t kid : kids) { c.merge_targets[kid] = size_t(t); moved_to[size_t(t)].insert(kid); }
		moved_to[base].clear();
	}
}

// ---- UMI merge: MergeUMIsStrategySimple.cpp:21-102, MergeUMIsStrategyAbstract.cpp:11-23, Cell.cpp:31-42, Gene.cpp:38-58 -----------
static std::string fix_n_random(const std::string &umi)
{
	std::string t(umi);
	for (char &ch : t)
		if (ch == 'N') ch = "ACGT"[rand() % 4];
	return t;
}

static void apply_umi_targets(Container &c, Cell &cell, size_t gene, const std::unordered_map<std::string, std::string> &targets)
{
	umis_t &umis = cell.genes.at(gene);
	for (auto const &t : targets)
	{
		if (t.second == t.first) continue;
		auto src = umis.find(c.umi_idx.index.at(t.first));
		if (src == umis.end()) throw std::runtime_error("Source UMI doesn't belong to the gene: " + t.first);
		auto dst = umis.emplace(c.umi_idx.add(t.second), src->second);
		if (!dst.second) { dst.first->second.reads += src->second.reads; dst.first->second.mark |= src->second.mark; }
		umis.erase(src);
		cell.total_umis--;
	}
}

static void merge_umis_simple(Container &c)
{
	for (auto &cell : c.cells)
	{
		if (!c.is_real(cell)) continue;
		for (auto &g : cell.genes)
		{
			std::unordered_set<std::string> bad;
			for (auto const &u : g.second)
			{
				const std::string &seq = c.umi_idx.values.at(u.first);
				if (seq.find('N') != std::string::npos) bad.insert(seq);
			}
			if (bad.empty()) continue;
			std::unordered_map<std::string, std::string> targets;
			for (auto const &b : bad)
			{
				int min_ed = std::numeric_limits<unsigned>::max(); // wraps to -1 exactly like MergeUMIsStrategySimple.cpp:73
				std::string best;
				long best_size = 0;
				for (auto const &u : g.second)
				{
					const std::string &seq = c.umi_idx.values.at(u.first);
					if (bad.count(seq)) continue;
					unsigned ed = hamming_distance(seq, b);
					if (ed < unsigned(min_ed) || (int(ed) == min_ed && long(u.second.reads) > best_size))
					{
						min_ed = int(ed); best = seq; best_size = long(u.second.reads);
					}
				}
				if (best.empty() || unsigned(min_ed) > c.p.max_umi_ed) targets[b] = fix_n_random(b);
				else targets[b] = best;
			}
			apply_umi_targets(c, cell, g.first, targets);
		}
	}
}

// MergeUMIsStrategyDirectional.cpp:18-116
static void merge_umis_directional(Container &c)
{
	struct W { std::string seq; size_t reads; };
	for (auto &cell : c.cells)
	{
		if (!c.is_real(cell)) continue;
		for (auto &g : cell.genes)
		{
			std::vector<W> umis;
			for (auto const &u : g.second) umis.push_back(W{c.umi_idx.values.at(u.first), u.second.reads});
			std::sort(umis.begin(), umis.end(), [](const W &a, const W &b) { return a.reads < b.reads; });
			std::unordered_map<std::string, std::string> targets;
			for (size_t s = 0; s < umis.size(); ++s)
			{
				const bool has_n = umis[s].seq.find('N') != std::string::npos;
				std::string target;
				unsigned min_ed = std::numeric_limits<unsigned>::max();
				for (long d = long(umis.size()) - 1; d > long(s); --d)
				{
					if (umis[s].reads * c.p.umi_mult > umis[size_t(d)].reads) break;
					unsigned ed = edit_distance(umis[s].seq, umis[size_t(d)].seq, true, c.p.max_umi_ed);
					if (ed > c.p.max_umi_ed) continue;
					if (ed < min_ed)
					{
						target = umis[size_t(d)].seq;
						if ((!has_n && ed <= 1) || ed == 0) break;
						min_ed = ed;
					}
				}
				if (has_n && target.empty()) target = fix_n_random(umis[s].seq);
				if (!target.empty()) targets[umis[s].seq] = target;
			}
			for (long i = long(umis.size()) - 1; i >= 0; --i)
			{
				auto it = targets.find(umis[size_t(i)].seq);
				if (it == targets.end()) continue;
				auto it2 = targets.find(it->second);
				if (it2 == targets.end()) continue;
				targets[umis[size_t(i)].seq] = it2->second;
			}
			if (!targets.empty()) apply_umi_targets(c, cell, g.first, targets);
		}
	}
}

static void merge_and_filter(Container &c, const Whitelist &wl) // CellsDataContainer.cpp:39-57
{
	if (!c.initialized) throw std::runtime_error("You must initialize container");
	merge_cells_stage(c, wl);
	if (c.p.umi_merge == "directional") merge_umis_directional(c);
	else merge_umis_simple(c);
	c.update_cell_sizes(c.min_after(), c.p.max_cells);
}
} // namespace port

// =====================================================================================================================
using namespace port;

static std::vector<int> marks_from_code(const std::string &code)
{
	std::vector<int> q;
	for (char ch : code)
	{
		switch (ch)
		{
			case 'e': q.push_back(2); break; case 'i': q.push_back(4); break; case 'E': q.push_back(3); break;
			case 'I': q.push_back(5); break; case 'B': q.push_back(6); break; case 'A': q.push_back(7); break;
			default: throw std::runtime_error(std::string("Unexpected gene match levels: ") + ch);
		}
	}
	return q;
}

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char **argv)
{
	try
	{
		Params p;
		std::string in, out;
		bool dump_umis = false, init_only = false;
		size_t limit = 0;
		for (int i = 1; i < argc; ++i)
		{
			std::string k = argv[i];
			auto next = [&]() -> std::string { if (i + 1 >= argc) throw std::runtime_error("missing value for " + k); return argv[++i]; };
			if (k == "--in") in = next(); else if (k == "--out") out = next(); else if (k == "--merge") p.merge = next();
			else if (k == "--barcodes") p.barcodes = next(); else if (k == "--barcodes-type") p.barcodes_type = next();
			else if (k == "--umi-merge") p.umi_merge = next(); else if (k == "--marks") p.marks = next();
			else if (k == "--min-genes-before") p.min_genes_before = std::stoul(next());
			else if (k == "--min-genes-after") p.min_genes_after = std::stoul(next());
			else if (k == "--max-cb-ed") p.max_cb_ed = unsigned(std::stoul(next()));
			else if (k == "--max-umi-ed") p.max_umi_ed = unsigned(std::stoul(next()));
			else if (k == "--min-frac") p.min_frac = std::stod(next());
			else if (k == "--max-merge-prob") p.max_merge_prob = std::stod(next());
			else if (k == "--max-real-merge-prob") p.max_real_merge_prob = std::stod(next());
			else if (k == "--umi-mult") p.umi_mult = std::stod(next());
			else if (k == "--max-cells") p.max_cells = std::stoi(next());
			else if (k == "--limit") limit = std::stoul(next());
			else if (k == "--reads-output") p.reads_output = true; else if (k == "--dump-umis") dump_umis = true;
			else if (k == "--init-only") init_only = true; else if (k == "--stream") {}
			else if (k == "--edit-distance")
			{   // pin helper: --edit-distance s1 s2 skip_n max_ed
				std::string a = next(), b = next(); bool sn = next() != "0"; unsigned me = unsigned(std::stoul(next()));
				std::cout << edit_distance(a, b, sn, me) << "\n";
				return 0;
			}
			else throw std::runtime_error("unknown argument " + k);
		}
		if (in.empty() || out.empty()) throw std::runtime_error("usage: dropest_port --in reads.{bin,tsv} --out out.dgeo [options]");

		Container c;
		c.p = p;
		c.query = marks_from_code(p.marks);
		Whitelist wl;
		if (p.merge == "real" || p.merge == "poisson_real") wl.load(p.barcodes, p.barcodes_type == "indrop");
		if (p.umi_merge != "directional") srand(42); // only MergeUMIsStrategySimple's ctor seeds rand() (MergeUMIsStrategySimple.cpp:18)

		double t_fill = 0;
		size_t n_reads = 0;
		const bool is_tsv = in.size() > 4 && in.substr(in.size() - 4) == ".tsv";
		if (is_tsv)
		{
			std::ifstream f(in);
			if (!f) throw std::runtime_error("can't open " + in);
			std::string line;
			std::vector<strs_t> rows;
			while (std::getline(f, line))
			{
				if (line.empty() || line[0] == '#') continue;
				strs_t t; size_t s = 0;
				while (true) { size_t e = line.find('\t', s); t.push_back(line.substr(s, e == std::string::npos ? e : e - s)); if (e == std::string::npos) break; s = e + 1; }
				if (t.size() < 5) throw std::runtime_error("bad tsv line: " + line);
				rows.push_back(t);
			}
			double t0 = now_s();
			for (auto const &t : rows) c.add_record(t[0], t[1], t[2] == "-" ? "" : t[2], std::stoi(t[4]));
			t_fill = now_s() - t0;
			n_reads = rows.size();
		}
		else
		{
			dge_io::ReadStream s = dge_io::read_packed(in);
			n_reads = limit ? std::min(limit, s.recs.size()) : s.recs.size();
			strs_t gnames(s.n_genes);
			for (uint32_t g = 0; g < s.n_genes; ++g) gnames[g] = s.gene_name(g);
			double t0 = now_s();
			for (size_t i = 0; i < n_reads; ++i)
			{
				const dge_io::Record16 &r = s.recs[i];
				const uint32_t gid = r.gene & 0xFFFFFFu;
				c.add_record(s.cb_of(r), s.umi_of(r),
				             gid == dge_io::NO_GENE ? std::string() : gnames.at(gid), int((r.gene >> 24) & 7), s.chr_name(i));
			}
			t_fill = now_s() - t0;
		}
		double t0 = now_s();
		c.set_initialized();
		double t_init = now_s() - t0;
		std::vector<int64_t> filtered_pre(c.filtered.begin(), c.filtered.end());
		t0 = now_s();
		if (!init_only) merge_and_filter(c, wl);
		else { c.merge_targets.clear(); }
		double t_merge = now_s() - t0;

		dge_io::Writer w;
		const size_t n = c.cells.size();
		strs_t barcodes(n);
		std::vector<uint8_t> flags(n);
		std::vector<int32_t> n_genes(n), umis_stat(n), reads_stat(n);
		std::vector<int64_t> req_genes(n), req_umis(n);
		for (size_t i = 0; i < n; ++i)
		{
			const Cell &cell = c.cells[i];
			barcodes[i] = cell.barcode;
			flags[i] = uint8_t((c.is_real(cell) ? 1 : 0) | (cell.merged ? 2 : 0) | (cell.excluded ? 4 : 0));
			n_genes[i] = int32_t(cell.genes.size()); umis_stat[i] = cell.total_umis; reads_stat[i] = cell.total_reads;
			req_genes[i] = int64_t(cell.requested_genes); req_umis[i] = int64_t(cell.requested_umis);
		}
		w.add_strings("cell_barcodes", barcodes);
		w.add("cell_flags", dge_io::U8, flags);
		w.add("cell_n_genes", dge_io::I32, n_genes);
		w.add("cell_umis_stat", dge_io::I32, umis_stat);
		w.add("cell_reads_stat", dge_io::I32, reads_stat);
		w.add("cell_req_genes", dge_io::I64, req_genes);
		w.add("cell_req_umis", dge_io::I64, req_umis);
		w.add("filtered_pre_merge", dge_io::I64, filtered_pre);
		w.add("filtered_cells", dge_io::I64, std::vector<int64_t>(c.filtered.begin(), c.filtered.end()));
		w.add("merge_targets", dge_io::I64, std::vector<int64_t>(c.merge_targets.begin(), c.merge_targets.end()));
		w.add_strings("gene_names", c.gene_idx.values);
		{   // get_stat_by_real_cells(CellChrStatType, ...), CellsDataContainer.cpp:292-307 + Stats::get / presented_chromosomes, Stats.cpp:50-73
			const char *names[3] = {"chr_exon", "chr_intron", "chr_intergenic"};
			for (int t = 0; t < 3; ++t)
			{
				strs_t cells_out, chrs;
				std::vector<int32_t> counts;
				for (auto const &cell : c.cells)
				{
					if (!c.is_real(cell) || cell.chr_stat[t].empty()) continue;
					for (size_t id : c.presented[t])
					{
						auto it = cell.chr_stat[t].find(id);
						counts.push_back(it == cell.chr_stat[t].end() ? 0 : it->second);
					}
					cells_out.push_back(cell.barcode);
				}
				for (size_t id : c.presented[t]) chrs.push_back(c.chr_names[id]);
				w.add_strings(std::string(names[t]) + "_cells", cells_out);
				w.add_strings(std::string(names[t]) + "_chrs", chrs);
				w.add(std::string(names[t]) + "_counts", dge_io::I32, counts);
			}
		}
		{   // cm: ResultsPrinter.cpp:334-361 (row order: first met while walking the per-cell unordered_map of Cell.cpp:54-68)
			std::vector<int64_t> col, gene, val;
			strs_t row_names;
			std::unordered_map<std::string, size_t> gene_ids;
			for (size_t k = 0; k < c.filtered.size(); ++k)
			{
				const Cell &cell = c.cells[c.filtered[k]];
				std::unordered_map<std::string, size_t> per_gene;
				for (auto const &g : cell.genes)
				{
					size_t v = c.requested_in_gene(g.second, p.reads_output);
					if (!v) continue;
					col.push_back(int64_t(k)); gene.push_back(int64_t(g.first)); val.push_back(int64_t(v));
					per_gene.emplace(c.gene_idx.values.at(g.first), v);
				}
				for (auto const &pg : per_gene)
					if (gene_ids.emplace(pg.first, gene_ids.size()).second) row_names.push_back(pg.first);
			}
			w.add("cm_col", dge_io::I64, col); w.add("cm_gene", dge_io::I64, gene); w.add("cm_val", dge_io::I64, val);
			w.add_strings("cm_row_names", row_names);
		}
		{   // cm_raw: ResultsPrinter.cpp:363-396
			std::vector<int64_t> col, gene, val, cellsv;
			strs_t row_names;
			std::unordered_map<std::string, size_t> gene_ids;
			size_t column = 0;
			for (size_t i = 0; i < n; ++i)
			{
				const Cell &cell = c.cells[i];
				if (!c.is_real(cell)) continue;
				cellsv.push_back(int64_t(i));
				for (auto const &g : cell.genes)
				{
					const std::string &name = c.gene_idx.values.at(g.first);
					if (gene_ids.emplace(name, gene_ids.size()).second) row_names.push_back(name);
					size_t v = g.second.size();
					if (p.reads_output) { v = 0; for (auto const &u : g.second) v += u.second.reads; }
					col.push_back(int64_t(column)); gene.push_back(int64_t(g.first)); val.push_back(int64_t(v));
				}
				++column;
			}
			w.add("cm_raw_cells", dge_io::I64, cellsv);
			w.add("cm_raw_col", dge_io::I64, col); w.add("cm_raw_gene", dge_io::I64, gene); w.add("cm_raw_val", dge_io::I64, val);
			w.add_strings("cm_raw_row_names", row_names);
		}
		if (dump_umis)
		{
			std::vector<int64_t> ucell, ugene, ucount;
			std::vector<uint8_t> umark;
			strs_t useq;
			for (size_t i = 0; i < n; ++i)
				for (auto const &g : c.cells[i].genes)
					for (auto const &u : g.second)
					{
						ucell.push_back(int64_t(i)); ugene.push_back(int64_t(g.first)); ucount.push_back(int64_t(u.second.reads));
						umark.push_back(uint8_t(u.second.mark)); useq.push_back(c.umi_idx.values.at(u.first));
					}
			w.add("umi_cell", dge_io::I64, ucell); w.add("umi_gene", dge_io::I64, ugene); w.add("umi_count", dge_io::I64, ucount);
			w.add("umi_mark", dge_io::U8, umark);
			w.add_strings("umi_seq", useq);
		}
		w.add_scalar_i64("n_reads", int64_t(n_reads));
		w.add_scalar_i64("n_cells", int64_t(n));
		w.add_scalar_i64("real_cells_number", int64_t(c.n_real));
		w.add_scalar_i64("intergenic_reads", int64_t(c.intergenic));
		w.add_scalar_i64("has_exon_reads", int64_t(c.has_exon));
		w.add_scalar_i64("has_intron_reads", int64_t(c.has_intron));
		w.add_scalar_i64("has_not_annotated_reads", int64_t(c.has_na));
		w.add_scalar_f64("t_decode_s", 0.0);
		w.add_scalar_f64("t_fill_s", t_fill);
		w.add_scalar_f64("t_init_s", t_init);
		w.add_scalar_f64("t_merge_s", t_merge);
		w.write(out);
		std::cerr << "dropest_port: " << n_reads << " reads, " << n << " cells, " << c.n_real << " real, " << c.filtered.size()
		          << " filtered; fill " << t_fill << " s, init " << t_init << " s, merge " << t_merge << " s\n";
		return 0;
	}
	catch (std::exception &e)
	{
		std::cerr << "dropest_port: ERROR: " << e.what() << "\n";
		return 1;
	}
}
