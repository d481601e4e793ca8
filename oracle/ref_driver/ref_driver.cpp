// oracle/ref_driver/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY (checker + CPU baseline; never shipped or called by the product).
//
// Drives the UNMODIFIED reference hot path (compiled in place from /root/reference, see oracle/Makefile) exactly the way
// dropest.cpp:239-254 does: construct strategies -> CellsDataContainer -> add_record per read -> set_initialized ->
// merge_and_filter, then dumps a canonical view of the container (DGEO0001, oracle/common/dge_io.h).
// Matrix assembly restates Estimation/ResultsPrinter.cpp:334-396 (that file needs Rcpp/Eigen and cannot be compiled here).
#include <Estimation/CellsDataContainer.h>
#include <Estimation/Merge/DummyMergeStrategy.h>
#include <Estimation/Merge/MergeAllMergeStrategy.h>
#include <Estimation/Merge/PoissonRealBarcodesMergeStrategy.h>
#include <Estimation/Merge/PoissonSimpleMergeStrategy.h>
#include <Estimation/Merge/RealBarcodesMergeStrategy.h>
#include <Estimation/Merge/SimpleMergeStrategy.h>
#include <Estimation/Merge/BarcodesParsing/ConstLengthBarcodesParser.h>
#include <Estimation/Merge/BarcodesParsing/InDropBarcodesParser.h>
#include <Estimation/Merge/UMIs/MergeUMIsStrategyDirectional.h>
#include <Estimation/Merge/UMIs/MergeUMIsStrategySimple.h>
#include <Tools/Logs.h>

#include "../common/dge_io.h"

#include <chrono>
#include <iostream>

using namespace Estimation;
using Mark = UMI::Mark;

namespace Tools
{
	// Replaces Tools/Logs.cpp (Boost.Log).  Signatures from Tools/Logs.h:15-20.
	void init_log(bool, bool, const std::string &, const std::string &) {}
	void init_test_logs(boost::log::trivial::severity_level) {}
	void trace_time(const std::string &, bool) {}
}

struct Args
{
	std::string in, out, merge = "none", barcodes, barcodes_type = "const", umi_merge = "simple", marks = "eEBA";
	size_t min_genes_before = 10, min_genes_after = 10;
	unsigned max_cb_ed = 2, max_umi_ed = 1;
	double min_frac = 0.2, max_merge_prob = 1e-4, max_real_merge_prob = 1e-7, umi_mult = 2;
	int max_cells = -1;
	bool reads_output = false, dump_umis = false, dump_rpupc = false, stream = false, init_only = false;
	size_t limit = 0;
};

static Args parse_args(int argc, char **argv)
{
	Args a;
	for (int i = 1; i < argc; ++i)
	{
		std::string k = argv[i];
		auto next = [&]() -> std::string {
			if (i + 1 >= argc) throw std::runtime_error("missing value for " + k);
			return argv[++i];
		};
		if (k == "--in") a.in = next();
		else if (k == "--out") a.out = next();
		else if (k == "--merge") a.merge = next();
		else if (k == "--barcodes") a.barcodes = next();
		else if (k == "--barcodes-type") a.barcodes_type = next();
		else if (k == "--umi-merge") a.umi_merge = next();
		else if (k == "--marks") a.marks = next();
		else if (k == "--min-genes-before") a.min_genes_before = std::stoul(next());
		else if (k == "--min-genes-after") a.min_genes_after = std::stoul(next());
		else if (k == "--max-cb-ed") a.max_cb_ed = unsigned(std::stoul(next()));
		else if (k == "--max-umi-ed") a.max_umi_ed = unsigned(std::stoul(next()));
		else if (k == "--min-frac") a.min_frac = std::stod(next());
		else if (k == "--max-merge-prob") a.max_merge_prob = std::stod(next());
		else if (k == "--max-real-merge-prob") a.max_real_merge_prob = std::stod(next());
		else if (k == "--umi-mult") a.umi_mult = std::stod(next());
		else if (k == "--max-cells") a.max_cells = std::stoi(next());
		else if (k == "--limit") a.limit = std::stoul(next());
		else if (k == "--reads-output") a.reads_output = true;
		else if (k == "--dump-umis") a.dump_umis = true;
		else if (k == "--dump-rpupc") a.dump_rpupc = true;
		else if (k == "--stream") a.stream = true;
		else if (k == "--init-only") a.init_only = true;
		else throw std::runtime_error("unknown argument " + k);
	}
	if (a.in.empty() || a.out.empty()) throw std::runtime_error("usage: dropest_ref --in reads.{bin,tsv} --out out.dgeo [options]");
	return a;
}

// Strategy selection follows Estimation/Merge/MergeStrategyFactory.cpp:61-126 (that file needs boost::property_tree).
static std::shared_ptr<Merge::BarcodesParsing::BarcodesParser> make_parser(const Args &a)
{
	using namespace Merge::BarcodesParsing;
	if (a.barcodes_type == "indrop") return std::make_shared<InDropBarcodesParser>(a.barcodes);
	if (a.barcodes_type == "const") return std::make_shared<ConstLengthBarcodesParser>(a.barcodes);
	throw std::runtime_error("Unexpected barcodes type: " + a.barcodes_type);
}

static std::shared_ptr<Merge::MergeStrategyAbstract> make_cb_strategy(const Args &a)
{
	using namespace Merge;
	if (a.merge == "none") return std::make_shared<DummyMergeStrategy>(a.min_genes_before, a.min_genes_after);
	if (a.merge == "all") return std::make_shared<MergeAllMergeStrategy>(a.min_genes_before, a.min_genes_after, a.max_cb_ed);
	if (a.merge == "simple") return std::make_shared<SimpleMergeStrategy>(a.min_genes_before, a.min_genes_after, a.max_cb_ed, a.min_frac);
	if (a.merge == "real") return std::make_shared<RealBarcodesMergeStrategy>(make_parser(a), a.min_genes_before, a.min_genes_after, a.max_cb_ed, a.min_frac);
	PoissonTargetEstimator est(a.max_merge_prob, a.max_real_merge_prob);
	if (a.merge == "poisson_simple") return std::make_shared<PoissonSimpleMergeStrategy>(est, a.min_genes_before, a.min_genes_after, a.max_cb_ed);
	if (a.merge == "poisson_real") return std::make_shared<PoissonRealBarcodesMergeStrategy>(est, make_parser(a), a.min_genes_before, a.min_genes_after, a.max_cb_ed);
	throw std::runtime_error("unknown merge type " + a.merge);
}

struct TextRead { std::string cb, umi, gene, chr, qual; int mark; };

static std::vector<TextRead> read_tsv(const std::string &fname)
{
	std::ifstream f(fname);
	if (!f) throw std::runtime_error("can't open " + fname);
	std::vector<TextRead> reads;
	std::string line;
	while (std::getline(f, line))
	{
		if (line.empty() || line[0] == '#') continue;
		std::vector<std::string> t;
		size_t s = 0;
		while (true)
		{
			size_t e = line.find('\t', s);
			t.push_back(line.substr(s, e == std::string::npos ? e : e - s));
			if (e == std::string::npos) break;
			s = e + 1;
		}
		if (t.size() < 5) throw std::runtime_error("bad tsv line: " + line);
		TextRead r{t[0], t[1], t[2] == "-" ? "" : t[2], t[3] == "-" ? "" : t[3], t.size() > 5 ? t[5] : "", std::stoi(t[4])};
		reads.push_back(r);
	}
	return reads;
}

static Mark mark_of(int bits)
{
	Mark m;
	if (bits & 1) m.add(Mark::HAS_NOT_ANNOTATED);
	if (bits & 2) m.add(Mark::HAS_EXONS);
	if (bits & 4) m.add(Mark::HAS_INTRONS);
	return m;
}

static int mark_bits(const Mark &m)
{
	return (m.check(Mark::HAS_NOT_ANNOTATED) ? 1 : 0) | (m.check(Mark::HAS_EXONS) ? 2 : 0) | (m.check(Mark::HAS_INTRONS) ? 4 : 0);
}

static double now_s()
{
	return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char **argv)
{
	try
	{
		Args a = parse_args(argc, argv);
		auto cb_strat = make_cb_strategy(a);
		std::shared_ptr<Merge::UMIs::MergeUMIsStrategyAbstract> umi_strat;
		if (a.umi_merge == "directional") umi_strat = std::make_shared<Merge::UMIs::MergeUMIsStrategyDirectional>(a.umi_mult, a.max_umi_ed);
		else umi_strat = std::make_shared<Merge::UMIs::MergeUMIsStrategySimple>(a.max_umi_ed);

		auto marks = Mark::get_by_code(a.marks);
		CellsDataContainer container(cb_strat, umi_strat, marks, false, a.max_cells);

		double t_fill = 0, t_decode = 0;
		size_t n_reads = 0;
		bool is_tsv = a.in.size() > 4 && a.in.substr(a.in.size() - 4) == ".tsv";
		if (is_tsv)
		{
			auto reads = read_tsv(a.in);
			std::vector<ReadInfo> infos;
			for (auto const &r : reads)
				infos.emplace_back(Tools::ReadParameters(r.cb, r.umi, "", r.qual), r.gene, r.chr, mark_of(r.mark));
			double t0 = now_s();
			for (auto const &ri : infos) container.add_record(ri);
			t_fill = now_s() - t0;
			n_reads = infos.size();
		}
		else
		{
			dge_io::ReadStream s = dge_io::read_packed(a.in);
			n_reads = a.limit ? std::min(a.limit, s.recs.size()) : s.recs.size();
			std::vector<std::string> gnames(s.n_genes);
			for (uint32_t g = 0; g < s.n_genes; ++g) gnames[g] = s.gene_name(g);
			auto decode = [&](size_t i) {
				const dge_io::Record16 &r = s.recs[i];
				uint32_t gid = r.gene & 0xFFFFFFu;
				return ReadInfo(Tools::ReadParameters(s.cb_of(r), s.umi_of(r), "", ""),
				                gid == dge_io::NO_GENE ? std::string() : gnames.at(gid), s.chr_name(i), mark_of((r.gene >> 24) & 7));
			};
			if (a.stream)
			{
				double t0 = now_s();
				for (size_t i = 0; i < n_reads; ++i) container.add_record(decode(i));
				t_fill = now_s() - t0;
			}
			else
			{
				// dropest.cpp materialises one ReadInfo per alignment before add_record; the BASELINE.md plan times
				// the add_record loop over pre-built ReadInfo objects, in chunks to bound memory.
				const size_t chunk = 4u << 20;
				for (size_t start = 0; start < n_reads; start += chunk)
				{
					size_t end = std::min(n_reads, start + chunk);
					double t0 = now_s();
					std::vector<ReadInfo> infos;
					infos.reserve(end - start);
					for (size_t i = start; i < end; ++i) infos.push_back(decode(i));
					double t1 = now_s();
					for (auto const &ri : infos) container.add_record(ri);
					double t2 = now_s();
					t_decode += t1 - t0;
					t_fill += t2 - t1;
				}
			}
		}

		double t0 = now_s();
		container.set_initialized();
		double t_init = now_s() - t0;

		std::vector<int64_t> filtered_pre(container.filtered_cells().begin(), container.filtered_cells().end());

		t0 = now_s();
		if (!a.init_only) container.merge_and_filter();
		double t_merge = now_s() - t0;

		dge_io::Writer w;
		const size_t n_cells = container.total_cells_number();
		std::vector<std::string> barcodes(n_cells);
		std::vector<uint8_t> flags(n_cells);
		std::vector<int32_t> n_genes(n_cells), umis_stat(n_cells), reads_stat(n_cells);
		std::vector<int64_t> req_genes(n_cells), req_umis(n_cells);
		for (size_t i = 0; i < n_cells; ++i)
		{
			auto const &c = container.cell(i);
			barcodes[i] = c.barcode();
			flags[i] = uint8_t((c.is_real() ? 1 : 0) | (c.is_merged() ? 2 : 0) | (c.is_excluded() ? 4 : 0));
			n_genes[i] = int32_t(c.size());
			umis_stat[i] = c.stats().get(Stats::TOTAL_UMIS_PER_CB);
			reads_stat[i] = c.stats().get(Stats::TOTAL_READS_PER_CB);
			req_genes[i] = int64_t(c.requested_genes_num());
			req_umis[i] = int64_t(c.requested_umis_num());
		}
		w.add_strings("cell_barcodes", barcodes);
		w.add("cell_flags", dge_io::U8, flags);
		w.add("cell_n_genes", dge_io::I32, n_genes);
		w.add("cell_umis_stat", dge_io::I32, umis_stat);
		w.add("cell_reads_stat", dge_io::I32, reads_stat);
		w.add("cell_req_genes", dge_io::I64, req_genes);
		w.add("cell_req_umis", dge_io::I64, req_umis);
		w.add("filtered_pre_merge", dge_io::I64, filtered_pre);
		w.add("filtered_cells", dge_io::I64, std::vector<int64_t>(container.filtered_cells().begin(), container.filtered_cells().end()));
		w.add("merge_targets", dge_io::I64, std::vector<int64_t>(container.merge_targets().begin(), container.merge_targets().end()));
		w.add_strings("gene_names", container.gene_indexer().values());

		// cm: ResultsPrinter.cpp:334-361.  Canonical triplets (col, gene_indexer id, value) + the reference's own row order.
		{
			std::vector<int64_t> col, gene, val;
			std::vector<std::string> row_names;
			std::unordered_map<std::string, size_t> gene_ids;
			for (size_t column_num = 0; column_num < container.filtered_cells().size(); ++column_num)
			{
				auto const &cell = container.cell(container.filtered_cells()[column_num]);
				for (auto const &g : cell.genes())
				{
					size_t v = g.second.number_of_requested_umis(container.gene_match_level(), a.reads_output);
					if (v == 0) continue;
					col.push_back(int64_t(column_num)); gene.push_back(int64_t(g.first)); val.push_back(int64_t(v));
				}
				for (auto const &upg : cell.requested_umis_per_gene(container.gene_match_level(), a.reads_output))
				{
					auto it = gene_ids.emplace(upg.first, gene_ids.size());
					if (it.second) row_names.push_back(upg.first);
				}
			}
			w.add("cm_col", dge_io::I64, col); w.add("cm_gene", dge_io::I64, gene); w.add("cm_val", dge_io::I64, val);
			w.add_strings("cm_row_names", row_names);
		}
		// cm_raw: ResultsPrinter.cpp:363-396.
		{
			std::vector<int64_t> col, gene, val, cells;
			std::vector<std::string> row_names;
			std::unordered_map<std::string, size_t> gene_ids;
			size_t column_num = 0;
			for (size_t cell_id = 0; cell_id < n_cells; ++cell_id)
			{
				auto const &cell = container.cell(cell_id);
				if (!cell.is_real()) continue;
				cells.push_back(int64_t(cell_id));
				for (auto const &g : cell.genes())
				{
					auto const &name = container.gene_indexer().get_value(g.first);
					auto it = gene_ids.emplace(name, gene_ids.size());
					if (it.second) row_names.push_back(name);
					col.push_back(int64_t(column_num)); gene.push_back(int64_t(g.first));
					val.push_back(int64_t(g.second.number_of_umis(a.reads_output)));
				}
				column_num++;
			}
			w.add("cm_raw_cells", dge_io::I64, cells);
			w.add("cm_raw_col", dge_io::I64, col); w.add("cm_raw_gene", dge_io::I64, gene); w.add("cm_raw_val", dge_io::I64, val);
			w.add_strings("cm_raw_row_names", row_names);
		}

		// per-chromosome statistics of the real cells: CellsDataContainer::get_stat_by_real_cells(CellChrStatType, ...) exactly as
		// ResultsPrinter::get_reads_per_chr_per_cell_info calls it (ResultsPrinter.cpp:140-166); counts are cell-major
		{
			const char *names[3] = {"chr_exon", "chr_intron", "chr_intergenic"};
			const Stats::CellChrStatType types[3] = {Stats::EXON_READS_PER_CHR_PER_CELL, Stats::INTRON_READS_PER_CHR_PER_CELL, Stats::INTERGENIC_READS_PER_CHR_PER_CELL};
			for (int t = 0; t < 3; ++t)
			{
				std::vector<std::string> cells, chrs;
				std::vector<int> counts;
				container.get_stat_by_real_cells(types[t], cells, chrs, counts);
				w.add_strings(std::string(names[t]) + "_cells", cells);
				w.add_strings(std::string(names[t]) + "_chrs", chrs);
				w.add(std::string(names[t]) + "_counts", dge_io::I32, std::vector<int32_t>(counts.begin(), counts.end()));
			}
		}

		if (a.dump_umis)
		{
			std::vector<int64_t> ucell, ugene, ucount;
			std::vector<uint8_t> umark;
			std::vector<std::string> useq;
			for (size_t i = 0; i < n_cells; ++i)
				for (auto const &g : container.cell(i).genes())
					for (auto const &u : g.second.umis())
					{
						ucell.push_back(int64_t(i)); ugene.push_back(int64_t(g.first)); ucount.push_back(int64_t(u.second.read_count()));
						umark.push_back(uint8_t(mark_bits(u.second.mark())));
						useq.push_back(container.umi_indexer().get_value(u.first));
					}
			w.add("umi_cell", dge_io::I64, ucell); w.add("umi_gene", dge_io::I64, ugene); w.add("umi_count", dge_io::I64, ucount);
			w.add("umi_mark", dge_io::U8, umark);
			w.add_strings("umi_seq", useq);
		}

		if (a.dump_rpupc)
		{
			// reads_per_umi_per_cell: the walk of ResultsPrinter::get_reads_per_umi_per_cell (ResultsPrinter.cpp:261-314; that file needs Rcpp) with
			// the container's own accessors -- filtered cells, requested_reads_per_umi_per_gene, UMI::mean_quality
			StringIndexer cell_indexer, gene_indexer;
			std::vector<int64_t> e_cell, e_gene, u_entry, u_reads, u_qlen;
			std::vector<std::string> u_seq;
			std::vector<double> u_q;
			for (auto container_cell_id : container.filtered_cells())
			{
				auto const &cur_cell = container.cell(container_cell_id);
				unsigned cell_id = cell_indexer.add(cur_cell.barcode());
				for (auto const &gene_rpus : cur_cell.requested_reads_per_umi_per_gene(container.gene_match_level()))
				{
					unsigned gene_id = gene_indexer.add(gene_rpus.first);
					for (auto const &umi_reads : gene_rpus.second)
					{
						auto mean_quality = cur_cell.at(gene_rpus.first).at(umi_reads.first).mean_quality();
						u_entry.push_back(int64_t(e_cell.size()));
						u_seq.push_back(umi_reads.first);
						u_reads.push_back(int64_t((unsigned) umi_reads.second));
						u_qlen.push_back(int64_t(mean_quality.size()));
						u_q.insert(u_q.end(), mean_quality.begin(), mean_quality.end());
					}
					e_cell.push_back(cell_id);
					e_gene.push_back(gene_id);
				}
			}
			w.add_strings("rp_cells", cell_indexer.values());
			w.add_strings("rp_genes", gene_indexer.values());
			w.add("rp_cell_indexes", dge_io::I64, e_cell); w.add("rp_gene_indexes", dge_io::I64, e_gene);
			w.add("rp_umi_entry", dge_io::I64, u_entry); w.add("rp_umi_reads", dge_io::I64, u_reads); w.add("rp_umi_qlen", dge_io::I64, u_qlen);
			w.add_strings("rp_umi_seq", u_seq);
			w.add("rp_umi_quality", dge_io::F64, u_q);
		}

		w.add_scalar_i64("n_reads", int64_t(n_reads));
		w.add_scalar_i64("n_cells", int64_t(n_cells));
		w.add_scalar_i64("real_cells_number", int64_t(container.real_cells_number()));
		w.add_scalar_i64("intergenic_reads", int64_t(container.intergenic_reads_num()));
		w.add_scalar_i64("has_exon_reads", int64_t(container.has_exon_reads_num()));
		w.add_scalar_i64("has_intron_reads", int64_t(container.has_intron_reads_num()));
		w.add_scalar_i64("has_not_annotated_reads", int64_t(container.has_not_annotated_reads_num()));
		w.add_scalar_f64("t_decode_s", t_decode);
		w.add_scalar_f64("t_fill_s", t_fill);
		w.add_scalar_f64("t_init_s", t_init);
		w.add_scalar_f64("t_merge_s", t_merge);
		w.write(a.out);

		std::cerr << "dropest_ref: " << n_reads << " reads, " << n_cells << " cells, " << container.real_cells_number() << " real, "
		          << container.filtered_cells().size() << " filtered; fill " << t_fill << " s, init " << t_init << " s, merge " << t_merge << " s\n";
		return 0;
	}
	catch (std::exception &e)
	{
		std::cerr << "dropest_ref: ERROR: " << e.what() << "\n";
		return 1;
	}
}
