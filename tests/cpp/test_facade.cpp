// tests/cpp/test_facade.cpp -- the reference's own container tests (Tests/TestEstimation.cpp), re-typed against the host-side
// mirror dropest_b200/host/Estimation.h.  Same calls, same expectations; Boost.Test is absent so checks are plain macros.
// Needs a CUDA device (run by tests/test_facade.py under -m gpu).  usage: test_facade <whitelist test_est> <out dir>
#include "../../dropest_b200/host/Estimation.h"

#include <iostream>

using namespace Estimation;
using Mark = UMI::Mark;

static int failures = 0;
#define CHECK(cond) do { if (!(cond)) { std::cerr << "CHECK failed: " #cond " (line " << __LINE__ << ")\n"; ++failures; } } while (0)
#define CHECK_EQUAL(a, b) do { auto _a = (a); auto _b = (b); if (!(_a == _b)) { std::cerr << "CHECK_EQUAL failed: " #a " == " #b " (" << _a << " vs " << _b << ", line " << __LINE__ << ")\n"; ++failures; } } while (0)
#define CHECK_THROW(expr, ex) do { bool _t = false; try { expr; } catch (ex &) { _t = true; } if (!_t) { std::cerr << "CHECK_THROW failed: " #expr " (line " << __LINE__ << ")\n"; ++failures; } } while (0)

static ReadInfo read_info(const std::string &cell_barcode, const std::string &umi, const std::string &gene,
                          const std::string &chr_name = "", const Mark &mark = Mark(Mark::HAS_EXONS))
{
	return ReadInfo(Tools::ReadParameters(cell_barcode, umi, "", umi), gene, chr_name, mark);
}

int main(int argc, char **argv)
{
	if (argc < 3) { std::cerr << "usage: test_facade <whitelist> <out dir>\n"; return 2; }
	const std::string whitelist = argv[1], out_dir = argv[2];
	try
	{
		// Fixture, Tests/TestEstimation.cpp:33-80
		auto barcodes_parser = std::shared_ptr<Merge::BarcodesParsing::BarcodesParser>(new Merge::BarcodesParsing::InDropBarcodesParser(whitelist));
		auto real_cb_strat = std::make_shared<Merge::RealBarcodesMergeStrategy>(barcodes_parser, 0, 0, 7, 0);
		auto umi_merge_strat = std::make_shared<Merge::UMIs::MergeUMIsStrategySimple>(1);
		auto any_mark = Mark::get_by_code(Mark::DEFAULT_CODE);
		CellsDataContainer container_full(real_cb_strat, umi_merge_strat, any_mark, false, -1, 0, 64);
		static const char *reads[][3] = {
			{"AAATTAGGTCCA", "AAACCT", "Gene1"}, {"AAATTAGGTCCA", "CCCCCT", "Gene2"}, {"AAATTAGGTCCA", "ACCCCT", "Gene3"},
			{"AAATTAGGTCCA", "ACCCCT", "Gene4"}, {"AAATTAGGTCCC", "CAACCT", "Gene1"}, {"AAATTAGGTCCC", "CAACCT", "Gene10"},
			{"AAATTAGGTCCC", "CAACCT", "Gene20"}, {"AAATTAGGTCCG", "CAACCT", "Gene1"}, {"AAATTAGGTCGG", "AAACCT", "Gene1"},
			{"AAATTAGGTCGG", "CCCCCT", "Gene2"}, {"CCCTTAGGTCCA", "CCATTC", "Gene3"}, {"CCCTTAGGTCCA", "CCCCCT", "Gene2"},
			{"CCCTTAGGTCCA", "ACCCCT", "Gene3"}, {"CAATTAGGTCCG", "CAACCT", "Gene1"}, {"CAATTAGGTCCG", "AAACCT", "Gene1"},
			{"CAATTAGGTCCG", "CCCCCT", "Gene2"}, {"AAAAAAAAAAAA", "CCCCCT", "Gene2"}};
		// chromosome names (not part of the reference fixture, which leaves them empty): one per gene family, for the per-chromosome Stats tables
		auto chr_of = [](const std::string &gene) { return gene == "Gene1" ? std::string("chr1") : gene == "Gene2" ? std::string("chr2") : std::string("chr3"); };
		for (auto const &r : reads) container_full.add_record(read_info(r[0], r[1], r[2], chr_of(r[2])));
		CHECK_THROW(container_full.merge_and_filter(), std::runtime_error); // "You must initialize container"
		container_full.set_initialized();
		CHECK_THROW(container_full.add_record(read_info("AAATTAGGTCCA", "AAACCT", "Gene1")), std::runtime_error);
		CHECK_THROW(container_full.set_initialized(), std::runtime_error);

		// testMergeByRealBarcodes, Tests/TestEstimation.cpp:237-280
		container_full.merge_and_filter();
		CHECK_EQUAL(container_full.total_cells_number(), size_t(7));
		CHECK_EQUAL(container_full.filtered_cells().size(), size_t(2));
		auto &cell0 = container_full.cell(container_full.filtered_cells()[0]);
		auto &cell1 = container_full.cell(container_full.filtered_cells()[1]);
		CHECK_EQUAL(cell0.size(), size_t(3));
		CHECK_EQUAL(cell1.size(), size_t(4));
		CHECK_EQUAL(cell0.at("Gene1").size(), size_t(1));
		CHECK_EQUAL(cell0.at("Gene1").at("CAACCT").read_count(), size_t(2));
		CHECK_EQUAL(cell1.at("Gene1").size(), size_t(2));
		CHECK_EQUAL(cell1.at("Gene1").at("AAACCT").read_count(), size_t(3));
		CHECK_EQUAL(cell1.at("Gene2").size(), size_t(1));
		CHECK_EQUAL(cell1.at("Gene2").at("CCCCCT").read_count(), size_t(4));
		CHECK_EQUAL(cell1.at("Gene3").size(), size_t(2));
		CHECK_EQUAL(cell1.at("Gene3").at("ACCCCT").read_count(), size_t(2));
		CHECK_EQUAL(cell1.at("Gene3").at("CCATTC").read_count(), size_t(1));
		CHECK(!container_full.cell(0).is_merged());
		CHECK(!container_full.cell(1).is_merged());
		CHECK(container_full.cell(2).is_merged());
		CHECK(container_full.cell(3).is_merged());
		CHECK(container_full.cell(4).is_merged());
		CHECK(container_full.cell(5).is_merged());
		CHECK(!container_full.cell(6).is_merged());
		size_t excluded_num = 0;
		for (size_t i = 0; i < container_full.total_cells_number(); ++i) excluded_num += container_full.cell(i).is_excluded();
		CHECK_EQUAL(excluded_num, size_t(1));
		CHECK_EQUAL(container_full.cell_id_by_cb("CCCTTAGGTCCA"), size_t(4));
		CHECK_THROW(container_full.cell_id_by_cb("GGGGGGGGGGGG"), std::out_of_range);
		CHECK_EQUAL(container_full.merge_targets()[5], size_t(0));
		CHECK_EQUAL(container_full.gene_indexer().get_value(0), std::string("Gene1"));
		CHECK_EQUAL(container_full.merge_type(), std::string("Real CBs"));

		// testUmiExclusion, Tests/TestEstimation.cpp:369-397 (Mark accumulation + exact-match query semantics)
		{
			CellsDataContainer container(real_cb_strat, umi_merge_strat, Mark::get_by_code("e"), false, -1, 0, 64);
			container.add_record(read_info("AAATTAGGTCCA", "AAACCT", "Gene1"));
			container.add_record(read_info("AAATTAGGTCCA", "CCCCCT", "Gene2"));
			container.add_record(read_info("AAATTAGGTCCA", "ACCCCT", "Gene3"));
			container.add_record(read_info("AAATTAGGTCCA", "ACCCCT", "Gene4"));
			container.add_record(read_info("AAATTAGGTCCA", "TTTTTT", "Gene3", "chr1", Mark(Mark::HAS_NOT_ANNOTATED)));
			container.add_record(read_info("AAATTAGGTCCA", "ACCCCT", "Gene4", "chr1", Mark(Mark::HAS_NOT_ANNOTATED)));
			container.set_initialized();
			container.merge_and_filter();
			CHECK(container.cell(0).at("Gene3").at("TTTTTT").mark().check(Mark::HAS_NOT_ANNOTATED));
			CHECK(container.cell(0).at("Gene4").at("ACCCCT").mark().check(Mark::HAS_NOT_ANNOTATED));
			auto requested = container.cell(0).requested_umis_per_gene(container.gene_match_level(), true);
			CHECK_EQUAL(requested.at("Gene3"), size_t(1));
			CHECK_THROW(requested.at("Gene4"), std::out_of_range);
			CHECK_EQUAL(container.cell(0).at("Gene4").at("ACCCCT").read_count(), size_t(2));
		}

		// testUMIMergeStrategyDirectional, Tests/TestEstimation.cpp:588-608: the per-segment entry point ...
		{
			using Strat = Merge::UMIs::MergeUMIsStrategyDirectional;
			auto directional = std::make_shared<Strat>();
			Strat::umi_vec_t umis;
			umis.emplace_back("AAA", 2); umis.emplace_back("AAC", 5); umis.emplace_back("AAT", 6);
			umis.emplace_back("AGT", 20); umis.emplace_back("CCC", 10); umis.emplace_back("TCC", 20);
			auto targets = directional->find_targets(umis);
			CHECK_EQUAL(targets.size(), size_t(3));
			CHECK_EQUAL(targets.at("AAA"), std::string("AGT"));
			CHECK_EQUAL(targets.at("AAT"), std::string("AGT"));
			CHECK_EQUAL(targets.at("CCC"), std::string("TCC"));
			// ... and the same six UMIs through the container (device path): AAA+AAT -> AGT (28 reads), CCC -> TCC (30), AAC stays
			CellsDataContainer container(std::make_shared<Merge::DummyMergeStrategy>(0, 0), directional, any_mark, false, -1, 0, 64);
			static const struct { const char *umi; int reads; } seg[] = {{"AAA", 2}, {"AAC", 5}, {"AAT", 6}, {"AGT", 20}, {"CCC", 10}, {"TCC", 20}};
			for (auto const &u : seg)
				for (int k = 0; k < u.reads; ++k) container.add_record(read_info("AAATTAGGTCCA", u.umi, "Gene1"));
			container.set_initialized();
			container.merge_and_filter();
			auto const &gene = container.cell(0).at("Gene1");
			CHECK_EQUAL(gene.size(), size_t(3));
			CHECK_EQUAL(gene.at("AGT").read_count(), size_t(28));
			CHECK_EQUAL(gene.at("TCC").read_count(), size_t(30));
			CHECK_EQUAL(gene.at("AAC").read_count(), size_t(5));
			CHECK_EQUAL(container.cell(0).umis_number(), size_t(3));
		}

		// MergeStrategyFactory::get_cb_strat (MergeStrategyFactory.cpp:61-89) + SimpleMergeStrategy through the container
		{
			Merge::MergeStrategyFactory factory;
			factory.min_genes_before_merge = 0; factory.min_genes_after_merge = 0;
			CHECK_EQUAL(factory.get_cb_strat(false, false)->merge_type(), std::string("No"));
			CHECK_EQUAL(factory.get_cb_strat(true, false)->merge_type(), std::string("Simple"));   // no barcodes file
			factory.merge_type = "all";
			CHECK_EQUAL(factory.get_cb_strat(true, false)->merge_type(), std::string("Merge all"));
			factory.merge_type = "";
			CellsDataContainer container(factory.get_cb_strat(true, false), umi_merge_strat, any_mark, false, -1, 0, 64);
			static const char *big[][2] = {{"AAAAAA", "Gene1"}, {"AAAAAC", "Gene1"}, {"AAAAAG", "Gene2"}, {"AAAAAT", "Gene3"}, {"AAAACA", "Gene4"}, {"AAAACC", "Gene5"}};
			static const char *small_[][2] = {{"AAAAAA", "Gene1"}, {"AAAAAG", "Gene2"}, {"AAAAAT", "Gene3"}, {"TTTTTT", "Gene5"}};
			for (auto const &r : big) container.add_record(read_info("AAATTAGGTCCA", r[0], r[1]));
			for (auto const &r : small_) container.add_record(read_info("AAATTAGGTCCC", r[0], r[1]));   // edit distance 1, shares 3 of its 4 UMI-genes
			for (auto const &r : small_) container.add_record(read_info("CCCTTAGGTCCC", r[0], r[1]));   // same content, edit distance 4: must stay
			container.set_initialized();
			container.merge_and_filter();
			CHECK_EQUAL(container.merge_type(), std::string("Simple"));
			CHECK_EQUAL(container.merge_targets().at(1), size_t(0));
			CHECK_EQUAL(container.merge_targets().at(2), size_t(2));
			CHECK(container.cell(1).is_merged());
			CHECK_EQUAL(container.cell(0).at("Gene1").at("AAAAAA").read_count(), size_t(2));
			CHECK_EQUAL(container.cell(0).size(), size_t(5));
		}

		// testUMIMergeStrategySimple, Tests/TestEstimation.cpp:505-540: UMIs containing N travel as indices into the container's N-string list
		// and are repaired after the barcode merge (nearest N-free UMI within max_umi_merge_edit_distance, else random bases)
		{
			CellsDataContainer container(real_cb_strat, umi_merge_strat, any_mark);
			static const char *r1[] = {"AAACCT", "AAACCT", "AAACCG", "AAACCN", "CCCCCT", "ACCCCT"};
			static const char *r2[] = {"TTTTTT", "TTTNNG", "TTGNNG", "ACCCCT", "NNNNNN"};
			for (auto u : r1) container.add_record(read_info("AAATTAGGTCCA", u, "Gene1"));
			for (auto u : r2) container.add_record(read_info("AAATTAGGTCCA", u, "Gene2"));
			container.add_record(read_info("AAATTAGGNCCA", "ACGTAC", "Gene3")); // a barcode with N is a cell of its own (N matches any base in the whitelist walk)
			container.set_initialized();
			CHECK_EQUAL(container.total_cells_number(), size_t(2));
			CHECK_EQUAL(container.cell(1).barcode(), std::string("AAATTAGGNCCA"));
			CHECK_EQUAL(container.cell(0).at("Gene1").size(), size_t(5));          // before the repair the N-UMI is a UMI of its own
			CHECK_EQUAL(container.cell(0).at("Gene1").at("AAACCN").read_count(), size_t(1));
			container.merge_and_filter();
			CHECK_EQUAL(container.cell(0).at("Gene1").size(), size_t(4));
			CHECK_EQUAL(container.cell(0).at("Gene2").size(), size_t(3));
			CHECK_EQUAL(container.cell(0).at("Gene1").at("AAACCT").read_count(), size_t(3)); // AAACCN -> AAACCT (distance 0 to both AAACCT and AAACCG, more reads wins)
			CHECK_EQUAL(container.cell(0).at("Gene1").at("AAACCG").read_count(), size_t(1));
			CHECK_EQUAL(container.cell(0).at("Gene1").at("CCCCCT").read_count(), size_t(1));
			CHECK_EQUAL(container.cell(0).at("Gene1").at("ACCCCT").read_count(), size_t(1));
			CHECK(container.cell(0).at("Gene2").has("TTTTTT"));
			CHECK(container.cell(0).at("Gene2").has("ACCCCT"));
			for (auto const &umi : container.cell(0).at("Gene2").umis())
				CHECK_EQUAL(container.umi_indexer().get_value(umi.first).find('N'), std::string::npos);
		}

		// variable-length barcodes (inDrop v1 / v2) are cells of their own, like barcodes with N; a UMI of another length is counted and skipped
		// (never fatal) and keeps its stream position
		{
			CellsDataContainer container(real_cb_strat, umi_merge_strat, any_mark);
			container.add_record(read_info("AAATTAGGTCCA", "AAACCT", "Gene1"));
			container.add_record(read_info("AAATTAGGTCC", "AAACCT", "Gene1"));    // 11-base barcode: kept
			container.add_record(read_info("CCCTTAGGTCCA", "AAACC", "Gene2"));    // 5-base UMI: skipped
			container.add_record(read_info("CCCTTAGGTCCA", "AAACCT", "Gene2"));
			container.add_record(read_info("AAATTAGGTCC", "AAACCG", "Gene1"));
			container.set_initialized();
			CHECK_EQUAL(container.skipped_length_reads(), uint64_t(1));
			CHECK_EQUAL(container.total_cells_number(), size_t(3));
			CHECK_EQUAL(container.cell(1).barcode(), std::string("AAATTAGGTCC"));
			CHECK_EQUAL(container.cell(1).at("Gene1").size(), size_t(2));
			CHECK_EQUAL(container.cell(2).barcode(), std::string("CCCTTAGGTCCA"));
			CHECK_EQUAL(container.cell(2).at("Gene2").at("AAACCT").read_count(), size_t(1));
			CHECK_EQUAL(container.cell_id_by_cb("AAATTAGGTCC"), size_t(1));
		}

		// -u with an N read (MergeUMIsStrategyDirectional.cpp:57-116): the N-UMI matches AAACCT at distance 0 (N is a wildcard) and merges into it
		{
			auto directional = std::make_shared<Merge::UMIs::MergeUMIsStrategyDirectional>(2, 1);
			CellsDataContainer container(real_cb_strat, directional, any_mark);
			container.add_record(read_info("AAATTAGGTCCA", "AAACCT", "Gene1"));
			container.add_record(read_info("AAATTAGGTCCA", "AAACNT", "Gene1"));
			container.add_record(read_info("AAATTAGGTCCA", "AAACCT", "Gene1"));
			container.add_record(read_info("AAATTAGGTCCA", "TTNTTT", "Gene2"));   // alone in its gene: renamed with random bases
			container.set_initialized();
			container.merge_and_filter();
			CHECK_EQUAL(container.skipped_n_reads(), uint64_t(0));
			CHECK_EQUAL(container.cell(0).at("Gene1").size(), size_t(1));
			CHECK_EQUAL(container.cell(0).at("Gene1").at("AAACCT").read_count(), size_t(3));
			CHECK_EQUAL(container.cell(0).at("Gene2").size(), size_t(1));
			for (auto const &umi : container.cell(0).at("Gene2").umis())
				CHECK_EQUAL(container.umi_indexer().get_value(umi.first).find('N'), std::string::npos);
		}

		// -M: MergeStrategyFactory::get_cb_poisson_strat (MergeStrategyFactory.cpp:91-103) + PoissonSimpleMergeStrategy through the container.
		// Same reads as above: the small cell shares 3 UMI-genes with the big one 1 substitution away; with only 7 distinct UMIs in the whole
		// container the expected random overlap is large (lambda ~ 0.6, P[X >= 3] ~ 0.02), so the thresholds are raised for this toy input;
		// the copy 3 substitutions away is outside max_merge_edit_distance and stays.
		{
			Merge::MergeStrategyFactory factory;
			factory.min_genes_before_merge = 0; factory.min_genes_after_merge = 0;
			factory.max_merge_prob = 0.5; factory.max_real_cb_merge_prob = 0.5;
			CHECK_EQUAL(factory.get_cb_strat(true, true)->merge_type(), std::string("Poisson Simple"));
			CellsDataContainer container(factory.get_cb_strat(true, true), umi_merge_strat, any_mark, false, -1, 0, 64);
			static const char *big[][2] = {{"AAAAAA", "Gene1"}, {"AAAAAC", "Gene1"}, {"AAAAAG", "Gene2"}, {"AAAAAT", "Gene3"}, {"AAAACA", "Gene4"}, {"AAAACC", "Gene5"}};
			static const char *small_[][2] = {{"AAAAAA", "Gene1"}, {"AAAAAG", "Gene2"}, {"AAAAAT", "Gene3"}, {"TTTTTT", "Gene5"}};
			for (auto const &r : big) container.add_record(read_info("AAATTAGGTCCA", r[0], r[1]));
			for (auto const &r : small_) container.add_record(read_info("AAATTAGGTCCC", r[0], r[1]));
			for (auto const &r : small_) container.add_record(read_info("CCCTTAGGTCCC", r[0], r[1]));
			container.set_initialized();
			container.merge_and_filter();
			CHECK_EQUAL(container.merge_type(), std::string("Poisson Simple"));
			CHECK_EQUAL(container.merge_targets().at(1), size_t(0));
			CHECK_EQUAL(container.merge_targets().at(2), size_t(2));
			CHECK(container.cell(1).is_merged());
		}

		// testEditDistance, Tests/TestTools.cpp:47-54 ; testReadParams :56-87 (codec part)
		CHECK_EQUAL(Tools::edit_distance("ATTTTC", "ATTTGC"), 1u);
		CHECK_EQUAL(Tools::edit_distance("ATTTTCC", "ATTTGNC"), 1u);
		CHECK_EQUAL(Tools::edit_distance("ATTTTCC", "ATTTGNC", false), 2u);
		CHECK_EQUAL(Tools::edit_distance("ATTTTCC", "ATTTGTC"), 2u);
		CHECK_EQUAL(Tools::edit_distance("ATTTTCC", "ATTTTCC"), 0u);
		auto rp = Tools::ReadParameters::parse_encoded_id("@111!ATTTGC#ATATC");
		CHECK_EQUAL(rp.cell_barcode(), std::string("ATTTGC"));
		CHECK_EQUAL(rp.umi(), std::string("ATATC"));
		CHECK_THROW(Tools::ReadParameters::parse_encoded_id("ATTTG#ATAT"), std::runtime_error);

		// Tools::CollisionsAdjuster (CollisionsAdjuster.cpp:12-49): values observed on the compiled reference (ref_pins.json)
		{
			Tools::CollisionsAdjuster adjuster;
			adjuster.init(std::vector<double>(4096, 1.0 / 4096));
			CHECK_EQUAL(adjuster.estimate_adjusted_gene_expression(100), size_t(101));
			CHECK_EQUAL(adjuster.estimate_adjusted_gene_expression(1000), size_t(1146));
			CHECK_EQUAL(adjuster.estimate_adjusted_gene_expression(3000), size_t(5400));
			CHECK_EQUAL(adjuster.estimate_adjusted_gene_expression(10), size_t(10));
		}

		// umi_distribution (CellsDataContainer.cpp:182-197) over the two filtered cells: AAATTAGGTCCA holds 6 (gene, UMI) entries, AAATTAGGTCCC 3
		{
			auto dist = container_full.umi_distribution();
			size_t total_entries = 0;
			for (auto const &kv : dist) total_entries += kv.second;
			CHECK_EQUAL(total_entries, size_t(9));
			CHECK_EQUAL(dist.at("CAACCT"), size_t(4));   // Gene1 of both cells, Gene10, Gene20
			CHECK_EQUAL(dist.at("ACCCCT"), size_t(2));   // Gene3 and Gene4 of AAATTAGGTCCA
		}

		// get_stat_by_real_cells(CellChrStatType, ...), CellsDataContainer.cpp:292-307: Stats::merge has added the merged cells' counters
		{
			CellsDataContainer::names_t cells, chrs;
			CellsDataContainer::counts_t counts;
			container_full.get_stat_by_real_cells(Stats::EXON_READS_PER_CHR_PER_CELL, cells, chrs, counts);
			CHECK_EQUAL(cells.size(), size_t(2));
			CHECK_EQUAL(chrs.size(), size_t(3));
			CHECK_EQUAL(counts.size(), size_t(6));
			if (cells.size() == 2 && chrs.size() == 3 && counts.size() == 6)
			{
				CHECK_EQUAL(cells[0], std::string("AAATTAGGTCCA"));
				CHECK_EQUAL(cells[1], std::string("AAATTAGGTCCC"));
				for (size_t k = 0; k < 3; ++k)
				{
					CHECK_EQUAL(counts[k], 4);                                   // AAATTAGGTCCA + its three merged barcodes: 4 reads on each chromosome
					CHECK_EQUAL(counts[3 + k], chrs[k] == "chr2" ? 0 : 2);       // AAATTAGGTCCC + AAATTAGGTCCG
				}
			}
			cells.clear(); chrs.clear(); counts.clear();
			container_full.get_stat_by_real_cells(Stats::INTERGENIC_READS_PER_CHR_PER_CELL, cells, chrs, counts);
			CHECK_EQUAL(cells.size() + chrs.size() + counts.size(), size_t(0));
		}

		// ResultsPrinter::save_results (ResultsPrinter.cpp:23-91): files consumed by dropReport / dropestr
		ResultsPrinter printer(true, false);
		printer.save_results(container_full, out_dir + "/cell.counts.rds");
		auto cm = printer.get_count_matrix(container_full, true);
		CHECK_EQUAL(cm.col_names.size(), size_t(2));
		CHECK_EQUAL(cm.col_names[0], std::string("AAATTAGGTCCC"));
		CHECK_EQUAL(cm.col_names[1], std::string("AAATTAGGTCCA"));
		double total = 0;
		for (double v : cm.x) total += v;
		CHECK_EQUAL(total, 3.0 + 6.0); // cell 1: 3 genes x 1 UMI ; merged cell 0: Gene1 2, Gene2 1, Gene3 2, Gene4 1
		// -V: ResultsPrinter::save_intron_exon_matrices (ResultsPrinter.cpp:455-474); every read of the fixture is exonic
		printer.save_intron_exon_matrices(container_full, out_dir + "/cell.counts.rds");
		auto exon = printer.get_count_matrix_filtered(container_full, UMI::Mark::get_by_code("e"));
		auto intron = printer.get_count_matrix_filtered(container_full, UMI::Mark::get_by_code("i"));
		CHECK_EQUAL(exon.x.size(), cm.x.size());
		CHECK_EQUAL(exon.row_names.size(), cm.row_names.size());
		CHECK_EQUAL(intron.x.size(), size_t(0));
		CHECK_EQUAL(intron.col_names.size(), size_t(2));
	}
	catch (std::exception &e)
	{
		std::cerr << "unexpected exception: " << e.what() << "\n";
		return 1;
	}
	std::cout << (failures ? "FAILED" : "OK") << " (" << failures << " failures)\n";
	return failures ? 1 : 0;
}
