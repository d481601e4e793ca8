// oracle/ref_driver/ref_pins.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Re-runs, against the compiled UNMODIFIED reference (oracle/_ref/libdropest_ref.a), the known-answer checks the reference's
// own Boost.Test suites hold for the hot path (Tests/TestEstimation.cpp, Tests/TestTools.cpp:47-54,
// Tests/TestEstimationMergeProbs.cpp), and prints the observed values as JSON.  Boost.Test is absent from this image, so the
// fixtures are re-typed here; white-box access uses the same `friend struct TestEstimator::...` hooks the reference declares.
// tests/golden/make_golden.py stores this program's output as tests/golden/ref_pins.json; tests/test_oracle_pins.py checks
// both the literal expectations written in the reference tests and (oracle port / product) == this output.
#include <Estimation/CellsDataContainer.h>
#include <Estimation/Merge/PoissonRealBarcodesMergeStrategy.h>
#include <Estimation/Merge/PoissonTargetEstimator.h>
#include <Estimation/Merge/RealBarcodesMergeStrategy.h>
#include <Estimation/Merge/BarcodesParsing/ConstLengthBarcodesParser.h>
#include <Estimation/Merge/BarcodesParsing/InDropBarcodesParser.h>
#include <Estimation/Merge/UMIs/MergeUMIsStrategyDirectional.h>
#include <Estimation/Merge/UMIs/MergeUMIsStrategySimple.h>
#include <Tools/CollisionsAdjuster.h>
#include <Tools/IndexedValue.h>
#include <Tools/Logs.h>
#include <Tools/UtilFunctions.h>

#include <iostream>

using namespace Estimation;
using Mark = UMI::Mark;

namespace Tools
{
	void init_log(bool, bool, const std::string &, const std::string &) {}
	void init_test_logs(boost::log::trivial::severity_level) {}
	void trace_time(const std::string &, bool) {}
}

static std::string DATA;

static ReadInfo read_info(const std::string &cb, const std::string &umi, const std::string &gene,
                          const std::string &chr = "", const Mark &mark = Mark(Mark::HAS_EXONS))
{
	return ReadInfo(Tools::ReadParameters(cb, umi, "", umi), gene, chr, mark);
}

template <class T> static std::string jlist(const std::vector<T> &v)
{
	std::ostringstream s;
	s << "[";
	for (size_t i = 0; i < v.size(); ++i) s << (i ? "," : "") << v[i];
	s << "]";
	return s.str();
}

static std::string jstrs(const std::vector<std::string> &v)
{
	std::ostringstream s;
	s << "[";
	for (size_t i = 0; i < v.size(); ++i) s << (i ? "," : "") << '"' << v[i] << '"';
	s << "]";
	return s.str();
}

// Fixture of Tests/TestEstimation.cpp:33-80
struct EstFixture
{
	std::shared_ptr<Merge::RealBarcodesMergeStrategy> real_cb_strat;
	std::shared_ptr<Merge::UMIs::MergeUMIsStrategySimple> umi_merge_strat;
	std::shared_ptr<CellsDataContainer> c;

	EstFixture()
	{
		auto parser = std::shared_ptr<Merge::BarcodesParsing::BarcodesParser>(
			new Merge::BarcodesParsing::InDropBarcodesParser(DATA + "/barcodes/test_est"));
		real_cb_strat = std::make_shared<Merge::RealBarcodesMergeStrategy>(parser, 0, 0, 7, 0);
		umi_merge_strat = std::make_shared<Merge::UMIs::MergeUMIsStrategySimple>(1);
		c = std::make_shared<CellsDataContainer>(real_cb_strat, umi_merge_strat, Mark::get_by_code(Mark::DEFAULT_CODE));
		static const char *reads[][3] = {
			{"AAATTAGGTCCA", "AAACCT", "Gene1"}, {"AAATTAGGTCCA", "CCCCCT", "Gene2"}, {"AAATTAGGTCCA", "ACCCCT", "Gene3"},
			{"AAATTAGGTCCA", "ACCCCT", "Gene4"}, {"AAATTAGGTCCC", "CAACCT", "Gene1"}, {"AAATTAGGTCCC", "CAACCT", "Gene10"},
			{"AAATTAGGTCCC", "CAACCT", "Gene20"}, {"AAATTAGGTCCG", "CAACCT", "Gene1"}, {"AAATTAGGTCGG", "AAACCT", "Gene1"},
			{"AAATTAGGTCGG", "CCCCCT", "Gene2"}, {"CCCTTAGGTCCA", "CCATTC", "Gene3"}, {"CCCTTAGGTCCA", "CCCCCT", "Gene2"},
			{"CCCTTAGGTCCA", "ACCCCT", "Gene3"}, {"CAATTAGGTCCG", "CAACCT", "Gene1"}, {"CAATTAGGTCCG", "AAACCT", "Gene1"},
			{"CAATTAGGTCCG", "CCCCCT", "Gene2"}, {"AAAAAAAAAAAA", "CCCCCT", "Gene2"}};
		for (auto const &r : reads) c->add_record(read_info(r[0], r[1], r[2]));
		c->set_initialized();
	}
};

namespace TestEstimator
{
	struct testBarcodesFile
	{
		static void run()
		{
			EstFixture f;
			auto cbs = f.real_cb_strat->_barcodes_parser->_barcodes;
			std::cout << "\"testBarcodesFile\": {\"part0\": " << jstrs(cbs[0]) << ", \"part1\": " << jstrs(cbs[1]) << "},\n";
		}
	};

	struct testUmigsIntersection
	{
		static void run()
		{
			EstFixture f;
			auto is = [&](const char *a, const char *b) {
				return Merge::RealBarcodesMergeStrategy::get_umigs_intersect_size(f.c->cell(f.c->cell_id_by_cb(a)), f.c->cell(f.c->cell_id_by_cb(b)));
			};
			std::vector<size_t> v{is("AAATTAGGTCCA", "CCCTTAGGTCCA"), is("AAATTAGGTCCC", "AAATTAGGTCCG"), is("AAATTAGGTCCA", "AAATTAGGTCCC")};
			std::cout << "\"testUmigsIntersection\": " << jlist(v) << ",\n";
		}
	};

	struct testFillDistances
	{
		static void run()
		{
			std::vector<std::string> cbs1{"AAT", "AAA", "CCT"};
			Merge::BarcodesParsing::InDropBarcodesParser parser("");
			Merge::BarcodesParsing::InDropBarcodesParser::barcode_parts_list_t barcodes{cbs1, cbs1};
			parser._barcode2_length = 3;
			parser._barcodes = barcodes;
			auto dists = parser.get_distances_to_barcode("ACTACT");
			std::cout << "\"testFillDistances\": {";
			for (int p = 0; p < 2; ++p)
			{
				std::vector<long> vals, inds;
				for (auto const &d : dists[p]) { vals.push_back(d.value); inds.push_back(long(d.index)); }
				std::cout << "\"values" << p << "\": " << jlist(vals) << ", \"index" << p << "\": " << jlist(inds) << (p ? "" : ", ");
			}
			std::cout << "},\n";
		}
	};

	struct testRealNeighboursCbs
	{
		static void run()
		{
			EstFixture f;
			auto names = [&](const char *cb) {
				std::vector<std::string> r;
				for (size_t id : f.real_cb_strat->get_real_neighbour_cbs(*f.c, f.c->cell_id_by_cb(cb))) r.push_back(f.c->cell(id).barcode());
				return r;
			};
			std::cout << "\"testRealNeighboursCbs\": {\"CAATTAGGTCCG\": " << jstrs(names("CAATTAGGTCCG"))
			          << ", \"AAATTAGGTCCC\": " << jstrs(names("AAATTAGGTCCC")) << "},\n";
		}
	};

	struct testRealNeighbours
	{
		static void run()
		{
			EstFixture f;
			std::vector<long> t;
			for (size_t i = 0; i < f.c->total_cells_number(); ++i) t.push_back(f.real_cb_strat->get_merge_target(*f.c, i));
			std::cout << "\"testRealNeighbours\": " << jlist(t) << ",\n";
		}
	};

	struct testConstLengthBarcodeParser
	{
		static void run()
		{
			Merge::BarcodesParsing::ConstLengthBarcodesParser indrop(DATA + "/barcodes/indrop_v3");
			indrop.init();
			Merge::BarcodesParsing::ConstLengthBarcodesParser tenx(DATA + "/barcodes/10x_aug_2016_split");
			tenx.init();
			auto d = tenx.get_distances_to_barcode("GGTGCGTAGCTAAACA");
			std::vector<size_t> il(indrop._barcode_lengths), tl(tenx._barcode_lengths);
			std::vector<size_t> isz{indrop._barcodes[0].size(), indrop._barcodes[1].size()}, tsz{tenx._barcodes[0].size(), tenx._barcodes[1].size()};
			std::cout << "\"testConstLengthBarcodeParser\": {\"indrop_lengths\": " << jlist(il) << ", \"indrop_sizes\": " << jlist(isz)
			          << ", \"tenx_lengths\": " << jlist(tl) << ", \"tenx_sizes\": " << jlist(tsz)
			          << ", \"tenx_min_dists\": [" << d[0][0].value << "," << d[1][0].value << "]"
			          << ", \"tenx_first\": [\"" << tenx._barcodes[0][0] << "\",\"" << tenx._barcodes[1][0] << "\"]"
			          << ", \"split_indrop\": " << jstrs(indrop.split_barcode("TAATGAGCACTAATGA")) << "},\n";
		}
	};

	struct testSplitBarcode {};

	struct testUMIMergeStrategyDirectional
	{
		static void run()
		{
			using Strat = Merge::UMIs::MergeUMIsStrategyDirectional;
			Strat strat;
			Strat::umi_vec_t umis;
			umis.emplace_back("AAA", 2); umis.emplace_back("AAC", 5); umis.emplace_back("AAT", 6);
			umis.emplace_back("AGT", 20); umis.emplace_back("CCC", 10); umis.emplace_back("TCC", 20);
			auto targets = strat.find_targets(umis);
			std::map<std::string, std::string> sorted(targets.begin(), targets.end());
			std::cout << "\"testUMIMergeStrategyDirectional\": {";
			bool first = true;
			for (auto const &t : sorted) { std::cout << (first ? "" : ", ") << '"' << t.first << "\": \"" << t.second << '"'; first = false; }
			std::cout << "},\n";
		}
	};
}

// Tests/TestEstimation.cpp:237-280
static void merge_by_real_barcodes()
{
	EstFixture f;
	f.c->merge_and_filter();
	auto &c = *f.c;
	std::vector<size_t> filt(c.filtered_cells().begin(), c.filtered_cells().end());
	std::vector<size_t> targets(c.merge_targets().begin(), c.merge_targets().end());
	std::vector<int> merged, excluded, sizes, umis_stat;
	for (size_t i = 0; i < c.total_cells_number(); ++i)
	{
		merged.push_back(c.cell(i).is_merged()); excluded.push_back(c.cell(i).is_excluded());
		sizes.push_back(int(c.cell(i).size())); umis_stat.push_back(int(c.cell(i).umis_number()));
	}
	auto &cell0 = c.cell(filt.at(0));
	auto &cell1 = c.cell(filt.at(1));
	std::vector<size_t> rc{cell0.at("Gene1").at("CAACCT").read_count(), cell1.at("Gene1").at("AAACCT").read_count(),
	                       cell1.at("Gene2").at("CCCCCT").read_count(), cell1.at("Gene3").at("ACCCCT").read_count(),
	                       cell1.at("Gene3").at("CCATTC").read_count()};
	std::vector<size_t> gs{cell0.size(), cell1.size(), cell0.at("Gene1").size(), cell1.at("Gene1").size(), cell1.at("Gene2").size(), cell1.at("Gene3").size()};
	std::cout << "\"testMergeByRealBarcodes\": {\"total_cells\": " << c.total_cells_number() << ", \"filtered\": " << jlist(filt)
	          << ", \"merge_targets\": " << jlist(targets) << ", \"merged\": " << jlist(merged) << ", \"excluded\": " << jlist(excluded)
	          << ", \"n_genes\": " << jlist(sizes) << ", \"umis_stat\": " << jlist(umis_stat)
	          << ", \"read_counts\": " << jlist(rc) << ", \"gene_sizes\": " << jlist(gs) << "},\n";
}

// Tests/TestEstimation.cpp:468-540
static void umi_merge_simple()
{
	EstFixture f;
	CellsDataContainer c(f.real_cb_strat, std::shared_ptr<Merge::UMIs::MergeUMIsStrategyAbstract>(f.umi_merge_strat), Mark::get_by_code(Mark::DEFAULT_CODE));
	static const char *reads[][2] = {{"AAACCT", "Gene1"}, {"AAACCT", "Gene1"}, {"AAACCG", "Gene1"}, {"AAACCN", "Gene1"}, {"CCCCCT", "Gene1"},
	                                 {"ACCCCT", "Gene1"}, {"TTTTTT", "Gene2"}, {"TTTNNG", "Gene2"}, {"TTGNNG", "Gene2"}, {"ACCCCT", "Gene2"}, {"NNNNNN", "Gene2"}};
	for (auto const &r : reads) c.add_record(read_info("AAATTAGGTCCA", r[0], r[1]));
	c.set_initialized();
	f.umi_merge_strat->merge(c);
	std::cout << "\"testUMIMergeStrategySimple\": {";
	for (const char *g : {"Gene1", "Gene2"})
	{
		std::map<std::string, size_t> umis;
		for (auto const &u : c.cell(0).at(g).umis()) umis[c.umi_indexer().get_value(u.first)] = u.second.read_count();
		std::cout << '"' << g << "\": {";
		bool first = true;
		for (auto const &u : umis) { std::cout << (first ? "" : ", ") << '"' << u.first << "\": " << u.second; first = false; }
		std::cout << "}" << (std::string(g) == "Gene1" ? ", " : "");
	}
	std::cout << ", \"umis_stat\": " << c.cell(0).umis_number() << "},\n";
}

static void edit_distances()
{
	std::vector<unsigned> v{Tools::edit_distance("ATTTTC", "ATTTGC"), Tools::edit_distance("ATTTTCC", "ATTTGNC"),
	                        Tools::edit_distance("ATTTTCC", "ATTTGNC", false), Tools::edit_distance("ATTTTCC", "ATTTGTC"),
	                        Tools::edit_distance("ATTTTCC", "ATTTTCC"),
	                        // banded behaviour pinned by SURVEY.md A6 probes
	                        Tools::edit_distance("ACGTACG", "ACGTACGT", true, 1), Tools::edit_distance("AAAA", "TTTT", true, 1),
	                        Tools::edit_distance("ACGTAC", "ACGAAC", true, 1), Tools::edit_distance("ACGTACGT", "ACGACGTT", true, 2)};
	std::cout << "\"testEditDistance\": " << jlist(v) << ",\n";
}

static void collisions_adjuster()
{
	// No reference test covers CollisionsAdjuster; values below are outputs of the compiled reference (SURVEY.md A9 probes).
	Tools::CollisionsAdjuster adj;
	adj.init(std::vector<double>(4096, 1.0 / 4096));
	std::vector<size_t> v;
	for (size_t s : {1, 10, 100, 500, 1000, 2000, 3000}) v.push_back(adj.estimate_adjusted_gene_expression(s));
	Tools::CollisionsAdjuster adj2;
	std::vector<double> p(256);
	double sum = 0;
	for (size_t i = 0; i < p.size(); ++i) { p[i] = 1.0 / (1 + i); sum += p[i]; }
	for (auto &x : p) x /= sum;
	adj2.init(p);
	std::vector<size_t> v2;
	for (size_t s : {1, 5, 20, 50, 100, 150}) v2.push_back(adj2.estimate_adjusted_gene_expression(s));
	std::cout << "\"collisionsAdjuster\": {\"uniform4096\": " << jlist(v) << ", \"zipf256\": " << jlist(v2) << "},\n";
}

namespace TestEstimatorMergeProbs
{
	// Fixture of Tests/TestEstimationMergeProbs.cpp:29-85
	struct PFixture
	{
		std::shared_ptr<Merge::PoissonRealBarcodesMergeStrategy> strat;
		std::shared_ptr<CellsDataContainer> c;
		PFixture()
		{
			auto parser = std::shared_ptr<Merge::BarcodesParsing::BarcodesParser>(
				new Merge::BarcodesParsing::InDropBarcodesParser(DATA + "/barcodes/test_est"));
			const Merge::PoissonTargetEstimator est(1e-4, 1e-7);
			strat = std::make_shared<Merge::PoissonRealBarcodesMergeStrategy>(est, parser, 0, 0, 7);
			c = std::make_shared<CellsDataContainer>(strat, std::make_shared<Merge::UMIs::MergeUMIsStrategySimple>(1),
			                                         Mark::get_by_code(Mark::DEFAULT_CODE), -1);
			static const char *reads[][3] = {
				{"AAATTAGGTCCA", "AAACCT", "Gene1"}, {"AAATTAGGTCCA", "CCCCCT", "Gene2"}, {"AAATTAGGTCCA", "ACCCCT", "Gene3"},
				{"AAATTAGGTCCC", "CAACCT", "Gene1"}, {"AAATTAGGTCCG", "CAACCT", "Gene1"}, {"AAATTAGGTCGG", "AAACCT", "Gene1"},
				{"AAATTAGGTCGG", "CCCCCT", "Gene2"}, {"CCCTTAGGTCCA", "CCATTC", "Gene3"}, {"CCCTTAGGTCCA", "CCCCCT", "Gene2"},
				{"CCCTTAGGTCCA", "ACCCCT", "Gene3"}, {"CAATTAGGTCCG", "CAACCT", "Gene1"}, {"CAATTAGGTCCG", "AAACCT", "Gene1"},
				{"CAATTAGGTCCG", "CCCCCT", "Gene2"}, {"CAATTAGGTCCG", "TTTTTT", "Gene2"}, {"CAATTAGGTCCG", "TTCTTT", "Gene2"},
				{"CCCCCCCCCCCC", "CAACCT", "Gene1"}, {"CCCCCCCCCCCC", "AAACCT", "Gene1"}, {"CCCCCCCCCCCC", "CCCCCT", "Gene2"},
				{"CCCCCCCCCCCC", "TTTTTT", "Gene2"}, {"CCCCCCCCCCCC", "TTCTTT", "Gene2"}, {"TAATTAGGTCCA", "AAAAAA", "Gene4"}};
			for (auto const &r : reads) c->add_record(read_info(r[0], r[1], r[2]));
			c->set_initialized();
		}
	};

	struct testIntersectionSizeEstimation
	{
		static void run()
		{
			PFixture f;
			Merge::PoissonTargetEstimator est(1e-4, 1e-7);
			est.init(f.c->umi_distribution());
			std::cout.precision(17);
			std::cout << "\"poisson\": {\"umi_distribution_size\": " << est._umi_distribution.size() << ", \"intersection_sizes\": [";
			const size_t pairs[][2] = {{1, 5}, {2, 5}, {3, 5}, {4, 5}, {5, 5}, {5, 3}};
			for (size_t i = 0; i < 6; ++i) std::cout << (i ? "," : "") << est.estimate_genes_intersection_size(pairs[i][0], pairs[i][1]);
			std::cout << "], \"probs\": [";
			const size_t cp[][2] = {{0, 1}, {1, 2}, {3, 4}, {5, 6}};
			for (size_t i = 0; i < 4; ++i) std::cout << (i ? "," : "") << est.estimate_intersection_prob(*f.c, cp[i][0], cp[i][1]).merge_probability;
			std::cout << "]";
		}
	};

	struct testPoissonMergeRejections
	{
		static void run()
		{
			PFixture f;
			f.strat->init(*f.c);
			std::vector<long> t;
			for (size_t i = 0; i < f.c->total_cells_number(); ++i) t.push_back(f.strat->get_merge_target(*f.c, i));
			std::cout << ", \"merge_targets_phase1\": " << jlist(t) << "}\n";
		}
	};
}

int main(int argc, char **argv)
{
	DATA = argc > 1 ? argv[1] : "/root/reference/data";
	std::cout << "{\n";
	TestEstimator::testBarcodesFile::run();
	TestEstimator::testUmigsIntersection::run();
	TestEstimator::testFillDistances::run();
	TestEstimator::testRealNeighboursCbs::run();
	TestEstimator::testRealNeighbours::run();
	TestEstimator::testConstLengthBarcodeParser::run();
	TestEstimator::testUMIMergeStrategyDirectional::run();
	merge_by_real_barcodes();
	umi_merge_simple();
	edit_distances();
	collisions_adjuster();
	TestEstimatorMergeProbs::testIntersectionSizeEstimation::run();
	TestEstimatorMergeProbs::testPoissonMergeRejections::run();
	std::cout << "}\n";
	return 0;
}
