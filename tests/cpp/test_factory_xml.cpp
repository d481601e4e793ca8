// CPU-only: MergeStrategyFactory::from_xml against the rules of the reference constructor (Estimation/Merge/MergeStrategyFactory.cpp:23-59):
// keys, defaults, the mandatory max_cb_merge_edit_distance, barcodes_file resolved against the configuration file, the -G override, and
// the strategy selection that follows from the values (:61-126).  No CUDA call is made: strategies are descriptors until a container uses them.
#include "../../dropest_b200/host/Estimation.h"

#include <iostream>

using namespace Estimation;

static int failures = 0;
#define CHECK(cond) do { if (!(cond)) { std::cerr << "CHECK failed: " #cond " (line " << __LINE__ << ")\n"; ++failures; } } while (0)
#define CHECK_EQUAL(a, b) do { auto _a = (a); auto _b = (b); if (!(_a == _b)) { std::cerr << "CHECK_EQUAL failed: " #a " == " #b " (" << _a << " vs " << _b << ", line " << __LINE__ << ")\n"; ++failures; } } while (0)

int main(int argc, char **argv)
{
	if (argc < 2) { std::cerr << "usage: test_factory_xml <dir with the fixture configs>\n"; return 2; }
	const std::string dir = argv[1];
	try
	{
		auto f = Merge::MergeStrategyFactory::from_xml(dir + "/10x_like.xml");
		CHECK_EQUAL(f.merge_type, std::string("none"));
		CHECK_EQUAL(f.barcodes_type, std::string("const"));
		CHECK_EQUAL(f.barcodes_filename, dir + "/../wl_synth_7x9.txt");   // ltrim + relative to the configuration file
		CHECK_EQUAL(f.min_genes_before_merge, size_t(20));
		CHECK_EQUAL(f.min_genes_after_merge, size_t(100));
		CHECK_EQUAL(f.max_merge_edit_distance, 3u);
		CHECK_EQUAL(f.max_umi_merge_edit_distance, 1u + 1u);
		CHECK_EQUAL(f.min_merge_fraction, 0.25);
		CHECK_EQUAL(f.max_merge_prob, 1e-5);
		CHECK_EQUAL(f.max_real_cb_merge_prob, 1e-7);
		CHECK_EQUAL(f.umi_merge_mult, 2.0);
		CHECK_EQUAL(f.get_cb_strat(false, false)->merge_type(), std::string("No"));
		CHECK_EQUAL(f.get_cb_strat(true, false)->merge_type(), std::string("Real CBs"));
		CHECK_EQUAL(f.get_cb_strat(true, true)->merge_type(), std::string("Poisson Real CBs"));
		CHECK_EQUAL(f.get_cb_strat(true, false)->min_genes_after_merge(), size_t(100));

		auto g = Merge::MergeStrategyFactory::from_xml(dir + "/10x_like.xml", 250); // -G 250
		CHECK_EQUAL(g.min_genes_after_merge, size_t(250));

		auto d = Merge::MergeStrategyFactory::from_xml(dir + "/dropseq_like.xml");
		CHECK_EQUAL(d.merge_type, std::string("all"));
		CHECK_EQUAL(d.barcodes_type, std::string("indrop"));  // defaults of the reference
		CHECK(d.barcodes_filename.empty());
		CHECK_EQUAL(d.min_genes_before_merge, size_t(10));
		CHECK_EQUAL(d.min_genes_after_merge, size_t(10));
		CHECK_EQUAL(d.min_merge_fraction, 0.2);
		CHECK_EQUAL(d.max_merge_prob, 1e-4);
		CHECK_EQUAL(d.umi_merge_mult, 3.0);
		CHECK_EQUAL(d.get_cb_strat(true, false)->merge_type(), std::string("Merge all"));

		bool thrown = false;
		try { Merge::MergeStrategyFactory::from_xml(dir + "/missing_distance.xml"); } catch (std::runtime_error &) { thrown = true; }
		CHECK(thrown);   // ptree::get without a default throws in the reference
		thrown = false;
		try { Merge::MergeStrategyFactory::from_xml(dir + "/does_not_exist.xml"); } catch (std::runtime_error &) { thrown = true; }
		CHECK(thrown);
	}
	catch (std::exception &e)
	{
		std::cerr << "unexpected exception: " << e.what() << "\n";
		return 1;
	}
	std::cout << (failures ? "FAILED" : "OK") << " (" << failures << " failures)\n";
	return failures ? 1 : 0;
}
