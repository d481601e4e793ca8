// oracle/ref_driver/ref_bam_flow.cpp -- TEST INFRASTRUCTURE ONLY (checker; never shipped or called by the product).
//
// The reference's UNMODIFIED BAM flow -- BamController::parse_bam_files (-b: BamProcessor writes "<name>.tagged.bam"), set_initialized,
// merge_and_filter, BamController::write_filtered_bam_files (-F: FilteringBamProcessor writes "<first name>.filtered.bam") -- exactly as
// dropest.cpp:239-254,303-310 drives it, compiled in place from /root/reference (BamController.cpp, BamProcessor.cpp, BamProcessorAbstract.cpp,
// FilteringBamProcessor.cpp, the parameter parsers, the container and the merge strategies).  BamTools is shimmed: the "BAM" files this
// driver reads and writes are TEXT (oracle/shim/api/BamReader.h, BamWriter.h); a test writes the same alignments as a real BAM for the
// product and compares the tags of what comes out.  Output files land in the working directory, as the reference does.
//   ref_bam_flow [--merge none|real|simple ...] [--barcodes F --barcodes-type const|indrop] [--min-genes-before N] [--min-genes-after N]
//                [--umi-merge simple|directional] [--filled 0|1] [--read-params "F F ..."] [--min-quality Q] [--type-tag T --intronic V --intergenic V --exonic V]
//                [--bam-output 0|1] [--filtered 0|1] alignments.txt ...
#include <Estimation/BamProcessing/BamController.h>
#include <Estimation/CellsDataContainer.h>
#include <Estimation/Merge/DummyMergeStrategy.h>
#include <Estimation/Merge/RealBarcodesMergeStrategy.h>
#include <Estimation/Merge/SimpleMergeStrategy.h>
#include <Estimation/Merge/BarcodesParsing/ConstLengthBarcodesParser.h>
#include <Estimation/Merge/BarcodesParsing/InDropBarcodesParser.h>
#include <Estimation/Merge/UMIs/MergeUMIsStrategyDirectional.h>
#include <Estimation/Merge/UMIs/MergeUMIsStrategySimple.h>
#include <Tools/Logs.h>
#include <Tools/ReadParameters.h>

#include <iostream>

using namespace Estimation;

namespace Tools
{
	// Replaces Tools/Logs.cpp (Boost.Log).  Signatures from Tools/Logs.h:15-20.
	void init_log(bool, bool, const std::string &, const std::string &) {}
	void init_test_logs(boost::log::trivial::severity_level) {}
	void trace_time(const std::string &, bool) {}
}

int main(int argc, char **argv)
{
	try
	{
		std::string merge = "none", barcodes, barcodes_type = "const", umi_merge = "simple", marks = "eEBA", read_params;
		int min_quality = 0;
		size_t min_before = 10, min_after = 10;
		unsigned max_cb_ed = 2, max_umi_ed = 1;
		double min_frac = 0.2, umi_mult = 2;
		bool filled = true, bam_output = false, filtered = true;
		boost::property_tree::ptree cfg;
		std::vector<std::string> files;
		for (int i = 1; i < argc; ++i)
		{
			const std::string k = argv[i];
			auto next = [&]() -> std::string {
				if (i + 1 >= argc) throw std::runtime_error("missing value for " + k);
				return argv[++i];
			};
			if (k == "--merge") merge = next();
			else if (k == "--barcodes") barcodes = next();
			else if (k == "--barcodes-type") barcodes_type = next();
			else if (k == "--umi-merge") umi_merge = next();
			else if (k == "--marks") marks = next();
			else if (k == "--min-genes-before") min_before = std::stoul(next());
			else if (k == "--min-genes-after") min_after = std::stoul(next());
			else if (k == "--max-cb-ed") max_cb_ed = unsigned(std::stoul(next()));
			else if (k == "--max-umi-ed") max_umi_ed = unsigned(std::stoul(next()));
			else if (k == "--min-frac") min_frac = std::stod(next());
			else if (k == "--umi-mult") umi_mult = std::stod(next());
			else if (k == "--filled") filled = next() == "1";
			else if (k == "--read-params") read_params = next();   // -r: file names separated by blanks
			else if (k == "--min-quality") min_quality = std::stoi(next());
			else if (k == "--bam-output") bam_output = next() == "1";
			else if (k == "--filtered") filtered = next() == "1";
			else if (k == "--type-tag") cfg.put("BamTags.Type.tag", next());
			else if (k == "--intronic") cfg.put("BamTags.Type.intronic", next());
			else if (k == "--intergenic") cfg.put("BamTags.Type.intergenic", next());
			else if (k == "--exonic") cfg.put("BamTags.Type.exonic", next());
			else if (k.compare(0, 2, "--") == 0) throw std::runtime_error("unknown argument " + k);
			else files.push_back(k);
		}
		// strategy selection as in Estimation/Merge/MergeStrategyFactory.cpp:61-126 (that file needs boost::property_tree's XML reader)
		std::shared_ptr<Merge::MergeStrategyAbstract> cb_strat;
		if (merge == "none") cb_strat = std::make_shared<Merge::DummyMergeStrategy>(min_before, min_after);
		else if (merge == "simple") cb_strat = std::make_shared<Merge::SimpleMergeStrategy>(min_before, min_after, max_cb_ed, min_frac);
		else if (merge == "real")
		{
			std::shared_ptr<Merge::BarcodesParsing::BarcodesParser> parser;
			if (barcodes_type == "indrop") parser = std::make_shared<Merge::BarcodesParsing::InDropBarcodesParser>(barcodes);
			else parser = std::make_shared<Merge::BarcodesParsing::ConstLengthBarcodesParser>(barcodes);
			cb_strat = std::make_shared<Merge::RealBarcodesMergeStrategy>(parser, min_before, min_after, max_cb_ed, min_frac);
		}
		else throw std::runtime_error("unknown merge type " + merge);
		std::shared_ptr<Merge::UMIs::MergeUMIsStrategyAbstract> umi_strat;
		if (umi_merge == "directional") umi_strat = std::make_shared<Merge::UMIs::MergeUMIsStrategyDirectional>(umi_mult, max_umi_ed);
		else umi_strat = std::make_shared<Merge::UMIs::MergeUMIsStrategySimple>(max_umi_ed);

		BamProcessing::BamController bam_controller(BamProcessing::BamTags(cfg), filled, read_params, "", false, Tools::ReadParameters::quality_to_phred(min_quality));
		// get_cells_container, dropest.cpp:239-254
		CellsDataContainer container(cb_strat, umi_strat, UMI::Mark::get_by_code(marks), filtered, -1);
		bam_controller.parse_bam_files(files, bam_output, container);
		container.set_initialized();
		container.merge_and_filter();
		if (filtered) bam_controller.write_filtered_bam_files(files, container); // dropest.cpp:307-310
		std::cout << "cells\t" << container.total_cells_number() << "\treal\t" << container.real_cells_number() << "\tfiltered\t"
		          << container.filtered_cells().size() << '\n';
		return 0;
	}
	catch (std::exception &e)
	{
		std::cerr << "ref_bam_flow: ERROR: " << e.what() << "\n";
		return 1;
	}
}
