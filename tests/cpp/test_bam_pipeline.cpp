// GPU: BAM file(s) -> BamIngest -> CellsDataContainer (device) -> filtered count matrix, printed as text for tests/test_bam_ingest.py.
//   test_bam_pipeline <whitelist (const type) or -> <min_genes_before> <min_genes_after> <type tag> <intronic> <intergenic> file...
#include "../../dropest_b200/host/BamIngest.h"

#include <chrono>
#include <iostream>

using namespace Estimation;

int main(int argc, char **argv)
{
	if (argc < 8) { std::cerr << "usage: see the source\n"; return 2; }
	try
	{
		Merge::MergeStrategyFactory factory;
		factory.barcodes_filename = std::string(argv[1]) == "-" ? std::string() : std::string(argv[1]);
		factory.barcodes_type = "const";
		factory.min_genes_before_merge = size_t(std::stoul(argv[2]));
		factory.min_genes_after_merge = size_t(std::stoul(argv[3]));
		BamProcessing::IngestParams p;
		p.tags.read_type = argv[4]; p.tags.intronic_read_value = argv[5]; p.tags.intergenic_read_value = argv[6];
		if (std::getenv("DGE_BAM_NAME_MODE")) p.filled_bam = false;                        // barcode / UMI from the read names
		if (const char *g = std::getenv("DGE_BAM_GENES")) p.genes_filename = g;             // -g
		std::vector<std::string> files(argv + 7, argv + argc);
		CellsDataContainer container(factory.get_cb_strat(true, false), factory.get_umi(false), UMI::Mark::get_by_code(UMI::Mark::DEFAULT_CODE), false, -1, 0, 1u << 15);
		BamProcessing::IngestStats st;
		const auto t0 = std::chrono::steady_clock::now();
		BamProcessing::parse_bam_files(files, p, container, st);
		const auto t1 = std::chrono::steady_clock::now();
		container.set_initialized();
		container.merge_and_filter();
		const auto t2 = std::chrono::steady_clock::now();
		std::cout << "timing\t" << std::chrono::duration<double>(t1 - t0).count() << "\t" << std::chrono::duration<double>(t2 - t1).count() << "\t"
		          << double(st.total_reads) / std::chrono::duration<double>(t2 - t0).count() / 1e6 << " M reads/s BAM -> matrix\n";
		std::cout << "stats\t" << st.total_reads << '\t' << st.cant_parse << '\t' << st.low_quality << '\t' << container.total_cells_number() << '\t'
		          << container.real_cells_number() << '\t' << container.intergenic_reads_num() << '\n';
		ResultsPrinter printer(false, false);
		auto cm = printer.get_count_matrix(container, true);
		for (auto const &c : cm.col_names) std::cout << "cell\t" << c << '\n';
		for (size_t col = 0; col + 1 < cm.p.size(); ++col)
			for (int k = cm.p[col]; k < cm.p[col + 1]; ++k) std::cout << "cm\t" << col << '\t' << cm.row_names[size_t(cm.i[size_t(k)])] << '\t' << long(cm.x[size_t(k)]) << '\n';
		CellsDataContainer::names_t cells, chrs;
		CellsDataContainer::counts_t counts;
		container.get_stat_by_real_cells(Stats::EXON_READS_PER_CHR_PER_CELL, cells, chrs, counts);
		for (size_t r = 0; r < cells.size(); ++r)
			for (size_t c = 0; c < chrs.size(); ++c)
				if (counts[r * chrs.size() + c]) std::cout << "exon\t" << cells[r] << '\t' << chrs[c] << '\t' << counts[r * chrs.size() + c] << '\n';
	}
	catch (std::exception &e)
	{
		std::cout << "error\t" << e.what() << '\n';
		return 1;
	}
	return 0;
}
