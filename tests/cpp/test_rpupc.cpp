// GPU: reads given as text (the oracle's tsv: barcode, UMI, gene or -, chromosome or -, mark bits, UMI quality) -> CellsDataContainer built with
// save_umi_qualities -> merge_and_filter -> ResultsPrinter::get_reads_per_umi_per_cell, printed for tests/test_rpupc.py; also writes the .rds.
//   test_rpupc <whitelist (const type) or -> <none|real|simple> <simple|directional> <min_genes_before> <min_genes_after> <reads.tsv> <out.rds>
#include "../../dropest_b200/host/Estimation.h"

#include <fstream>
#include <iostream>
#include <sstream>

using namespace Estimation;

int main(int argc, char **argv)
{
	if (argc < 8) { std::cerr << "usage: see the source\n"; return 2; }
	try
	{
		Merge::MergeStrategyFactory factory;
		factory.barcodes_filename = std::string(argv[1]) == "-" ? std::string() : std::string(argv[1]);
		factory.barcodes_type = "const";
		const std::string merge = argv[2];
		factory.min_genes_before_merge = size_t(std::stoul(argv[4]));
		factory.min_genes_after_merge = size_t(std::stoul(argv[5]));
		CellsDataContainer container(factory.get_cb_strat(merge != "none", false), factory.get_umi(std::string(argv[3]) == "directional"),
		                             UMI::Mark::get_by_code(UMI::Mark::DEFAULT_CODE), false, -1, 0, 1u << 12, false, true);
		std::ifstream in(argv[6]);
		std::string line;
		while (std::getline(in, line))
		{
			if (line.empty() || line[0] == '#') continue;
			std::vector<std::string> t;
			std::istringstream ls(line);
			std::string tok;
			while (std::getline(ls, tok, '\t')) t.push_back(tok);
			if (t.size() < 5) throw std::runtime_error("bad tsv line: " + line);
			const int bits = std::stoi(t[4]);
			UMI::Mark mark;
			if (bits & 1) mark.add(UMI::Mark::HAS_NOT_ANNOTATED);
			if (bits & 2) mark.add(UMI::Mark::HAS_EXONS);
			if (bits & 4) mark.add(UMI::Mark::HAS_INTRONS);
			container.add_record(ReadInfo(Tools::ReadParameters(t[0], t[1], "", t.size() > 5 ? t[5] : std::string()), t[2] == "-" ? std::string() : t[2],
			                              t[3] == "-" ? std::string() : t[3], mark));
		}
		container.set_initialized();
		container.merge_and_filter();
		ResultsPrinter printer(false, false, false, true);
		auto r = printer.get_reads_per_umi_per_cell(container);
		for (auto const &c : r.cells) std::cout << "cell\t" << c << '\n';
		for (auto const &g : r.genes) std::cout << "gene\t" << g << '\n';
		for (size_t k = 0; k < r.reads_per_umi.size(); ++k)
		{
			std::cout << "entry\t" << r.cell_indexes[k] << '\t' << r.gene_indexes[k] << '\n';
			auto const &e = r.reads_per_umi[k];
			for (size_t u = 0; u < e.umis.size(); ++u)
			{
				std::cout << "umi\t" << k << '\t' << e.umis[u] << '\t' << e.reads[u] << '\t';
				for (size_t q = 0; q < e.mean_quality[u].size(); ++q) std::cout << (q ? "," : "") << e.mean_quality[u][q];
				std::cout << '\n';
			}
		}
		printer.save_results(container, argv[7]);
		return 0;
	}
	catch (std::exception &e)
	{
		std::cout << "error\t" << e.what() << '\n';
		return 1;
	}
}
