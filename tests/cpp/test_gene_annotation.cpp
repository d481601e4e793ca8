// CPU-only tool behind tests/test_gene_annotation.py: dropest_b200/host/GeneAnnotation answering the same queries, in the same output format,
// as oracle/ref_driver/ref_gtf.cpp does with the reference's own code.
#include "../../dropest_b200/host/GeneAnnotation.h"

#include <fstream>
#include <iostream>

int main(int argc, char **argv)
{
	if (argc < 3) { std::cerr << "usage: test_gene_annotation <genes file> <queries.tsv>\n"; return 2; }
	try
	{
		Tools::GeneAnnotation::RefGenesContainer container(argv[1]);
		std::ifstream q(argv[2]);
		std::string chr;
		unsigned long start, end;
		while (q >> chr >> start >> end)
		{
			try
			{
				auto res = container.get_gene_info(chr, start, end);
				if (res.empty()) { std::cout << "-\n"; continue; }
				bool first = true;
				for (auto const &r : res) { std::cout << (first ? "" : ",") << r.gene_name << ':' << int(r.type); first = false; }
				std::cout << '\n';
			}
			catch (Tools::GeneAnnotation::RefGenesContainer::ChrNotFoundException &) { std::cout << "!chr\n"; }
		}
		std::cout << "#has_introns " << (container.has_introns() ? 1 : 0) << '\n';
	}
	catch (std::exception &e)
	{
		std::cout << "#error " << e.what() << '\n';
		return 1;
	}
	return 0;
}
