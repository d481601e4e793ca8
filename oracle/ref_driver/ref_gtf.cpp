// oracle/ref_driver/ref_gtf.cpp -- TEST INFRASTRUCTURE ONLY.
// The reference's UNMODIFIED gene annotation (Tools/GeneAnnotation/RefGenesContainer.cpp, GtfRecord.cpp, Interval.cpp, compiled in place)
// answering position queries:  ref_gtf <genes.gtf[.gz] | genes.bed[.gz]> <queries.tsv>
// queries.tsv: "chr<TAB>start<TAB>end" per line (0-based, end exclusive).  Output per query: "gene:type,gene:type,..." in the order of the
// reference's result set (type 1 = intron, 2 = exon), "-" when empty, "!chr" when the chromosome is unknown; then "#has_introns <0|1>".
#include <Tools/GeneAnnotation/RefGenesContainer.h>

#include <fstream>
#include <iostream>
#include <sstream>

int main(int argc, char **argv)
{
	if (argc < 3) { std::cerr << "usage: ref_gtf <genes file> <queries.tsv>\n"; return 2; }
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
