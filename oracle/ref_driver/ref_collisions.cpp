// ref_collisions.cpp -- drives the UNMODIFIED reference Tools::CollisionsAdjuster (Tools/CollisionsAdjuster.cpp, compiled in place by
// oracle/Makefile).  Test infrastructure only.
// usage: ref_collisions <probabilities.f64> <max_gene_expression>     (raw little-endian doubles in, adjusted sizes 1..max out, one per line)
#include <Tools/CollisionsAdjuster.h>

#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <vector>

int main(int argc, char **argv)
{
	if (argc < 3) { std::cerr << "usage: ref_collisions <probabilities.f64> <max_gene_expression>\n"; return 2; }
	FILE *f = fopen(argv[1], "rb");
	if (!f) { std::cerr << "cannot open " << argv[1] << "\n"; return 1; }
	std::vector<double> p;
	double x;
	while (fread(&x, sizeof(double), 1, f) == 1) p.push_back(x);
	fclose(f);
	const size_t max_size = size_t(atol(argv[2]));
	Tools::CollisionsAdjuster adj;
	adj.init(p, max_size);
	for (size_t s = 1; s <= max_size; ++s) std::cout << adj.estimate_adjusted_gene_expression(s) << "\n";
	return 0;
}
