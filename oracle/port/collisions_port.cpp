// collisions_port.cpp -- CPU restatement of Tools::CollisionsAdjuster (reference Tools/CollisionsAdjuster.cpp:12-49) and Tools::fpow
// (Tools/UtilFunctions.cpp:13-30).  TEST INFRASTRUCTURE: only tests/ may execute it.  Pinned by tests/test_collisions.py against
// tests/golden/collisions.npz (outputs of the compiled, unmodified reference: oracle/_ref/ref_collisions).
// usage: collisions_port <probabilities.f64> <max_gene_expression>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <vector>

static double fpow(double base, long exp) // UtilFunctions.cpp:13-30
{
	if (exp == 1) return base;
	double result = 1;
	while (exp)
	{
		if (exp & 1) result *= base;
		exp >>= 1;
		base *= base;
	}
	return result;
}

int main(int argc, char **argv)
{
	if (argc < 3) { std::cerr << "usage: collisions_port <probabilities.f64> <max_gene_expression>\n"; return 2; }
	FILE *f = fopen(argv[1], "rb");
	if (!f) { std::cerr << "cannot open " << argv[1] << "\n"; return 1; }
	std::vector<double> p;
	double x;
	while (fread(&x, sizeof(double), 1, f) == 1) p.push_back(x);
	fclose(f);
	const size_t max_size = size_t(atol(argv[2]));
	// CollisionsAdjuster::init + update_adjusted_sizes, CollisionsAdjuster.cpp:12-39
	std::vector<double> neg_prod(p.size(), 1);
	double sum_collisions = 0;
	size_t last_total = 0;
	for (size_t s = 1; s <= max_size; ++s)
	{
		const size_t total = s + size_t(sum_collisions);
		double new_umi_prob = 0;
		for (size_t i = 0; i < p.size(); ++i)
		{
			neg_prod[i] *= fpow(1 - p[i], long(total - last_total));
			new_umi_prob += p[i] * (1 - neg_prod[i]);
		}
		last_total = total;
		const double collision_num = 1.0 / (1.0 - new_umi_prob) - 1.0;
		sum_collisions += collision_num;
		std::cout << std::lround(s + sum_collisions) << "\n";
	}
	return 0;
}
