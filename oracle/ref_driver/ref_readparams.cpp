// oracle/ref_driver/ref_readparams.cpp -- TEST INFRASTRUCTURE ONLY.
// The reference's UNMODIFIED read-parameter / gene assignment code (Estimation/BamProcessing/ReadParamsParser.cpp, FilledBamParamsParser.cpp,
// BamTags.cpp, compiled in place with the BamAlignment shim) applied to alignments given as text:
//   ref_readparams <genes file or -> <filled 0|1> <min_barcode_quality> <gene_in_chr 0|1> <type tag or -> <intronic or -> <intergenic or -> <alignments.tsv>
// alignments.tsv: name <TAB> chr <TAB> position <TAB> CIGAR <TAB> TAG:TYPE:VALUE ...   (like SAM, optional fields only as needed)
// Output per alignment: "barcode umi gene mark quality_ok" or "!params" (get_read_params failed) / "!chr" (ChrNotFoundException).
#include <Estimation/BamProcessing/FilledBamParamsParser.h>
#include <Estimation/BamProcessing/ReadParamsParser.h>
#include <Tools/ReadParameters.h>

#include <fstream>
#include <iostream>
#include <sstream>

using namespace Estimation;

static int mark_bits(const UMI::Mark &m)
{
	return (m.check(UMI::Mark::HAS_NOT_ANNOTATED) ? 1 : 0) | (m.check(UMI::Mark::HAS_EXONS) ? 2 : 0) | (m.check(UMI::Mark::HAS_INTRONS) ? 4 : 0);
}

int main(int argc, char **argv)
{
	if (argc < 9) { std::cerr << "usage: see the source\n"; return 2; }
	try
	{
		auto opt = [](const char *s) { return std::string(s) == "-" ? std::string() : std::string(s); };
		const std::string genes = opt(argv[1]);
		const bool filled = std::string(argv[2]) == "1";
		const int min_quality = std::stoi(argv[3]);
		const bool gene_in_chr = std::string(argv[4]) == "1";
		boost::property_tree::ptree cfg;
		if (!opt(argv[5]).empty()) cfg.put("BamTags.Type.tag", argv[5]);
		if (!opt(argv[6]).empty()) cfg.put("BamTags.Type.intronic", argv[6]);
		if (!opt(argv[7]).empty()) cfg.put("BamTags.Type.intergenic", argv[7]);
		BamProcessing::BamTags tags(cfg);
		std::shared_ptr<BamProcessing::ReadParamsParser> parser;
		if (filled) parser = std::make_shared<BamProcessing::FilledBamParamsParser>(genes, tags, gene_in_chr, Tools::ReadParameters::quality_to_phred(min_quality));
		else parser = std::make_shared<BamProcessing::ReadParamsParser>(genes, tags, gene_in_chr);

		std::ifstream in(argv[8]);
		std::string line;
		while (std::getline(in, line))
		{
			if (line.empty()) continue;
			std::vector<std::string> f;
			std::istringstream ls(line);
			std::string tok;
			while (std::getline(ls, tok, '\t')) f.push_back(tok);
			BamTools::BamAlignment al;
			al.Name = f.at(0);
			const std::string chr = f.at(1);
			al.Position = std::stoi(f.at(2));
			{
				const std::string &cg = f.at(3);
				size_t i = 0;
				while (i < cg.size() && cg != "*")
				{
					size_t j = i;
					while (j < cg.size() && isdigit(cg[j])) ++j;
					al.CigarData.emplace_back(cg[j], uint32_t(std::stoul(cg.substr(i, j - i))));
					i = j + 1;
				}
			}
			for (size_t k = 4; k < f.size(); ++k)
				if (f[k].size() >= 5) al.Tags[f[k].substr(0, 2)] = std::make_pair(f[k][3], f[k].substr(5));
			Tools::ReadParameters rp;
			if (!parser->get_read_params(al, rp)) { std::cout << "!params\n"; continue; }
			std::string gene;
			UMI::Mark mark;
			try { mark = parser->get_gene(chr, al, gene); }
			catch (Tools::GeneAnnotation::RefGenesContainer::ChrNotFoundException &) { std::cout << "!chr\n"; continue; }
			std::cout << rp.cell_barcode() << ' ' << rp.umi() << ' ' << (gene.empty() ? "-" : gene) << ' ' << mark_bits(mark) << ' ' << (rp.pass_quality_threshold() ? 1 : 0) << '\n';
		}
	}
	catch (std::exception &e)
	{
		std::cout << "#error " << e.what() << '\n';
		return 1;
	}
	return 0;
}
