// CPU-only tool behind tests/test_bam_ingest.py: reads BAM files with dropest_b200/host/BamIngest and prints one line per accepted read
// (barcode, UMI, gene, chromosome, mark bits, barcode quality) followed by the counters.  No container, no CUDA call.
//   test_bam_ingest <filled 0|1> <min_barcode_quality> <gene_in_chr 0|1> <type tag or -> <intronic value or -> <intergenic value or -> <threads> file...
//   environment DGE_BAM_GENES=<annotation.gtf[.gz] | .bed[.gz]>: gene and mark from the annotation (-g) instead of the gene tag
//   environment DGE_BAM_PACKED=1: the bulk path (parse_batch_packed) instead of one ReadInfo per read; three more columns (packable, packed barcode, packed UMI)
//   environment DGE_BAM_READ_PARAMS="<file> <file> ...": barcode / UMI by read name from droptag's read-parameter files (-r)
#include "../../dropest_b200/host/BamIngest.h"

#include <chrono>
#include <cstdlib>
#include <iostream>

using namespace Estimation;

int main(int argc, char **argv)
{
	if (argc < 9) { std::cerr << "usage: see the source\n"; return 2; }
	try
	{
		BamProcessing::IngestParams p;
		p.filled_bam = std::string(argv[1]) == "1";
		p.min_barcode_quality = std::stoi(argv[2]);
		p.gene_in_chromosome_name = std::string(argv[3]) == "1";
		auto opt = [](const char *s) { return std::string(s) == "-" ? std::string() : std::string(s); };
		p.tags.read_type = opt(argv[4]); p.tags.intronic_read_value = opt(argv[5]); p.tags.intergenic_read_value = opt(argv[6]);
		p.threads = unsigned(std::stoi(argv[7]));
		if (const char *g = std::getenv("DGE_BAM_GENES")) p.genes_filename = g;
		if (const char *r = std::getenv("DGE_BAM_READ_PARAMS")) p.read_param_filenames = r; // -r (with filled = 0)
		std::vector<std::string> files(argv + 8, argv + argc);
		BamProcessing::IngestStats st;
		if (std::getenv("DGE_BAM_COUNT_ONLY"))
		{   // parse rate without the printing
			size_t n = 0, bytes = 0;
			const auto t0 = std::chrono::steady_clock::now();
			BamProcessing::for_each_read(files, p, st, [&](const ReadInfo &ri) { ++n; bytes += ri.gene.size() + ri.params.umi().size(); });
			const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
			std::cout << "#count\t" << n << '\t' << bytes << '\t' << dt << " s\t" << double(n) / dt / 1e6 << " M reads/s\n";
		}
		else if (std::getenv("DGE_BAM_PACKED"))
		{   // the bulk path (parse_batch_packed), printed like the one-read path prints its ReadInfo
			if (!BamProcessing::packed_path_applies(p)) throw std::runtime_error("the packed path does not apply to these parameters");
			if (!p.genes && !p.genes_filename.empty()) p.genes = std::make_shared<const Tools::GeneAnnotation::RefGenesContainer>(p.genes_filename);
			for (auto const &file : files)
			{
				BamProcessing::BamReader reader(file, p.threads);
				std::vector<BamProcessing::BamReader::RecordView> views;
				BamProcessing::PackedBatch batch;
				while (true)
				{
					reader.next_batch(views, 50000);
					if (views.empty()) break;
					BamProcessing::parse_batch_packed(views, reader.reference_names(), p, batch, p.threads);
					for (size_t k = 0; k < batch.status.size(); ++k)
					{
						using S = BamProcessing::ParsedRead;
						switch (S::Status(batch.status[k]))
						{
						case S::SKIPPED: ++st.skipped_unmapped_or_secondary; break;
						case S::NO_CHROMOSOME: ++st.cant_parse; break;
						case S::CANT_PARSE: ++st.total_reads; ++st.cant_parse; break;
						case S::LOW_QUALITY: ++st.total_reads; ++st.low_quality; break;
						case S::OK:
						{
							++st.total_reads;
							const PackedRead &r = batch.reads[k];
							auto str = [](const char *q, size_t n) { return n ? std::string(q, n) : std::string("-"); };
							std::cout << std::string(r.cb, r.cb_len) << '\t' << std::string(r.umi, r.umi_len) << '\t' << str(r.gene, r.gene_len) << '\t'
							          << reader.reference_names()[size_t(r.chromosome)] << '\t' << int(r.mark_bits) << '\t' << str(r.cb_quality, r.cb_quality_len) << '\t'
							          << str(r.umi_quality, r.umi_quality_len) << '\t' << int(r.packable) << '\t' << r.cb_packed << '\t' << r.umi_packed << '\n';
							break;
						}
						}
					}
				}
			}
		}
		else BamProcessing::for_each_read(files, p, st, [](const ReadInfo &ri) {
			std::cout << ri.params.cell_barcode() << '\t' << ri.params.umi() << '\t' << (ri.gene.empty() ? "-" : ri.gene) << '\t' << ri.chromosome_name << '\t'
			          << ri.umi_mark.bits() << '\t' << (ri.params.cell_barcode_quality().empty() ? "-" : ri.params.cell_barcode_quality()) << '\t'
			          << (ri.params.umi_quality().empty() ? "-" : ri.params.umi_quality()) << '\n';
		});
		std::cout << "#stats\t" << st.total_reads << '\t' << st.cant_parse << '\t' << st.low_quality << '\t' << st.skipped_unmapped_or_secondary << '\n';
	}
	catch (std::exception &e)
	{
		std::cout << "#error\t" << e.what() << '\n';
		return 1;
	}
	return 0;
}
