// The HOST side of BAM -> container (BamProcessing::parse_bam_files: background loader, producer thread, hand-over thread) against the no-op
// device stand-in tests/cpp/stub_device.c: counters, the number of reads that reach the device entry points and a digest of everything they
// carry (packed key, gene word, stream position, chromosome id per read), bulk path == one-read path,
// errors thrown on the loader / producer threads arriving at the caller.  CPU only; nothing is computed (see stub_device.c).
// usage: test_ingest_pipeline_host <threads> <name_mode 0|1> <min_quality> <bam>...
#include "../../dropest_b200/host/BamIngest.h"

#include <iostream>

extern "C" unsigned long long stub_reads_seen(void);
extern "C" unsigned long long stub_digest(void);

using namespace Estimation;

int main(int argc, char **argv)
{
	if (argc < 5) { std::cerr << "usage: see the source\n"; return 2; }
	try
	{
		Merge::MergeStrategyFactory factory;
		factory.barcodes_type = "const";
		BamProcessing::IngestParams p;
		p.threads = unsigned(std::stoi(argv[1]));
		p.filled_bam = std::string(argv[2]) == "0";
		p.min_barcode_quality = std::stoi(argv[3]);
		p.tags.read_type = "XF"; p.tags.intronic_read_value = "INTRONIC"; p.tags.intergenic_read_value = "INTERGENIC";
		if (const char *g = std::getenv("DGE_BAM_GENES")) p.genes_filename = g;
		std::vector<std::string> files(argv + 4, argv + argc);
		CellsDataContainer container(factory.get_cb_strat(true, false), factory.get_umi(false), UMI::Mark::get_by_code(UMI::Mark::DEFAULT_CODE), false, -1, 0, 1u << 12);
		BamProcessing::IngestStats st;
		BamProcessing::parse_bam_files(files, p, container, st);
		container.set_initialized(); // hands over the last batch
		std::cout << "stats\t" << st.total_reads << '\t' << st.cant_parse << '\t' << st.low_quality << '\t' << st.skipped_unmapped_or_secondary << '\t'
		          << stub_reads_seen() << '\t' << container.skipped_n_reads() << '\t' << container.skipped_length_reads() << '\t' << stub_digest() << std::endl;
	}
	catch (std::exception &e)
	{
		std::cout << "error\t" << e.what() << std::endl;
		std::_Exit(1); // the container's destructor would talk to the stand-in device: nothing to release
	}
	std::_Exit(0);
}
