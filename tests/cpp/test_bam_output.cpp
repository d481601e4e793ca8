// The BAM writers (dropest_b200/host/BamOutput), driven for tests/test_bam_output.py.
//   test_bam_output cpu <out dir> <threads> <type tag or -> <intronic or -> <intergenic or -> file...
//       CPU only: every accepted read of every file is written to "<out dir>/<name>.tagged.bam" with the tags of
//       BamProcessorAbstract::save_alignment and made-up corrections (barcode reversed; the UMI itself when it starts with 'A', else none).
//   test_bam_output gpu <out dir> <whitelist (const type; '-' = no barcode merge)> <min_genes_before> <min_genes_after> <simple|directional> <type tag or -> <intronic or ->
//                       <intergenic or -> <bam output 0|1> file...
//       the whole flow on the GPU: parse_bam_files (-b) -> set_initialized -> merge_and_filter -> write_filtered_bam_files (-F)
#include "../../dropest_b200/host/BamOutput.h"

#include <iostream>

using namespace Estimation;

int main(int argc, char **argv)
{
	auto opt = [](const char *s) { return std::string(s) == "-" ? std::string() : std::string(s); };
	try
	{
		const std::string mode = argc > 1 ? argv[1] : "";
		if (mode == "cpu" && argc >= 8)
		{
			BamProcessing::IngestParams p;
			p.output_dir = argv[2];
			p.threads = unsigned(std::stoul(argv[3]));
			p.tags.read_type = opt(argv[4]); p.tags.intronic_read_value = opt(argv[5]); p.tags.intergenic_read_value = opt(argv[6]);
			std::vector<std::string> files(argv + 7, argv + argc);
			BamProcessing::IngestStats st;
			std::unique_ptr<BamProcessing::BamWriter> writer;
			std::vector<BamProcessing::BamWriter::TagEdit> edits;
			size_t written = 0;
			BamProcessing::for_each_alignment(files, p, st, true,
				[&](const std::string &file, const BamProcessing::BamReader &reader) {
					if (writer) { written += writer->written(); writer->close(); }
					writer.reset(new BamProcessing::BamWriter(BamProcessing::result_bam_name(file, ".tagged.bam", p.output_dir), reader.header_text(),
					                                          reader.reference_names(), reader.reference_lengths(), p.threads));
				},
				[&](const ReadInfo &ri, const BamProcessing::BamReader::RecordView *view) {
					const std::string &cb = ri.params.cell_barcode(), &umi = ri.params.umi();
					BamProcessing::tag_edits(p.tags, ri, std::string(cb.rbegin(), cb.rend()), umi[0] == 'A' ? umi : std::string(), edits);
					writer->save_alignment(view->raw, view->raw_bytes, edits);
				});
			if (writer) { written += writer->written(); writer->close(); }
			std::cout << "stats\t" << st.total_reads << '\t' << st.cant_parse << '\t' << st.low_quality << '\t' << written << '\n';
			return 0;
		}
		if (mode == "gpu" && argc >= 12)
		{
			Merge::MergeStrategyFactory factory;
			factory.barcodes_filename = opt(argv[3]);
			factory.barcodes_type = "const";
			factory.min_genes_before_merge = size_t(std::stoul(argv[4]));
			factory.min_genes_after_merge = size_t(std::stoul(argv[5]));
			const bool directional = std::string(argv[6]) == "directional";
			BamProcessing::IngestParams p;
			p.output_dir = argv[2];
			p.tags.read_type = opt(argv[7]); p.tags.intronic_read_value = opt(argv[8]); p.tags.intergenic_read_value = opt(argv[9]);
			const bool bam_output = std::string(argv[10]) == "1";
			std::vector<std::string> files(argv + 11, argv + argc);
			CellsDataContainer container(factory.get_cb_strat(!factory.barcodes_filename.empty(), false), factory.get_umi(directional), UMI::Mark::get_by_code(UMI::Mark::DEFAULT_CODE),
			                             true, -1, 0, 1u << 12);
			BamProcessing::IngestStats st;
			BamProcessing::parse_bam_files(files, p, container, st, bam_output);
			container.set_initialized();
			container.merge_and_filter();
			BamProcessing::FilteredBamStats fs;
			BamProcessing::write_filtered_bam_files(files, p, container, fs);
			std::cout << "stats\t" << st.total_reads << '\t' << container.total_cells_number() << '\t' << container.real_cells_number() << '\t'
			          << container.filtered_cells().size() << '\t' << fs.written_reads << '\t' << fs.wrong_genes << '\t' << fs.wrong_umis << '\t' << fs.file_name << '\n';
			return 0;
		}
		std::cerr << "usage: see the source\n";
		return 2;
	}
	catch (std::exception &e)
	{
		std::cout << "error\t" << e.what() << '\n';
		return 1;
	}
}
