// vcf2eds — CLI contract of the reference tool (src/cpp/tools/vcf2eds.cpp:28-226): options, extension and reference
// checks, output naming, stdout lines, the statistics block and the [Performance] line; the work runs on the GPU.
#include <filesystem>
#include <fstream>

#include "cli_common.hpp"
#include "edsparser/transforms/eds_transforms.hpp"
#include "edsparser/transforms/vcf_transforms.hpp"

using namespace edsparser;
namespace fs = std::filesystem;

static void usage() {
    std::cout << "vcf2eds - Transform VCF (Variant Call Format) to EDS\n\n"
                 "Transform VCF (Variant Call Format) to EDS/l-EDS:\n"
                 "  -h [ --help ]                     Show help message\n"
                 "  -i [ --input ] arg                Input VCF file (.vcf)\n"
                 "  -r [ --reference ] arg            Reference FASTA file\n"
                 "  -o [ --output ] arg               Output EDS file (default: <input>.eds)\n"
                 "  -s [ --sources ] arg              Output source file (default: <output>.seds)\n"
                 "  -l [ --context-length ] arg (=0)  Create l-EDS with minimum context length (0 = regular EDS)\n"
                 "  --device arg (=0)                 CUDA device to run on (B200 build)\n"
                 "  --gpus arg (=1)                   Slice the record lines over this many GPUs of the node,\n"
                 "                                    starting at --device (B200 build)\n\n"
                 "One path per sample column (1-based, VCF order); both alleles of a diploid genotype count.\n"
                 "SNPs, indels, <DEL>, <INS> and multi-allelic sites are supported; overlapping records are merged\n"
                 "into one symbol; other symbolic alleles and malformed lines are skipped and counted.\n\n"
                 "OUTPUT:\n"
                 "  Regular EDS:     <input_base>.eds, <input_base>.seds\n"
                 "  l-EDS (with -l): <input_base>_l<N>.leds, <input_base>_l<N>.seds (VCF -> EDS -> LINEAR merge)\n\n";
}

static int run(int argc, char** argv) {
    Timer timer;
    timer.start();
    try {
        const cli::Args args(argc, argv, {{"help", 'h', false}, {"input", 'i', true}, {"reference", 'r', true}, {"output", 'o', true},
                                         {"sources", 's', true}, {"context-length", 'l', true}, {"device", 0, true}, {"gpus", 0, true}});
        if (args.has("help")) {
            usage();
            cli::print_performance(timer);
            return 0;
        }
        args.require("input");
        args.require("reference");
        const fs::path input_file = args.get("input");
        const fs::path reference_file = args.get("reference");
        const fs::path output_file = args.has("output") ? fs::path(args.get("output")) : fs::path();
        const fs::path sources_file = args.has("sources") ? fs::path(args.get("sources")) : fs::path();
        const unsigned long l_arg = args.has("context-length") ? args.to_uint("context-length") : 0;
        if (l_arg > 0xfffffffful) throw std::invalid_argument("the argument for option '--context-length' is invalid");
        const Length context_length = (Length)l_arg;
        const int first_device = args.has("device") ? (int)args.to_uint("device") : 0;
        b200::set_device(first_device);
        if (args.has("gpus")) {
            const unsigned long n = args.to_uint("gpus");
            if (n < 1 || n > 64) throw std::invalid_argument("the argument for option '--gpus' is invalid");
            std::vector<int> devices;
            for (unsigned long i = 0; i < n; ++i) devices.push_back(first_device + (int)i);
            b200::set_devices(devices);
        }

        if (input_file.extension() != ".vcf") {
            std::cerr << "Error: Input file must be a VCF file (.vcf)\n";
            std::cerr << "Got: " << input_file << "\n";
            cli::print_performance(timer);
            return 1;
        }
        if (!fs::exists(reference_file)) {
            std::cerr << "Error: Reference FASTA file not found: " << reference_file << "\n";
            cli::print_performance(timer);
            return 1;
        }
        std::ifstream vcf_in(input_file, std::ios::binary);
        if (!vcf_in) throw std::runtime_error("Failed to open VCF file: " + input_file.string());
        std::ifstream fasta_in(reference_file, std::ios::binary);
        if (!fasta_in) throw std::runtime_error("Failed to open reference FASTA file: " + reference_file.string());

        const bool create_leds = context_length > 0;
        if (create_leds) {
            std::cout << "VCF → l-EDS transformation (l=" << context_length << ")\n";
            std::cout << "  Using two-stage pipeline: VCF→EDS→l-EDS\n";
        } else {
            std::cout << "VCF → EDS transformation\n";
        }
        std::cout << "  Input: " << input_file << "\n";
        std::cout << "  Reference: " << reference_file << "\n";

        VCFStats stats;
        const auto result = create_leds ? parse_vcf_to_leds_streaming(vcf_in, fasta_in, context_length, &stats)
                                        : parse_vcf_to_eds_streaming(vcf_in, fasta_in, &stats);
        vcf_in.close();
        fasta_in.close();

        fs::path eds_path, seds_path;
        if (create_leds) {
            const std::string base_name = input_file.stem().string();
            const std::string suffix = "_l" + std::to_string(context_length);
            eds_path = output_file.empty() ? input_file.parent_path() / (base_name + suffix + ".leds") : output_file;
            seds_path = sources_file.empty() ? eds_path.parent_path() / (base_name + suffix + ".seds") : sources_file;
        } else {
            eds_path = output_file.empty() ? input_file.parent_path() / (input_file.stem().string() + ".eds") : output_file;
            seds_path = sources_file.empty() ? eds_path.parent_path() / (eds_path.stem().string() + ".seds") : sources_file;
        }
        std::ofstream eds_out(eds_path, std::ios::binary);
        if (!eds_out) throw std::runtime_error("Failed to open output file: " + eds_path.string());
        eds_out << result.first;
        eds_out.close();
        std::ofstream seds_out(seds_path, std::ios::binary);
        if (!seds_out) throw std::runtime_error("Failed to open sources file: " + seds_path.string());
        seds_out << result.second;
        seds_out.close();

        std::cout << "Transformation complete!\n";
        std::cout << "  Output: " << eds_path << "\n";
        std::cout << "  Sources: " << seds_path << "\n\n";
        std::cout << "Variant Processing Statistics:\n";
        std::cout << "  Total variants read:        " << stats.total_variants << "\n";
        std::cout << "  Successfully processed:     " << stats.processed_variants << "\n";
        std::cout << "  Skipped (malformed):        " << stats.skipped_malformed << "\n";
        std::cout << "  Skipped (unsupported SV):   " << stats.skipped_unsupported_sv << "\n";
        std::cout << "  Total skipped:              " << stats.total_skipped() << "\n";
        std::cout << "  Variant groups created:     " << stats.variant_groups << "\n";
        if (stats.total_variants > 0)
            std::cout << "  Success rate:               " << std::fixed << std::setprecision(1)
                      << (100.0 * stats.processed_variants) / stats.total_variants << "%\n";
        std::cout << "\n";
        cli::print_performance(timer);
        return 0;
    } catch (const std::exception& e) {
        std::cerr << "Error: " << e.what() << "\n";
        cli::print_performance(timer);
        return 1;
    }
}

int main(int argc, char** argv) { cli::finish(run(argc, argv)); }
