// eds2leds — CLI contract of the reference tool (src/cpp/tools/eds2leds.cpp:12-218): sources given => LINEAR
// (phasing-aware) merging, else CARTESIAN; compact output unless --full; --threads is accepted and reported
// but the merge runs on the GPU.
#include <filesystem>
#include <fstream>
#include <vector>

#include "cli_common.hpp"
#include "edsparser/transforms/eds_transforms.hpp"

using namespace edsparser;
namespace fs = std::filesystem;

static void usage() {
    std::cout << "eds2leds - Transform EDS to l-EDS (length-constrained EDS)\n\n"
                 "Transform EDS to l-EDS (length-constrained EDS):\n"
                 "  -h [ --help ]                Show help message\n"
                 "  -i [ --input ] arg           Input EDS file (.eds)\n"
                 "  -o [ --output ] arg          Output l-EDS file (default: <input>_l<N>.leds)\n"
                 "  -l [ --context-length ] arg  Minimum context length\n"
                 "  -s [ --sources ] arg         Input source file (.seds) for linear (phasing-aware) merging\n"
                 "  --full                       Use full output format with brackets on all symbols (default: compact)\n"
                 "  -t [ --threads ] arg (=1)    Number of threads for parallel processing (accepted; the merge runs on the GPU)\n"
                 "  --device arg (=0)            CUDA device to run on (B200 build)\n"
                 "  --gpus arg (=1)              Merge symbol ranges on this many GPUs of the node, starting at --device\n"
                 "  --max-output-bytes arg (=0)  Refuse (do not truncate) a merge whose output would exceed this (0 = no limit)\n\n"
                 "MERGING METHODS (auto-detected):\n"
                 "  LINEAR:    used when --sources/-s is provided; keeps only combinations with a common source\n"
                 "  CARTESIAN: used when no source file is provided; cross-product of alternatives\n\n";
}

static int run(int argc, char** argv) {
    Timer timer;
    timer.start();
    try {
        const cli::Args args(argc, argv, {{"help", 'h', false}, {"input", 'i', true}, {"output", 'o', true},
                                         {"context-length", 'l', true}, {"sources", 's', true}, {"full", 0, false},
                                         {"threads", 't', true}, {"device", 0, true}, {"gpus", 0, true}, {"max-output-bytes", 0, true}});
        if (args.has("help")) {
            usage();
            cli::print_performance(timer);
            return 0;
        }
        args.require("input");
        args.require("context-length");
        const fs::path input_file = args.get("input");
        fs::path output_file = args.has("output") ? fs::path(args.get("output")) : fs::path();
        const fs::path sources_file = args.has("sources") ? fs::path(args.get("sources")) : fs::path();
        const unsigned long l_arg = args.to_uint("context-length");
        if (l_arg > 0xfffffffful) throw std::invalid_argument("the argument for option '--context-length' is invalid");
        const Length context_length = (Length)l_arg;
        const bool compact_mode = !args.has("full");
        long num_threads = 1;
        if (args.has("threads")) {
            try {
                num_threads = std::stol(args.get("threads"));
            } catch (const std::exception&) {
                throw std::invalid_argument("the argument ('" + args.get("threads") + "') for option '--threads' is invalid");
            }
        }
        const int first_device = args.has("device") ? (int)args.to_uint("device") : 0;
        b200::set_device(first_device);
        if (args.has("gpus")) {
            const unsigned long n = args.to_uint("gpus");
            if (n < 1 || n > 64) throw std::invalid_argument("the argument for option '--gpus' is invalid");
            std::vector<int> devices;
            for (unsigned long i = 0; i < n; ++i) devices.push_back(first_device + (int)i);
            b200::set_devices(devices);
        }
        if (args.has("max-output-bytes")) b200::set_max_output_bytes(args.to_uint("max-output-bytes"));

        if (input_file.extension() != ".eds") {
            std::cerr << "Error: Input file must be an EDS file (.eds)\n";
            std::cerr << "Got: " << input_file << "\n";
            cli::print_performance(timer);
            return 1;
        }
        if (num_threads < 1) {
            std::cerr << "Error: Number of threads must be >= 1\n";
            cli::print_performance(timer);
            return 1;
        }
        if (context_length == 0) {
            std::cerr << "Error: Context length must be > 0\n";
            cli::print_performance(timer);
            return 1;
        }
        if (output_file.empty())
            output_file = input_file.parent_path() / (input_file.stem().string() + "_l" + std::to_string(context_length) + ".leds");

        std::cout << "EDS → l-EDS transformation\n";
        std::cout << "  Input: " << input_file << "\n";
        std::cout << "  Output: " << output_file << "\n";
        std::cout << "  Context length: " << context_length << "\n";
        if (!sources_file.empty()) std::cout << "  Sources: " << sources_file << "\n";
        std::cout << "  Output mode: " << (compact_mode ? "compact" : "full") << "\n";
        std::cout << "  Threads: " << num_threads << (num_threads == 1 ? " (sequential)" : " (parallel)") << "\n";

        std::ifstream input(input_file, std::ios::binary);
        if (!input) throw std::runtime_error("Cannot open input file: " + input_file.string());
        std::ofstream output(output_file, std::ios::binary);
        if (!output) throw std::runtime_error("Cannot open output file: " + output_file.string());

        if (!sources_file.empty()) {
            std::ifstream sources_in(sources_file, std::ios::binary);
            if (!sources_in) throw std::runtime_error("Cannot open sources file: " + sources_file.string());
            fs::path output_sources = output_file;
            output_sources.replace_extension(".seds");
            std::ofstream sources_out(output_sources, std::ios::binary);
            if (!sources_out) throw std::runtime_error("Cannot create output sources file: " + output_sources.string());
            std::cout << "  Output sources: " << output_sources << "\n";
            eds_to_leds_linear(input, output, context_length, &sources_in, &sources_out, (size_t)num_threads, compact_mode);
        } else {
            eds_to_leds_cartesian(input, output, context_length, (size_t)num_threads, compact_mode);
        }
        std::cout << "Transformation complete!\n";
        cli::print_performance(timer);
        return 0;
    } catch (const std::exception& e) {
        std::cerr << "Error: " << e.what() << "\n";
        cli::print_performance(timer);
        return 1;
    }
}

int main(int argc, char** argv) { cli::finish(run(argc, argv)); }
