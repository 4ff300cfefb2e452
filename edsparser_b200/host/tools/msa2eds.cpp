// msa2eds — CLI contract of the reference tool (src/cpp/tools/msa2eds.cpp:12-190): options, extension check,
// output naming, stdout lines and the [Performance] line on stderr; the transform itself runs on the GPU.
#include <filesystem>
#include <fstream>

#include "cli_common.hpp"
#include "edsparser/transforms/eds_transforms.hpp"
#include "edsparser/transforms/msa_transforms.hpp"

using namespace edsparser;
namespace fs = std::filesystem;

static void usage() {
    std::cout << "msa2eds - Transform MSA (Multiple Sequence Alignment) to EDS\n\n"
                 "Transform MSA (Multiple Sequence Alignment) to EDS/l-EDS:\n"
                 "  -h [ --help ]                     Show help message\n"
                 "  -i [ --input ] arg                Input MSA file (.msa) in FASTA format with gaps as '-'\n"
                 "  -o [ --output ] arg               Output EDS file (default: <input>.eds)\n"
                 "  -s [ --sources ] arg              Output source file (default: <output>.seds)\n"
                 "  -l [ --context-length ] arg (=0)  Create l-EDS with minimum context length (0 = regular EDS)\n"
                 "  --device arg (=0)                 CUDA device to run on (B200 build)\n"
                 "  --gpus arg (=1)                   Shard the alignment's columns over this many GPUs of the node,\n"
                 "                                    starting at --device (NCCL exchange of the output offsets)\n\n"
                 "OUTPUT:\n"
                 "  Regular EDS:     <input_base>.eds, <input_base>.seds\n"
                 "  l-EDS (with -l): <input_base>_l<N>.leds, <input_base>_l<N>.seds\n\n";
}

static int run(int argc, char** argv) {
    Timer timer;
    timer.start();
    try {
        const cli::Args args(argc, argv, {{"help", 'h', false}, {"input", 'i', true}, {"output", 'o', true},
                                         {"sources", 's', true}, {"context-length", 'l', true}, {"device", 0, true}, {"gpus", 0, true}});
        if (args.has("help")) {
            usage();
            cli::print_performance(timer);
            return 0;
        }
        args.require("input");
        const fs::path input_file = args.get("input");
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

        if (input_file.extension() != ".msa") {
            std::cerr << "Error: Input file must be an MSA file (.msa)\n";
            std::cerr << "Got: " << input_file << "\n";
            cli::print_performance(timer);
            return 1;
        }
        std::ifstream msa_in(input_file, std::ios::binary);
        if (!msa_in) throw std::runtime_error("Failed to open input file: " + input_file.string());

        const bool create_leds = context_length > 0;
        if (create_leds)
            std::cout << "MSA → l-EDS transformation (l=" << context_length << ")\n";
        else
            std::cout << "MSA → EDS transformation\n";
        std::cout << "  Input: " << input_file << "\n";

        const auto result = create_leds ? parse_msa_to_leds_streaming(msa_in, context_length) : parse_msa_to_eds_streaming(msa_in);
        msa_in.close();

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
        std::cout << "  Sources: " << seds_path << "\n";
        cli::print_performance(timer);
        return 0;
    } catch (const std::exception& e) {
        std::cerr << "Error: " << e.what() << "\n";
        cli::print_performance(timer);
        return 1;
    }
}

int main(int argc, char** argv) { cli::finish(run(argc, argv)); }
