// TEST INFRASTRUCTURE ONLY (oracle/). Driver that links the UNMODIFIED
// reference library (compiled in place from /root/reference by oracle/Makefile
// into oracle/_ref/) and makes exactly the library calls the reference's CLI
// mains make (tools/msa2eds.cpp:124-131, tools/eds2leds.cpp:176-196,
// tools/vcf2eds.cpp:143-151). The reference's mains need Boost.program_options,
// which this image lacks; nothing else is replaced.
//
//   ref_driver msa2eds  <in.msa> <l> <out.eds> <out.seds>
//   ref_driver eds2leds <in.eds> <in.seds|-> <l> <out.leds> <out.seds|-> <threads> <compact 0|1>
//   ref_driver vcf2eds  <in.vcf> <ref.fa> <l> <out.eds> <out.seds>
//
// Each command prints "seconds=<wall seconds of the library call>" on stdout.
// Errors: prints "Error: <what>" on stderr, exit code 1 (as the CLIs do).
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>

#include "transforms/eds_transforms.hpp"
#include "transforms/msa_transforms.hpp"
#include "transforms/vcf_transforms.hpp"

namespace {

double now_s() {
    using clk = std::chrono::steady_clock;
    return std::chrono::duration<double>(clk::now().time_since_epoch()).count();
}

void spill(const std::string& path, const std::string& bytes) {
    std::ofstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot write " + path);
    f << bytes;
}

int usage() {
    std::cerr << "usage: ref_driver msa2eds|eds2leds|vcf2eds ... (see oracle/ref_driver.cpp)\n";
    return 2;
}

}  // namespace

int main(int argc, char** argv) {
    if (argc < 2) return usage();
    std::string cmd = argv[1];
    try {
        if (cmd == "msa2eds" && argc == 6) {
            std::ifstream in(argv[2]);
            if (!in) throw std::runtime_error(std::string("cannot open ") + argv[2]);
            size_t l = std::strtoull(argv[3], nullptr, 10);
            double t0 = now_s();
            auto res = l ? edsparser::parse_msa_to_leds_streaming(in, l)
                         : edsparser::parse_msa_to_eds_streaming(in);
            double t1 = now_s();
            spill(argv[4], res.first);
            spill(argv[5], res.second);
            std::cout << "seconds=" << (t1 - t0) << "\n";
            return 0;
        }
        if (cmd == "eds2leds" && argc == 9) {
            std::ifstream in(argv[2]);
            if (!in) throw std::runtime_error(std::string("cannot open ") + argv[2]);
            bool linear = std::strcmp(argv[3], "-") != 0;
            edsparser::Length l = static_cast<edsparser::Length>(std::strtoul(argv[4], nullptr, 10));
            size_t threads = std::strtoull(argv[7], nullptr, 10);
            bool compact = std::atoi(argv[8]) != 0;
            std::ostringstream out, sout;
            double t0 = now_s();
            if (linear) {
                std::ifstream sin(argv[3]);
                if (!sin) throw std::runtime_error(std::string("cannot open ") + argv[3]);
                edsparser::eds_to_leds_linear(in, out, l, &sin, &sout, threads, compact);
            } else {
                edsparser::eds_to_leds_cartesian(in, out, l, threads, compact);
            }
            double t1 = now_s();
            spill(argv[5], out.str());
            if (linear && std::strcmp(argv[6], "-") != 0) spill(argv[6], sout.str());
            std::cout << "seconds=" << (t1 - t0) << "\n";
            return 0;
        }
        if (cmd == "vcf2eds" && argc == 7) {
            std::ifstream vcf(argv[2]);
            std::ifstream fa(argv[3]);
            if (!vcf || !fa) throw std::runtime_error("cannot open vcf/fasta input");
            size_t l = std::strtoull(argv[4], nullptr, 10);
            edsparser::VCFStats stats;
            double t0 = now_s();
            auto res = l ? edsparser::parse_vcf_to_leds_streaming(vcf, fa, l, &stats)
                         : edsparser::parse_vcf_to_eds_streaming(vcf, fa, &stats);
            double t1 = now_s();
            spill(argv[5], res.first);
            spill(argv[6], res.second);
            std::cout << "seconds=" << (t1 - t0) << "\n";
            std::cout << "stats total=" << stats.total_variants << " processed=" << stats.processed_variants
                      << " malformed=" << stats.skipped_malformed << " sv=" << stats.skipped_unsupported_sv
                      << " groups=" << stats.variant_groups << "\n";
            return 0;
        }
    } catch (const std::exception& e) {
        std::cerr << "Error: " << e.what() << "\n";
        return 1;
    }
    return usage();
}
