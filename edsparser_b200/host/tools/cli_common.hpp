// Option parsing shared by the tools. The reference uses boost::program_options (absent here); this accepts the
// same spellings: --name value, --name=value, -n value, -nvalue, boolean switches.
#ifndef EDSPARSER_B200_CLI_COMMON_HPP
#define EDSPARSER_B200_CLI_COMMON_HPP

#include <iomanip>
#include <iostream>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "edsparser/common.hpp"

namespace cli {

struct Option {
    std::string long_name;
    char short_name;  // 0 = none
    bool takes_value;
};

class Args {
   public:
    Args(int argc, char** argv, const std::vector<Option>& options) {
        for (int i = 1; i < argc; ++i) {
            const std::string a = argv[i];
            const Option* opt = nullptr;
            std::string value;
            bool has_value = false;
            if (a.rfind("--", 0) == 0) {
                const size_t eq = a.find('=');
                const std::string name = a.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
                for (const Option& o : options)
                    if (o.long_name == name) opt = &o;
                if (!opt) throw std::invalid_argument("unrecognised option '" + a + "'");
                if (eq != std::string::npos) {
                    value = a.substr(eq + 1);
                    has_value = true;
                }
            } else if (a.size() >= 2 && a[0] == '-') {
                for (const Option& o : options)
                    if (o.short_name && o.short_name == a[1]) opt = &o;
                if (!opt) throw std::invalid_argument("unrecognised option '" + a + "'");
                if (a.size() > 2) {
                    value = a.substr(2);
                    has_value = true;
                }
            } else {
                throw std::invalid_argument("too many positional options have been specified on the command line");
            }
            if (opt->takes_value && !has_value) {
                if (i + 1 >= argc) throw std::invalid_argument("the required argument for option '--" + opt->long_name + "' is missing");
                value = argv[++i];
            }
            values_[opt->long_name] = value;
        }
    }
    bool has(const std::string& name) const { return values_.count(name) != 0; }
    std::string get(const std::string& name) const { return values_.at(name); }
    void require(const std::string& name) const {
        if (!has(name)) throw std::invalid_argument("the option '--" + name + "' is required but missing");
    }
    unsigned long to_uint(const std::string& name) const {
        const std::string v = get(name);
        size_t used = 0;
        unsigned long out = 0;
        try {
            if (!v.empty() && v[0] == '-') throw std::invalid_argument(v);
            out = std::stoul(v, &used);
        } catch (const std::exception&) {
            used = 0;
        }
        if (used != v.size() || v.empty()) throw std::invalid_argument("the argument ('" + v + "') for option '--" + name + "' is invalid");
        return out;
    }

   private:
    std::map<std::string, std::string> values_;
};

// "[Performance] Runtime: X.XXs | Peak Memory: Y.Y MB" on stderr, as every reference tool prints on every exit path
inline void print_performance(edsparser::Timer& timer) {
    timer.stop();
    const double memory_mb = edsparser::get_peak_memory_mb();
    std::cerr << "[Performance] Runtime: " << std::fixed << std::setprecision(2) << timer.elapsed_seconds() << "s";
    if (memory_mb > 0.0) std::cerr << " | Peak Memory: " << std::fixed << std::setprecision(1) << memory_mb << " MB";
    std::cerr << "\n";
}

// End of a tool: everything the tool wrote is flushed (its file streams are closed by now), then the process leaves
// WITHOUT running static destructors — tearing the CUDA context down costs 0.3 - 0.4 s of a run whose device work takes
// milliseconds, and the operating system reclaims it all anyway.
[[noreturn]] inline void finish(int rc) {
    std::cout.flush();
    std::cerr.flush();
    fflush(nullptr);
    _Exit(rc);
}

}  // namespace cli
#endif
