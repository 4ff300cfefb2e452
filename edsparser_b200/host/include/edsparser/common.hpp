// Host-side mirror of the reference's common.hpp (src/cpp/lib/common.hpp:12-72): the typedefs, format
// constants, Timer and get_peak_memory_mb that the transforms API and the CLI tools use.
#ifndef EDSPARSER_COMMON_HPP
#define EDSPARSER_COMMON_HPP

#include <chrono>
#include <cstdint>
#include <string>
#include <vector>

namespace edsparser {

constexpr const char* VERSION = "1.0.0-b200";

using String = std::string;
using StringSet = std::vector<String>;
using Position = uint64_t;
using Length = uint32_t;

constexpr char SET_OPEN = '{';
constexpr char SET_CLOSE = '}';
constexpr char SET_SEPARATOR = ',';

constexpr const char* EXT_MSA = ".msa";
constexpr const char* EXT_EDS = ".eds";
constexpr const char* EXT_SEDS = ".seds";
constexpr const char* EXT_LEDS = ".leds";

// wall-clock timer behind the "[Performance] Runtime" line of every tool (common.cpp:10-41)
class Timer {
   public:
    void start() {
        begin_ = clock::now();
        running_ = true;
    }
    void stop() {
        end_ = clock::now();
        running_ = false;
    }
    double elapsed_seconds() const {
        return std::chrono::duration<double>((running_ ? clock::now() : end_) - begin_).count();
    }
    double elapsed_milliseconds() const { return elapsed_seconds() * 1e3; }
    double elapsed_microseconds() const { return elapsed_seconds() * 1e6; }

   private:
    using clock = std::chrono::steady_clock;
    clock::time_point begin_{}, end_{};
    bool running_ = false;
};

// peak resident set size in MB (VmHWM of /proc/self/status), 0 when unavailable (common.cpp:44-77)
double get_peak_memory_mb();

}  // namespace edsparser
#endif
