// Host layer: the reference's transforms API on top of the C ABI (include/edsparser_b200.h). Status codes
// are turned back into the exception classes the reference throws. No compute happens here.
#include <fstream>
#include <iterator>
#include <mutex>
#include <sstream>
#include <stdexcept>

#include "../../../include/edsparser_b200.h"
#include "edsparser/common.hpp"
#include "edsparser/formats/eds.hpp"
#include "edsparser/transforms/eds_transforms.hpp"
#include "edsparser/transforms/msa_transforms.hpp"

namespace edsparser {

double get_peak_memory_mb() {
    std::ifstream status("/proc/self/status");
    std::string line;
    while (std::getline(status, line)) {
        if (line.compare(0, 6, "VmHWM:") == 0) {
            std::istringstream in(line.substr(6));
            double value = 0;
            std::string unit;
            in >> value >> unit;
            return unit == "kB" ? value / 1024.0 : value;
        }
    }
    return 0.0;
}

namespace {

int g_device = 0;
uint64_t g_budget = 0;

[[noreturn]] void rethrow(eds_status rc) {
    const std::string what = eds_last_error();
    switch (rc) {
        case EDS_ERR_INVALID_ARGUMENT: throw std::invalid_argument(what);
        case EDS_ERR_OUT_OF_RANGE: throw std::out_of_range(what);
        default: throw std::runtime_error(what);
    }
}

// one context per thread: the library's contexts are not thread-safe, the reference's functions are re-entrant
struct Session {
    eds_ctx* ctx = nullptr;
    int device = -1;
    ~Session() { eds_ctx_destroy(ctx); }
    eds_ctx* get() {
        if (!ctx || device != g_device) {
            eds_ctx_destroy(ctx);
            ctx = nullptr;
            const eds_status rc = eds_ctx_create(g_device, nullptr, &ctx);
            if (rc != EDS_OK) rethrow(rc);
            device = g_device;
        }
        return ctx;
    }
};
thread_local Session t_session;

std::string slurp(std::istream& in) { return std::string(std::istreambuf_iterator<char>(in), std::istreambuf_iterator<char>()); }

struct HostBuf {
    eds_buffer b{nullptr, 0};
    ~HostBuf() { eds_buffer_free_host(&b); }
    std::string str() const { return std::string(reinterpret_cast<const char*>(b.data), b.bytes); }
};

std::pair<std::string, std::string> msa_transform(std::istream& in, uint32_t l, int leds) {
    const std::string file = slurp(in);
    HostBuf e, s;
    const eds_status rc = eds_msa_transform_host(t_session.get(), reinterpret_cast<const uint8_t*>(file.data()), file.size(), l, leds,
                                                 &e.b, &s.b, nullptr);
    if (rc != EDS_OK) rethrow(rc);
    return {e.str(), s.str()};
}

void merge(std::istream& input, std::ostream& output, Length l, std::istream* phasing_in, std::ostream* phasing_out, bool compact) {
    if (l == 0) throw std::invalid_argument("context_length must be > 0 for l-EDS transformation");
    const std::string eds = slurp(input);
    std::string seds;
    if (phasing_in) seds = slurp(*phasing_in);
    HostBuf o, so;
    uint32_t rounds = 0;
    const eds_status rc = eds_leds_merge_host(t_session.get(), reinterpret_cast<const uint8_t*>(eds.data()), eds.size(),
                                              phasing_in ? reinterpret_cast<const uint8_t*>(seds.data()) : nullptr, seds.size(), l,
                                              compact ? 1 : 0, g_budget, &o.b, &so.b, &rounds);
    if (rc != EDS_OK) rethrow(rc);
    output.write(reinterpret_cast<const char*>(o.b.data), (std::streamsize)o.b.bytes);
    if (phasing_in && phasing_out) phasing_out->write(reinterpret_cast<const char*>(so.b.data), (std::streamsize)so.b.bytes);
}

}  // namespace

namespace b200 {
void set_device(int device) { g_device = device; }
void set_max_output_bytes(uint64_t bytes) { g_budget = bytes; }
}  // namespace b200

std::pair<std::string, std::string> parse_msa_to_eds_streaming(std::istream& msa_stream) { return msa_transform(msa_stream, 0, 0); }

std::pair<std::string, std::string> parse_msa_to_leds_streaming(std::istream& msa_stream, size_t context_length) {
    if (context_length > 0xffffffffull) context_length = 0xffffffffull;  // longer than any alignment the device holds
    return msa_transform(msa_stream, (uint32_t)context_length, 1);
}

void eds_to_leds_linear(std::istream& input, std::ostream& output, Length context_length, std::istream* phasing_input,
                        std::ostream* phasing_output, size_t /*num_threads*/, bool compact) {
    // without a phasing stream the reference's linear entry point degenerates to all combinations kept
    merge(input, output, context_length, phasing_input, phasing_output, compact);
}

void eds_to_leds_cartesian(std::istream& input, std::ostream& output, Length context_length, size_t /*num_threads*/, bool compact) {
    merge(input, output, context_length, nullptr, nullptr, compact);
}

EDS::EDS(std::istream& eds_stream) : text_(slurp(eds_stream)) {}
EDS::EDS(std::istream& eds_stream, std::istream& sources_stream)
    : text_(slurp(eds_stream)), sources_(slurp(sources_stream)), has_sources_(true) {}
EDS::EDS(const std::string& eds_text) : text_(eds_text) {}
EDS::EDS(const std::string& eds_text, const std::string& sources_text) : text_(eds_text), sources_(sources_text), has_sources_(true) {}

bool is_leds(const EDS& eds, Length context_length) {
    if (context_length == 0) return true;  // eds_transforms.cpp:440-442
    int answer = 0;
    const eds_status rc = eds_is_leds_host(t_session.get(), reinterpret_cast<const uint8_t*>(eds.text().data()), eds.text().size(),
                                           context_length, &answer);
    if (rc != EDS_OK) rethrow(rc);
    return answer != 0;
}

}  // namespace edsparser
