// Host layer: the reference's transforms API on top of the C ABI (include/edsparser_b200.h). Status codes
// are turned back into the exception classes the reference throws. No compute happens here.
#include <cctype>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <iterator>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/edsparser_b200.h"
#include "edsparser/common.hpp"
#include "edsparser/formats/eds.hpp"
#include "edsparser/transforms/eds_transforms.hpp"
#include "edsparser/transforms/msa_transforms.hpp"
#include "edsparser/transforms/vcf_transforms.hpp"

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
std::vector<int> g_devices;  // more than one entry: msa2eds is column-sharded over these devices (eds_group)

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

// one group per thread and device list, like the per-thread context
struct GroupSession {
    eds_group* group = nullptr;
    std::vector<int> devices;
    ~GroupSession() { eds_group_destroy(group); }
    eds_group* get() {
        if (!group || devices != g_devices) {
            eds_group_destroy(group);
            group = nullptr;
            const eds_status rc = eds_group_create(g_devices.data(), (int)g_devices.size(), &group);
            if (rc != EDS_OK) rethrow(rc);
            devices = g_devices;
        }
        return group;
    }
};
thread_local GroupSession t_group;

std::pair<std::string, std::string> msa_transform(std::istream& in, uint32_t l, int leds) {
    const std::string file = slurp(in);
    if (g_devices.size() > 1) {
        // column-sharded over the devices: every device transforms its column range (+ halo), the byte counts are
        // all-gathered over NCCL inside the library and every device writes its slice of the one result
        eds_buffer e{nullptr, 0}, s{nullptr, 0};
        const eds_status rc = eds_group_msa_transform_host(t_group.get(), reinterpret_cast<const uint8_t*>(file.data()), file.size(), l, leds,
                                                           0, &e, &s, nullptr);
        if (rc != EDS_OK) rethrow(rc);
        std::pair<std::string, std::string> out{std::string(reinterpret_cast<const char*>(e.data), e.bytes),
                                                std::string(reinterpret_cast<const char*>(s.data), s.bytes)};
        eds_buffer_free_host(&e);
        eds_buffer_free_host(&s);
        return out;
    }
    eds_buffer e{nullptr, 0}, s{nullptr, 0};  // views into pinned memory kept by the context: nothing to free
    const eds_status rc = eds_msa_transform_host_view(t_session.get(), reinterpret_cast<const uint8_t*>(file.data()), file.size(), l, leds,
                                                      &e, &s, nullptr);
    if (rc != EDS_OK) rethrow(rc);
    return {std::string(reinterpret_cast<const char*>(e.data), e.bytes), std::string(reinterpret_cast<const char*>(s.data), s.bytes)};
}

void merge(std::istream& input, std::ostream& output, Length l, std::istream* phasing_in, std::ostream* phasing_out, bool compact) {
    if (l == 0) throw std::invalid_argument("context_length must be > 0 for l-EDS transformation");
    const std::string eds = slurp(input);
    std::string seds;
    if (phasing_in) seds = slurp(*phasing_in);
    eds_buffer o{nullptr, 0}, so{nullptr, 0};  // views into pinned memory kept by the context: nothing to free
    uint32_t rounds = 0;
    const eds_status rc = eds_leds_merge_host_view(t_session.get(), reinterpret_cast<const uint8_t*>(eds.data()), eds.size(),
                                                   phasing_in ? reinterpret_cast<const uint8_t*>(seds.data()) : nullptr, seds.size(), l,
                                                   compact ? 1 : 0, g_budget, &o, &so, &rounds);
    if (rc != EDS_OK) rethrow(rc);
    output.write(reinterpret_cast<const char*>(o.data), (std::streamsize)o.bytes);
    if (phasing_in && phasing_out) phasing_out->write(reinterpret_cast<const char*>(so.data), (std::streamsize)so.bytes);
}

// "Warning: Skipping variant at CHROM:POS - Unsupported structural variant type: X" (vcf_transforms.cpp:299-305) for a
// record the device skipped; the fields are split the way parse_vcf_line does (:257-279).
void warn_skipped_sv(const std::string& vcf, uint64_t line_at) {
    size_t end = vcf.find('\n', line_at);
    if (end == std::string::npos) end = vcf.size();
    const std::string line = vcf.substr(line_at, end - line_at);
    std::vector<std::string> f;
    size_t at = 0;
    while (at < line.size()) {
        size_t d = line.find('\t', at);
        if (d == std::string::npos) d = line.size();
        if (d > at) f.push_back(line.substr(at, d - at));
        at = d + 1;
    }
    if (f.size() < 5) {
        f.clear();
        std::istringstream ss(line);
        std::string tok;
        while (ss >> tok) f.push_back(tok);
    }
    if (f.size() < 5) return;
    unsigned long long pos = 0;
    try {
        pos = std::stoull(f[1]);
    } catch (...) {
        return;
    }
    std::istringstream alts(f[4]);
    std::string a;
    while (std::getline(alts, a, ',')) {
        if (a.size() >= 2 && a.front() == '<' && a.back() == '>') {
            const std::string kind = a.substr(1, a.size() - 2);
            if (kind != "DEL" && kind != "INS") {
                std::cerr << "Warning: Skipping variant at " << f[0] << ":" << pos << " - Unsupported structural variant type: " << kind
                          << std::endl;
                return;
            }
        }
    }
}

std::pair<std::string, std::string> vcf_transform(std::istream& vcf_stream, std::istream& fasta_stream, uint32_t l, VCFStats* stats) {
    // the reference reads the FASTA first (vcf_transforms.cpp:683), so its errors win; both streams end up consumed
    const std::string fasta = slurp(fasta_stream);
    const std::string vcf = slurp(vcf_stream);
    eds_buffer e{nullptr, 0}, s{nullptr, 0};  // views into pinned memory kept by the context: nothing to free
    eds_vcf_stats st{};
    uint64_t* sv = nullptr;
    uint64_t n_sv = 0;
    const eds_status rc = eds_vcf_transform_host_view(t_session.get(), reinterpret_cast<const uint8_t*>(vcf.data()), vcf.size(),
                                                      reinterpret_cast<const uint8_t*>(fasta.data()), fasta.size(), l, &e, &s, &st, &sv, &n_sv);
    for (uint64_t i = 0; i < n_sv; ++i) warn_skipped_sv(vcf, sv[i]);
    std::free(sv);
    if (stats && (rc == EDS_OK || st.total_variants)) {  // the reference fills the counters before the merge can throw
        stats->total_variants += st.total_variants;
        stats->processed_variants += st.processed_variants;
        stats->skipped_malformed += st.skipped_malformed;
        stats->skipped_unsupported_sv += st.skipped_unsupported_sv;
        stats->variant_groups = st.variant_groups;
    }
    if (rc != EDS_OK) rethrow(rc);
    return {std::string(reinterpret_cast<const char*>(e.data), e.bytes), std::string(reinterpret_cast<const char*>(s.data), s.bytes)};
}

}  // namespace

namespace b200 {
void set_device(int device) { g_device = device; }
void set_devices(const std::vector<int>& devices) {
    g_devices = devices;
    if (!devices.empty()) g_device = devices[0];
}
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

std::pair<std::string, std::string> parse_vcf_to_eds_streaming(std::istream& vcf_stream, std::istream& fasta_stream, VCFStats* stats) {
    return vcf_transform(vcf_stream, fasta_stream, 0, stats);
}

std::pair<std::string, std::string> parse_vcf_to_leds_streaming(std::istream& vcf_stream, std::istream& fasta_stream,
                                                                size_t context_length, VCFStats* stats) {
    if (context_length == 0) {  // eds_to_leds_linear refuses l = 0 after the EDS has been built (vcf_transforms.cpp:742-752)
        vcf_transform(vcf_stream, fasta_stream, 0, stats);
        throw std::invalid_argument("context_length must be > 0 for l-EDS transformation");
    }
    if (context_length > 0xffffffffull) context_length = 0xffffffffull;
    return vcf_transform(vcf_stream, fasta_stream, (uint32_t)context_length, stats);
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
