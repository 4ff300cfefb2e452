// Host layer: the reference's transforms API on top of the C ABI (include/edsparser_b200.h). Status codes
// are turned back into the exception classes the reference throws. No compute happens here.
#include <cctype>
#include <cstdlib>
#include <filesystem>
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

// everything from the stream's position to its end. A seekable stream (the tools' ifstreams) is read with ONE read() of
// the known size: character-by-character iteration moves a gigabyte at a few hundred MB/s
std::string slurp(std::istream& in) {
    const std::istream::pos_type at = in.tellg();
    if (at != std::istream::pos_type(-1) && in.seekg(0, std::ios::end)) {
        const std::istream::pos_type end = in.tellg();
        in.seekg(at);
        if (end != std::istream::pos_type(-1) && end >= at && in) {
            std::string out((size_t)(end - at), '\0');
            in.read(&out[0], (std::streamsize)out.size());
            out.resize((size_t)in.gcount());
            in.peek();  // as after an iteration to the end: eofbit set
            return out;
        }
    }
    in.clear();
    return std::string(std::istreambuf_iterator<char>(in), std::istreambuf_iterator<char>());
}


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
    if (g_devices.size() > 1 && g_budget == 0) {
        // symbol ranges over the devices of the group (cut inside long conserved symbols, verified, glued back together)
        eds_buffer ge{nullptr, 0}, gs{nullptr, 0};
        uint32_t r = 0, used = 0;
        const eds_status grc = eds_group_leds_merge_host(t_group.get(), reinterpret_cast<const uint8_t*>(eds.data()), eds.size(),
                                                         phasing_in ? reinterpret_cast<const uint8_t*>(seds.data()) : nullptr, seds.size(), l,
                                                         compact ? 1 : 0, &ge, &gs, &r, &used);
        if (grc != EDS_OK) rethrow(grc);
        output.write(reinterpret_cast<const char*>(ge.data), (std::streamsize)ge.bytes);
        if (phasing_in && phasing_out) phasing_out->write(reinterpret_cast<const char*>(gs.data), (std::streamsize)gs.bytes);
        eds_buffer_free_host(&ge);
        eds_buffer_free_host(&gs);
        return;
    }
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
    const bool sharded = g_devices.size() > 1;  // slices of the record lines over the devices of the group (malloc'd results)
    const eds_status rc =
        sharded ? eds_group_vcf_transform_host(t_group.get(), reinterpret_cast<const uint8_t*>(vcf.data()), vcf.size(),
                                               reinterpret_cast<const uint8_t*>(fasta.data()), fasta.size(), l, &e, &s, &st, &sv, &n_sv, nullptr)
                : eds_vcf_transform_host_view(t_session.get(), reinterpret_cast<const uint8_t*>(vcf.data()), vcf.size(),
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
    std::pair<std::string, std::string> out{std::string(reinterpret_cast<const char*>(e.data), e.bytes),
                                            std::string(reinterpret_cast<const char*>(s.data), s.bytes)};
    if (sharded) {
        eds_buffer_free_host(&e);
        eds_buffer_free_host(&s);
    }
    return out;
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

// ---- class EDS: a host-side view of what the device parsed (eds_parse_host) ------------------------------------------
void EDS::build(const std::string& eds_text, const std::string* seds_text) {
    eds_parsed p;
    const eds_status rc =
        eds_parse_host(t_session.get(), reinterpret_cast<const uint8_t*>(eds_text.data()), eds_text.size(),
                       seds_text ? reinterpret_cast<const uint8_t*>(seds_text->data()) : nullptr, seds_text ? seds_text->size() : 0, &p);
    if (rc != EDS_OK) rethrow(rc);
    struct Release {
        eds_parsed* p;
        ~Release() { eds_parsed_free(p); }
    } release{&p};
    n_ = p.n_symbols;
    m_ = p.n_strings;
    N_ = p.total_chars;
    is_empty_ = p.n_symbols == 0;
    sets_.assign(n_, StringSet());
    Metadata& md = metadata_;
    md = Metadata();
    md.symbol_sizes.resize(n_);
    md.cum_set_sizes.resize(n_);
    md.is_degenerate.resize(n_);
    md.string_lengths.resize(m_);
    md.cum_common_positions.assign(1, 0);
    md.cum_degenerate_counts.assign(1, 0);
    const char* text = reinterpret_cast<const char*>(p.text);
    Position common = 0;
    int degenerate_strings = 0;
    for (size_t i = 0; i < n_; ++i) {
        const uint32_t first = p.sym_first[i], count = p.sym_first[i + 1] - first;
        md.symbol_sizes[i] = count;
        md.cum_set_sizes[i] = first;
        md.is_degenerate[i] = count > 1;
        StringSet& set = sets_[i];
        set.reserve(count);
        for (uint32_t j = first; j < first + count; ++j) {
            set.emplace_back(text + p.str_start[j], p.str_end[j] - p.str_start[j]);
            md.string_lengths[j] = p.str_end[j] - p.str_start[j];
        }
        if (count > 1)
            degenerate_strings += (int)count;
        else
            common += md.string_lengths[first];
        md.cum_common_positions.push_back(common);
        md.cum_degenerate_counts.push_back(degenerate_strings);
    }
    if (is_empty_) {
        md.cum_common_positions.clear();
        md.cum_degenerate_counts.clear();
    }
    md.min_context_length = p.min_context_length;
    md.max_context_length = p.max_context_length;
    md.avg_context_length = p.num_context_blocks ? (double)p.sum_context_length / (double)p.num_context_blocks : 0.0;
    md.num_degenerate_symbols = p.num_degenerate_symbols;
    md.num_common_chars = p.num_common_chars;
    md.total_change_size = p.total_change_size;
    md.num_empty_strings = p.num_empty_strings;
    has_sources_ = p.has_sources != 0;
    sources_.clear();
    if (has_sources_) {
        sources_.resize(m_);
        for (size_t j = 0; j < m_; ++j) sources_[j].insert(p.src_ids + p.src_off[j], p.src_ids + p.src_off[j + 1]);
        md.num_paths = p.num_paths;
        md.max_paths_per_string = p.max_paths_per_string;
        md.avg_paths_per_string = m_ ? (double)p.total_paths / (double)m_ : 0.0;
    }
}

EDS::EDS(std::istream& eds_stream) { build(slurp(eds_stream), nullptr); }
EDS::EDS(std::istream& eds_stream, std::istream& seds_stream) {
    const std::string e = slurp(eds_stream), s = slurp(seds_stream);
    build(e, &s);
}
EDS::EDS(const std::string& eds_string) { build(eds_string, nullptr); }
EDS::EDS(const std::string& eds_string, const std::string& seds_string) { build(eds_string, &seds_string); }

EDS EDS::from_string(const std::string& eds_string) { return EDS(eds_string); }
EDS EDS::from_string(const std::string& eds_string, const std::string& seds_string) { return EDS(eds_string, seds_string); }

namespace {
std::string read_file(const std::filesystem::path& path, const char* what) {
    std::ifstream in(path, std::ios::binary);
    if (!in) throw std::runtime_error(std::string("Failed to open ") + what + " file: " + path.string());
    return slurp(in);
}
}  // namespace

EDS EDS::load(const std::filesystem::path& path, StoringMode mode) {
    EDS out(read_file(path, "EDS"));
    out.mode_ = mode;
    return out;
}

EDS EDS::load(const std::filesystem::path& eds_path, const std::filesystem::path& seds_path, StoringMode mode) {
    EDS out(read_file(eds_path, "EDS"), read_file(seds_path, "sEDS"));
    out.mode_ = mode;
    return out;
}

EDS::Statistics EDS::get_statistics() const {
    const Metadata& md = metadata_;
    return Statistics{md.min_context_length, md.max_context_length,  md.avg_context_length,   md.num_degenerate_symbols,
                      md.num_common_chars,   md.total_change_size,   md.num_empty_strings,    md.num_paths,
                      md.max_paths_per_string, md.avg_paths_per_string};
}

void EDS::print_statistics(std::ostream& os) const {
    const Statistics st = get_statistics();
    const char* rule = "========================================\n";
    auto row = [&os](const char* label, auto value) {
        std::string l(label);
        l.resize(32, ' ');  // the reference pads its labels to column 32
        os << l << value << "\n";
    };
    os << rule << "EDS Statistics\n" << rule << "Structure:\n";
    row("  Number of sets (n):", n_);
    row("  Total characters (N):", N_);
    row("  Total strings (m):", m_);
    row("  Degenerate symbols:", st.num_degenerate_symbols);
    row("  Regular symbols:", n_ - st.num_degenerate_symbols);
    os << "\nContext Lengths:\n";
    row("  Minimum:", st.min_context_length);
    row("  Maximum:", st.max_context_length);
    row("  Average:", st.avg_context_length);
    os << "\nVariations:\n";
    row("  Total change size:", st.total_change_size);
    row("  Common characters:", st.num_common_chars);
    row("  Empty strings:", st.num_empty_strings);
    os << "\n";
    if (has_sources_)
        os << "Sources: Loaded (" << sources_.size() << " strings with source info)\n";
    else
        os << "Sources: Not loaded\n";
    os << rule;
}

void EDS::print(std::ostream& os) const {
    if (mode_ == StoringMode::METADATA_ONLY)
        throw std::runtime_error("Cannot print EDS in METADATA_ONLY mode. Load with StoringMode::FULL to access string data for printing.");
    if (is_empty_) {
        os << "(empty EDS)\n";
        return;
    }
    os << "EDS with " << n_ << " sets, " << m_ << " total strings:\n";
    for (size_t i = 0; i < n_; ++i) {
        os << "Set " << i << ": {";
        const char* sep = "";
        for (const String& str : sets_[i]) {
            os << sep;
            if (str.empty()) os << "ε"; else os << '"' << str << '"';
            sep = ", ";
        }
        os << "}" << (metadata_.is_degenerate[i] ? " [degenerate]" : "") << "\n";
    }
}

namespace {
// the EDS text of a set list: every symbol braced (FULL) or only the degenerate ones (COMPACT), eds.cpp:600-631
void write_sets(std::ostream& os, const std::vector<StringSet>& sets, const std::vector<bool>& degenerate, bool full) {
    for (size_t i = 0; i < sets.size(); ++i) {
        const bool braced = full || degenerate[i];
        if (braced) os << SET_OPEN;
        for (size_t j = 0; j < sets[i].size(); ++j) {
            if (j) os << SET_SEPARATOR;
            os << sets[i][j];
        }
        if (braced) os << SET_CLOSE;
    }
}
void write_sources(std::ostream& os, const std::vector<std::set<int>>& sources) {
    for (const std::set<int>& ids : sources) {
        os << SET_OPEN;
        const char* sep = "";
        for (int id : ids) {
            os << sep << id;
            sep = ",";
        }
        os << SET_CLOSE;
    }
}
}  // namespace

void EDS::save(std::ostream& os, OutputFormat format) const {
    if (mode_ == StoringMode::METADATA_ONLY)
        throw std::runtime_error("Cannot save EDS in METADATA_ONLY mode. Load with StoringMode::FULL to access string data for saving.");
    write_sets(os, sets_, metadata_.is_degenerate, format == OutputFormat::FULL);
    os << "\n";
}

void EDS::save(const std::filesystem::path& path, OutputFormat format) const {
    std::ofstream out(path);
    if (!out) throw std::runtime_error("Failed to open file for writing: " + path.string());
    save(out, format);
}

void EDS::save_sources(std::ostream& os) const {
    if (!has_sources_) throw std::runtime_error("Cannot save sources: no sources loaded");
    write_sources(os, sources_);
    os << "\n";
}

void EDS::save_sources(const std::filesystem::path& path) const {
    std::ofstream out(path);
    if (!out) throw std::runtime_error("Failed to open file for writing: " + path.string());
    save_sources(out);
}

std::string EDS::text() const {
    std::ostringstream os;
    write_sets(os, sets_, metadata_.is_degenerate, true);
    return os.str();
}

std::string EDS::sources_text() const {
    std::ostringstream os;
    write_sources(os, sources_);
    return os.str();
}

void EDS::load_sources(const std::string& seds_string) {
    const StoringMode mode = mode_;
    build(text(), &seds_string);  // the device checks the sources against this EDS (count, syntax) as it parses them
    mode_ = mode;
}
void EDS::load_sources(std::istream& is) { load_sources(slurp(is)); }
void EDS::load_sources(const std::filesystem::path& path) { load_sources(read_file(path, "sEDS")); }

const std::vector<StringSet>& EDS::get_sets() const {
    if (mode_ == StoringMode::METADATA_ONLY)
        throw std::runtime_error("Cannot access sets in METADATA_ONLY mode. Use read_symbol(pos) for on-demand access, or load with StoringMode::FULL");
    return sets_;
}

StringSet EDS::read_symbol(Position pos) const {
    if (pos >= n_) throw std::out_of_range("Position " + std::to_string(pos) + " out of range");
    return sets_[pos];
}

EDS EDS::merge_adjacent(size_t pos1, size_t pos2) const {
    if (pos2 != pos1 + 1)  // eds.cpp:1429-1434
        throw std::invalid_argument("Positions must be adjacent: pos2 (" + std::to_string(pos2) + ") must equal pos1 + 1 (" +
                                    std::to_string(pos1 + 1) + ")");
    if (pos1 >= n_ || pos2 >= n_)  // eds.cpp:1437-1442
        throw std::out_of_range("Position out of range: pos1=" + std::to_string(pos1) + ", pos2=" + std::to_string(pos2) +
                                ", n=" + std::to_string(n_));
    const std::string e = text(), s = has_sources_ ? sources_text() : std::string();
    eds_buffer oe{nullptr, 0}, os{nullptr, 0};
    const eds_status rc = eds_merge_adjacent_host(t_session.get(), reinterpret_cast<const uint8_t*>(e.data()), e.size(),
                                                  has_sources_ ? reinterpret_cast<const uint8_t*>(s.data()) : nullptr, s.size(), pos1, &oe, &os);
    if (rc != EDS_OK) rethrow(rc);
    const std::string me(reinterpret_cast<const char*>(oe.data), oe.bytes), ms(reinterpret_cast<const char*>(os.data), os.bytes);
    eds_buffer_free_host(&oe);
    eds_buffer_free_host(&os);
    return has_sources_ ? EDS(me, ms) : EDS(me);
}

bool is_leds(const EDS& eds, Length context_length) {
    if (context_length == 0) return true;  // eds_transforms.cpp:440-442
    int answer = 0;
    const std::string text = eds.text();
    const eds_status rc = eds_is_leds_host(t_session.get(), reinterpret_cast<const uint8_t*>(text.data()), text.size(), context_length, &answer);
    if (rc != EDS_OK) rethrow(rc);
    return answer != 0;
}

}  // namespace edsparser
