// ============================================================================
// TEST INFRASTRUCTURE ONLY — CPU restatement ("oracle") of the EDSParser hot
// path. Nothing in the product (edsparser_b200/, include/) may link, import or
// execute this file; only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs use it, and only as the checker.
//
// Parity status: PINNED. This restatement is checked byte-for-byte against
//   * the reference's own golden strings (tests/cpp/test_msa.cpp:20-229,
//     tests/cpp/test_merge.cpp:81-259,359-371, data/eds/*_l<N>.eds, data/vcf/*.{eds,seds}), and
//   * outputs of the UNMODIFIED reference library built into oracle/_ref/
//     (tests/golden/*.json, produced by tests/golden/make_golden.py),
// see tests/test_oracle.py.
//
// Third-party arithmetic: the reference calls SDSL (simongog/sdsl-lite, version
// unpinned, not vendored) for bit_vector + select_0/select_1. select_b(k) is
// "position of the k-th bit equal to b, k>=1"; the run walk below restates that
// with plain loops.
//
// All file:line citations are relative to /root/reference/src/cpp/lib/.
// ============================================================================
#include <algorithm>
#include <cctype>
#include <climits>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <iterator>
#include <map>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace oracle {

// ---------------------------------------------------------------------------
// MSA -> EDS / l-EDS          (transforms/msa_transforms.cpp)
// ---------------------------------------------------------------------------

struct MsaScan {
    std::string first_row;             // row 0 residues incl. '-'  (:64)
    std::vector<long long> row_begin;  // byte offset after each header line (:60)
    size_t rows = 0;
    long long wrap = -1;               // length of the first data line (:65-67)
    std::vector<uint8_t> conserved;    // B, size C+1 (:56,:75-77,:84)
};

// Pass 1 (:36-90). Emulates std::getline over an in-memory copy of the file.
static MsaScan msa_scan(const std::string& file) {
    MsaScan s;
    size_t at = 0;
    size_t col = 0;
    bool have_bits = false;
    while (at < file.size()) {
        size_t nl = file.find('\n', at);
        bool hit_eof = (nl == std::string::npos);
        size_t line_end = hit_eof ? file.size() : nl;
        const char* line = file.data() + at;
        size_t len = line_end - at;
        at = hit_eof ? file.size() : nl + 1;
        if (len == 0) continue;  // blank lines are skipped (:47-49)
        if (line[0] == '>') {
            if (s.rows == 1) {
                s.conserved.assign(s.first_row.size() + 1, 1);
                have_bits = true;
            }
            col = 0;
            s.rows++;
            // tellg() after a getline that ran into EOF is -1 (failbit/eofbit)
            s.row_begin.push_back(hit_eof ? -1 : static_cast<long long>(at));
        } else if (s.rows == 1) {
            s.first_row.append(line, len);
            if (s.wrap == -1) s.wrap = static_cast<long long>(len);
        } else {
            if (!have_bits) throw std::runtime_error("oracle: residue line before any header (undefined in the reference)");
            for (size_t j = 0; j < len; ++j) {
                if (col >= s.first_row.size())
                    throw std::runtime_error("oracle: row longer than the first row (undefined in the reference)");
                if (line[j] != s.first_row[col] || line[j] == '-') s.conserved[col] = 0;
                ++col;
            }
        }
    }
    if (!have_bits || s.first_row.empty())
        throw std::runtime_error("oracle: fewer than 2 rows or empty first row (undefined in the reference)");
    size_t C = s.first_row.size();
    s.conserved[C] = s.conserved[C - 1] ^ 1;  // sentinel (:84)
    return s;
}

// Pass 2a (:101-115): a symbol starts at 0 and wherever B changes.
static std::vector<uint8_t> starts_plain(const std::vector<uint8_t>& B) {
    std::vector<uint8_t> H(B.size(), 0);
    H[0] = 1;
    for (size_t i = 1; i < B.size(); ++i) H[i] = (B[i] != B[i - 1]);
    return H;
}

// Pass 2b (:133-190): walk maximal runs of B. A conserved run is "standalone"
// when it is at least l long or touches either end; it opens a symbol, and so
// does whatever run follows it. Everything else is glued to its left.
static std::vector<uint8_t> starts_merged(const std::vector<uint8_t>& B, size_t l, size_t C) {
    std::vector<uint8_t> H(B.size(), 0);
    bool after_standalone = false;
    size_t i = 0;
    while (i < C) {
        size_t j = i;
        while (B[j] == B[i]) ++j;  // sentinel B[C] != B[C-1] stops the walk (select_0/select_1)
        if (B[i]) {
            bool standalone = (j - i >= l) || i == 0 || j == C;
            if (standalone) {
                H[i] = 1;
                after_standalone = true;
            } else {
                if (after_standalone) H[i] = 1;
                after_standalone = false;
            }
        } else {
            if (after_standalone) {
                H[i] = 1;
                after_standalone = false;
            }
        }
        i = j;
    }
    H[0] = 1;
    return H;
}

// Pass 3 (:200-324). The seek+read per (symbol,row) is emulated on the memory
// copy, including the reused read buffer (:219,:276-286).
static void msa_emit(const std::string& file, const MsaScan& s, const std::vector<uint8_t>& H,
                     std::string& eds, std::string& seds) {
    const size_t C = s.first_row.size();
    const size_t lw = static_cast<size_t>(s.wrap);
    std::vector<size_t> opens;
    for (size_t i = 0; i < C; ++i)
        if (H[i]) opens.push_back(i);
    std::vector<char> window(C + C / lw + 10, '\0');
    for (size_t k = 0; k < opens.size(); ++k) {
        size_t lo = opens[k];
        size_t hi = (k + 1 < opens.size()) ? opens[k + 1] : C;
        bool all_conserved = true;
        for (size_t i = lo; i < hi && all_conserved; ++i) all_conserved = s.conserved[i] != 0;
        eds.push_back('{');
        if (all_conserved) {
            for (size_t i = lo; i < hi; ++i)
                if (s.first_row[i] != '-') eds.push_back(s.first_row[i]);
            seds += "{0}";
        } else {
            std::map<std::string, std::set<int>> carriers;
            std::vector<std::string> order;
            size_t span = hi - lo;
            size_t want = span + ((lo % lw) + span) / lw;  // residues + newlines crossed (:272-273)
            for (size_t r = 0; r < s.rows; ++r) {
                long long from = s.row_begin[r] + static_cast<long long>(lo + lo / lw);  // (:268-269)
                if (from >= 0 && static_cast<size_t>(from) < file.size()) {
                    size_t got = std::min(want, file.size() - static_cast<size_t>(from));
                    std::memcpy(window.data(), file.data() + from, got);
                }
                std::string hap;
                for (size_t q = 0; q < want && window[q] != '\0'; ++q)
                    if (window[q] != '\n' && window[q] != '-') hap.push_back(window[q]);
                if (!carriers.count(hap)) order.push_back(hap);
                carriers[hap].insert(static_cast<int>(r) + 1);
            }
            for (size_t v = 0; v < order.size(); ++v) {
                if (v) eds.push_back(',');
                eds += order[v];
                seds.push_back('{');
                bool first = true;
                for (int id : carriers[order[v]]) {
                    if (!first) seds.push_back(',');
                    seds += std::to_string(id);
                    first = false;
                }
                seds.push_back('}');
            }
        }
        eds.push_back('}');
    }
}

// parse_msa_to_eds_streaming (:334-345) when leds == false,
// parse_msa_to_leds_streaming (:351-365) otherwise.
static void msa_to_eds(const std::string& file, bool leds, size_t l, std::string& eds, std::string& seds) {
    MsaScan s = msa_scan(file);
    std::vector<uint8_t> H = leds ? starts_merged(s.conserved, l, s.first_row.size()) : starts_plain(s.conserved);
    msa_emit(file, s, H, eds, seds);
}

// ---------------------------------------------------------------------------
// EDS text model               (formats/eds.cpp)
// ---------------------------------------------------------------------------

struct Eds {
    std::vector<std::vector<std::string>> sym;  // sets_
    std::vector<std::vector<int>> src;          // sources_, one sorted unique id list per string
    bool with_src = false;
    size_t strings() const {
        size_t m = 0;
        for (auto& s : sym) m += s.size();
        return m;
    }
};

static std::string drop_space(const std::string& in) {
    std::string out;
    out.reserve(in.size());
    for (unsigned char c : in)
        if (!std::isspace(c)) out.push_back(static_cast<char>(c));
    return out;
}

// normalize_eds_format (formats/eds.cpp:831-881): bare text at depth 0 becomes its own set.
static std::string brace_bare_text(const std::string& in) {
    std::string out, bare;
    int depth = 0;
    for (char c : in) {
        if (c == '{') {
            if (!bare.empty() && depth == 0) {
                out += '{' + bare + '}';
                bare.clear();
            }
            out.push_back(c);
            ++depth;
        } else if (c == '}') {
            out.push_back(c);
            --depth;
        } else if (depth > 0) {
            out.push_back(c);
        } else {
            bare.push_back(c);
        }
    }
    if (!bare.empty() && depth == 0) out += '{' + bare + '}';
    return out;
}

// EDS::parse (formats/eds.cpp:39-155)
static void parse_eds(const std::string& raw, Eds& e) {
    std::string t = drop_space(raw);
    e.sym.clear();
    if (t.empty()) return;
    t = brace_bare_text(t);
    size_t p = 0;
    while (p < t.size()) {
        if (t[p] != '{') throw std::runtime_error("Expected '{' at position " + std::to_string(p));
        ++p;
        std::vector<std::string> alts;
        std::string cur;
        while (p < t.size() && t[p] != '}') {
            if (t[p] == ',') {
                alts.push_back(cur);
                cur.clear();
            } else {
                cur.push_back(t[p]);
            }
            ++p;
        }
        alts.push_back(cur);
        if (p >= t.size()) throw std::runtime_error("Expected '}' at position " + std::to_string(p));
        ++p;
        e.sym.push_back(std::move(alts));
    }
}

// EDS::parse_sources (formats/eds.cpp:268-355)
static void parse_seds(const std::string& raw, Eds& e) {
    std::string t = drop_space(raw);
    if (t.empty()) throw std::runtime_error("sEDS input is empty");
    e.src.clear();
    size_t p = 0;
    auto to_id = [](const std::string& digits) {
        // std::stoi semantics: overflow -> std::out_of_range("stoi")
        long long v = 0;
        for (char c : digits) {
            v = v * 10 + (c - '0');
            if (v > INT_MAX) throw std::out_of_range("stoi");
        }
        return static_cast<int>(v);
    };
    while (p < t.size()) {
        if (t[p] != '{') throw std::runtime_error("sEDS: Expected '{' at position " + std::to_string(p));
        ++p;
        std::set<int> ids;
        std::string num;
        while (p < t.size() && t[p] != '}') {
            if (t[p] == ',') {
                if (!num.empty()) {
                    ids.insert(to_id(num));
                    num.clear();
                }
            } else if (t[p] >= '0' && t[p] <= '9') {
                num.push_back(t[p]);
            } else {
                throw std::runtime_error("sEDS: Invalid character '" + std::string(1, t[p]) + "' at position " +
                                         std::to_string(p));
            }
            ++p;
        }
        if (!num.empty()) ids.insert(to_id(num));
        if (p >= t.size()) throw std::runtime_error("sEDS: Expected '}' at position " + std::to_string(p));
        ++p;
        if (ids.empty()) throw std::runtime_error("sEDS: Empty path set at string " + std::to_string(e.src.size()));
        e.src.emplace_back(ids.begin(), ids.end());
    }
    if (e.src.size() != e.strings())
        throw std::runtime_error("sEDS: Source count (" + std::to_string(e.src.size()) +
                                 ") does not match EDS cardinality (" + std::to_string(e.strings()) + ")");
    e.with_src = true;
}

// EDS::save (formats/eds.cpp:600-631) and save_sources (:641-659)
static std::string eds_text(const Eds& e, bool compact) {
    std::string out;
    for (auto& alts : e.sym) {
        bool braces = !compact || alts.size() > 1;
        if (braces) out.push_back('{');
        for (size_t i = 0; i < alts.size(); ++i) {
            if (i) out.push_back(',');
            out += alts[i];
        }
        if (braces) out.push_back('}');
    }
    out.push_back('\n');
    return out;
}

static std::string seds_text(const Eds& e) {
    std::string out;
    for (auto& ids : e.src) {
        out.push_back('{');
        for (size_t i = 0; i < ids.size(); ++i) {
            if (i) out.push_back(',');
            out += std::to_string(ids[i]);
        }
        out.push_back('}');
    }
    out.push_back('\n');
    return out;
}

// ---------------------------------------------------------------------------
// l-EDS merge rounds           (transforms/eds_transforms.cpp, formats/eds.cpp:1425-1695)
// ---------------------------------------------------------------------------

static bool short_inner_solid(const Eds& e, size_t i, uint32_t l) {
    // non-degenerate, not first, not last, single string shorter than l
    return e.sym[i].size() == 1 && i > 0 && i + 1 < e.sym.size() && e.sym[i][0].size() < l;
}

// is_leds (transforms/eds_transforms.cpp:439-468)
static bool leds_holds(const Eds& e, uint32_t l) {
    if (l == 0) return true;
    for (size_t i = 0; i < e.sym.size(); ++i) {
        if (short_inner_solid(e, i, l)) return false;
        if (i + 1 < e.sym.size() && e.sym[i].size() > 1 && e.sym[i + 1].size() > 1) return false;
    }
    return true;
}

// select_independent_merge_pairs (:46-107): greedy, left to right, disjoint.
static std::vector<size_t> pick_pairs(const Eds& e, uint32_t l) {
    std::vector<size_t> left;
    size_t n = e.sym.size();
    if (n < 2) return left;
    std::vector<char> taken(n, 0);
    for (size_t i = 0; i + 1 < n; ++i) {
        if (taken[i] || taken[i + 1]) continue;
        bool go = short_inner_solid(e, i, l) || short_inner_solid(e, i + 1, l) ||
                  (e.sym[i].size() > 1 && e.sym[i + 1].size() > 1);
        if (go) {
            left.push_back(i);
            taken[i] = taken[i + 1] = 1;
        }
    }
    return left;
}

static bool universal(const std::vector<int>& ids) { return !ids.empty() && ids.front() == 0; }  // sorted: 0 is first

// Source combination of EDS::merge_adjacent (formats/eds.cpp:1479-1500).
static std::vector<int> meet(const std::vector<int>& a, const std::vector<int>& b) {
    bool ua = universal(a), ub = universal(b);
    if (ua && ub) return {0};
    if (ua) return b;
    if (ub) return a;
    std::vector<int> out;
    std::set_intersection(a.begin(), a.end(), b.begin(), b.end(), std::back_inserter(out));
    return out;
}

// One round = pick_pairs + merge_multiple_pairs (:120-196) + reconstruct_eds (:207-296).
// The reference re-serialises and re-parses after every round; on this model that is the
// identity, so the round is applied in place. `first` = string index of each symbol's
// first alternative (cum_set_sizes).
static void apply_round(Eds& e, const std::vector<size_t>& pairs, size_t max_out_bytes = 0) {
    std::vector<size_t> first(e.sym.size() + 1, 0);
    for (size_t i = 0; i < e.sym.size(); ++i) first[i + 1] = first[i] + e.sym[i].size();
    std::vector<std::vector<std::string>> nsym;
    std::vector<std::vector<int>> nsrc;
    size_t next_pair = 0;
    for (size_t i = 0; i < e.sym.size(); ++i) {
        if (next_pair < pairs.size() && pairs[next_pair] == i) {
            const auto& L = e.sym[i];
            const auto& Rr = e.sym[i + 1];
            if (max_out_bytes && L.size() * Rr.size() > max_out_bytes) throw std::length_error("oracle: output budget exceeded");
            std::vector<std::string> merged;
            size_t kept_before = nsrc.size();
            for (size_t a = 0; a < L.size(); ++a)
                for (size_t b = 0; b < Rr.size(); ++b) {
                    if (e.with_src) {
                        std::vector<int> both = meet(e.src[first[i] + a], e.src[first[i + 1] + b]);
                        if (both.empty()) continue;
                        nsrc.push_back(std::move(both));
                    }
                    merged.push_back(L[a] + Rr[b]);
                }
            if (e.with_src && nsrc.size() == kept_before)
                throw std::runtime_error("Merging positions " + std::to_string(i) + " and " + std::to_string(i + 1) +
                                         " results in empty set (no valid source intersections)");
            nsym.push_back(std::move(merged));
            ++next_pair;
            ++i;  // the right symbol is consumed
        } else {
            if (e.with_src)
                for (size_t a = 0; a < e.sym[i].size(); ++a) nsrc.push_back(std::move(e.src[first[i] + a]));
            nsym.push_back(std::move(e.sym[i]));
        }
    }
    e.sym.swap(nsym);
    if (e.with_src) e.src.swap(nsrc);
}

// eds_to_leds_linear (:313-373) when seds != nullptr, eds_to_leds_cartesian (:381-426) otherwise.
// max_out_bytes (0 = unlimited) is an oracle-side guard against cartesian blow-up.
static void eds_to_leds(const std::string& eds_in, const std::string* seds_in, uint32_t l, bool compact,
                        std::string& eds_out, std::string& seds_out, size_t max_out_bytes) {
    if (l == 0) throw std::invalid_argument("context_length must be > 0 for l-EDS transformation");
    Eds e;
    parse_eds(eds_in, e);
    if (seds_in) parse_seds(*seds_in, e);
    size_t round = 0;
    const size_t kMaxRounds = 10000;
    while (round < kMaxRounds) {
        if (leds_holds(e, l)) break;
        std::vector<size_t> pairs = pick_pairs(e, l);
        if (pairs.empty()) break;
        apply_round(e, pairs, max_out_bytes);
        if (max_out_bytes) {
            size_t bytes = 0;
            for (auto& s : e.sym)
                for (auto& a : s) bytes += a.size() + 1;
            if (bytes > max_out_bytes) throw std::length_error("oracle: output budget exceeded");
        }
        ++round;
    }
    if (round >= kMaxRounds) throw std::runtime_error("Maximum iterations reached without convergence");
    eds_out = eds_text(e, compact);
    if (e.with_src) seds_out = seds_text(e);
}


// ---------------------------------------------------------------------------
// VCF + FASTA -> EDS / l-EDS      (transforms/vcf_transforms.cpp)
// ---------------------------------------------------------------------------

struct VcfCounters {  // VCFStats, vcf_transforms.hpp:24-36
    size_t total = 0, processed = 0, malformed = 0, unsupported_sv = 0, groups = 0;
};

// std::getline over an in-memory file: yields the bytes up to the next '\n'; fails only when nothing is left.
struct LineReader {
    const std::string& buf;
    size_t at = 0;
    bool eof_hit = false;  // eofbit: the last getline ran into the end without seeing '\n'
    explicit LineReader(const std::string& b) : buf(b) {}
    bool next(std::string& line) {
        if (eof_hit || at >= buf.size()) { eof_hit = true; return false; }
        size_t nl = buf.find('\n', at);
        if (nl == std::string::npos) { line = buf.substr(at); at = buf.size(); eof_hit = true; }
        else { line = buf.substr(at, nl - at); at = nl + 1; }
        return true;
    }
};

struct FastaIndex {  // FASTAMetadata :20-25
    size_t n_bases = 0, wrap = 0, first_base_at = 0;
};

// parse_fasta_metadata :51-86
static FastaIndex fasta_index(const std::string& fa) {
    LineReader rd(fa);
    std::string line;
    if (!rd.next(line) || line.empty() || line[0] != '>')
        throw std::runtime_error("Invalid FASTA format: expected header line starting with '>'");
    FastaIndex ix;
    // tellg() after a getline that hit EOF fails (sentry sets failbit) and so does the next getline (:72)
    ix.first_base_at = rd.at;
    if (!rd.next(line)) throw std::runtime_error("FASTA file is empty");
    ix.wrap = line.size();
    ix.n_bases = line.size();
    while (rd.next(line)) {
        if (line.empty()) continue;
        if (line[0] == '>') break;
        ix.n_bases += line.size();
    }
    return ix;
}

// read_fasta_region :98-129
static std::string fasta_slice(const std::string& fa, const FastaIndex& ix, size_t from, size_t n) {
    if (from >= ix.n_bases) return "";
    if (from + n > ix.n_bases) n = ix.n_bases - from;
    if (ix.wrap == 0) throw std::runtime_error("oracle: FASTA line width 0 (the reference divides by zero here)");
    std::string out;
    size_t p = ix.first_base_at + from + from / ix.wrap;
    while (out.size() < n && p < fa.size()) {
        char c = fa[p++];
        if (c != '\n' && c != '\r') out.push_back(c);
    }
    return out;
}

struct VcfRecord {  // VCFVariant :27-33
    std::string chrom, ref;
    size_t pos = 0;
    std::vector<std::string> alts;
    std::vector<std::vector<int>> gts;
};

// std::getline(ss, tok, delim) tokenisation: a trailing empty piece is not produced
static std::vector<std::string> split_on(const std::string& s, char delim) {
    std::vector<std::string> out;
    size_t at = 0;
    while (at < s.size()) {
        size_t d = s.find(delim, at);
        if (d == std::string::npos) { out.push_back(s.substr(at)); break; }
        out.push_back(s.substr(at, d - at));
        at = d + 1;
    }
    return out;
}
static std::vector<std::string> split_space(const std::string& s) {  // operator>> tokens
    std::vector<std::string> out;
    size_t i = 0;
    while (i < s.size()) {
        while (i < s.size() && std::isspace(static_cast<unsigned char>(s[i]))) ++i;
        size_t b = i;
        while (i < s.size() && !std::isspace(static_cast<unsigned char>(s[i]))) ++i;
        if (i > b) out.push_back(s.substr(b, i - b));
    }
    return out;
}

// parse_genotype :190-216
static std::vector<int> gt_alleles(const std::string& gt) {
    char d = gt.find('/') != std::string::npos ? '/' : '|';
    std::vector<int> out;
    for (const std::string& piece : split_on(gt, d)) {
        if (piece == ".") continue;
        try { out.push_back(std::stoi(piece)); } catch (...) {}
    }
    return out;
}

enum class LineKind { RECORD, HEADER, MALFORMED, SV };

// parse_vcf_line :232-326 (+ parse_alt_field :142-176)
static LineKind vcf_line(const std::string& line, VcfRecord& r, std::vector<std::string>* warnings) {
    if (line.empty() || line[0] == '#') return LineKind::HEADER;  // n_samples from #CHROM is never used (:558)
    std::vector<std::string> f;
    for (std::string& t : split_on(line, '\t'))
        if (!t.empty()) f.push_back(std::move(t));
    if (f.size() < 5) f = split_space(line);
    if (f.size() < 5) return LineKind::MALFORMED;
    r.chrom = f[0];
    try { r.pos = std::stoull(f[1]); } catch (...) { return LineKind::MALFORMED; }
    r.ref = f[3];
    for (const std::string& a : split_on(f[4], ',')) {
        if (!a.empty() && a.front() == '<' && a.back() == '>') {
            std::string kind = a.substr(1, a.size() - 2);
            if (kind == "DEL") r.alts.push_back("");
            else if (kind == "INS") r.alts.push_back(r.ref);
            else {
                if (warnings)
                    warnings->push_back("Warning: Skipping variant at " + r.chrom + ":" + std::to_string(r.pos) +
                                        " - Unsupported structural variant type: " + kind);
                return LineKind::SV;
            }
        } else {
            r.alts.push_back(a);  // an empty piece ("A,,C") reads as a plain empty allele
        }
    }
    for (size_t i = 9; i < f.size(); ++i) r.gts.push_back(gt_alleles(f[i].substr(0, f[i].find(':'))));
    return LineKind::RECORD;
}

// apply_variant_to_span :356-390
static std::string with_allele(const std::string& span, size_t span_from, const VcfRecord& r, int allele) {
    if (allele < 1 || allele > static_cast<int>(r.alts.size())) return span;
    size_t off = (r.pos - 1) - span_from;
    std::string out = span.substr(0, off);  // throws std::out_of_range when the span was cut short by the FASTA end
    out += r.alts[allele - 1];
    if (off + r.ref.size() < span.size()) out += span.substr(off + r.ref.size());
    return out;
}

// parse_vcf_to_eds_streaming :677-729 = group_overlapping_variants :482-534 + merge_variant_group :396-476 +
// generate_eds_from_variants :554-668
static void vcf_to_eds(const std::string& vcf, const std::string& fa, std::string& eds, std::string& seds,
                       VcfCounters& st, std::vector<std::string>* warnings) {
    FastaIndex ix = fasta_index(fa);
    std::vector<VcfRecord> recs;
    {
        LineReader rd(vcf);
        std::string line;
        while (rd.next(line)) {
            VcfRecord r;
            switch (vcf_line(line, r, warnings)) {
                case LineKind::HEADER: break;
                case LineKind::MALFORMED: ++st.total; ++st.malformed; break;
                case LineKind::SV: ++st.total; ++st.unsupported_sv; break;
                case LineKind::RECORD: ++st.total; ++st.processed; recs.push_back(std::move(r)); break;
            }
        }
    }
    // :715-718 — std::sort, unstable: the same call on the same sequence reproduces the reference's order of ties
    std::sort(recs.begin(), recs.end(), [](const VcfRecord& a, const VcfRecord& b) { return a.pos < b.pos; });

    size_t cursor = 0;
    auto common = [&](size_t upto) {
        std::string text = fasta_slice(fa, ix, cursor, upto - cursor);
        if (!text.empty()) { eds += '{' + text + '}'; seds += "{0}"; }
    };
    size_t i = 0;
    while (i < recs.size()) {
        size_t from = recs[i].pos - 1, to = from + recs[i].ref.size(), j = i + 1;
        while (j < recs.size() && recs[j].pos - 1 < to) { to = std::max(to, recs[j].pos - 1 + recs[j].ref.size()); ++j; }
        ++st.groups;
        std::string span = fasta_slice(fa, ix, from, to - from);
        // haplotype list: the reference span, then every ALT of every record applied alone; equal strings once
        std::vector<std::string> haps{span};
        auto hap_id = [&](const std::string& h) -> int {
            auto it = std::find(haps.begin(), haps.end(), h);
            return it == haps.end() ? -1 : static_cast<int>(it - haps.begin());
        };
        for (size_t v = i; v < j; ++v)
            for (size_t a = 0; a < recs[v].alts.size(); ++a) {
                std::string h = with_allele(span, from, recs[v], static_cast<int>(a) + 1);
                if (hap_id(h) < 0) haps.push_back(h);
            }
        size_t n_samples = recs[i].gts.size();
        std::vector<std::set<int>> carriers(haps.size());
        for (size_t s = 0; s < n_samples; ++s) {
            std::set<int> mine;
            for (size_t v = i; v < j; ++v) {
                if (s >= recs[v].gts.size()) continue;
                for (int a : recs[v].gts[s]) {
                    int h = hap_id(with_allele(span, from, recs[v], a));
                    if (h >= 0) mine.insert(h);
                }
            }
            if (mine.empty()) mine.insert(0);
            for (int h : mine) carriers[h].insert(static_cast<int>(s) + 1);
        }
        if (from > cursor) { common(from); cursor = from; }
        eds += '{';
        if (n_samples == 0) {  // :603-615
            for (size_t h = 0; h < haps.size(); ++h) { if (h) eds += ','; eds += haps[h]; }
            eds += '}';
            seds += "{0}";
        } else {
            bool first = true;
            for (size_t h = 0; h < haps.size(); ++h) {
                if (carriers[h].empty()) continue;  // :619-624
                if (!first) eds += ',';
                first = false;
                eds += haps[h];
                seds += '{';
                bool f1 = true;
                for (int id : carriers[h]) { if (!f1) seds += ','; f1 = false; seds += std::to_string(id); }
                seds += '}';
            }
            eds += '}';
        }
        cursor = from + span.size();  // group.end_pos :404
        i = j;
    }
    if (cursor < ix.n_bases) common(ix.n_bases);
}

// parse_vcf_to_leds_streaming :735-755 — LINEAR merge, one thread, compact
static void vcf_to_leds(const std::string& vcf, const std::string& fa, size_t l, std::string& eds, std::string& seds,
                        VcfCounters& st, std::vector<std::string>* warnings) {
    std::string e0, s0;
    vcf_to_eds(vcf, fa, e0, s0, st, warnings);
    eds_to_leds(e0, &s0, static_cast<uint32_t>(l), true, eds, seds, 0);
}

}  // namespace oracle

// ---------------------------------------------------------------------------
// C ABI for ctypes (tests/, bench.py cpu_baseline). status: 0 ok, 1 runtime_error,
// 2 invalid_argument, 3 out_of_range, 4 other.
// ---------------------------------------------------------------------------
namespace {
char* dup_bytes(const std::string& s) {
    char* p = static_cast<char*>(std::malloc(s.size() + 1));
    std::memcpy(p, s.data(), s.size());
    p[s.size()] = 0;
    return p;
}
int fail(const std::exception& ex, char* err, size_t cap) {
    if (err && cap) std::snprintf(err, cap, "%s", ex.what());
    if (dynamic_cast<const std::invalid_argument*>(&ex)) return 2;
    if (dynamic_cast<const std::out_of_range*>(&ex)) return 3;
    if (dynamic_cast<const std::runtime_error*>(&ex)) return 1;
    return 4;
}
}  // namespace

extern "C" {

int oracle_msa2eds(const char* text, size_t n, int leds, size_t l, char** eds, size_t* eds_n, char** seds,
                   size_t* seds_n, char* err, size_t errcap) {
    try {
        std::string e, s;
        oracle::msa_to_eds(std::string(text, n), leds != 0, l, e, s);
        *eds = dup_bytes(e);
        *eds_n = e.size();
        *seds = dup_bytes(s);
        *seds_n = s.size();
        return 0;
    } catch (const std::exception& ex) {
        return fail(ex, err, errcap);
    }
}

// Conserved-column bit vector B incl. the sentinel; out must hold C+1 bytes. Returns C or -1.
long long oracle_msa_conserved(const char* text, size_t n, unsigned char* out, size_t cap, char* err, size_t errcap) {
    try {
        oracle::MsaScan s = oracle::msa_scan(std::string(text, n));
        if (s.conserved.size() > cap) throw std::runtime_error("oracle: output too small");
        std::memcpy(out, s.conserved.data(), s.conserved.size());
        return static_cast<long long>(s.first_row.size());
    } catch (const std::exception& ex) {
        fail(ex, err, errcap);
        return -1;
    }
}

int oracle_eds2leds(const char* eds, size_t n, const char* seds, size_t sn, unsigned l, int compact,
                    size_t max_out_bytes, char** out, size_t* out_n, char** sout, size_t* sout_n, char* err,
                    size_t errcap) {
    try {
        std::string e, s;
        std::string seds_s = seds ? std::string(seds, sn) : std::string();
        oracle::eds_to_leds(std::string(eds, n), seds ? &seds_s : nullptr, l, compact != 0, e, s, max_out_bytes);
        *out = dup_bytes(e);
        *out_n = e.size();
        *sout = dup_bytes(s);
        *sout_n = s.size();
        return 0;
    } catch (const std::exception& ex) {
        return fail(ex, err, errcap);
    }
}

// stats: total, processed, malformed, unsupported_sv, groups. warnings (optional): '\n'-joined stderr lines.
int oracle_vcf2eds(const char* vcf, size_t vn, const char* fa, size_t fn, size_t l, char** eds, size_t* eds_n,
                   char** seds, size_t* seds_n, unsigned long long* stats, char** warnings, char* err, size_t errcap) {
    try {
        std::string e, s;
        oracle::VcfCounters st;
        std::vector<std::string> warn;
        if (l) oracle::vcf_to_leds(std::string(vcf, vn), std::string(fa, fn), l, e, s, st, &warn);
        else oracle::vcf_to_eds(std::string(vcf, vn), std::string(fa, fn), e, s, st, &warn);
        *eds = dup_bytes(e);
        *eds_n = e.size();
        *seds = dup_bytes(s);
        *seds_n = s.size();
        if (stats) {
            stats[0] = st.total; stats[1] = st.processed; stats[2] = st.malformed; stats[3] = st.unsupported_sv;
            stats[4] = st.groups;
        }
        if (warnings) {
            std::string w;
            for (auto& x : warn) w += x + "\n";
            *warnings = dup_bytes(w);
        }
        return 0;
    } catch (const std::exception& ex) {
        return fail(ex, err, errcap);
    }
}

void oracle_free(char* p) { std::free(p); }

}  // extern "C"

#ifdef EDS_ORACLE_MAIN
// eds_oracle msa2eds <in.msa> <l> <out.eds> <out.seds>
// eds_oracle eds2leds <in.eds> <in.seds|-> <l> <out.leds> <out.seds|-> <compact>
// eds_oracle vcf2eds <in.vcf> <ref.fa> <l> <out.eds> <out.seds>
static std::string slurp(const char* path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error(std::string("cannot open ") + path);
    std::ostringstream ss;
    ss << f.rdbuf();
    return ss.str();
}
static void spill(const char* path, const std::string& s) {
    std::ofstream f(path, std::ios::binary);
    f << s;
}
int main(int argc, char** argv) {
    try {
        std::string cmd = argc > 1 ? argv[1] : "";
        if (cmd == "msa2eds" && argc == 6) {
            size_t l = std::strtoull(argv[3], nullptr, 10);
            std::string e, s;
            oracle::msa_to_eds(slurp(argv[2]), l > 0, l, e, s);
            spill(argv[4], e);
            spill(argv[5], s);
            return 0;
        }
        if (cmd == "eds2leds" && argc == 8) {
            bool linear = std::strcmp(argv[3], "-") != 0;
            std::string seds = linear ? slurp(argv[3]) : std::string();
            std::string e, s;
            oracle::eds_to_leds(slurp(argv[2]), linear ? &seds : nullptr,
                                static_cast<uint32_t>(std::strtoul(argv[4], nullptr, 10)), std::atoi(argv[7]) != 0, e, s, 0);
            spill(argv[5], e);
            if (linear && std::strcmp(argv[6], "-") != 0) spill(argv[6], s);
            return 0;
        }
        if (cmd == "vcf2eds" && argc == 7) {
            size_t l = std::strtoull(argv[4], nullptr, 10);
            std::string e, s;
            oracle::VcfCounters st;
            std::vector<std::string> warn;
            if (l) oracle::vcf_to_leds(slurp(argv[2]), slurp(argv[3]), l, e, s, st, &warn);
            else oracle::vcf_to_eds(slurp(argv[2]), slurp(argv[3]), e, s, st, &warn);
            for (auto& w : warn) std::cerr << w << "\n";
            spill(argv[5], e);
            spill(argv[6], s);
            std::cout << "stats total=" << st.total << " processed=" << st.processed << " malformed=" << st.malformed
                      << " sv=" << st.unsupported_sv << " groups=" << st.groups << "\n";
            return 0;
        }
        std::cerr << "usage: eds_oracle msa2eds|eds2leds|vcf2eds ...\n";
        return 2;
    } catch (const std::exception& ex) {
        std::cerr << "Error: " << ex.what() << "\n";
        return 1;
    }
}
#endif
