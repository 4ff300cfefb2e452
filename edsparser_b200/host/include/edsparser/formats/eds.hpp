// class EDS with the public surface of the reference (src/cpp/lib/formats/eds.hpp:26-169) for the transform path:
// construction from streams / strings / files (± sources), the query members, get_sets / get_is_degenerate /
// get_sources / read_symbol, Metadata and Statistics, save / save_sources / load_sources, merge_adjacent.
// Parsing (EDS::parse + normaliser, parse_sources), the statistics and merge_adjacent run on the GPU through
// eds_parse_host / eds_merge_adjacent_host (include/edsparser_b200.h); this class holds what they return.
// Not part of this repo's scope (off the transform path): generate_patterns, extract,
// check_position, and METADATA_ONLY streaming from the file (the mode is accepted and recorded; the strings are
// held in memory all the same, get_sets() refuses as in the reference).
#ifndef EDSPARSER_FORMATS_EDS_HPP
#define EDSPARSER_FORMATS_EDS_HPP

#include <filesystem>
#include <fstream>
#include <iostream>
#include <set>
#include <string>
#include <vector>

#include "../common.hpp"

namespace edsparser {

class EDS {
   public:
    enum class StoringMode { FULL, METADATA_ONLY };
    enum class OutputFormat { FULL, COMPACT };

    EDS() = default;
    explicit EDS(std::istream& eds_stream);
    EDS(std::istream& eds_stream, std::istream& seds_stream);
    explicit EDS(const std::string& eds_string);
    EDS(const std::string& eds_string, const std::string& seds_string);

    static EDS load(const std::filesystem::path& path, StoringMode mode = StoringMode::FULL);
    static EDS load(const std::filesystem::path& eds_path, const std::filesystem::path& seds_path, StoringMode mode = StoringMode::FULL);
    static EDS from_string(const std::string& eds_string);
    static EDS from_string(const std::string& eds_string, const std::string& seds_string);

    ~EDS() = default;
    EDS(const EDS&) = delete;
    EDS& operator=(const EDS&) = delete;
    EDS(EDS&&) = default;
    EDS& operator=(EDS&&) = default;

    bool empty() const { return is_empty_; }
    size_t length() const { return n_; }       // number of sets
    size_t size() const { return N_; }         // total characters
    size_t cardinality() const { return m_; }  // total number of strings
    bool has_sources() const { return has_sources_; }
    StoringMode get_storing_mode() const { return mode_; }

    struct Metadata {
        std::vector<std::streampos> base_positions;  // not maintained here (file streaming is out of scope): empty
        std::vector<Length> symbol_sizes;
        std::vector<Length> string_lengths;
        std::vector<Length> cum_set_sizes;
        std::vector<bool> is_degenerate;
        Length min_context_length = 0;
        Length max_context_length = 0;
        double avg_context_length = 0.0;
        size_t num_degenerate_symbols = 0;
        size_t num_common_chars = 0;
        size_t total_change_size = 0;
        size_t num_empty_strings = 0;
        size_t num_paths = 0;
        size_t max_paths_per_string = 0;
        double avg_paths_per_string = 0.0;
        std::vector<Position> cum_common_positions;
        std::vector<int> cum_degenerate_counts;
    };

    struct Statistics {
        Length min_context_length;
        Length max_context_length;
        double avg_context_length;
        size_t num_degenerate_symbols;
        size_t num_common_chars;
        size_t total_change_size;
        size_t num_empty_strings;
        size_t num_paths;
        size_t max_paths_per_string;
        double avg_paths_per_string;
    };

    const Metadata& get_metadata() const { return metadata_; }
    Statistics get_statistics() const;
    void print_statistics(std::ostream& os = std::cout) const;

    void print(std::ostream& os = std::cout) const;
    void save(std::ostream& os, OutputFormat format = OutputFormat::FULL) const;
    void save(const std::filesystem::path& path, OutputFormat format = OutputFormat::FULL) const;
    void save_sources(std::ostream& os) const;
    void save_sources(const std::filesystem::path& path) const;

    void load_sources(std::istream& is);
    void load_sources(const std::filesystem::path& path);
    void load_sources(const std::string& seds_string);

    // Merge two adjacent symbols: all combinations without sources, the combinations whose source sets meet with them
    // (a set containing 0 is universal). Returns a new EDS; *this is unchanged.
    EDS merge_adjacent(size_t pos1, size_t pos2) const;

    const std::vector<StringSet>& get_sets() const;  // throws in METADATA_ONLY mode, as the reference does
    const std::vector<bool>& get_is_degenerate() const { return metadata_.is_degenerate; }
    const std::vector<std::set<int>>& get_sources() const { return sources_; }

    StringSet read_symbol(Position pos) const;
    Length get_symbol_size(Position pos) const { return metadata_.symbol_sizes[pos]; }
    Length get_string_length(size_t string_id) const { return metadata_.string_lengths[string_id]; }

    // the serialised forms this object was last built from / would be saved as (FULL dialect, no trailing newline)
    std::string text() const;
    std::string sources_text() const;

   private:
    void build(const std::string& eds_text, const std::string* seds_text);

    bool is_empty_ = true;
    size_t n_ = 0, N_ = 0, m_ = 0;
    StoringMode mode_ = StoringMode::FULL;
    Metadata metadata_;
    std::vector<StringSet> sets_;
    bool has_sources_ = false;
    std::vector<std::set<int>> sources_;
};

}  // namespace edsparser
#endif
