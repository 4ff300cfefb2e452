// The slice of the reference's EDS class (src/cpp/lib/formats/eds.hpp:26-169) that the hot path needs on the
// host: an EDS (+ optional source sets) held as text, handed to the device for every operation. The
// statistics / pattern / position-query members of the reference class are out of this repo's scope
// (SURVEY.md §2 row 9).
#ifndef EDSPARSER_FORMATS_EDS_HPP
#define EDSPARSER_FORMATS_EDS_HPP

#include <istream>
#include <ostream>
#include <string>

#include "../common.hpp"

namespace edsparser {

class EDS {
   public:
    enum class OutputFormat { FULL, COMPACT };

    EDS() = default;
    explicit EDS(std::istream& eds_stream);
    EDS(std::istream& eds_stream, std::istream& sources_stream);
    explicit EDS(const std::string& eds_text);
    EDS(const std::string& eds_text, const std::string& sources_text);

    EDS(const EDS&) = delete;
    EDS& operator=(const EDS&) = delete;
    EDS(EDS&&) = default;
    EDS& operator=(EDS&&) = default;

    bool has_sources() const { return has_sources_; }
    bool empty() const { return text_.empty(); }
    const std::string& text() const { return text_; }
    const std::string& sources_text() const { return sources_; }

   private:
    std::string text_, sources_;
    bool has_sources_ = false;
};

}  // namespace edsparser
#endif
