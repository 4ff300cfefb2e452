// Same signatures as the reference's transforms/vcf_transforms.hpp:24-79; the work runs on the GPU through
// eds_vcf_transform_host (include/edsparser_b200.h). Output strings and VCFStats are identical to the reference's;
// its "Warning: Skipping variant at ..." lines for unsupported symbolic alleles are printed to std::cerr as well.
#ifndef EDSPARSER_TRANSFORMS_VCF_TRANSFORMS_HPP
#define EDSPARSER_TRANSFORMS_VCF_TRANSFORMS_HPP

#include <iostream>
#include <string>
#include <utility>

#include "../common.hpp"

namespace edsparser {

struct VCFStats {
    size_t total_variants = 0;          // record lines read (headers excluded)
    size_t processed_variants = 0;      // records that entered the EDS
    size_t skipped_malformed = 0;       // fewer than five fields or an unreadable POS
    size_t skipped_unsupported_sv = 0;  // symbolic ALT other than <DEL> / <INS>
    size_t variant_groups = 0;          // symbols created (overlapping records merged)
    size_t total_skipped() const { return skipped_malformed + skipped_unsupported_sv; }
};

// VCF + reference FASTA -> {EDS text, SEDS text}; one path id per sample column, 1-based. Both streams are read to
// their end. Only the first FASTA record is used; CHROM is ignored.
std::pair<std::string, std::string> parse_vcf_to_eds_streaming(std::istream& vcf_stream, std::istream& fasta_stream,
                                                               VCFStats* stats = nullptr);

// The same followed by the LINEAR source-aware merge to context_length (compact output, trailing newline).
std::pair<std::string, std::string> parse_vcf_to_leds_streaming(std::istream& vcf_stream, std::istream& fasta_stream,
                                                                size_t context_length, VCFStats* stats = nullptr);

}  // namespace edsparser
#endif
