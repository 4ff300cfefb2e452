// Same signatures as the reference's transforms/msa_transforms.hpp:27,37-39; the work runs on the GPU
// through eds_msa_transform_host (include/edsparser_b200.h). Output strings are byte-identical.
#ifndef EDSPARSER_TRANSFORMS_MSA_TRANSFORMS_HPP
#define EDSPARSER_TRANSFORMS_MSA_TRANSFORMS_HPP

#include <iostream>
#include <string>
#include <utility>

#include "../common.hpp"

namespace edsparser {

// MSA (gapped FASTA) -> {EDS text, SEDS text}. The stream is read to its end.
std::pair<std::string, std::string> parse_msa_to_eds_streaming(std::istream& msa_stream);

// MSA -> {l-EDS text, SEDS text}: conserved runs shorter than context_length are absorbed into the
// neighbouring variable symbols while the boundaries are being built.
std::pair<std::string, std::string> parse_msa_to_leds_streaming(std::istream& msa_stream, size_t context_length);

}  // namespace edsparser
#endif
