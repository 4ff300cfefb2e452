// Same signatures as the reference's transforms/eds_transforms.hpp:29-57; the work runs on the GPU through
// eds_leds_merge_host. num_threads is kept for source compatibility and ignored (the reference's output does
// not depend on it either).
#ifndef EDSPARSER_TRANSFORMS_EDS_TRANSFORMS_HPP
#define EDSPARSER_TRANSFORMS_EDS_TRANSFORMS_HPP

#include <iostream>
#include <vector>

#include "../common.hpp"
#include "../formats/eds.hpp"

namespace edsparser {

void eds_to_leds_linear(std::istream& input, std::ostream& output, Length context_length,
                        std::istream* phasing_input = nullptr, std::ostream* phasing_output = nullptr,
                        size_t num_threads = 1, bool compact = true);

void eds_to_leds_cartesian(std::istream& input, std::ostream& output, Length context_length, size_t num_threads = 1,
                           bool compact = true);

bool is_leds(const EDS& eds, Length context_length);

// B200 additions (not in the reference): device selection and an output budget for CARTESIAN merging,
// which has no bound in the reference. 0 = unlimited.
namespace b200 {
void set_device(int device);
// msa2eds column-sharded over these devices of one node (one entry: the same as set_device)
void set_devices(const std::vector<int>& devices);
void set_max_output_bytes(uint64_t bytes);
}  // namespace b200

}  // namespace edsparser
#endif
