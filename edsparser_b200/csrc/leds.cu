// l-EDS merge on the GPU (placeholder until the kernels land).
#include "leds.h"

namespace edsb {

LedsPipeline::LedsPipeline(eds_ctx* ctx) : ctx_(ctx) {}
LedsPipeline::~LedsPipeline() {}

void LedsPipeline::merge_host(const uint8_t*, uint64_t, const uint8_t*, uint64_t, uint32_t, bool, uint64_t, eds_buffer*,
                              eds_buffer*, uint32_t*) {
    (void)ctx_;
    throw std::runtime_error("eds_leds_merge_host: not implemented yet");
}

}  // namespace edsb
