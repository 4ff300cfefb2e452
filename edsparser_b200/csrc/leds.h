// Host-side launcher of the l-EDS merge kernels (leds.cu): eds_to_leds_linear / eds_to_leds_cartesian.
#pragma once
#include <functional>
#include <stdexcept>

#include "ctx.h"

namespace edsb {

struct BudgetError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

class LedsPipeline {
   public:
    explicit LedsPipeline(eds_ctx* ctx);
    ~LedsPipeline();
    LedsPipeline(const LedsPipeline&) = delete;
    LedsPipeline& operator=(const LedsPipeline&) = delete;

    // Host text in, malloc'd host text out.
    void merge_host(const uint8_t* eds_in, uint64_t eds_bytes, const uint8_t* seds_in, uint64_t seds_bytes, uint32_t l,
                    bool compact, uint64_t max_output_bytes, eds_buffer* leds_out, eds_buffer* seds_out,
                    uint32_t* rounds_out, int* check_only = nullptr, bool input_on_device = false,
                    const std::function<uint8_t*(int, uint64_t)>& sink = {}, eds_parsed* parse_only = nullptr,
                    const uint64_t* single_pair = nullptr, uint32_t* edge_unmerged = nullptr);
    // edge_unmerged != nullptr: bit 0 / bit 1 = the first / last symbol of the result is an original, unmerged string
    // parse_only != nullptr: stop after the ingest (EDS::parse / parse_sources) and export the index + statistics
    // single_pair != nullptr: merge exactly the symbols (*single_pair, *single_pair + 1) — EDS::merge_adjacent — and emit
    // sink(which, bytes): where result `which` (0 l-EDS, 1 SEDS) goes instead of a fresh malloc'd buffer
    // check_only != nullptr: stop after the first round's pair selection; *check_only = 1 iff no pair exists
    // input_on_device: eds_in / seds_in are device pointers (the VCF front end hands its output over in HBM)

    // genrandomeds-shaped EDS + SEDS text generated in device memory (buffers owned by the pipeline until the next call)
    void genrandomeds(uint64_t n, uint32_t ppm, uint32_t paths, uint64_t seed, eds_buffer* eds_out, eds_buffer* seds_out);

   private:
    struct Bufs;
    eds_ctx* ctx_;
    Bufs* bufs_;
    uint32_t present_dirty_words_ = 0xffffffffu;  // words of the id bitmap that may be non-zero
};

}  // namespace edsb
