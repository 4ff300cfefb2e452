// Host-side launcher of the VCF front end (vcf.cu): parse_vcf_to_eds_streaming / parse_vcf_to_leds_streaming
// (draessld/EDSParser src/cpp/lib/transforms/vcf_transforms.cpp:677-755) behind eds_vcf_transform_*.
#pragma once
#include <stdexcept>
#include <vector>

#include "ctx.h"

namespace edsb {

// Input outside the domain where the reference is well defined (SURVEY.md C.4 + DESIGN.md §4b).
struct BadVcf : std::runtime_error {
    using std::runtime_error::runtime_error;
};

class VcfPipeline {
   public:
    explicit VcfPipeline(eds_ctx* ctx);
    ~VcfPipeline();
    VcfPipeline(const VcfPipeline&) = delete;
    VcfPipeline& operator=(const VcfPipeline&) = delete;

    // VCF + FASTA text already in device memory (16-byte aligned, readable up to the next 16-byte boundary)
    // -> EDS / SEDS text in ctx-owned device memory. sv_lines (optional) receives the byte offsets of the lines
    // skipped for an unsupported symbolic ALT, in file order (the caller prints the reference's warnings).
    void transform_device(const uint8_t* vcf, uint64_t vcf_bytes, const uint8_t* fasta, uint64_t fasta_bytes,
                          eds_buffer* eds_out, eds_buffer* seds_out, eds_vcf_stats* stats,
                          std::vector<uint64_t>* sv_lines);

   private:
    struct Bufs;
    eds_ctx* ctx_;
    Bufs* bufs_;
    uint32_t words_hint_ = 0;  // sample-bitset width of the previous call
    uint64_t id_text_n_ = 0;   // entries of the id -> decimal text table built so far
};

}  // namespace edsb
