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

// Sharded run (shard.cu): every device transforms a slice of the record lines. The tie order of records that share a
// position comes from ONE sort over ALL records (the reference's std::sort is unstable: vcf_transforms.cpp:715-718), and a
// slice only renders the reference bases of its own range, so the slices meet in the middle of the transform.
struct VcfShardHook {
    virtual ~VcfShardHook() {}
    // Called exactly once per transform, after the slice's records are known (genotype kernel in flight). pos: POS of
    // the slice's records in file order; max_end: largest 0-based record end. Blocks until every slice has arrived.
    // Out: perm (empty = file order) = the slice's records in the global order, indices local to the slice;
    // [ref_lo, ref_hi) = the reference bases this slice renders. Throws when the slices cannot be joined.
    virtual void exchange(const std::vector<uint64_t>& pos, uint64_t max_end, uint64_t n_bases, std::vector<uint32_t>* perm,
                          uint64_t* ref_lo, uint64_t* ref_hi) = 0;
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
                          std::vector<uint64_t>* sv_lines, VcfShardHook* hook = nullptr);

   private:
    struct Bufs;
    eds_ctx* ctx_;
    Bufs* bufs_;
    uint32_t words_hint_ = 0;  // sample-bitset width of the previous call
    uint64_t id_text_n_ = 0;   // entries of the id -> decimal text table built so far
};

}  // namespace edsb
