// Host-side launcher of the MSA -> EDS / l-EDS kernels (msa.cu).
#pragma once
#include <stdexcept>

#include "ctx.h"
#include "msa_kernels.cuh"
#include "scan_fused.h"

namespace edsb {

struct BadMsa : std::runtime_error {
    using std::runtime_error::runtime_error;
};
struct HaloError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

class MsaPipeline {
   public:
    explicit MsaPipeline(eds_ctx* ctx);
    ~MsaPipeline();
    MsaPipeline(const MsaPipeline&) = delete;
    MsaPipeline& operator=(const MsaPipeline&) = delete;

    // Outputs point into buffers owned by this object (valid until the next call).
    void transform(const eds_msa_view& view, uint32_t l, int leds, eds_buffer* eds_out, eds_buffer* seds_out,
                   eds_msa_stats* stats);
    void conserved_bits(const eds_msa_view& view, uint8_t* out_bits, uint64_t out_bytes);

   private:
    uint32_t partitions() const;
    void prepare(const eds_msa_view& view, uint32_t l, int leds);
    void bind(MsaBufs& b);
    void launch_scan(const MsaBufs& b, bool allow_fused);
    void plan_fused();
    void run_once(MsaBufs& b);

    eds_ctx* ctx_;
    MsaGeom geom_;
    MsaStatus* h_status_ = nullptr;  // pinned
    std::vector<uint64_t> h_rows_;
    DevBuf d_rows_, d_mism_, d_vbits_, d_tbits_, d_rank_, d_refc_, d_part_, d_varcol_, d_runs_, d_sym_, d_stash_,
        d_altid_, d_leadmask_, d_symmeta_, d_eds_, d_seds_, d_ws_, d_status_, d_rowbits_, d_seen_, d_id_text_;
    // k_scan_fused (scan_fused.cuh): geometry of the last prepare(); on = false -> k_scan + k_stash
    struct FzPlan {
        bool on = false;
        uint32_t S = 0, NC = 1, RG = 0, slot_pitch = 0, regions = 0, capc = 0, PW = 1, DW = 1, T = 32, H = 1, CW = 8;
        bool direct = false;  // k_scan_fused with rows that bypass the ring
        bool l2 = false;  // k_scan_l2 (caches as the stage) instead of k_scan_fused (shared-memory ring)
        size_t smem = 0;
    } fz_;
    FzParams fzp_;
    std::vector<unsigned char> h_fz_;
    DevBuf d_fz_rows_, d_fz_tmp_, d_fz_col_, d_fz_cnt_;
    size_t fz_attr_smem_ = 0;      // largest dynamic shared memory size set on k_scan_fused so far
    uint32_t fz_occ_key_ = 0, fz_occ_regions_ = 0;  // cached cluster occupancy (key = NC, S, slot_pitch)
    uint64_t sym_plan_key_ = ~0ull;  // launch plan of the per-symbol kernels: occupancies looked up once per row count
    int sym_occ_[6] = {1, 1, 1, 1, 1, 1};
    uint64_t id_text_n_ = 0;  // entries of the id -> text table built so far
    uint32_t cap_var_ = 0, cap_runs_ = 0;
    uint64_t cap_eds_ = 0, cap_seds_ = 0;
};

void msa_synth(eds_ctx* ctx, uint32_t n_rows, uint64_t total_cols, uint32_t lw, uint64_t col_begin, uint64_t col_count,
               uint64_t seed, uint32_t variable_ppm, eds_msa_view* view);

}  // namespace edsb
