// The library context behind eds_ctx (include/edsparser_b200.h): device, stream, launch knobs,
// per-kernel event timing, and the per-path pipelines with their grow-only device buffers.
#pragma once
#include <string>
#include <vector>

#include "../../include/edsparser_b200.h"
#include "common.cuh"

namespace edsb {
class MsaPipeline;
class LedsPipeline;
class VcfPipeline;

// One event pair per launch when profiling is on; resolved to milliseconds after the stream syncs.
struct KernelClock {
    bool on = false;
    cudaStream_t stream = nullptr;
    std::vector<const char*> names;
    std::vector<cudaEvent_t> pool;  // 2 per launch
    std::vector<float> ms;
    uint32_t launches = 0;

    void reset() {
        names.clear();
        ms.clear();
        launches = 0;
    }
    cudaStream_t cur = nullptr;  // stream of the launch being timed
    void begin(const char* name, cudaStream_t on_stream = nullptr) {
        ++launches;
        cur = on_stream ? on_stream : stream;
        if (!on) return;
        size_t i = names.size();
        names.push_back(name);
        while (pool.size() < 2 * (i + 1)) {
            cudaEvent_t e;
            EDSB_CUDA(cudaEventCreate(&e));
            pool.push_back(e);
        }
        EDSB_CUDA(cudaEventRecord(pool[2 * i], cur));
    }
    void end() {
        if (!on) return;
        EDSB_CUDA(cudaEventRecord(pool[2 * (names.size() - 1) + 1], cur));
    }
    void resolve() {
        ms.assign(names.size(), 0.f);
        if (!on) return;
        for (size_t i = 0; i < names.size(); ++i) EDSB_CUDA(cudaEventElapsedTime(&ms[i], pool[2 * i], pool[2 * i + 1]));
    }
    void release() {
        for (cudaEvent_t e : pool) cudaEventDestroy(e);
        pool.clear();
    }
};
}  // namespace edsb

struct eds_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    // side streams + events: independent kernels of one transform overlap (fork/join by events, no host sync)
    cudaStream_t aux[2] = {nullptr, nullptr};
    cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool serial = false;  // EDSB_DEBUG_SERIAL=1: everything on the main stream
    // k_scan_fused (scan + variable-column gather in one pass, scan_fused.cuh); EDSB_FUSED=0 selects k_scan + k_stash
    bool fused = true;
    uint32_t fused_min_rows = 32;  // shallower alignments: a tile is too small to pay for the staging
    uint32_t fused_nc = 0;         // EDSB_FUSED_NC: force the cluster size (0 = by row count)
    uint32_t fused_stages = 0;     // EDSB_FUSED_STAGES: cap the ring depth (0 = what fits)
    uint32_t fused_pw = 0;         // EDSB_FUSED_PW: producer warps (0 = default)
    uint32_t fused_pwb = 0;        // EDSB_FUSED_PWB: mode 2, producer warps that issue bulk copies (0 = half)
    uint32_t fused_bulk_pct = 0;   // EDSB_FUSED_BULK_PCT: mode 2, share of the rows fed by bulk copies (0 = 50)
    uint32_t fused_t = 0;          // EDSB_FUSED_T: 16 or 32 chunks per tile (0 = default)
    uint32_t fused_dw = 0;         // EDSB_FUSED_DW: duty warps (0 = default)
    uint32_t fused_pair = 0;       // EDSB_FUSED_PAIR=1: two adjacent tiles per bulk copy (1040 bytes; measured slower: half the ring slots)
    uint32_t fused_l2 = 0;         // EDSB_FUSED_L2=1: k_scan_l2 (plain loads, L1 / L2 as the stage; measured slower) instead of k_scan_fused (TMA ring in shared memory)
    uint32_t fused_cw = 0;         // EDSB_FUSED_CW: consumer warps per CTA of k_scan_l2 (0 = default)
    uint32_t fused_direct = 0;     // EDSB_FUSED_DIRECT: rows per CTA of k_scan_fused that bypass the ring (<= 32)
    uint32_t fused_split = 0;      // EDSB_FUSED_SPLIT=1: k_scan_fused with one `full` barrier per stage and producer warp
    uint32_t group_cta = 1;        // EDSB_DEBUG_GROUP_CTA: k_group2 gives a symbol a whole block when the symbols are few (1), never (0), always (2)
    uint32_t peer_headroom = 0;    // bytes of shared memory per SM the fused scan leaves free (set by eds_comm_create: the all-gather's
                                   // kernel must fit beside the persistent scan CTAs, or it holds an SM's CTA back until its peers arrive)
    uint32_t fused_probe = 0;      // EDSB_FUSED_PROBE: timing probes of k_scan_fused (scan_fused.h); results are not valid
    uint32_t fused_mode = 0;       // EDSB_FUSED_MODE: 0 = one bulk copy (TMA) per row (measured faster), 1 = 16-byte cp.async per lane
    int sm_count = 148;
    size_t smem_optin = 227 * 1024;
    uint32_t partitions = 0;          // 0 = default
    uint32_t scan_blocks_per_sm = 0;  // 0 = default
    uint32_t scan_row_slices = 0;     // tests (EDSB_DEBUG_ROW_SLICES): force k_scan's row split; 0 = automatic
    int tuple_off = 0;   // tests (EDSB_DEBUG_TUPLE_OFF): every multi-column symbol on the hashed row path
    // single-column symbols (EDSB_DEBUG_NARROW_OFF): 0 (default) = lane per symbol up to 160 rows, rows across lanes (a lane
    // group per symbol) beyond; 1 = rows across lanes at every row count (config 2: k_emit_var 0.116 -> 0.068 ms but
    // k_group 0.026 -> 0.044 and the three emit kernels crowd each other: step 0.545 -> 0.585 ms, profiles/r02_b);
    // 2 = every symbol on the hashed warp path (tests)
    int narrow_off = 0;
    uint64_t hash_mask = ~0ull;       // tests narrow it (EDSB_HASH_MASK) to exercise the exact-compare fallback
    edsb::KernelClock clock;
    edsb::MsaPipeline* msa = nullptr;
    edsb::LedsPipeline* leds = nullptr;
    edsb::VcfPipeline* vcf = nullptr;
    edsb::DevBuf vcf_in[2];   // eds_vcf_transform_host: the .vcf and .fa bytes on the device
    edsb::DevBuf vcf_out[2];  // EDS / SEDS text of the last VCF transform
    void* host_out[2] = {nullptr, nullptr};  // pinned host copies of the last result (eds_vcf_transform_host_view), grow-only
    size_t host_out_cap[2] = {0, 0};
    // synthetic alignment (eds_msa_synth_device)
    edsb::DevBuf synth_text;
    edsb::DevBuf file_buf;  // eds_msa_transform_host: the .msa bytes on the device
    std::vector<uint64_t> synth_rows;
};
