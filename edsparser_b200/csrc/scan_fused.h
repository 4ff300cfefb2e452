// Launch geometry of k_scan_fused (scan_fused.cuh), shared by the kernel and the host launcher.
#pragma once
#include "msa_kernels.cuh"
#include "pipe.cuh"

namespace edsb {

// T = 16-byte chunks per tile (32: lane = chunk; 16: two rows per warp instruction, twice the ring depth);
// a staged row takes 16 T + 16 bytes (the aligned superset of its T chunks)
constexpr uint32_t kFzMaxNC = 8;                 // portable cluster size
constexpr uint32_t kFzGroupRows = 128;           // most rows of one CTA (4 per lane in the gather)
constexpr uint32_t kFzMaxPW = 4;                 // most producer warps (more gain nothing: profiles/r02_a)
constexpr uint32_t kFzMaxDW = 4;                 // most duty warps
constexpr uint32_t kFzMaxDirect = 32;            // most rows of a CTA that bypass the ring (4 per consumer warp)
#ifdef EDSB_EMU
constexpr int kFzCW = 2;
#else
#ifndef EDSB_FZ_CW
#define EDSB_FZ_CW 8
#endif
constexpr int kFzCW = EDSB_FZ_CW;  // consumer warps
#endif

struct FzParams {
    const unsigned long long* pack;  // [NC][slot_pitch]: (address of the vector holding p-space byte 0) | byte shift; slot 0 = row 0
    const uint32_t* meta;            // [NC][8]: slots, class boundaries cls[0..4] (slot indices), first row, rows
    const uint16_t* info;            // [NC][RG]: local row -> slot << 4 | shift (0xffff: no such row)
    uint16_t* mism16;
    uint8_t* tmp_stash;              // [regions * capc][Rp]
    unsigned long long* tmp_col;     // [regions * capc]: p-space byte position of the slot's column
    uint32_t* region_count;          // [regions]
    uint32_t S, NC, RG, slot_pitch, n_tiles, capc, all_aligned;
    uint32_t T;     // chunks per tile (the kernel instantiation)
    uint32_t PW;    // producer warps
    uint32_t DW;    // duty warps (mask + gather), tiles in rotation
    uint32_t probe; // timing probes (EDSB_FUSED_PROBE; results are NOT valid): 1 = consumers skip the compare, 2 = duty warps skip everything (no variable columns), 4 = duty warps skip the gather
    uint32_t mode;  // 0: one bulk copy (TMA) per row, 1: 16-byte cp.async per lane, 2: both (first PWB warps bulk, slots [0, n_bulk))
    uint32_t PWB, n_bulk;
    uint32_t split;        // one `full` barrier per stage and producer warp instead of one per stage
    uint32_t n_direct;     // the last n_direct slots of a CTA are not staged: the consumers load them straight into registers
    uint32_t stage_slots;  // rows a stage holds (= slot_pitch when nothing is direct)
    long long tile_lo_ok, tile_hi_ok;  // tiles [lo, hi) can be fetched with bulk copies (every vector inside the buffer)
};

inline size_t fz_smem_bytes(uint32_t T, uint32_t S, uint32_t NC, uint32_t RG, uint32_t slot_pitch, uint32_t DW, uint32_t stage_slots) {
    size_t b = (size_t)S * stage_slots * (16u * T + 16u);     // stages
    b += (size_t)S * 32 * 4;                                  // red16
    b += NC > 1 ? (size_t)2 * S * NC * 32 * 4 : 0;            // mask_in
    b += (size_t)slot_pitch * 8;                              // s_pack
    b += ((size_t)RG * 2 + 15) & ~(size_t)15;                 // s_info
    b += (((size_t)slot_pitch + 3) & ~(size_t)3) * 4;         // s_off16
    b += (size_t)DW * 16 * T * 2;                             // s_vpos
    b += (size_t)(4 * S + S * kFzMaxPW) * sizeof(Mbar) + 16;  // barriers: full [S][kFzMaxPW], empty, red_full [S], maskbar [2S]
    return b;
}

}  // namespace edsb
