// Device-side structures shared by the MSA kernels and their host launcher.
#pragma once
#include <stdint.h>

namespace edsb {

constexpr uint32_t kCommonFlag = 0x80000000u;  // bit 31 of a run/symbol entry: conserved run / common symbol
constexpr uint32_t kColMask = 0x7fffffffu;
constexpr uint32_t kEmptySlot = 0xffffffffu;

// Geometry of a column window in device memory (see eds_msa_view in include/edsparser_b200.h).
// p-space: byte positions of a row segment shifted so that row 0 is 16-byte aligned,
// p = (u - u_begin) + a0 with u = c + c / lw the position of column c inside a wrapped row.
struct MsaGeom {
    const uint8_t* text;      // device, 16-byte aligned
    const uint64_t* row_off;  // device, R entries: byte offset of the first held residue of each row
    uint64_t n_vec;           // readable 16-byte vectors in text
    uint64_t total_cols;
    uint64_t col_begin;  // first held column (global)
    uint64_t u_begin;    // col_begin + col_begin / lw
    uint64_t row_bytes;  // bytes per held row segment (first..last held residue, newlines inside)
    uint64_t sum_id_width;  // sum of decimal_width(i), i = 1..R
    uint64_t hash_mask;     // all ones; tests narrow it to force hash collisions
    long long d_min_vec, d_max_vec;  // min / max over rows of floor((row_off[r] - a0) / 16)
    uint32_t R;
    uint32_t Rp;  // R rounded up to a multiple of 32 (row pitch of the column stash)
    uint32_t lw;
    uint32_t ncols;   // held columns
    uint32_t own_lo;  // owned range, window-relative columns
    uint32_t own_hi;
    uint32_t a0;        // row_off[0] & 15
    uint32_t n_chunks;  // 16-byte chunks covering [0, a0 + row_bytes) in p-space
    uint32_t n_words;   // ceil(ncols / 32)
    uint32_t l;
    uint32_t leds;   // 1: l-EDS boundaries (build_leds_boundaries), 0: plain EDS
    uint32_t alt32;  // alternative ids stored as uint32 (R > 65535) instead of uint16
    uint32_t cls[5];       // k_scan: rows of word-shift class c are row_pack[cls[c]..cls[c+1])
    uint32_t all_aligned;  // every row is 16-byte congruent with row 0
};

enum MsaAbort : uint32_t {
    kAbortNone = 0,
    kAbortVarCap = 1,
    kAbortRunsCap = 2,
    kAbortEdsCap = 3,
    kAbortSedsCap = 4,
};

enum MsaBad : uint32_t {
    kBadNewlineLayout = 1,  // a line break is missing / misplaced in some row
    kBadResidueByte = 2,    // '\n' where a residue is expected
};

// Device-resident status block, copied to the host once per call.
struct MsaStatus {
    uint32_t n_var;   // variable columns in the window
    uint32_t n_runs;  // runs in the window
    uint32_t n_syms;  // symbols opening anywhere in the window
    uint32_t k_lo, k_hi;  // owned symbols are [k_lo, k_hi)
    uint32_t n_var_syms;  // owned variable symbols
    uint32_t n_varsyms_window;  // variable symbols in the window
    uint32_t v_lo, v_hi;        // owned slice of varsym[]
    uint32_t n_hard, n_wide;    // entries of hardlist / widelist
    uint32_t n_easy;            // entries of easylist
    uint32_t n_emit2;           // entries of emitlist
    uint32_t lead_lo, lead_hi;  // window columns of conserved text continuing a lower shard's symbol
    uint32_t lead_close;        // 1: that symbol's '}' belongs to this shard
    uint32_t tail_open;         // 1: the last owned symbol is conserved and is closed by a higher shard
    uint32_t abort;
    uint32_t bad_msa;
    uint32_t halo_fail;
    uint32_t fz_need;  // k_scan_fused: most variable columns any cluster found (its region of the temporary stash overflowed)
    uint64_t need_var, need_runs, need_eds, need_seds;  // capacities wanted by the stage that aborted
    unsigned long long eds_total, seds_total, n_alts;
    uint64_t first_open_col;
};

struct MsaBufs {
    uint32_t* mism;      // p-space mismatch bits (uint16 per chunk, read back as uint32 words)
    uint32_t* vbits;     // window columns: 1 = variable (not conserved)
    uint32_t* tbits;     // 1 = a run starts at this column
    uint32_t* rankdir;   // variable columns before each 32-column word
    uint8_t* refc;       // row 0 in column space (no newlines), n_words * 32 bytes
    uint2* part_cnt;     // per partition {variable columns, run starts}
    unsigned long long* part_sym;  // per run partition: symbols opened | variable symbols opened << 32
    unsigned long long* part_sz;  // per symbol partition: {eds bytes, seds bytes}
    uint32_t* varcol;    // window column of the k-th variable column
    uint32_t* runs;      // run k: start column | kCommonFlag; runs[n_runs] = ncols
    uint32_t* sym;       // symbol k: start column | kCommonFlag; sym[n_syms] = ncols
    uint32_t* varsym;    // indices of the variable symbols, ascending
    uint32_t* widelist;  // multi-column variable symbols (any order), queued by k_group
    uint32_t* easylist;  // ... of which the tuple formulation takes (k_group3 / k_emit3)
    uint32_t* hardlist;  // ... and the rest: hashed row path (k_group2)
    uint32_t* emitlist;  // symbols k_emit2 renders row by row: hardlist, and with few rows easylist's too
    uint8_t* stash;      // stash[k * Rp + r] = residue of row r at the k-th variable column
    void* altid;         // altid[slot0 * Rp + r]: alternative index of row r in the symbol whose first variable column is slot0
    uint32_t* leadmask;  // leadmask[slot0 * Rp/32 + r/32]: rows that introduce an alternative
    // rows-across-lanes path (many rows): k_group leaves the rows of each of the (at most 8) alternatives of a
    // single-column symbol as bitsets, rowbits[(slot0 * 8 + a) * Rp/32 + r/32], and the residue list, seen[slot0 * 3 ..]
    // = {lo, hi, n}; k_emit_var renders them (idlist.cuh). Null when the lane-per-symbol path is in use.
    uint32_t* rowbits;
    uint32_t* seen;
    const unsigned long long* id_text;  // id -> decimal text table, R + 1 entries
    uint32_t* sym_nalts;
    unsigned long long* sym_edsz;
    unsigned long long* eds_off;
    unsigned long long* seds_off;
    uint8_t* eds_out;
    uint8_t* seds_out;
    uint8_t* group_ws;  // global scratch for k_group / k_emit_var when R is too large for shared memory
    MsaStatus* status;
    uint32_t cap_var, cap_runs;
    uint64_t cap_eds, cap_seds;
};

}  // namespace edsb
