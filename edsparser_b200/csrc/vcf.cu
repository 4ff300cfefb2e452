// VCF + FASTA -> EDS / SEDS on the GPU: the front end of vcf2eds
// (draessld/EDSParser src/cpp/lib/transforms/vcf_transforms.cpp:51-729) behind eds_vcf_transform_*.
//
// The reference tokenises every line with stringstreams, keeps a vector<vector<int>> of genotypes per site and
// builds std::map<string, set<int>> per group. Here the text stays where it is in HBM and everything is an
// index into it:
//   * FASTA: header / wrap width / next-record / newline census by three small kernels; base q of the first
//     record lives at seq_start + q + q / wrap (read_fasta_region :98-129);
//   * VCF line index: newline count + scatter over 16-byte vectors (two-launch scan);
//   * k_head: thread per line, the first nine fields only (POS, REF, ALT, where the sample columns start);
//   * k_gt: warp per record streams the sample columns with 16-byte loads; token starts come from byte masks and a
//     warp prefix sum, every GT is parsed by the lane that owns its first byte, and the carriers of every allele are
//     collected as one bitset per (record, allele) in shared memory;
//   * records in position order (std::sort is unstable: ties are ordered by the same std::sort call on the host,
//     SURVEY.md C.4, overlapped with k_gt), groups of overlapping records by a max-scan (:503-519), haplotype
//     de-duplication per group without materialising strings (:416-435), carrier bitsets per haplotype (:438-473,
//     :587-624), output sizes, two offset scans, and two emit kernels.
// Everything is integer/byte work; the output is byte-identical to the reference's (tests/test_vcf_*.py).
#include "vcf.h"

#include <string.h>

#include <algorithm>
#include <functional>
#include <string>
#include <utility>
#include <vector>

#include "idlist.cuh"
#include "scan.cuh"

namespace edsb {

namespace {

typedef unsigned long long u64;
constexpr u64 kNone64 = ~0ull;
constexpr uint32_t kNone32 = 0xffffffffu;
constexpr uint32_t kSvListCap = 1u << 20;

enum FaErr : uint32_t { kFaNoHeader = 1, kFaEmpty = 2, kFaZeroWidth = 4, kFaCarriage = 8 };
enum VcfErr : uint32_t { kVcfPosZero = 1, kVcfPastEnd = 2 };
enum LineKind : uint8_t { kHeader = 0, kRecord = 1, kMalformed = 2, kSv = 3 };

struct VcfStatus {
    uint32_t fa_err, err;
    u64 fa_seq_start, fa_lw, fa_rec_end, fa_nl, fa_last_base, fa_bad_nl;
    u64 n_malformed, n_sv;
    u64 bad_line;    // smallest line offset of a record outside the supported domain
    u64 max_tokens;  // largest number of sample columns seen by k_gt
    uint32_t unsorted, pad;
    u64 first_tokens;
    u64 max_end;     // largest 0-based end of a record (sharded runs: no group may span a cut)
};

__device__ __forceinline__ bool is_space(uint8_t c) { return c == ' ' || (c >= 9 && c <= 13); }
__device__ __forceinline__ bool is_digit(uint8_t c) { return c >= (uint8_t)'0' && c <= (uint8_t)'9'; }
__device__ __forceinline__ uint4 load16(const uint8_t* t, u64 chunk) { return __ldg(reinterpret_cast<const uint4*>(t) + chunk); }
__device__ __forceinline__ u64 gtid() { return (u64)blockIdx.x * blockDim.x + threadIdx.x; }
__device__ __forceinline__ u64 gthreads() { return (u64)gridDim.x * blockDim.x; }

// ---------------------------------------------------------------------------------------------------------
// FASTA (parse_fasta_metadata :51-86)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ u64 warp_find_byte(const uint8_t* t, u64 from, u64 n, uint8_t c) {
    const unsigned lane = threadIdx.x & 31;
    for (u64 base = from; base < n; base += 32) {
        const u64 p = base + lane;
        const unsigned b = __ballot_sync(0xffffffffu, p < n && t[p] == c);
        if (b) return base + (u64)(__ffs((int)b) - 1);
    }
    return n;
}

// one warp: header line, start and width of the first sequence line
__global__ void k_fa_head(const uint8_t* fa, u64 n, VcfStatus* st) {
    uint32_t err = 0;
    u64 seq_start = 0, lw = 0;
    if (n == 0 || fa[0] != (uint8_t)'>') {
        err = kFaNoHeader;  // :56-58
    } else {
        const u64 nl1 = warp_find_byte(fa, 0, n, (uint8_t)'\n');
        if (nl1 + 1 >= n) {
            err = kFaEmpty;  // :72-74 (a header without '\n' leaves the stream failed, same message)
        } else {
            seq_start = nl1 + 1;
            lw = warp_find_byte(fa, seq_start, n, (uint8_t)'\n') - seq_start;
            if (lw == 0) err = kFaZeroWidth;
        }
    }
    if (threadIdx.x == 0) {
        st->fa_err = err;
        st->fa_seq_start = seq_start;
        st->fa_lw = lw;
        st->fa_rec_end = n;
        st->fa_nl = 0;
        st->fa_last_base = 0;
        st->fa_bad_nl = kNone64;
    }
}

// start of the next record: the first "\n>" at or after the end of the first sequence line (:79-83)
__global__ void k_fa_next_record(const uint8_t* fa, u64 n, VcfStatus* st) {
    if (st->fa_err) return;
    const u64 lo = st->fa_seq_start + st->fa_lw, n_chunks = (n + 15) / 16;
    for (u64 c = lo / 16 + gtid(); c < n_chunks; c += gthreads()) {
        const uint4 v = load16(fa, c);
        const u64 base = c * 16;
        const uint32_t in = n - base >= 16 ? 0xffffu : low_bits((uint32_t)(n - base));
        const uint32_t nl = eq_bytes16(v, 0x0a0a0a0au) & in, gt = eq_bytes16(v, 0x3e3e3e3eu) & in;
        uint32_t pair = nl & (gt >> 1);
        if ((nl & 0x8000u) && base + 16 < n && fa[base + 16] == (uint8_t)'>') pair |= 0x8000u;
        while (pair) {
            const u64 p = base + (u64)(__ffs((int)pair) - 1);
            pair &= pair - 1;
            if (p >= lo) {
                atomicMin(&st->fa_rec_end, p + 1);
                break;
            }
        }
    }
}

// newline census of the first record: count, last base, first newline off the wrap grid, '\r'
__global__ void k_fa_measure(const uint8_t* fa, VcfStatus* st) {
    if (st->fa_err) return;
    const u64 lo = st->fa_seq_start, hi = st->fa_rec_end, lw = st->fa_lw;
    u64 cnt = 0, last = 0, bad = kNone64;
    uint32_t cr = 0;
    for (u64 c = lo / 16 + gtid(); c * 16 < hi; c += gthreads()) {
        const uint4 v = load16(fa, c);
        const u64 base = c * 16;
        uint32_t in = 0xffffu;
        if (base < lo) in &= ~low_bits((uint32_t)(lo - base));
        if (hi - base < 16) in &= low_bits((uint32_t)(hi - base));
        uint32_t nl = eq_bytes16(v, 0x0a0a0a0au) & in;
        cr |= eq_bytes16(v, 0x0d0d0d0du) & in;
        cnt += (u64)__popc(nl);
        const uint32_t bases = in & ~nl;
        if (bases) last = base + (u64)(32 - __clz((int)bases));  // position + 1
        while (nl) {
            const u64 p = base + (u64)(__ffs((int)nl) - 1);
            nl &= nl - 1;
            if ((p - lo) % (lw + 1) != lw && p < bad) bad = p;
        }
    }
    // one set of atomics per warp
    cnt = warp_sum(cnt);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const u64 l2 = __shfl_xor_sync(0xffffffffu, last, d), b2 = __shfl_xor_sync(0xffffffffu, bad, d);
        last = l2 > last ? l2 : last;
        bad = b2 < bad ? b2 : bad;
        cr |= __shfl_xor_sync(0xffffffffu, cr, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (cnt) atomicAdd(&st->fa_nl, cnt);
        if (last) atomicMax(&st->fa_last_base, last);
        if (bad != kNone64) atomicMin(&st->fa_bad_nl, bad);
        if (cr) atomicOr(&st->fa_err, (uint32_t)kFaCarriage);
    }
}

struct Fasta {
    const uint8_t* t;
    u64 seq_start, lw;
    __device__ __forceinline__ uint8_t at(u64 q) const { return t[seq_start + q + q / lw]; }
};

// ---------------------------------------------------------------------------------------------------------
// VCF lines
// ---------------------------------------------------------------------------------------------------------
// Line index. Lines are long (one per site, thousands of sample columns), so newlines are rare: every warp streams a
// contiguous run of 16-byte vectors, counts its newlines (k_nl_count), the host turns the per-warp counts into
// starting line numbers (it needs the total anyway to size the index), and k_nl_write streams the same run again and
// only falls into the prefix-sum path for the few tiles that do contain a newline. Line i + 1 starts after '\n' i.
__device__ __forceinline__ uint32_t nl_mask(const uint8_t* t, u64 n, u64 c) {
    const uint32_t in = n - c * 16 >= 16 ? 0xffffu : low_bits((uint32_t)(n - c * 16));
    return eq_bytes16(load16(t, c), 0x0a0a0a0au) & in;
}

__device__ __forceinline__ void nl_run(u64 n_chunks, u64& lo, u64& hi) {
    const u64 warp = gtid() >> 5, n_warps = gthreads() >> 5;
    const u64 per = ((n_chunks + n_warps - 1) / n_warps + 31) / 32 * 32;
    lo = warp * per < n_chunks ? warp * per : n_chunks;
    hi = lo + per < n_chunks ? lo + per : n_chunks;
}

__global__ void k_nl_count(const uint8_t* t, u64 n, u64 n_chunks, u64* counts) {
    u64 lo, hi;
    nl_run(n_chunks, lo, hi);
    const unsigned lane = threadIdx.x & 31;
    uint32_t cnt = 0;
    for (u64 c = lo + lane; c < hi; c += 32) cnt += (uint32_t)__popc(nl_mask(t, n, c));
    cnt = warp_sum(cnt);
    if (lane == 0) counts[gtid() >> 5] = cnt;
}

__global__ void k_nl_write(const uint8_t* t, u64 n, u64 n_chunks, const u64* first_line, u64* line_start) {
    u64 lo, hi;
    nl_run(n_chunks, lo, hi);
    const unsigned lane = threadIdx.x & 31;
    u64 line = first_line[gtid() >> 5];  // newlines before this run
    for (u64 c0 = lo; c0 < hi; c0 += 32) {
        const u64 c = c0 + lane;
        uint32_t m = c < hi ? nl_mask(t, n, c) : 0u;
        if (!__any_sync(0xffffffffu, m != 0)) continue;
        const uint32_t cnt = (uint32_t)__popc(m), incl = warp_inclusive_scan(cnt);
        u64 at = line + (incl - cnt);
        while (m) {
            line_start[++at] = c * 16 + (u64)__ffs((int)m);
            m &= m - 1;
        }
        line += __shfl_sync(0xffffffffu, incl, 31);
    }
}

struct LineHead {
    u64 pos, ref_off, alt_off, gt_off;
    uint32_t ref_len, alt_len, nalts;
    uint8_t kind, mode;
    uint16_t pad;
};

struct Rec {
    u64 pos, ref_off, alt_off, gt_off, line_start, line_end, row0, n_gt;
    uint32_t ref_len, alt_len, nalts, mode;
};

// std::getline(ss, piece, delim) over [b, e): a trailing empty piece is not produced
template <typename F>
__device__ __forceinline__ void for_each_piece(const uint8_t* t, u64 b, u64 e, uint8_t delim, F f) {
    u64 at = b;
    while (at < e) {
        u64 d = at;
        while (d < e && t[d] != delim) ++d;
        f(at, d);
        at = d + 1;
    }
}

// ALT piece (:150-172): 0 = plain allele, 1 = <DEL>, 2 = <INS>, 3 = unsupported symbolic allele
__device__ __forceinline__ int alt_class(const uint8_t* t, u64 b, u64 e) {
    if (e - b < 2 || t[b] != (uint8_t)'<' || t[e - 1] != (uint8_t)'>') return 0;
    if (e - b == 5) {
        if (t[b + 1] == 'D' && t[b + 2] == 'E' && t[b + 3] == 'L') return 1;
        if (t[b + 1] == 'I' && t[b + 2] == 'N' && t[b + 3] == 'S') return 2;
    }
    return 3;
}

// std::stoull on [b, e): leading white space, optional sign, at least one digit; the rest is ignored
__device__ __forceinline__ bool parse_stoull(const uint8_t* t, u64 b, u64 e, u64& out) {
    while (b < e && is_space(t[b])) ++b;
    bool neg = false;
    if (b < e && (t[b] == (uint8_t)'+' || t[b] == (uint8_t)'-')) {
        neg = t[b] == (uint8_t)'-';
        ++b;
    }
    if (b >= e || !is_digit(t[b])) return false;
    u64 v = 0;
    for (; b < e && is_digit(t[b]); ++b) {
        const u64 d = (u64)(t[b] - (uint8_t)'0');
        if (v > (kNone64 - d) / 10) return false;  // out_of_range -> malformed line (:288-293)
        v = v * 10 + d;
    }
    out = neg ? (0ull - v) : v;
    return true;
}

// First ten fields of a line under one delimiter rule (tab, or any white space: :262-279). Returns the number of
// fields seen, counting stops when the tenth starts. b/e: bounds of fields 0..4, f9: start of field 9.
__device__ __forceinline__ uint32_t head_fields(const uint8_t* t, u64 ls, u64 le, bool ws, u64* b, u64* e, u64& f9) {
    uint32_t n = 0;
    bool in = false;
    for (u64 p = ls; p < le; ++p) {
        const uint8_t c = t[p];
        const bool d = ws ? is_space(c) : c == (uint8_t)'\t';
        if (!d && !in) {
            if (n < 5) b[n] = p;
            in = true;
            if (++n == 10) {
                f9 = p;
                return n;
            }
        } else if (d && in) {
            if (n <= 5) e[n - 1] = p;
            in = false;
        }
    }
    if (in && n <= 5) e[n - 1] = le;
    return n;
}

// parse_vcf_line :232-326 without the genotypes: thread per line
__global__ void k_head(const uint8_t* t, const u64* line_start, u64 n_lines, LineHead* heads, u64* sv_list, VcfStatus* st) {
    for (u64 i = gtid(); i < n_lines; i += gthreads()) {
        const u64 ls = line_start[i], le = line_start[i + 1] - 1;
        LineHead h;
        memset(&h, 0, sizeof(h));
        h.kind = kHeader;
        if (le > ls && t[ls] != (uint8_t)'#') {
            u64 b[5], e[5], f9 = 0;
            uint32_t nf = head_fields(t, ls, le, false, b, e, f9);
            if (nf < 5) {
                f9 = 0;
                nf = head_fields(t, ls, le, true, b, e, f9);
                h.mode = 1;
            }
            u64 pos = 0;
            if (nf < 5 || !parse_stoull(t, b[1], e[1], pos)) {
                h.kind = kMalformed;
                atomicAdd(&st->n_malformed, 1ull);
            } else {
                uint32_t nalts = 0;
                bool sv = false;
                for_each_piece(t, b[4], e[4], (uint8_t)',', [&](u64 pb, u64 pe) {
                    ++nalts;
                    if (alt_class(t, pb, pe) == 3) sv = true;
                });
                if (sv) {
                    h.kind = kSv;
                    const u64 at = atomicAdd(&st->n_sv, 1ull);
                    if (at < kSvListCap) sv_list[at] = ls;
                } else {
                    h.kind = kRecord;
                    h.pos = pos;
                    h.ref_off = b[3];
                    h.ref_len = (uint32_t)(e[3] - b[3]);
                    h.alt_off = b[4];
                    h.alt_len = (uint32_t)(e[4] - b[4]);
                    h.nalts = nalts;
                    h.gt_off = nf >= 10 ? f9 : 0;
                }
            }
        }
        heads[i] = h;
    }
}

struct RecFn {  // compaction of the record lines, file order
    const LineHead* heads;
    const u64* line_start;
    Rec* recs;
    __device__ u64 value(u64 i) const { return heads[i].kind == kRecord ? 1ull : 0ull; }
    __device__ void apply(u64 i, u64 prefix, u64 v) const {
        if (!v) return;
        const LineHead h = heads[i];
        Rec r;
        r.pos = h.pos;
        r.ref_off = h.ref_off;
        r.alt_off = h.alt_off;
        r.gt_off = h.gt_off;
        r.line_start = line_start[i];
        r.line_end = line_start[i + 1] - 1;
        r.row0 = 0;
        r.n_gt = 0;
        r.ref_len = h.ref_len;
        r.alt_len = h.alt_len;
        r.nalts = h.nalts;
        r.mode = h.mode;
        recs[prefix] = r;
    }
};

struct RowFn {  // one carrier bitset per (record, allele incl. REF)
    Rec* recs;
    __device__ u64 value(u64 k) const { return (u64)recs[k].nalts + 1ull; }
    __device__ void apply(u64 k, u64 prefix, u64) const { recs[k].row0 = prefix; }
};

// ALT strings of every record (:142-176): <DEL> = empty, <INS> = the REF field
__global__ void k_alleles(const uint8_t* t, const Rec* recs, u64 n_rec, u64* al_off, uint32_t* al_len) {
    for (u64 k = gtid(); k < n_rec; k += gthreads()) {
        const Rec r = recs[k];
        u64 row = r.row0;
        al_off[row] = 0;
        al_len[row] = 0;
        for_each_piece(t, r.alt_off, r.alt_off + r.alt_len, (uint8_t)',', [&](u64 pb, u64 pe) {
            ++row;
            const int cls = alt_class(t, pb, pe);
            if (cls == 1) {
                al_off[row] = pb;
                al_len[row] = 0;
            } else if (cls == 2) {
                al_off[row] = r.ref_off;
                al_len[row] = r.ref_len;
            } else {
                al_off[row] = pb;
                al_len[row] = (uint32_t)(pe - pb);
            }
        });
    }
}

// position order already strict? records inside the reference?
__global__ void k_rec_check(const Rec* recs, u64 n_rec, u64* pos_out, VcfStatus* st) {
    const u64 n_bases = st->fa_err ? 0 : st->fa_rec_end - st->fa_seq_start - st->fa_nl;
    u64 my_end = 0;
    for (u64 k = gtid(); k < n_rec; k += gthreads()) {
        const u64 pos = recs[k].pos;
        pos_out[k] = pos;
        if (pos) my_end = max(my_end, pos - 1 + (u64)recs[k].ref_len);
        if (k && pos <= recs[k - 1].pos) st->unsorted = 1;
        uint32_t err = 0;
        if (pos == 0) err = kVcfPosZero;
        else if (pos - 1 > n_bases || (u64)recs[k].ref_len > n_bases - (pos - 1)) err = kVcfPastEnd;
        if (err) {
            atomicOr(&st->err, err);
            atomicMin(&st->bad_line, recs[k].line_start);
        }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) my_end = max(my_end, __shfl_xor_sync(0xffffffffu, my_end, d));
    if ((threadIdx.x & 31) == 0 && my_end) atomicMax(&st->max_end, my_end);
}

// ---------------------------------------------------------------------------------------------------------
// Sample columns
// ---------------------------------------------------------------------------------------------------------
// One warp walks [from, le) in 512-byte tiles. tok(starts, base, first_index) runs per lane with the 16-bit mask of
// fields that start in the lane's 16-byte vector and the index of the first of them; tile_end() is warp-uniform.
template <typename Tok, typename TileEnd>
__device__ __forceinline__ u64 warp_fields(const uint8_t* t, u64 text_chunks, u64 from, u64 le, bool ws, Tok tok, TileEnd tile_end) {
    const unsigned lane = threadIdx.x & 31;
    u64 count = 0;
    uint32_t carry = 0;  // the byte before this tile belongs to a field
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    u64 c0 = from / 16;
    uint4 v = (c0 + lane) * 16 < le ? load16(t, c0 + lane) : zero;
    for (; c0 * 16 < le; c0 += 32) {
        const u64 c = c0 + lane, base = c * 16;
        // the vector after mine (fields that run past my 16 bytes) and my vector of the next tile, both in flight early
        const uint4 vn = (base < le && c + 1 < text_chunks) ? load16(t, c + 1) : zero;
        const uint4 v_next = (c + 32) * 16 < le ? load16(t, c + 32) : zero;
        uint32_t nd = 0;
        if (base < le) {
            uint32_t dl = eq_bytes16(v, 0x09090909u);
            if (ws)
                dl |= eq_bytes16(v, 0x20202020u) | eq_bytes16(v, 0x0b0b0b0bu) | eq_bytes16(v, 0x0c0c0c0cu) |
                      eq_bytes16(v, 0x0d0d0d0du);
            uint32_t in = 0xffffu;
            if (base < from) in &= ~low_bits((uint32_t)(from - base));
            if (le - base < 16) in &= low_bits((uint32_t)(le - base));
            nd = ~dl & in;
        }
        uint32_t prev = __shfl_up_sync(0xffffffffu, nd, 1) >> 15;
        if (lane == 0) prev = carry;
        const uint32_t starts = nd & ~((nd << 1) | prev) & 0xffffu;
        const uint32_t cnt = (uint32_t)__popc(starts);
        const uint32_t incl = warp_inclusive_scan(cnt);
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        tok(starts, base, count + (u64)(incl - cnt), v, vn);
        tile_end();
        carry = __shfl_sync(0xffffffffu, nd, 31) >> 15;
        count += total;
        v = v_next;
    }
    return count;
}

__global__ void k_first_fields(const uint8_t* t, u64 text_chunks, const Rec* recs, VcfStatus* st) {
    const Rec r = recs[0];
    u64 n = 0;
    if (r.gt_off)
        n = warp_fields(t, text_chunks, r.gt_off, r.line_end, r.mode != 0, [](uint32_t, u64, u64, uint4, uint4) {}, []() {});
    if (threadIdx.x == 0) st->first_tokens = n;
}

// parse_genotype :190-216 on the sample field that starts at p; add(a) receives each allele index that resolves
// to REF (a = 0: index 0, negative, or past the ALT list — apply_variant_to_span :363-370) or to ALT a
template <typename Add>
__device__ __forceinline__ void parse_gt(const uint8_t* t, u64 p, u64 le, bool ws, uint32_t nalts, Add add) {
    u64 e = p;
    bool slash = false;
    for (; e < le; ++e) {
        const uint8_t c = t[e];
        if (c == (uint8_t)':' || (ws ? is_space(c) : c == (uint8_t)'\t')) break;
        if (c == (uint8_t)'/') slash = true;
    }
    for_each_piece(t, p, e, slash ? (uint8_t)'/' : (uint8_t)'|', [&](u64 b, u64 pe) {
        if (pe - b == 1 && t[b] == (uint8_t)'.') return;
        while (b < pe && is_space(t[b])) ++b;  // std::stoi
        bool neg = false;
        if (b < pe && (t[b] == (uint8_t)'+' || t[b] == (uint8_t)'-')) {
            neg = t[b] == (uint8_t)'-';
            ++b;
        }
        if (b >= pe || !is_digit(t[b])) return;
        u64 v = 0;
        for (; b < pe && is_digit(t[b]); ++b) {
            v = v * 10 + (u64)(t[b] - (uint8_t)'0');
            if (v > 0x80000000ull) v = 0x80000001ull;
        }
        if (v > 0x80000000ull || (!neg && v > 0x7fffffffull)) return;  // out of int range: stoi throws, allele ignored
        add((!neg && v >= 1 && v <= (u64)nalts) ? (uint32_t)v : 0u);
    });
}

// bits |= v in the warp's shared-memory slice (no result needed: a reduction, not an atomic with return) or,
// for records whose rows do not fit the slice, in global memory
template <bool kShared>
__device__ __forceinline__ void or_into(uint32_t* p, uint32_t v) {
#ifdef EDSB_EMU
    atomicOr(p, v);
#else
    if (kShared)
        asm volatile("red.shared.or.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
    else
        atomicOr(p, v);
#endif
}

// The sample columns of one record: carriers of every allele, dst[a * W + s / 32] bit s % 32.
template <bool kShared>
__device__ __forceinline__ u64 gt_record(const uint8_t* t, u64 text_chunks, const Rec& r, uint32_t W, uint32_t* dst) {
    const uint32_t cap = W * 32u;
    const bool ws = r.mode != 0;
    const uint32_t nalts = r.nalts;
    const u64 le = r.line_end;
    return warp_fields(
        t, text_chunks, r.gt_off, le, ws,
        [&](uint32_t starts, u64 base, u64 first, uint4 v, uint4 vn) {
            if (!starts) return;
            const uint32_t b0 = (uint32_t)(__ffs((int)starts) - 1);
            // the usual vector: four diploid fields "d|d\t" (or "d/d\t") at one phase, checked a word at a time from
            // registers; REF and first-ALT carriers leave as one 4-bit group each (two reductions when it straddles words)
            if (!ws && starts == ((0x1111u << b0) & 0xffffu) && first + 3 < cap) {
                const uint4 x = realign16(v, vn, b0);
                const uint32_t xs[4] = {x.x, x.y, x.z, x.w};
                bool regular = true;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t frame = xs[i] & 0xff00ff00u;
                    regular = regular && (frame == 0x09007c00u || frame == 0x09002f00u) && (xs[i] & 0x00f000f0u) == 0x00300030u &&
                              (((xs[i] & 0x000f000fu) + 0x00060006u) & 0x00100010u) == 0u;
                }
                if (regular) {
                    uint32_t ref4 = 0, alt4 = 0, more = 0;
                    if (nalts == 1u) {  // biallelic site: a digit other than 1 is REF (0, or past the ALT list)
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const uint32_t dd = xs[i] & 0x000f000fu;
                            ref4 |= (uint32_t)(dd != 0x00010001u) << i;
                            alt4 |= (uint32_t)((dd & 0xffffu) == 1u || (dd >> 16) == 1u) << i;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            uint32_t d0 = xs[i] & 0xfu, d1 = (xs[i] >> 16) & 0xfu;
                            d0 = d0 <= nalts ? d0 : 0u;
                            d1 = d1 <= nalts ? d1 : 0u;
                            ref4 |= (uint32_t)(d0 == 0u || d1 == 0u) << i;
                            alt4 |= (uint32_t)(d0 == 1u || d1 == 1u) << i;
                            more |= (d0 | d1) >> 1;
                        }
                    }
                    const uint32_t w = (uint32_t)(first >> 5), sh = (uint32_t)(first & 31);
                    if (ref4) or_into<kShared>(&dst[w], ref4 << sh);
                    if (alt4) or_into<kShared>(&dst[W + w], alt4 << sh);
                    if (sh > 28u) {
                        if (ref4 >> (32u - sh)) or_into<kShared>(&dst[w + 1], ref4 >> (32u - sh));
                        if (alt4 >> (32u - sh)) or_into<kShared>(&dst[W + w + 1], alt4 >> (32u - sh));
                    }
                    if (more) {  // third and later alleles
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const uint32_t d0 = xs[i] & 0xfu, d1 = (xs[i] >> 16) & 0xfu;
                            const u64 smp = first + i;
                            if (d0 >= 2u && d0 <= nalts) or_into<kShared>(&dst[(size_t)d0 * W + (smp >> 5)], 1u << (smp & 31));
                            if (d1 >= 2u && d1 <= nalts) or_into<kShared>(&dst[(size_t)d1 * W + (smp >> 5)], 1u << (smp & 31));
                        }
                    }
                    return;
                }
            }
            u64 s = first;
            while (starts) {
                const uint32_t b = (uint32_t)(__ffs((int)starts) - 1);
                const u64 p = base + b;
                starts &= starts - 1;
                if (s < cap) {
                    const uint32_t w = (uint32_t)(s >> 5), bit = 1u << (s & 31);
                    // one diploid field straight from registers; anything else byte by byte
                    const uint32_t x = realign16(v, vn, b).x;
                    const uint32_t d0 = (x & 0xffu) - (uint32_t)'0', sep = (x >> 8) & 0xffu;
                    const uint32_t d1 = ((x >> 16) & 0xffu) - (uint32_t)'0', term = x >> 24;
                    if (!ws && d0 <= 9u && d1 <= 9u && (sep == (uint32_t)'|' || sep == (uint32_t)'/') && p + 3 <= le &&
                        (p + 3 == le || term == (uint32_t)'\t' || term == (uint32_t)':')) {
                        const uint32_t a0 = d0 <= nalts ? d0 : 0u, a1 = d1 <= nalts ? d1 : 0u;
                        or_into<kShared>(&dst[(size_t)a0 * W + w], bit);
                        if (a1 != a0) or_into<kShared>(&dst[(size_t)a1 * W + w], bit);
                    } else {
                        parse_gt(t, p, le, ws, nalts, [&](uint32_t a) { or_into<kShared>(&dst[(size_t)a * W + w], bit); });
                    }
                }
                ++s;
            }
        },
        []() {});
}

// Warp per record: the carrier bitsets of the record's alleles are built in the warp's shared-memory slice with
// reductions (red.shared.or) and stored once; records whose rows do not fit the slice work in global memory.
__global__ void k_gt(const uint8_t* t, u64 text_chunks, Rec* recs, u64 n_rec, uint32_t W, uint32_t smem_words, uint32_t* bits,
                     VcfStatus* st) {
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    uint32_t* slice = reinterpret_cast<uint32_t*>(EDSB_DYN_SMEM()) + (size_t)wid * smem_words;
    for (u64 k = (u64)blockIdx.x * wpb + wid; k < n_rec; k += (u64)gridDim.x * wpb) {
        const Rec r = recs[k];
        const u64 need = ((u64)r.nalts + 1) * W;
        uint32_t* out = bits + r.row0 * W;
        const bool in_smem = need <= smem_words;
        uint32_t* dst = in_smem ? slice : out;
        for (u64 i = lane; i < need; i += 32) dst[i] = 0;
        __syncwarp();
        u64 n_fields = 0;
        if (r.gt_off) n_fields = in_smem ? gt_record<true>(t, text_chunks, r, W, slice) : gt_record<false>(t, text_chunks, r, W, out);
        __syncwarp();
        if (in_smem)
            for (u64 i = lane; i < need; i += 32) out[i] = slice[i];
        if (lane == 0) {
            recs[k].n_gt = n_fields;
            if (n_fields > *reinterpret_cast<volatile u64*>(&st->max_tokens)) atomicMax(&st->max_tokens, n_fields);
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------
// Groups of overlapping records (group_overlapping_variants :482-534), records in position order
// ---------------------------------------------------------------------------------------------------------
struct SpanFn {  // max-scan of the ends: a record opens a group iff it starts at or after every earlier end
    const Rec* recs;
    const uint32_t* rec_of;
    u64* incl_max;
    uint8_t* opens;
    __device__ u64 value(u64 s) const {
        const Rec& r = recs[rec_of[s]];
        return r.pos - 1 + r.ref_len;
    }
    __device__ void apply(u64 s, u64 prefix, u64 v) const {
        opens[s] = (s == 0 || recs[rec_of[s]].pos - 1 >= prefix) ? 1 : 0;
        incl_max[s] = prefix > v ? prefix : v;
    }
};

// group id of every record; haplotype slots: one per group (the reference span) + one per ALT, in the order
// merge_variant_group generates them (:416-435)
struct GroupFn {
    const Rec* recs;
    const uint32_t* rec_of;
    const uint8_t* opens;
    uint32_t* gid;
    u64* slot_base;
    uint32_t* g_first;
    __device__ u64 value(u64 s) const { return (u64)opens[s] | ((u64)recs[rec_of[s]].nalts << 32); }
    __device__ void apply(u64 s, u64 prefix, u64) const {
        const uint32_t g = (uint32_t)prefix + opens[s] - 1u;
        gid[s] = g;
        slot_base[s] = (prefix >> 32) + g;
        if (opens[s]) g_first[g] = (uint32_t)s;
    }
};

__global__ void k_group_info(const Rec* recs, const uint32_t* rec_of, const u64* incl_max, uint32_t* g_first,
                             uint32_t n_groups, u64 n_rec, u64 n_bases, u64* g_from, u64* g_to) {
    for (u64 g = gtid(); g <= n_groups; g += gthreads()) {
        if (g == n_groups) {
            g_first[g] = (uint32_t)n_rec;
            g_from[g] = g_to[g] = n_bases;
        } else {
            const uint32_t first = g_first[g];
            const uint32_t last = (g + 1 < n_groups ? g_first[g + 1] : (uint32_t)n_rec) - 1u;
            g_from[g] = recs[rec_of[first]].pos - 1;
            g_to[g] = incl_max[last];
        }
    }
}

// A haplotype of a group = the reference span with one allele of one record substituted
// (apply_variant_to_span :356-390); strings are never built, characters are computed on demand.
struct Hap {
    u64 from, alt_at;
    uint32_t span_len, off, ref_len, alt_len;
    bool plain;
    __device__ __forceinline__ uint32_t len() const { return plain ? span_len : span_len - ref_len + alt_len; }
    __device__ __forceinline__ uint8_t at(uint32_t i, const Fasta& fa, const uint8_t* t) const {
        if (plain || i < off) return fa.at(from + i);
        if (i < off + alt_len) return t[alt_at + (i - off)];
        return fa.at(from + (i - alt_len + ref_len));
    }
};

struct GroupView {
    const Rec* recs;
    const uint32_t* rec_of;
    const u64* slot_base;
    const u64* al_off;
    const uint32_t* al_len;
    const u64* g_from;
    const u64* g_to;
    __device__ __forceinline__ Hap hap(uint32_t g, uint32_t s, uint32_t a) const {
        const Rec& r = recs[rec_of[s]];
        Hap h;
        h.from = g_from[g];
        h.span_len = (uint32_t)(g_to[g] - g_from[g]);
        h.plain = a == 0;
        h.off = (uint32_t)(r.pos - 1 - h.from);
        h.ref_len = r.ref_len;
        h.alt_at = al_off[r.row0 + a];
        h.alt_len = al_len[r.row0 + a];
        return h;
    }
};

__device__ __forceinline__ bool same_hap(const Hap& x, const Hap& y, const Fasta& fa, const uint8_t* t) {
    const uint32_t n = x.len();
    if (n != y.len()) return false;
    for (uint32_t i = 0; i < n; ++i)
        if (x.at(i, fa, t) != y.at(i, fa, t)) return false;
    return true;
}

// thread per group: equal haplotype strings share the slot of their first occurrence (:429-433)
__global__ void k_haps(const uint8_t* t, Fasta fa, GroupView gv, const uint32_t* g_first, uint32_t n_groups, u64* canon,
                       uint32_t* hap_len) {
    for (u64 g64 = gtid(); g64 < n_groups; g64 += gthreads()) {
        const uint32_t g = (uint32_t)g64, first = g_first[g], last = g_first[g + 1] - 1u;
        const u64 ref_slot = gv.slot_base[first];
        const Hap ref = gv.hap(g, first, 0);
        canon[ref_slot] = ref_slot;
        hap_len[ref_slot] = ref.len();
        for (uint32_t s = first; s <= last; ++s) {
            const uint32_t na = gv.recs[gv.rec_of[s]].nalts;
            for (uint32_t a = 1; a <= na; ++a) {
                const u64 slot = gv.slot_base[s] + a;
                const Hap h = gv.hap(g, s, a);
                const uint32_t n = h.len();
                u64 c = slot;
                if (same_hap(h, ref, fa, t)) {
                    c = ref_slot;
                } else {
                    for (uint32_t s2 = first; s2 <= s && c == slot; ++s2) {
                        const uint32_t n2 = s2 == s ? a - 1 : gv.recs[gv.rec_of[s2]].nalts;
                        for (uint32_t a2 = 1; a2 <= n2; ++a2) {
                            const u64 slot2 = gv.slot_base[s2] + a2;
                            if (canon[slot2] == slot2 && hap_len[slot2] == n && same_hap(h, gv.hap(g, s2, a2), fa, t)) {
                                c = slot2;
                                break;
                            }
                        }
                    }
                }
                canon[slot] = c;
                hap_len[slot] = n;
            }
        }
    }
}

// Warp per group: carriers per haplotype slot (:438-473, :587-599), which slots are printed (:619-624), sizes.
__global__ void k_combine(const Rec* recs, const uint32_t* rec_of, const u64* slot_base, const uint32_t* g_first,
                          uint32_t n_groups, const uint32_t* bits, uint32_t W, const u64* canon, const uint32_t* hap_len,
                          uint32_t* slot_bits, uint32_t* slot_seds, uint8_t* kept, u64* g_eds, u64* g_seds) {
    const unsigned lane = threadIdx.x & 31;
    const u64 warp = gtid() >> 5, n_warps = gthreads() >> 5;
    for (u64 g = warp; g < n_groups; g += n_warps) {
        const uint32_t first = g_first[g], last = g_first[g + 1] - 1u;
        const u64 ref_slot = slot_base[first];
        const u64 end_slot = slot_base[last] + recs[rec_of[last]].nalts + 1;
        const u64 n_s = recs[rec_of[first]].n_gt;  // :407
        for (uint32_t w = lane; w < W; w += 32) {
            const u64 lo = (u64)w * 32u;
            const uint32_t mask = n_s >= lo + 32u ? 0xffffffffu : (n_s <= lo ? 0u : low_bits((uint32_t)(n_s - lo)));
            for (u64 sl = ref_slot; sl < end_slot; ++sl) slot_bits[sl * W + w] = 0;
            uint32_t any = 0;
            for (uint32_t s = first; s <= last; ++s) {
                const Rec& r = recs[rec_of[s]];
                for (uint32_t a = 0; a <= r.nalts; ++a) {
                    const uint32_t v = bits[(r.row0 + a) * W + w] & mask;
                    if (!v) continue;
                    const u64 c = a == 0 ? ref_slot : canon[slot_base[s] + a];
                    slot_bits[c * W + w] |= v;
                    any |= v;
                }
            }
            slot_bits[ref_slot * W + w] |= mask & ~any;  // :466-468
        }
        __syncwarp();
        u64 eds = 2, seds = 0;
        uint32_t n_kept = 0;
        for (u64 sl = ref_slot; sl < end_slot; ++sl) {
            uint32_t bytes = 0, cnt = 0;
            for (uint32_t w = lane; w < W; w += 32) {
                const uint32_t v = slot_bits[sl * W + w];
                bytes += word_id_bytes(w, v);
                cnt += (uint32_t)__popc(v);
            }
            bytes = warp_sum(bytes);
            cnt = warp_sum(cnt);
            const bool keep = canon[sl] == sl && (n_s == 0 || cnt > 0);
            const uint32_t sb = (n_s != 0 && cnt > 0) ? bytes + 1u : 0u;  // '{' + "id," each, the last ',' is the '}'
            if (lane == 0) {
                kept[sl] = keep ? 1 : 0;
                slot_seds[sl] = sb;
            }
            if (keep) {
                eds += hap_len[sl] + (n_kept ? 1u : 0u);
                seds += sb;
                ++n_kept;
            }
        }
        if (n_s == 0) seds = 3;  // "{0}" :603-615
        if (lane == 0) {
            g_eds[g] = eds;
            g_seds[g] = seds;
        }
    }
}

// offsets: entry g = the common text in front of group g (if any) followed by group g; entry n_groups = the tail
struct OutFn {
    const u64* g_from;
    const u64* g_to;
    const u64* g_bytes;
    uint32_t n_groups, common_extra;  // 2 for "{}" around the text, 3 for "{0}"
    uint32_t text;                    // 1: count the text itself
    u64* off;
    u64 ref_lo;                       // first reference base of this run (0 unless sharded)
    __device__ u64 value(u64 g) const {
        const u64 clen = g_from[g] - (g ? g_to[g - 1] : ref_lo);
        return (clen ? (text ? clen : 0ull) + common_extra : 0ull) + (g < n_groups ? g_bytes[g] : 0ull);
    }
    __device__ void apply(u64 g, u64 prefix, u64) const { off[g] = prefix; }
};

// common text: thread per 16 reference positions, group lookup by binary search, then a linear walk
__global__ void k_emit_ref(Fasta fa, u64 ref_lo, u64 n_bases, const u64* g_from, const u64* g_to, uint32_t n_groups,
                           const u64* eds_off, uint8_t* out) {
    for (u64 q0 = ref_lo + gtid() * 16; q0 < n_bases; q0 += gthreads() * 16) {
        // g = number of groups that start at or before q0
        uint32_t lo = 0, hi = n_groups;
        while (lo < hi) {
            const uint32_t mid = lo + (hi - lo) / 2;
            if (g_from[mid] <= q0) lo = mid + 1;
            else hi = mid;
        }
        uint32_t g = lo;
        const u64 q1 = q0 + 16 < n_bases ? q0 + 16 : n_bases;
        for (u64 q = q0; q < q1; ++q) {
            while (g < n_groups && g_from[g] <= q) ++g;
            const u64 cursor = g ? g_to[g - 1] : ref_lo;
            if (q >= cursor) out[eds_off[g] + 1 + (q - cursor)] = fa.at(q);
        }
    }
}

// warp per entry: braces and "{0}" of the common text, then the group: {hap,hap,...} and {ids}{ids}...
__global__ void k_emit_groups(const uint8_t* t, Fasta fa, GroupView gv, const uint32_t* g_first, uint32_t n_groups,
                              const u64* canon, const uint32_t* hap_len, const uint8_t* kept, const uint32_t* slot_bits,
                              const uint32_t* slot_seds, uint32_t W, uint32_t stage_bytes, const u64* id_text,
                              const u64* eds_off, const u64* seds_off, uint8_t* out, uint8_t* sout, u64 ref_lo) {
    const unsigned lane = threadIdx.x & 31;
    const u64 warp = gtid() >> 5, n_warps = gthreads() >> 5;
    uint8_t* const stage = EDSB_DYN_SMEM() + (size_t)(threadIdx.x >> 5) * stage_bytes;
    for (u64 g64 = warp; g64 <= n_groups; g64 += n_warps) {
        const uint32_t g = (uint32_t)g64;
        u64 eo = eds_off[g], so = seds_off[g];
        const u64 clen = gv.g_from[g] - (g ? gv.g_to[g - 1] : ref_lo);
        if (clen) {
            if (lane == 0) {
                out[eo] = (uint8_t)'{';
                out[eo + 1 + clen] = (uint8_t)'}';
                sout[so] = (uint8_t)'{';
                sout[so + 1] = (uint8_t)'0';
                sout[so + 2] = (uint8_t)'}';
            }
            eo += clen + 2;
            so += 3;
        }
        if (g == n_groups) continue;
        const uint32_t first = g_first[g], last = g_first[g + 1] - 1u;
        const u64 n_s = gv.recs[gv.rec_of[first]].n_gt;
        if (lane == 0) out[eo] = (uint8_t)'{';
        ++eo;
        if (n_s == 0 && lane == 0) {
            sout[so] = (uint8_t)'{';
            sout[so + 1] = (uint8_t)'0';
            sout[so + 2] = (uint8_t)'}';
        }
        bool first_hap = true;
        for (uint32_t s = first; s <= last; ++s) {
            const uint32_t na = gv.recs[gv.rec_of[s]].nalts;
            for (uint32_t a = (s == first ? 0u : 1u); a <= na; ++a) {
                const u64 slot = gv.slot_base[s] + a;
                if (!kept[slot]) continue;
                if (!first_hap) {
                    if (lane == 0) out[eo] = (uint8_t)',';
                    ++eo;
                }
                first_hap = false;
                const Hap h = gv.hap(g, s, a);
                const uint32_t n = hap_len[slot];
                for (uint32_t i = lane; i < n; i += 32) out[eo + i] = h.at(i, fa, t);
                eo += n;
                const uint32_t sb = slot_seds[slot];
                if (!sb) continue;
                warp_render_id_list(stage, slot_bits + slot * W, W, id_text, sout, so, sb);
                const u64 end = so + sb;
                so = end;
            }
        }
        if (lane == 0) out[eo] = (uint8_t)'}';
    }
}

__global__ void k_iota(uint32_t* p, u64 n) {
    for (u64 i = gtid(); i < n; i += gthreads()) p[i] = (uint32_t)i;
}

}  // namespace

// =============================================================================================================
struct VcfPipeline::Bufs {
    DevBuf d[30];
    ~Bufs() {
        for (DevBuf& b : d) b.release();
    }
};

VcfPipeline::VcfPipeline(eds_ctx* ctx) : ctx_(ctx), bufs_(new Bufs()) {}
VcfPipeline::~VcfPipeline() { delete bufs_; }

void VcfPipeline::transform_device(const uint8_t* vcf, uint64_t vcf_bytes, const uint8_t* fasta, uint64_t fasta_bytes,
                                   eds_buffer* eds_out, eds_buffer* seds_out, eds_vcf_stats* stats,
                                   std::vector<uint64_t>* sv_lines, VcfShardHook* hook) {
    if ((reinterpret_cast<uintptr_t>(vcf) | reinterpret_cast<uintptr_t>(fasta)) & 15u)
        throw std::invalid_argument("eds_vcf_transform_device: buffers must be 16-byte aligned");
    cudaStream_t s = ctx_->stream;
    KernelClock& clk = ctx_->clock;
    clk.reset();
    const uint32_t sms = (uint32_t)ctx_->sm_count;
    const unsigned P = std::max(1u, std::min(ctx_->partitions ? ctx_->partitions : sms * 4u, 4096u));
#ifdef EDSB_EMU
    const uint32_t G = 2u, B = kScanBlock;
#else
    const uint32_t G = sms * 8u, B = kScanBlock;
#endif
    Bufs& B_ = *bufs_;
    DevBuf &d_status = B_.d[0], &d_part = B_.d[1], &d_line_start = B_.d[2], &d_heads = B_.d[3], &d_recs = B_.d[4],
           &d_sv = B_.d[5], &d_pos = B_.d[6], &d_al_off = B_.d[7], &d_al_len = B_.d[8], &d_bits = B_.d[9],
           &d_rec_of = B_.d[10], &d_incl = B_.d[11], &d_opens = B_.d[12], &d_gid = B_.d[13], &d_slot_base = B_.d[14],
           &d_gfirst = B_.d[15], &d_gfrom = B_.d[16], &d_gto = B_.d[17], &d_canon = B_.d[18], &d_hap_len = B_.d[19],
           &d_slot_bits = B_.d[20], &d_slot_seds = B_.d[21], &d_kept = B_.d[22], &d_geds = B_.d[23], &d_gseds = B_.d[24],
           &d_eds_off = B_.d[25], &d_seds_off = B_.d[26], &d_nl = B_.d[27], &d_id_text = B_.d[28];
    DevBuf& d_out = ctx_->vcf_out[0];
    DevBuf& d_sout = ctx_->vcf_out[1];

#define VCF_LAUNCH(name, kernel, grid, block, smem, ...)          \
    do {                                                           \
        clk.begin(name);                                           \
        EDSB_LAUNCH(kernel, grid, block, smem, s, __VA_ARGS__);    \
        clk.end();                                                 \
    } while (0)
#define VCF_SCAN(name, Op, n, fn)                 \
    do {                                          \
        clk.begin(name);                          \
        device_scan<Op>(s, P, (n), (fn), part);   \
        clk.end();                                \
        ++clk.launches;                           \
    } while (0)

    d_status.reserve(sizeof(VcfStatus));
    d_part.reserve((size_t)(P + 1) * 8);
    VcfStatus* st = d_status.as<VcfStatus>();
    u64* part = d_part.as<u64>();
    VcfStatus hst;
    memset(&hst, 0, sizeof(hst));
    hst.bad_line = kNone64;
    EDSB_CUDA(cudaMemcpyAsync(st, &hst, sizeof(hst), cudaMemcpyHostToDevice, s));
    auto total_of = [&]() -> u64 {
        u64 v = 0;
        EDSB_CUDA(cudaMemcpyAsync(&v, part + P, 8, cudaMemcpyDeviceToHost, s));
        EDSB_CUDA(cudaStreamSynchronize(s));
        return v;
    };
    auto status_now = [&]() {
        EDSB_CUDA(cudaMemcpyAsync(&hst, st, sizeof(hst), cudaMemcpyDeviceToHost, s));
        EDSB_CUDA(cudaStreamSynchronize(s));
        EDSB_CUDA(cudaGetLastError());
    };

    // ---- FASTA census + VCF line index ------------------------------------------------------------------------
    VCF_LAUNCH("k_fa_head", k_fa_head, 1, 32, 0, fasta, (u64)fasta_bytes, st);
    VCF_LAUNCH("k_fa_next_record", k_fa_next_record, G, B, 0, fasta, (u64)fasta_bytes, st);
    VCF_LAUNCH("k_fa_measure", k_fa_measure, G, B, 0, fasta, st);

    const u64 n_chunks = (vcf_bytes + 15) / 16;
    const uint32_t nl_blocks = std::max<uint32_t>(1u, (uint32_t)std::min<u64>((u64)G / 2u, (n_chunks + 32u * (B / 32u) - 1) / (32u * (B / 32u))));
    const uint32_t nl_warps = nl_blocks * (B / 32u);
    d_nl.reserve((size_t)nl_warps * 8);
    VCF_LAUNCH("k_nl_count", k_nl_count, nl_blocks, B, 0, vcf, (u64)vcf_bytes, n_chunks, d_nl.as<u64>());
    std::vector<u64> h_nl(nl_warps);
    EDSB_CUDA(cudaMemcpyAsync(h_nl.data(), d_nl.p, (size_t)nl_warps * 8, cudaMemcpyDeviceToHost, s));
    EDSB_CUDA(cudaStreamSynchronize(s));
    u64 n_nl = 0;
    for (u64& v : h_nl) {
        const u64 mine = v;
        v = n_nl;
        n_nl += mine;
    }
    uint8_t last_byte = '\n';
    if (vcf_bytes) EDSB_CUDA(cudaMemcpyAsync(&last_byte, vcf + vcf_bytes - 1, 1, cudaMemcpyDeviceToHost, s));
    status_now();
    // FASTA errors come first, as in the reference (:683)
    if (hst.fa_err & kFaNoHeader) throw std::runtime_error("Invalid FASTA format: expected header line starting with '>'");
    if (hst.fa_err & kFaEmpty) throw std::runtime_error("FASTA file is empty");
    if (hst.fa_err & kFaZeroWidth) throw BadVcf("FASTA: the first sequence line is empty (the reference divides by a line width of 0)");
    if (hst.fa_err & kFaCarriage) throw BadVcf("FASTA: carriage returns are not supported (the reference counts them as bases)");
    if (hst.fa_bad_nl != kNone64 && hst.fa_bad_nl < hst.fa_last_base)
        throw BadVcf("FASTA: sequence lines must all have the width of the first one (byte " + std::to_string(hst.fa_bad_nl) +
                     "); the reference addresses bases as start + q + q / width");
    const u64 n_bases = hst.fa_rec_end - hst.fa_seq_start - hst.fa_nl;
    const Fasta fa{fasta, hst.fa_seq_start, hst.fa_lw};

    const bool open_tail = vcf_bytes && last_byte != '\n';
    const u64 n_lines = n_nl + (open_tail ? 1 : 0);
    if (n_lines >= 0xfffffff0ull) throw std::invalid_argument("eds_vcf_transform: more than 2^32 lines");
    d_line_start.reserve((size_t)(n_lines + 2) * 8);
    u64* line_start = d_line_start.as<u64>();
    {
        const u64 zero = 0, tail = vcf_bytes + 1;
        EDSB_CUDA(cudaMemcpyAsync(line_start, &zero, 8, cudaMemcpyHostToDevice, s));
        if (open_tail) EDSB_CUDA(cudaMemcpyAsync(line_start + n_lines, &tail, 8, cudaMemcpyHostToDevice, s));
        EDSB_CUDA(cudaMemcpyAsync(d_nl.p, h_nl.data(), (size_t)nl_warps * 8, cudaMemcpyHostToDevice, s));
        VCF_LAUNCH("k_nl_write", k_nl_write, nl_blocks, B, 0, vcf, (u64)vcf_bytes, n_chunks, d_nl.as<u64>(), line_start);
    }

    // ---- heads, records ----------------------------------------------------------------------------------------
    d_heads.reserve((size_t)(n_lines + 1) * sizeof(LineHead));
    d_sv.reserve((size_t)std::min<u64>(n_lines + 1, kSvListCap) * 8);
    LineHead* heads = d_heads.as<LineHead>();
    u64 n_rec = 0, n_rows = 0;
    if (n_lines) {
        VCF_LAUNCH("k_head", k_head, G, B, 0, vcf, line_start, n_lines, heads, d_sv.as<u64>(), st);
        d_recs.reserve((size_t)(n_lines + 1) * sizeof(Rec));
        VCF_SCAN("records", OpSum64, n_lines, (RecFn{heads, line_start, d_recs.as<Rec>()}));
        n_rec = total_of();
    }
    Rec* recs = d_recs.as<Rec>();
    if (n_rec) {
        VCF_SCAN("rows", OpSum64, n_rec, (RowFn{recs}));
        d_pos.reserve((size_t)n_rec * 8);
        VCF_LAUNCH("k_rec_check", k_rec_check, G, B, 0, recs, n_rec, d_pos.as<u64>(), st);
        VCF_LAUNCH("k_first_fields", k_first_fields, 1, 32, 0, vcf, n_chunks, recs, st);
        n_rows = total_of();
    }
    status_now();
    if (stats) {
        stats->total_variants = n_rec + hst.n_malformed + hst.n_sv;
        stats->processed_variants = n_rec;
        stats->skipped_malformed = hst.n_malformed;
        stats->skipped_unsupported_sv = hst.n_sv;
        stats->n_lines = n_lines;
        stats->n_alleles = n_rows;
    }
    if (sv_lines) {
        const u64 listed = std::min<u64>(hst.n_sv, kSvListCap);
        sv_lines->resize(listed);
        if (listed) {
            EDSB_CUDA(cudaMemcpyAsync(sv_lines->data(), d_sv.p, (size_t)listed * 8, cudaMemcpyDeviceToHost, s));
            EDSB_CUDA(cudaStreamSynchronize(s));
            std::sort(sv_lines->begin(), sv_lines->end());
        }
    }
    if (hst.err) {
        const std::string where = " (VCF line at byte " + std::to_string(hst.bad_line) + ")";
        if (hst.bad_line != kNone64 && (hst.err & kVcfPosZero))
            throw BadVcf("VCF: POS 0 is outside the reference's domain (it underflows to 2^64 - 1 there)" + where);
        throw BadVcf("VCF: a record extends past the end of the FASTA sequence (" + std::to_string(n_bases) + " bases)" + where);
    }
    if (n_rows >= 0xfffffff0ull) throw std::invalid_argument("eds_vcf_transform: more than 2^32 alleles");

    uint32_t n_groups = 0;
    u64 n_slots = 0, max_samples = 0;
    uint32_t W = 1;
    uint32_t host_sorted = 0, retries = 0;
    u64 ref_lo = 0, ref_hi = n_bases;
    if (hook && !n_rec) {
        std::vector<uint32_t> none;
        uint64_t lo = 0, hi = n_bases;
        hook->exchange(std::vector<uint64_t>(), 0, n_bases, &none, &lo, &hi);
        ref_lo = lo;
        ref_hi = hi;
    }
    if (n_rec) {
        // ---- sample columns (main stream) || position order (side stream + host) -------------------------------
        d_al_off.reserve((size_t)n_rows * 8);
        d_al_len.reserve((size_t)n_rows * 4);
        VCF_LAUNCH("k_alleles", k_alleles, G, B, 0, vcf, recs, n_rec, d_al_off.as<u64>(), d_al_len.as<uint32_t>());
        W = std::max<uint32_t>(1u, std::max<uint32_t>(words_hint_, (uint32_t)((hst.first_tokens + 31) / 32)));
        d_rec_of.reserve((size_t)n_rec * 4);
        uint32_t* rec_of = d_rec_of.as<uint32_t>();
        // k_gt sizes its bitsets from the first record's sample count; a wider record makes it run again
        auto run_gt = [&](const std::function<void()>& while_it_runs) {
            for (;;) {
                d_bits.reserve((size_t)n_rows * W * 4 + 16);
                // shared-memory slice per warp: room for REF + 3 ALT rows (wider records work in global memory)
                const uint32_t wpb = 8;
                uint32_t smem_words = std::max<uint32_t>(4u * W, 64u);
                while ((size_t)smem_words * wpb * 4 > 40u * 1024u && smem_words > 2u * W) smem_words /= 2;
                if ((size_t)smem_words * wpb * 4 > 40u * 1024u) smem_words = 8;  // everything in global memory
                const uint32_t grid = (uint32_t)std::min<u64>((n_rec + wpb - 1) / wpb, (u64)sms * 32u);
                VCF_LAUNCH("k_gt", k_gt, grid, wpb * 32, (size_t)smem_words * wpb * 4, vcf, n_chunks, recs, n_rec, W, smem_words,
                           d_bits.as<uint32_t>(), st);
                while_it_runs();
                status_now();
                if (hst.max_tokens <= (u64)W * 32u) break;
                W = (uint32_t)((hst.max_tokens + 31) / 32);
                ++retries;
            }
        };
        if (hook) {
            // sharded: the order comes from the one sort over every slice's records, made where the slices meet
            std::vector<uint64_t> h_pos(n_rec);
            EDSB_CUDA(cudaMemcpyAsync(h_pos.data(), d_pos.p, (size_t)n_rec * 8, cudaMemcpyDeviceToHost, s));
            EDSB_CUDA(cudaStreamSynchronize(s));
            std::vector<uint32_t> h_perm;
            bool met = false;
            run_gt([&]() {
                if (met) return;
                met = true;
                uint64_t lo = 0, hi = n_bases;
                hook->exchange(h_pos, hst.max_end, n_bases, &h_perm, &lo, &hi);
                ref_lo = lo;
                ref_hi = hi;
            });
            if (h_perm.empty()) {
                VCF_LAUNCH("k_iota", k_iota, G, B, 0, rec_of, n_rec);
            } else {
                host_sorted = 1;
                EDSB_CUDA(cudaMemcpyAsync(rec_of, h_perm.data(), (size_t)n_rec * 4, cudaMemcpyHostToDevice, s));
                EDSB_CUDA(cudaStreamSynchronize(s));
            }
        } else if (hst.unsorted) {
            // std::sort is unstable (:715-718): the order of records that share a position is whatever that very
            // call leaves, so the same call (same comparator, same sequence) is made on the host while k_gt runs
            host_sorted = 1;
            std::vector<u64> h_pos(n_rec);
            EDSB_CUDA(cudaMemcpyAsync(h_pos.data(), d_pos.p, (size_t)n_rec * 8, cudaMemcpyDeviceToHost, s));
            EDSB_CUDA(cudaStreamSynchronize(s));
            std::vector<std::pair<u64, uint32_t>> keyed(n_rec);
            for (u64 k = 0; k < n_rec; ++k) keyed[k] = {h_pos[k], (uint32_t)k};
            std::vector<uint32_t> h_perm;
            run_gt([&]() {
                if (!h_perm.empty()) return;
                std::sort(keyed.begin(), keyed.end(),
                          [](const std::pair<u64, uint32_t>& a, const std::pair<u64, uint32_t>& b) { return a.first < b.first; });
                h_perm.resize(n_rec);
                for (u64 k = 0; k < n_rec; ++k) h_perm[k] = keyed[k].second;
            });
            EDSB_CUDA(cudaMemcpyAsync(rec_of, h_perm.data(), (size_t)n_rec * 4, cudaMemcpyHostToDevice, s));
            EDSB_CUDA(cudaStreamSynchronize(s));  // h_perm goes out of scope
        } else {
            run_gt([]() {});
            VCF_LAUNCH("k_iota", k_iota, G, B, 0, rec_of, n_rec);
        }
        words_hint_ = W;
        max_samples = hst.max_tokens;

        // ---- groups ----------------------------------------------------------------------------------------------
        d_incl.reserve((size_t)n_rec * 8);
        d_opens.reserve((size_t)n_rec);
        d_gid.reserve((size_t)n_rec * 4);
        d_slot_base.reserve((size_t)n_rec * 8);
        d_gfirst.reserve((size_t)(n_rec + 2) * 4);
        VCF_SCAN("spans", OpMax64, n_rec, (SpanFn{recs, rec_of, d_incl.as<u64>(), d_opens.as<uint8_t>()}));
        VCF_SCAN("groups", OpSum64, n_rec,
                 (GroupFn{recs, rec_of, d_opens.as<uint8_t>(), d_gid.as<uint32_t>(), d_slot_base.as<u64>(), d_gfirst.as<uint32_t>()}));
        const u64 tot = total_of();
        n_groups = (uint32_t)tot;
        n_slots = (tot >> 32) + n_groups;
    }
    d_gfrom.reserve((size_t)(n_groups + 2) * 8);
    d_gto.reserve((size_t)(n_groups + 2) * 8);
    d_gfirst.reserve((size_t)(n_groups + 2) * 4);
    d_geds.reserve((size_t)(n_groups + 2) * 8);
    d_gseds.reserve((size_t)(n_groups + 2) * 8);
    d_eds_off.reserve((size_t)(n_groups + 2) * 8);
    d_seds_off.reserve((size_t)(n_groups + 2) * 8);
    uint32_t* rec_of = d_rec_of.as<uint32_t>();
    u64 *g_from = d_gfrom.as<u64>(), *g_to = d_gto.as<u64>();
    VCF_LAUNCH("k_group_info", k_group_info, G, B, 0, recs, rec_of, d_incl.as<u64>(), d_gfirst.as<uint32_t>(), n_groups, n_rec,
               ref_hi, g_from, g_to);
    const GroupView gv{recs, rec_of, d_slot_base.as<u64>(), d_al_off.as<u64>(), d_al_len.as<uint32_t>(), g_from, g_to};
    if (n_groups) {
        d_canon.reserve((size_t)n_slots * 8);
        d_hap_len.reserve((size_t)n_slots * 4);
        d_slot_bits.reserve((size_t)n_slots * W * 4 + 16);
        d_slot_seds.reserve((size_t)n_slots * 4);
        d_kept.reserve((size_t)n_slots);
        VCF_LAUNCH("k_haps", k_haps, G, B, 0, vcf, fa, gv, d_gfirst.as<uint32_t>(), n_groups, d_canon.as<u64>(), d_hap_len.as<uint32_t>());
        VCF_LAUNCH("k_combine", k_combine, G, B, 0, recs, rec_of, d_slot_base.as<u64>(), d_gfirst.as<uint32_t>(), n_groups,
                   d_bits.as<uint32_t>(), W, d_canon.as<u64>(), d_hap_len.as<uint32_t>(), d_slot_bits.as<uint32_t>(),
                   d_slot_seds.as<uint32_t>(), d_kept.as<uint8_t>(), d_geds.as<u64>(), d_gseds.as<u64>());
    }
    VCF_SCAN("eds_offsets", OpSum64, (u64)n_groups + 1, (OutFn{g_from, g_to, d_geds.as<u64>(), n_groups, 2u, 1u, d_eds_off.as<u64>(), ref_lo}));
    const u64 eds_total = total_of();
    VCF_SCAN("seds_offsets", OpSum64, (u64)n_groups + 1, (OutFn{g_from, g_to, d_gseds.as<u64>(), n_groups, 3u, 0u, d_seds_off.as<u64>(), ref_lo}));
    const u64 seds_total = total_of();

    // ---- emit ------------------------------------------------------------------------------------------------------
    d_out.reserve(eds_total + 16);
    d_sout.reserve(seds_total + 16);
    if (ref_hi > ref_lo)
        VCF_LAUNCH("k_emit_ref", k_emit_ref, G, B, 0, fa, ref_lo, ref_hi, g_from, g_to, n_groups, d_eds_off.as<u64>(), d_out.as<uint8_t>());
    {
        const uint32_t stage_bytes = id_list_stage_bytes((u64)W * 32u);
        const uint32_t wpb = std::max<uint32_t>(1u, std::min<uint32_t>(B / 32u, (40u * 1024u) / stage_bytes));
        const uint32_t grid = (uint32_t)std::min<u64>(((u64)n_groups + 1 + wpb - 1) / wpb, (u64)G * (B / 32u) / wpb);
        if (id_text_n_ < (u64)W * 32u + 1) {
            id_text_n_ = (u64)W * 32u + 1;
            d_id_text.reserve((size_t)id_text_n_ * 8);
            VCF_LAUNCH("k_id_text", k_id_text, G, B, 0, d_id_text.as<u64>(), id_text_n_);
        }
        VCF_LAUNCH("k_emit_groups", k_emit_groups, grid, wpb * 32, (size_t)stage_bytes * wpb, vcf, fa, gv, d_gfirst.as<uint32_t>(),
                   n_groups, d_canon.as<u64>(), d_hap_len.as<uint32_t>(), d_kept.as<uint8_t>(), d_slot_bits.as<uint32_t>(),
                   d_slot_seds.as<uint32_t>(), W, stage_bytes, d_id_text.as<u64>(), d_eds_off.as<u64>(), d_seds_off.as<u64>(),
                   d_out.as<uint8_t>(), d_sout.as<uint8_t>(), ref_lo);
    }
    EDSB_CUDA(cudaStreamSynchronize(s));
    EDSB_CUDA(cudaGetLastError());
    clk.resolve();
    eds_out->data = d_out.as<uint8_t>();
    eds_out->bytes = eds_total;
    seds_out->data = d_sout.as<uint8_t>();
    seds_out->bytes = seds_total;
    if (stats) {
        stats->variant_groups = n_groups;
        stats->n_haplotype_slots = n_slots;
        stats->n_samples_max = max_samples;
        stats->n_bases = n_bases;
        stats->eds_bytes = eds_total;
        stats->seds_bytes = seds_total;
        stats->gpu_launches = clk.launches;
        stats->host_sorted = host_sorted;
        stats->retries = retries;
    }
#undef VCF_LAUNCH
#undef VCF_SCAN
}

}  // namespace edsb
