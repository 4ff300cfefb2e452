// l-EDS merge on the GPU: eds_to_leds_linear / eds_to_leds_cartesian
// (draessld/EDSParser src/cpp/lib/transforms/eds_transforms.cpp:313-426) behind eds_leds_merge_host.
//
// Reference behaviour being reproduced byte for byte:
//   EDS::parse + normalize_eds_format   formats/eds.cpp:39-155, 831-881   -> k_part_* with StripFn/DepthFn/EventFn
//   EDS::parse_sources                  formats/eds.cpp:268-355           -> SetFn, RankFn, k_bits_fill, k_universal
//   is_leds / select_independent_merge_pairs   eds_transforms.cpp:439-468 / 46-107  -> k_cand + RunFn (max-scan)
//   merge_multiple_pairs / EDS::merge_adjacent eds_transforms.cpp:120-196, eds.cpp:1425-1695 -> k_kept, k_merge_write
//   reconstruct_eds                     eds_transforms.cpp:207-296        -> k_rebuild (no text round trip)
//   EDS::save / save_sources            formats/eds.cpp:600-659           -> k_expand, k_emit_eds, k_emit_seds
//
// Formulation (nothing is taken from the reference's code, which rebuilds the whole EDS object per pair):
//  * every original string is a LEAF of an alternative pool {left, right, len, source bitset}; a merge round
//    appends one pool entry per surviving (left alt, right alt) combination — a merge TREE, no string is
//    copied until the final emit;
//  * sources are dense bitsets (ids ranked by value, so ascending bit = ascending id). A set containing 0
//    is the reference's universal marker: it is stored as ALL ONES, which makes the {0}-aware intersection
//    of eds.cpp:1481-1500 a plain AND and "kept" a non-zero test;
//  * the greedy left-to-right pair selection is "1st, 3rd, 5th ... candidate of every run of consecutive
//    candidates": a max-scan for the run start plus a parity test;
//  * CARTESIAN is the same machinery with zero-width bitsets (everything is kept).
#include "leds.h"

#include <string.h>

#include <algorithm>
#include <functional>
#include <string>
#include <vector>

#include "scan.cuh"

namespace edsb {

namespace {

constexpr uint32_t kNone = 0xffffffffu;
constexpr uint32_t kMaxPathId = 1u << 26;  // presence bitmap of 8 MB
constexpr uint32_t kBigLeaf = 96;          // leaves longer than this are copied by a whole block

enum LedsErr : uint32_t {
    kErrEdsSyntax = 1,
    kErrSedsSyntax = 2,
    kErrSedsBigId = 4,
    kErrSedsOverflow = 8,
};

struct LedsStatus {
    uint32_t err;
    uint32_t empty_set;    // smallest set index with no ids (kNone if none)
    uint32_t empty_merge;  // smallest left position of a selected pair that keeps nothing (kNone if none)
    uint32_t n_big;        // entries of the big-copy list
    uint32_t max_id;       // largest path id seen
    uint32_t pad;
    unsigned long long eds_total, seds_total;
};

__device__ __forceinline__ bool is_space(uint8_t c) { return c == ' ' || (c >= 9 && c <= 13); }
__device__ __forceinline__ bool is_digit(uint8_t c) { return c >= (uint8_t)'0' && c <= (uint8_t)'9'; }

// ---- whitespace strip (eds.cpp:46, :272): stream compaction ---------------------------------------
struct StripFn {
    const uint8_t* in;
    uint8_t* out;
    __device__ unsigned long long value(unsigned long long i) const { return is_space(in[i]) ? 0ull : 1ull; }
    __device__ void apply(unsigned long long i, unsigned long long prefix, unsigned long long v) const {
        if (v) out[prefix] = in[i];
    }
};

// White-space census of a text, 16 bytes per load: how many white-space bytes and where the first one is. Files end with
// a newline and hold none inside, so the usual answer is "only at the end": the text is then used where it lies, minus
// its tail, and the compaction pass (StripFn: read + write of every byte) is skipped.
struct SpaceCensus {
    unsigned long long count, first;
};

__device__ __forceinline__ uint32_t space_bytes4(uint32_t w) {  // bit k = byte k is ' ' or 9..13
    // quick reject: no byte below 0x21 (the haszero trick on w - 0x21..)
    if (!((w - 0x21212121u) & ~w & 0x80808080u)) return 0u;
    uint32_t m = 0;
    for (uint32_t k = 0; k < 4u; ++k) m |= (is_space((uint8_t)(w >> (8u * k))) ? 1u : 0u) << k;
    return m;
}

__global__ void __launch_bounds__(256) k_space_census(const uint8_t* t, unsigned long long n, SpaceCensus* out) {
    unsigned long long cnt = 0, first = ~0ull;
    const unsigned long long n16 = n / 16u;
    const uint4* v = reinterpret_cast<const uint4*>(t);
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (unsigned long long)gridDim.x * blockDim.x) {
        const uint4 x = ldg_nc(v + i);
        const uint32_t m = space_bytes4(x.x) | (space_bytes4(x.y) << 4) | (space_bytes4(x.z) << 8) | (space_bytes4(x.w) << 12);
        if (m) {
            cnt += (unsigned long long)__popc(m);
            first = min(first, i * 16u + (unsigned long long)(__ffs((int)m) - 1));
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (unsigned long long i = n16 * 16u; i < n; ++i)
            if (is_space(t[i])) {
                ++cnt;
                first = min(first, i);
            }
    cnt = warp_sum(cnt);
    for (int d = 16; d > 0; d >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, d));
    if ((threadIdx.x & 31) == 0 && (cnt || first != ~0ull)) {
        atomicAdd(&out->count, cnt);
        atomicMin(&out->first, first);
    }
}

// ---- EDS text: brace depth -------------------------------------------------------------------------
struct DepthFn {
    const uint8_t* t;
    uint8_t* depth;  // depth BEFORE the character: 0 outside a set, 1 inside
    LedsStatus* st;
    __device__ unsigned long long value(unsigned long long i) const {
        return (t[i] == (uint8_t)'{' ? 1ull : 0ull) | (t[i] == (uint8_t)'}' ? 1ull << 32 : 0ull);
    }
    __device__ void apply(unsigned long long i, unsigned long long prefix, unsigned long long) const {
        const long long d = (long long)(prefix & 0xffffffffull) - (long long)(prefix >> 32);
        const uint8_t c = t[i];
        if ((c == (uint8_t)'{' && d != 0) || (c == (uint8_t)'}' && d != 1) || d < 0 || d > 1) atomicOr(&st->err, (uint32_t)kErrEdsSyntax);
        depth[i] = (uint8_t)(d == 1);
    }
};

// ---- EDS text: strings and symbols ------------------------------------------------------------------
// A string starts after '{' or ',' and at the first character of a run of bare text (compact format,
// normalize_eds_format); a symbol starts at '{' and at the start of a bare run.
struct EventFn {
    const uint8_t* t;
    const uint8_t* depth;
    uint32_t* str_start;
    uint32_t* str_end;
    uint32_t* sym_first;
    __device__ __forceinline__ void classify(unsigned long long i, bool& e0, bool& e1) const {
        const uint8_t c = t[i];
        e1 = c == (uint8_t)'{' || c == (uint8_t)',';
        const bool bare = c != (uint8_t)'{' && c != (uint8_t)'}' && depth[i] == 0;
        e0 = bare && (i == 0 || t[i - 1] == (uint8_t)'}');
    }
    __device__ unsigned long long value(unsigned long long i) const {
        bool e0, e1;
        classify(i, e0, e1);
        const bool symstart = t[i] == (uint8_t)'{' || e0;
        return (unsigned long long)(e0 + e1) | (symstart ? 1ull << 32 : 0ull);
    }
    __device__ void apply(unsigned long long i, unsigned long long prefix, unsigned long long) const {
        bool e0, e1;
        classify(i, e0, e1);
        uint32_t j = (uint32_t)prefix;
        const uint32_t si = (uint32_t)(prefix >> 32);
        const uint8_t c = t[i];
        if (c == (uint8_t)'{' || e0) sym_first[si] = j;
        if (e0) {
            str_start[j] = (uint32_t)i;
            if (j > 0) str_end[j - 1] = (uint32_t)i - 1u;  // the previous string closed with the '}' before us
            ++j;
        }
        if (e1) {
            str_start[j] = (uint32_t)i + 1u;
            if (j > 0) str_end[j - 1] = (c == (uint8_t)',') ? (uint32_t)i : (uint32_t)i - ((i > 0 && t[i - 1] == (uint8_t)'}') ? 1u : 0u);
        }
    }
};

// Depth and events in ONE scan (two packed sums): a = strings | symbols << 32, b = opens | closes << 32. The depth
// before a character is opens - closes of its prefix; nothing is written per byte.
struct EdsFn {
    const uint8_t* t;
    uint32_t* str_start;
    uint32_t* str_end;
    uint32_t* sym_first;
    LedsStatus* st;
    __device__ U64x2 value(unsigned long long i) const {
        const uint8_t c = t[i];
        const bool open = c == (uint8_t)'{', close = c == (uint8_t)'}';
        // a bare run starts right after a '}' or at the start of the text (depth is 0 there in any well-formed text;
        // a malformed one is reported by the depth check below before the events are used)
        const bool e0 = !open && !close && (i == 0 || t[i - 1] == (uint8_t)'}');
        const bool e1 = open || c == (uint8_t)',';
        return U64x2{(unsigned long long)(e0 + e1) | ((open || e0) ? 1ull << 32 : 0ull),
                     (open ? 1ull : 0ull) | (close ? 1ull << 32 : 0ull)};
    }
    __device__ void apply(unsigned long long i, U64x2 prefix, U64x2) const {
        const uint8_t c = t[i];
        const long long d = (long long)(prefix.b & 0xffffffffull) - (long long)(prefix.b >> 32);
        if ((c == (uint8_t)'{' && d != 0) || (c == (uint8_t)'}' && d != 1) || d < 0 || d > 1) atomicOr(&st->err, (uint32_t)kErrEdsSyntax);
        const bool open = c == (uint8_t)'{', close = c == (uint8_t)'}';
        const bool e0 = !open && !close && (i == 0 || t[i - 1] == (uint8_t)'}');
        const bool e1 = open || c == (uint8_t)',';
        uint32_t j = (uint32_t)prefix.a;
        const uint32_t si = (uint32_t)(prefix.a >> 32);
        if (open || e0) sym_first[si] = j;
        if (e0) {
            str_start[j] = (uint32_t)i;
            if (j > 0) str_end[j - 1] = (uint32_t)i - 1u;  // the previous string closed with the '}' before us
            ++j;
        }
        if (e1) {
            str_start[j] = (uint32_t)i + 1u;
            if (j > 0) str_end[j - 1] = (c == (uint8_t)',') ? (uint32_t)i : (uint32_t)i - ((i > 0 && t[i - 1] == (uint8_t)'}') ? 1u : 0u);
        }
    }
};

__global__ void k_close_strings(const uint8_t* t, uint32_t n, uint32_t* str_end, uint32_t n_str, uint32_t* sym_first, uint32_t n_sym) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        if (n_str) str_end[n_str - 1] = n - ((n > 0 && t[n - 1] == (uint8_t)'}') ? 1u : 0u);
        sym_first[n_sym] = n_str;
    }
}

// ---- SEDS text ---------------------------------------------------------------------------------------
// Local grammar checks are enough: '{' only at the start or after '}', '}' / ',' / digits only inside.
struct SetFn {
    const uint8_t* s;
    unsigned long long n;
    uint32_t* present;  // bitmap over id values
    LedsStatus* st;
    __device__ __forceinline__ bool numstart(unsigned long long i) const { return is_digit(s[i]) && (i == 0 || !is_digit(s[i - 1])); }
    __device__ unsigned long long value(unsigned long long i) const {
        return (s[i] == (uint8_t)'{' ? 1ull : 0ull) | (numstart(i) ? 1ull << 32 : 0ull);
    }
    // the number that starts at i; false when std::stoi would throw or the id is beyond the device path's range
    __device__ __forceinline__ bool parse(unsigned long long i, uint32_t& out, uint32_t& why) const {
        unsigned long long v = 0;
        uint32_t digits = 0;
        for (unsigned long long k = i; k < n && is_digit(s[k]); ++k) {
            if (digits < 12) v = v * 10ull + (unsigned long long)(s[k] - (uint8_t)'0');
            ++digits;
        }
        out = 0;
        if (digits > 10 || v > 2147483647ull) {
            why = kErrSedsOverflow;  // std::stoi -> std::out_of_range
            return false;
        }
        if (v >= kMaxPathId) {
            why = kErrSedsBigId;
            return false;
        }
        out = (uint32_t)v;
        return true;
    }
    __device__ void apply(unsigned long long i, unsigned long long prefix, unsigned long long) const {
        const uint8_t c = s[i];
        const uint8_t p = i ? s[i - 1] : (uint8_t)'}';
        bool ok;
        if (c == (uint8_t)'{')
            ok = p == (uint8_t)'}';
        else if (c == (uint8_t)'}' || c == (uint8_t)',' || is_digit(c))
            ok = i > 0 && (p == (uint8_t)'{' || p == (uint8_t)',' || is_digit(p));
        else
            ok = false;
        if (i + 1 == n && c != (uint8_t)'}') ok = false;
        if (!ok) atomicOr(&st->err, (uint32_t)kErrSedsSyntax);
        if (!numstart(i)) return;
        uint32_t v, why = 0;
        if (!parse(i, v, why)) atomicOr(&st->err, why);
        // few distinct ids, millions of mentions: look before the atomic (a stale read only costs a redundant atomic)
        const uint32_t bit = 1u << (v & 31u);
        if (!(*reinterpret_cast<volatile uint32_t*>(&present[v >> 5]) & bit)) atomicOr(&present[v >> 5], bit);
        if ((uint32_t)v > *reinterpret_cast<volatile uint32_t*>(&st->max_id)) atomicMax(&st->max_id, (uint32_t)v);
    }
};

// Second pass over the same text once the ids are ranked: every number sets its bit in the dense bitset of its set
// (the index of the set is the number of '{' before it, minus one). No per-number arrays: the text is the list.
struct BitsFn {
    SetFn f;
    const uint32_t* rank;
    uint32_t* bits;
    uint32_t Wd, n_sets;
    __device__ unsigned long long value(unsigned long long i) const { return f.value(i); }
    __device__ void apply(unsigned long long i, unsigned long long prefix, unsigned long long) const {
        if (!f.numstart(i)) return;
        const uint32_t set = (uint32_t)prefix - 1u;
        uint32_t v, why;
        if (set >= n_sets || !f.parse(i, v, why)) return;  // reported after the first pass
        const uint32_t dense = rank[v >> 5] + (uint32_t)__popc(f.present[v >> 5] & low_bits(v & 31u));
        uint32_t* w = &bits[(size_t)set * Wd + (dense >> 5)];
        const uint32_t bit = 1u << (dense & 31u);
        if (!(*reinterpret_cast<volatile uint32_t*>(w) & bit)) atomicOr(w, bit);
    }
};

struct RankFn {
    const uint32_t* present;
    uint32_t* rank;
    __device__ unsigned long long value(unsigned long long w) const { return (unsigned long long)__popc(present[w]); }
    __device__ void apply(unsigned long long w, unsigned long long prefix, unsigned long long) const { rank[w] = (uint32_t)prefix; }
};

__global__ void k_id_table(const uint32_t* present, const uint32_t* rank, uint32_t n_words, uint32_t* id_of) {
    for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += gridDim.x * blockDim.x) {
        uint32_t k = rank[w];
        for (uint32_t bits = present[w]; bits; bits &= bits - 1) id_of[k++] = w * 32u + (uint32_t)__ffs((int)bits) - 1u;
    }
}

// A set that contains id 0 is universal (eds.cpp:1481-1487): store it as all ones. Sets with no id at
// all are an error ("Empty path set at string k", eds.cpp:336-338).
__global__ void k_universal(uint32_t* bits, uint32_t Wd, uint32_t n_sets, uint32_t P, uint32_t has_zero, LedsStatus* st) {
    for (uint32_t set = blockIdx.x * blockDim.x + threadIdx.x; set < n_sets; set += gridDim.x * blockDim.x) {
        uint32_t* b = bits + (size_t)set * Wd;
        uint32_t any = 0;
        for (uint32_t w = 0; w < Wd; ++w) any |= b[w];
        if (!any) {
            atomicMin(&st->empty_set, set);
        } else if (has_zero && (b[0] & 1u)) {
            for (uint32_t w = 0; w < Wd; ++w) b[w] = (w + 1 == Wd) ? low_bits(P - 32u * w) : 0xffffffffu;
        }
    }
}

// ---- the symbol table of a round and the alternative pool -------------------------------------------
struct SymTab {
    uint32_t* begin;  // first pool entry of the symbol's alternatives
    uint32_t* count;  // number of alternatives
};

struct Pool {
    uint32_t* left;   // kNone for a leaf
    uint32_t* right;  // leaf: index of the original string
    uint32_t* len;
    uint32_t* bits;   // Wd words per entry
};

__global__ void k_init(const uint32_t* str_start, const uint32_t* str_end, const uint32_t* sym_first, uint32_t n_str,
                       uint32_t n_sym, Pool pool, SymTab tab) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (uint32_t j = tid; j < n_str; j += nth) {
        pool.left[j] = kNone;
        pool.right[j] = j;
        pool.len[j] = str_end[j] - str_start[j];
    }
    for (uint32_t i = tid; i < n_sym; i += nth) {
        tab.begin[i] = sym_first[i];
        tab.count[i] = sym_first[i + 1] - sym_first[i];
    }
}

// short interior non-degenerate symbol (eds_transforms.cpp:75-97, 448-458)
__device__ __forceinline__ bool short_solid(const SymTab& t, const Pool& pool, uint32_t i, uint32_t n, uint32_t l) {
    return t.count[i] == 1u && i > 0 && i + 1 < n && pool.len[t.begin[i]] < l;
}

__global__ void k_cand(SymTab t, Pool pool, uint32_t n, uint32_t l, uint8_t* cand) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        bool c = false;
        if (i + 1 < n)
            c = short_solid(t, pool, i, n, l) || short_solid(t, pool, i + 1, n, l) || (t.count[i] > 1u && t.count[i + 1] > 1u);
        cand[i] = c ? 1 : 0;
    }
}

// Max-scan of "index after the last non-candidate": the exclusive prefix at a candidate is the start of
// its run; the greedy scan of select_independent_merge_pairs keeps the 1st, 3rd, ... pair of each run.
struct RunFn {
    const uint8_t* cand;
    uint8_t* sel;
    __device__ unsigned long long value(unsigned long long i) const { return cand[i] ? 0ull : i + 1ull; }
    __device__ void apply(unsigned long long i, unsigned long long prefix, unsigned long long) const {
        sel[i] = (cand[i] && ((i - prefix) & 1ull) == 0ull) ? 1 : 0;
    }
};

struct PairListFn {
    const uint8_t* sel;
    uint32_t* pairs_before;
    uint32_t* pair_list;
    __device__ unsigned long long value(unsigned long long i) const { return sel[i]; }
    __device__ void apply(unsigned long long i, unsigned long long prefix, unsigned long long v) const {
        pairs_before[i] = (uint32_t)prefix;
        if (v) pair_list[prefix] = (uint32_t)i;
    }
};

__device__ __forceinline__ bool meets(const uint32_t* a, const uint32_t* b, uint32_t Wd) {
    for (uint32_t w = 0; w < Wd; ++w)
        if (a[w] & b[w]) return true;
    return Wd == 0;  // no sources: CARTESIAN keeps everything
}

// kept[q] = alternatives the merged symbol of pair q will have. Pairs with few combinations (a conserved symbol
// against a variant site: the bulk of every round) take one thread each; the rest a warp, lanes over combinations.
constexpr unsigned long long kSmallPair = 8;

__global__ void k_kept(SymTab t, Pool pool, const uint32_t* pair_list, uint32_t n_pairs, uint32_t Wd,
                       unsigned long long* kept, LedsStatus* st) {
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n_pairs; q += gridDim.x * blockDim.x) {
        const uint32_t i = pair_list[q];
        const uint32_t ab = t.begin[i], na = t.count[i], bb = t.begin[i + 1], nb = t.count[i + 1];
        const unsigned long long combos = (unsigned long long)na * nb;
        if (combos > kSmallPair) continue;
        unsigned long long cnt = 0;
        if (Wd == 0) {
            cnt = combos;
        } else {
            for (uint32_t a = 0; a < na; ++a)
                for (uint32_t b = 0; b < nb; ++b)
                    cnt += meets(pool.bits + (size_t)(ab + a) * Wd, pool.bits + (size_t)(bb + b) * Wd, Wd) ? 1u : 0u;
        }
        kept[q] = cnt;
        if (cnt == 0) atomicMin(&st->empty_merge, i);
    }
    const uint32_t lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (uint32_t q = blockIdx.x * wpb + (threadIdx.x >> 5); q < n_pairs; q += gridDim.x * wpb) {
        const uint32_t i = pair_list[q];
        const uint32_t ab = t.begin[i], na = t.count[i], bb = t.begin[i + 1], nb = t.count[i + 1];
        const unsigned long long combos = (unsigned long long)na * nb;
        if (combos <= kSmallPair) continue;
        unsigned long long cnt = 0;
        if (Wd == 0) {
            cnt = combos;
        } else {
            for (unsigned long long c = lane; c < combos; c += 32) {
                const uint32_t a = (uint32_t)(c / nb), bq = (uint32_t)(c % nb);
                cnt += meets(pool.bits + (size_t)(ab + a) * Wd, pool.bits + (size_t)(bb + bq) * Wd, Wd) ? 1u : 0u;
            }
            cnt = warp_sum(cnt);
        }
        if (lane == 0) {
            kept[q] = cnt;
            if (cnt == 0) atomicMin(&st->empty_merge, i);
        }
    }
}

struct KeptFn {
    const unsigned long long* kept;
    unsigned long long* off;
    __device__ unsigned long long value(unsigned long long q) const { return kept[q]; }
    __device__ void apply(unsigned long long q, unsigned long long prefix, unsigned long long) const { off[q] = prefix; }
};

// New pool entries of every selected pair, i-major / j-minor (eds.cpp:1459-1468, 1640-1644).
__device__ __forceinline__ void write_entry(const Pool& pool, size_t e, uint32_t x, uint32_t y, uint32_t Wd) {
    pool.left[e] = x;
    pool.right[e] = y;
    pool.len[e] = pool.len[x] + pool.len[y];
    for (uint32_t w = 0; w < Wd; ++w) pool.bits[e * Wd + w] = pool.bits[(size_t)x * Wd + w] & pool.bits[(size_t)y * Wd + w];
}

__global__ void k_merge_write(SymTab t, Pool pool, const uint32_t* pair_list, uint32_t n_pairs, uint32_t Wd,
                              const unsigned long long* off, uint32_t pool_n) {
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n_pairs; q += gridDim.x * blockDim.x) {  // thread per small pair
        const uint32_t i = pair_list[q];
        const uint32_t ab = t.begin[i], na = t.count[i], bb = t.begin[i + 1], nb = t.count[i + 1];
        if ((unsigned long long)na * nb > kSmallPair) continue;
        size_t e = (size_t)(pool_n + off[q]);
        for (uint32_t a = 0; a < na; ++a)
            for (uint32_t b = 0; b < nb; ++b)
                if (meets(pool.bits + (size_t)(ab + a) * Wd, pool.bits + (size_t)(bb + b) * Wd, Wd)) write_entry(pool, e++, ab + a, bb + b, Wd);
    }
    const uint32_t lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (uint32_t q = blockIdx.x * wpb + (threadIdx.x >> 5); q < n_pairs; q += gridDim.x * wpb) {  // warp per large pair
        const uint32_t i = pair_list[q];
        const uint32_t ab = t.begin[i], na = t.count[i], bb = t.begin[i + 1], nb = t.count[i + 1];
        const unsigned long long combos = (unsigned long long)na * nb;
        if (combos <= kSmallPair) continue;
        unsigned long long base = pool_n + off[q];
        for (unsigned long long c0 = 0; c0 < combos; c0 += 32) {
            const unsigned long long c = c0 + lane;
            bool keep = false;
            uint32_t a = 0, bq = 0;
            if (c < combos) {
                a = (uint32_t)(c / nb);
                bq = (uint32_t)(c % nb);
                keep = meets(pool.bits + (size_t)(ab + a) * Wd, pool.bits + (size_t)(bb + bq) * Wd, Wd);
            }
            const uint32_t m = __ballot_sync(0xffffffffu, keep);
            if (keep) write_entry(pool, (size_t)(base + (unsigned long long)__popc(m & lanemask_lt())), ab + a, bb + bq, Wd);
            base += (unsigned long long)__popc(m);
        }
    }
}

// reconstruct_eds without the text round trip: the next round's symbol table
__global__ void k_rebuild(SymTab cur, SymTab next, const uint8_t* sel, const uint32_t* pairs_before, uint32_t n,
                          const unsigned long long* kept, const unsigned long long* off, uint32_t pool_n) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (i > 0 && sel[i - 1]) continue;  // right half of a merged pair
        const uint32_t q = pairs_before[i], to = i - q;
        if (sel[i]) {
            next.begin[to] = pool_n + (uint32_t)off[q];
            next.count[to] = (uint32_t)kept[q];
        } else {
            next.begin[to] = cur.begin[i];
            next.count[to] = cur.count[i];
        }
    }
}


// ---- EDS::Statistics (eds.cpp:361-505) as one reduction over the symbols, and the sources as id lists ------------
struct EdsStatsDev {
    unsigned long long total_chars, num_degenerate, num_common_chars, total_change_size, num_empty_strings, sum_ctx,
        num_ctx_blocks, max_paths_per_string, total_paths;
    uint32_t min_ctx, max_ctx;
};

__global__ void k_eds_stats(SymTab t, Pool pool, uint32_t n_sym, EdsStatsDev* out) {
    unsigned long long chars = 0, ndeg = 0, common = 0, change = 0, empty = 0, sctx = 0, nctx = 0;
    uint32_t mn = 0xffffffffu, mx = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_sym; i += gridDim.x * blockDim.x) {
        const uint32_t b0 = t.begin[i], cnt = t.count[i];
        if (cnt > 1u) {
            ++ndeg;
            change += cnt - 1u;
        } else {
            const uint32_t len = pool.len[b0];
            mn = min(mn, len);
            mx = max(mx, len);
            sctx += len;
            ++nctx;
            common += len;
        }
        for (uint32_t a = 0; a < cnt; ++a) {
            const uint32_t len = pool.len[b0 + a];
            chars += len;
            empty += len == 0u;
        }
    }
    // few partial results per thread: one atomic each (a statistics call, not the hot path)
    atomicAdd(&out->total_chars, chars);
    atomicAdd(&out->num_degenerate, ndeg);
    atomicAdd(&out->num_common_chars, common);
    atomicAdd(&out->total_change_size, change);
    atomicAdd(&out->num_empty_strings, empty);
    atomicAdd(&out->sum_ctx, sctx);
    atomicAdd(&out->num_ctx_blocks, nctx);
    atomicMin(&out->min_ctx, mn);
    atomicMax(&out->max_ctx, mx);
}

struct SrcCountFn {
    const uint32_t* bits;
    uint32_t Wd;
    unsigned long long* off;
    EdsStatsDev* st;
    __device__ unsigned long long value(unsigned long long j) const {
        unsigned long long c = 0;
        for (uint32_t w = 0; w < Wd; ++w) c += (unsigned long long)__popc(bits[j * Wd + w]);
        return c;
    }
    __device__ void apply(unsigned long long j, unsigned long long prefix, unsigned long long v) const {
        off[j] = prefix;
        atomicMax(&st->max_paths_per_string, v);
    }
};

__global__ void k_src_ids(const uint32_t* bits, uint32_t Wd, uint32_t n_str, const unsigned long long* off, const uint32_t* id_of,
                          int32_t* ids) {
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n_str; j += gridDim.x * blockDim.x) {
        unsigned long long at = off[j];
        for (uint32_t w = 0; w < Wd; ++w)
            for (uint32_t b = bits[(size_t)j * Wd + w]; b; b &= b - 1) ids[at++] = (int32_t)id_of[w * 32u + (uint32_t)__ffs((int)b) - 1u];
    }
}

__global__ void k_select_one(uint8_t* sel, uint32_t n, uint32_t pos) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) sel[i] = i == pos ? 1 : 0;
}

// ---- emit ---------------------------------------------------------------------------------------------
constexpr uint32_t kFirst = 1u, kLast = 2u, kBraced = 4u;

struct AltCountFn {
    SymTab t;
    unsigned long long* falt_off;
    __device__ unsigned long long value(unsigned long long i) const { return t.count[i]; }
    __device__ void apply(unsigned long long i, unsigned long long prefix, unsigned long long) const { falt_off[i] = prefix; }
};

// final alternative f -> (pool entry, position flags); warp per symbol, lanes over its alternatives
__global__ void k_expand(SymTab t, uint32_t n, const unsigned long long* falt_off, uint32_t compact, uint32_t* falt_pool,
                         uint8_t* falt_flags) {
    const uint32_t lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (uint32_t i = blockIdx.x * wpb + (threadIdx.x >> 5); i < n; i += gridDim.x * wpb) {
        const uint32_t cnt = t.count[i], b0 = t.begin[i];
        const unsigned long long f0 = falt_off[i];
        const uint32_t braced = (!compact || cnt > 1u) ? kBraced : 0u;
        for (uint32_t a = lane; a < cnt; a += 32) {
            falt_pool[f0 + a] = b0 + a;
            falt_flags[f0 + a] = (uint8_t)((a == 0 ? kFirst : 0u) | (a + 1 == cnt ? kLast : 0u) | braced);
        }
    }
}

__device__ __forceinline__ uint32_t eds_cost(uint32_t len, uint32_t flags) {
    // '{' with the first alternative, then the string, then ',' or (last) '}' when braced
    return len + ((flags & kFirst) && (flags & kBraced) ? 1u : 0u) + ((flags & kLast) ? ((flags & kBraced) ? 1u : 0u) : 1u);
}

struct EdsOffFn {
    const uint32_t* falt_pool;
    const uint8_t* falt_flags;
    const uint32_t* len;
    unsigned long long* eds_off;
    __device__ unsigned long long value(unsigned long long f) const { return eds_cost(len[falt_pool[f]], falt_flags[f]); }
    __device__ void apply(unsigned long long f, unsigned long long prefix, unsigned long long) const { eds_off[f] = prefix; }
};

__device__ __forceinline__ bool is_universal(const uint32_t* b, uint32_t has_zero) { return has_zero && (b[0] & 1u); }

// An alternative that was never merged prints the set it was given, verbatim (sorted, duplicates gone):
// "{0,7}" stays "{0,7}". A merged one prints {0} when universal (eds.cpp:1481-1487), else its intersection.
__device__ __forceinline__ const uint32_t* printable_bits(const Pool& pool, const uint32_t* raw_bits, uint32_t n_leaf,
                                                          uint32_t e, uint32_t Wd, uint32_t has_zero, bool& universal) {
    if (e < n_leaf) {
        universal = false;
        return raw_bits + (size_t)e * Wd;
    }
    const uint32_t* b = pool.bits + (size_t)e * Wd;
    universal = is_universal(b, has_zero);
    return b;
}

__global__ void k_seds_sizes(const uint32_t* falt_pool, unsigned long long n_falt, Pool pool, const uint32_t* raw_bits,
                             uint32_t n_leaf, uint32_t Wd, uint32_t has_zero, const uint32_t* id_of, uint32_t* seds_sz) {
    for (unsigned long long f = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; f < n_falt; f += (unsigned long long)gridDim.x * blockDim.x) {
        bool universal;
        const uint32_t* b = printable_bits(pool, raw_bits, n_leaf, falt_pool[f], Wd, has_zero, universal);
        uint32_t sz = 3;  // "{0}"
        if (!universal) {
            sz = 1;  // '{', then per id: digits + (',' or '}')
            for (uint32_t w = 0; w < Wd; ++w)
                for (uint32_t bits = b[w]; bits; bits &= bits - 1) sz += decimal_width(id_of[w * 32u + (uint32_t)__ffs((int)bits) - 1u]) + 1u;
        }
        seds_sz[f] = sz;
    }
}

struct SedsOffFn {
    const uint32_t* seds_sz;
    unsigned long long* seds_off;
    __device__ unsigned long long value(unsigned long long f) const { return seds_sz[f]; }
    __device__ void apply(unsigned long long f, unsigned long long prefix, unsigned long long) const { seds_off[f] = prefix; }
};

struct BigCopy {
    unsigned long long dst;
    uint32_t src, len;
};

// One thread per final alternative: iterative in-order walk of its merge tree, leaves copied from the
// stripped input text. Long leaves are deferred to k_big_copy.
__global__ void k_emit_eds(const uint32_t* falt_pool, const uint8_t* falt_flags, unsigned long long n_falt, Pool pool,
                           const uint32_t* str_start, const uint8_t* text, const unsigned long long* eds_off, uint8_t* out,
                           uint32_t* stack_ws, uint32_t stack_depth, BigCopy* big, uint32_t big_cap, LedsStatus* st) {
    const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t* stack = stack_ws + tid * stack_depth;
    for (unsigned long long f = tid; f < n_falt; f += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t flags = falt_flags[f];
        unsigned long long at = eds_off[f];
        if ((flags & kFirst) && (flags & kBraced)) out[at++] = '{';
        uint32_t sp = 0;
        stack[sp++] = falt_pool[f];
        while (sp) {
            const uint32_t e = stack[--sp];
            const uint32_t lft = pool.left[e];
            if (lft == kNone) {
                const uint32_t src = str_start[pool.right[e]], n = pool.len[e];
                if (n > kBigLeaf) {
                    const uint32_t slot = atomicAdd(&st->n_big, 1u);
                    if (slot < big_cap) big[slot] = BigCopy{at, src, n};
                } else {
                    for (uint32_t k = 0; k < n; ++k) out[at + k] = text[src + k];
                }
                at += n;
            } else {
                stack[sp++] = pool.right[e];  // right is emitted after left
                stack[sp++] = lft;
            }
        }
        if (flags & kLast) {
            if (flags & kBraced) out[at] = '}';
        } else {
            out[at] = ',';
        }
    }
}

__global__ void k_big_copy(const BigCopy* big, uint32_t n_big, const uint8_t* text, uint8_t* out) {
    for (uint32_t q = blockIdx.x; q < n_big; q += gridDim.x) {
        const BigCopy c = big[q];
        for (uint32_t k = threadIdx.x; k < c.len; k += blockDim.x) out[c.dst + k] = text[c.src + k];
    }
}

__global__ void k_emit_seds(const uint32_t* falt_pool, unsigned long long n_falt, Pool pool, const uint32_t* raw_bits,
                            uint32_t n_leaf, uint32_t Wd, uint32_t has_zero, const uint32_t* id_of,
                            const unsigned long long* seds_off, uint8_t* out) {
    for (unsigned long long f = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; f < n_falt; f += (unsigned long long)gridDim.x * blockDim.x) {
        bool universal;
        const uint32_t* b = printable_bits(pool, raw_bits, n_leaf, falt_pool[f], Wd, has_zero, universal);
        uint8_t* dst = out + seds_off[f];
        if (universal) {
            dst[0] = '{';
            dst[1] = '0';
            dst[2] = '}';
            continue;
        }
        *dst++ = '{';
        bool first = true;
        for (uint32_t w = 0; w < Wd; ++w)
            for (uint32_t bits = b[w]; bits; bits &= bits - 1) {
                const uint32_t id = id_of[w * 32u + (uint32_t)__ffs((int)bits) - 1u], wd = decimal_width(id);
                if (!first) *dst++ = ',';
                first = false;
                write_decimal(dst, id, wd);
                dst += wd;
            }
        *dst = '}';
    }
}

__global__ void k_newline(uint8_t* a, unsigned long long at_a, uint8_t* b, unsigned long long at_b) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        a[at_a] = '\n';
        if (b) b[at_b] = '\n';
    }
}

// ---- host-side error texts (error path only; formats/eds.cpp:80-82, 123-131, 278-349) ---------------
std::string strip_host(const uint8_t* p, uint64_t n) {
    std::string out;
    out.reserve(n);
    for (uint64_t i = 0; i < n; ++i) {
        const uint8_t c = p[i];
        if (!(c == ' ' || (c >= 9 && c <= 13))) out.push_back((char)c);
    }
    return out;
}

// what EDS::parse reports for a text the device found malformed
std::string explain_eds_error(const std::string& t) {
    std::string norm, bare;
    int depth = 0;
    for (char c : t) {
        if (c == '{') {
            if (!bare.empty() && depth == 0) {
                norm += '{' + bare + '}';
                bare.clear();
            }
            norm.push_back(c);
            ++depth;
        } else if (c == '}') {
            norm.push_back(c);
            --depth;
        } else if (depth > 0) {
            norm.push_back(c);
        } else {
            bare.push_back(c);
        }
    }
    if (!bare.empty() && depth == 0) norm += '{' + bare + '}';
    size_t p = 0;
    while (p < norm.size()) {
        if (norm[p] != '{') return "Expected '{' at position " + std::to_string(p);
        ++p;
        while (p < norm.size() && norm[p] != '}') ++p;
        if (p >= norm.size()) return "Expected '}' at position " + std::to_string(p);
        ++p;
    }
    return "malformed EDS text (nested braces are not supported)";
}

std::string explain_seds_error(const std::string& t, uint64_t n_strings, bool& out_of_range) {
    out_of_range = false;
    size_t p = 0;
    uint64_t sets = 0;
    while (p < t.size()) {
        if (t[p] != '{') return "sEDS: Expected '{' at position " + std::to_string(p);
        ++p;
        bool any = false;
        std::string num;
        auto flush = [&]() -> bool {
            if (num.empty()) return true;
            any = true;
            if (num.size() > 10 || std::stoull(num) > 2147483647ull) return false;
            num.clear();
            return true;
        };
        while (p < t.size() && t[p] != '}') {
            if (t[p] == ',') {
                if (!flush()) {
                    out_of_range = true;
                    return "stoi";
                }
            } else if (t[p] >= '0' && t[p] <= '9') {
                num.push_back(t[p]);
            } else {
                return "sEDS: Invalid character '" + std::string(1, t[p]) + "' at position " + std::to_string(p);
            }
            ++p;
        }
        if (!flush()) {
            out_of_range = true;
            return "stoi";
        }
        if (p >= t.size()) return "sEDS: Expected '}' at position " + std::to_string(p);
        ++p;
        if (!any) return "sEDS: Empty path set at string " + std::to_string(sets);
        ++sets;
    }
    if (sets != n_strings)
        return "sEDS: Source count (" + std::to_string(sets) + ") does not match EDS cardinality (" + std::to_string(n_strings) + ")";
    return "sEDS: path id too large for the device path (ids must be below 2^26)";
}

// grow-only device array that keeps its contents
template <typename T>
struct Keep {
    T* p = nullptr;
    size_t cap = 0;
    void ensure(size_t n, size_t live, cudaStream_t s) {
        if (n <= cap) return;
        const size_t want = n + n / 2 + 64;
        T* q = nullptr;
        cudaError_t e = cudaMalloc(&q, want * sizeof(T));
        if (e != cudaSuccess) {
            cudaGetLastError();
            throw BudgetError("the merged EDS does not fit in device memory (the CARTESIAN expansion has no bound in the reference either)");
        }
        if (p && live) EDSB_CUDA(cudaMemcpyAsync(q, p, live * sizeof(T), cudaMemcpyDeviceToDevice, s));
        if (p) {
            EDSB_CUDA(cudaStreamSynchronize(s));
            cudaFree(p);
        }
        p = q;
        cap = want;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

}  // namespace

struct LedsPipeline::Bufs {
    DevBuf d[40];
    Keep<uint32_t> k[4];
    ~Bufs() {
        for (DevBuf& b : d) b.release();
        for (auto& x : k) x.release();
    }
};

LedsPipeline::LedsPipeline(eds_ctx* ctx) : ctx_(ctx), bufs_(new Bufs()) {}
LedsPipeline::~LedsPipeline() { delete bufs_; }

void LedsPipeline::merge_host(const uint8_t* eds_in, uint64_t eds_bytes, const uint8_t* seds_in, uint64_t seds_bytes,
                              uint32_t l, bool compact, uint64_t max_output_bytes, eds_buffer* leds_out,
                              eds_buffer* seds_out, uint32_t* rounds_out, int* check_only, bool input_on_device,
                              const std::function<uint8_t*(int, uint64_t)>& sink, eds_parsed* parse_only,
                              const uint64_t* single_pair, uint32_t* edge_unmerged) {
    if (l == 0 && !parse_only && !single_pair)
        throw std::invalid_argument("context_length must be > 0 for l-EDS transformation");  // eds_transforms.cpp:322-324
    // string offsets into the EDS text are 32-bit (Length is uint32 in the reference too); the SEDS text has no such bound
    // (config 5: 13 GB of source sets for 0.1 GB of EDS)
    if (eds_bytes >= 0xfffffff0ull) throw std::invalid_argument("eds_leds_merge_host: the EDS text must be below 4 GiB (Length is uint32 in the reference too)");
    const bool linear = seds_in != nullptr;
    cudaStream_t s = ctx_->stream;
    const cudaMemcpyKind in_kind = input_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    // the error explanations re-read the input on the host (rare path)
    std::vector<uint8_t> h_in_copy;
    auto host_view = [&](const uint8_t* p, uint64_t n) -> const uint8_t* {
        if (!input_on_device) return p;
        h_in_copy.resize(n ? n : 1);
        if (n) EDSB_CUDA(cudaMemcpy(h_in_copy.data(), p, n, cudaMemcpyDeviceToHost));
        return h_in_copy.data();
    };
    KernelClock& clk = ctx_->clock;
    clk.reset();
    const uint32_t sms = (uint32_t)ctx_->sm_count;
    const unsigned P = std::max(1u, std::min(ctx_->partitions ? ctx_->partitions : sms * 8u, 4096u));
#ifdef EDSB_EMU
    const uint32_t G = 2u, B = kScanBlock;  // few OS threads per emulated launch
#else
    const uint32_t G = sms * 8u, B = kScanBlock;  // grid-stride launches
#endif

    // grow-only device buffers owned by the pipeline: a second call of similar size allocates nothing
    Bufs& B_ = *bufs_;
    DevBuf &d_raw = B_.d[0], &d_text = B_.d[1], &d_part = B_.d[3], &d_status = B_.d[4], &d_sraw = B_.d[5],
           &d_stext = B_.d[6], &d_str_start = B_.d[7], &d_str_end = B_.d[8], &d_sym_first = B_.d[9], &d_present = B_.d[12], &d_rank = B_.d[13], &d_idof = B_.d[14], &d_rawbits = B_.d[15],
           &d_cand = B_.d[16], &d_sel = B_.d[17], &d_pairs_before = B_.d[18], &d_pair_list = B_.d[19], &d_kept = B_.d[20],
           &d_off = B_.d[21], &d_falt_off = B_.d[22], &d_falt_pool = B_.d[23], &d_falt_flags = B_.d[24], &d_eds_off = B_.d[25],
           &d_seds_sz = B_.d[26], &d_seds_off = B_.d[27], &d_out = B_.d[28], &d_sout = B_.d[29], &d_stack = B_.d[30],
           &d_big = B_.d[31];
    DevBuf* d_tab = &B_.d[32];  // 4 entries
    Keep<uint32_t>&p_left = B_.k[0], &p_right = B_.k[1], &p_len = B_.k[2], &p_bits = B_.k[3];

    d_part.reserve((size_t)(P + 1) * 16);  // (the fused EDS scan carries two sums per partition)
    unsigned long long* part = d_part.as<unsigned long long>();
    d_status.reserve(sizeof(LedsStatus));
    LedsStatus* st = d_status.as<LedsStatus>();
    LedsStatus hst;
    memset(&hst, 0, sizeof(hst));
    hst.empty_set = hst.empty_merge = kNone;
    EDSB_CUDA(cudaMemcpyAsync(st, &hst, sizeof(hst), cudaMemcpyHostToDevice, s));

    auto total_of = [&]() -> unsigned long long {
        unsigned long long v = 0;
        EDSB_CUDA(cudaMemcpyAsync(&v, part + P, 8, cudaMemcpyDeviceToHost, s));
        EDSB_CUDA(cudaStreamSynchronize(s));
        return v;
    };
    auto status_now = [&]() {
        EDSB_CUDA(cudaMemcpyAsync(&hst, st, sizeof(hst), cudaMemcpyDeviceToHost, s));
        EDSB_CUDA(cudaStreamSynchronize(s));
        EDSB_CUDA(cudaGetLastError());
    };
#define LEDS_SCAN(name, Op, n, fn)                    \
    do {                                              \
        clk.begin(name);                              \
        device_scan<Op>(s, P, (n), (fn), part);       \
        clk.end();                                    \
        ++clk.launches;                               \
    } while (0)
#define LEDS_LAUNCH(name, kernel, grid, block, ...)        \
    do {                                                   \
        clk.begin(name);                                   \
        EDSB_LAUNCH(kernel, grid, block, 0, s, __VA_ARGS__); \
        clk.end();                                         \
    } while (0)

    // ---- EDS text: strip, parse ----------------------------------------------------------------------
    // input already in HBM (the VCF front end's output): read it where it is
    const uint8_t* raw = eds_in;
    if (!input_on_device) {
        d_raw.reserve(eds_bytes + 16);
        if (eds_bytes) EDSB_CUDA(cudaMemcpyAsync(d_raw.p, eds_in, eds_bytes, in_kind, s));
        raw = d_raw.as<uint8_t>();
    }
    // white space only at the end (or none): the text is used where it lies; else it is compacted (eds.cpp:46)
    DevBuf& d_census = B_.d[31];  // (the emit's big-copy list: not in use yet)
    d_census.reserve(2 * sizeof(SpaceCensus) + sizeof(BigCopy));
    auto census = [&](const char* name, const uint8_t* t, uint64_t bytes, SpaceCensus& h) {
        SpaceCensus* d = d_census.as<SpaceCensus>();
        h.count = 0;
        h.first = ~0ull;
        EDSB_CUDA(cudaMemcpyAsync(d, &h, sizeof(h), cudaMemcpyHostToDevice, s));
        if (bytes) LEDS_LAUNCH(name, k_space_census, G, 256, t, (unsigned long long)bytes, d);
        EDSB_CUDA(cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, s));
        EDSB_CUDA(cudaStreamSynchronize(s));
    };
    SpaceCensus ce;
    census("census_eds", raw, eds_bytes, ce);
    uint8_t* text;
    uint32_t n;
    if (ce.count == 0 || ce.first + ce.count == eds_bytes) {
        text = const_cast<uint8_t*>(raw);  // (read only from here on)
        n = (uint32_t)(ce.count ? ce.first : eds_bytes);
    } else {
        d_text.reserve(eds_bytes + 16);
        text = d_text.as<uint8_t>();
        LEDS_SCAN("strip_eds", OpSum64, eds_bytes, (StripFn{raw, text}));
        n = (uint32_t)total_of();
    }

    uint32_t n_str = 0, n_sym = 0;
    if (n) {
        // depth (validation) and events (strings, symbols) in one scan over the text: two packed sums
        d_str_start.reserve((size_t)(n + 2) * 4);
        d_str_end.reserve((size_t)(n + 2) * 4);
        d_sym_first.reserve((size_t)(n + 2) * 4);
        U64x2* part2 = reinterpret_cast<U64x2*>(part);
        clk.begin("eds_scan");
        device_scan<OpSum64x2>(s, P, n, EdsFn{text, d_str_start.as<uint32_t>(), d_str_end.as<uint32_t>(), d_sym_first.as<uint32_t>(), st}, part2);
        clk.end();
        ++clk.launches;
        U64x2 tot2{0, 0};
        EDSB_CUDA(cudaMemcpyAsync(&tot2, part2 + P, sizeof(tot2), cudaMemcpyDeviceToHost, s));
        status_now();
        if ((hst.err & kErrEdsSyntax) || (uint32_t)tot2.b != (uint32_t)(tot2.b >> 32))
            throw std::runtime_error(explain_eds_error(strip_host(host_view(eds_in, eds_bytes), eds_bytes)));
        n_str = (uint32_t)tot2.a;
        n_sym = (uint32_t)(tot2.a >> 32);
        LEDS_LAUNCH("k_close_strings", k_close_strings, 1, 32, text, n, d_str_end.as<uint32_t>(), n_str, d_sym_first.as<uint32_t>(), n_sym);
    }

    // ---- SEDS text: sets -> dense bitsets --------------------------------------------------------------
    uint32_t Wd = 0, n_paths = 0, has_zero = 0;
    uint32_t* id_of = nullptr;
    if (linear) {
        const uint8_t* sraw = seds_in;
        if (!input_on_device) {
            d_sraw.reserve(seds_bytes + 16);
            if (seds_bytes) EDSB_CUDA(cudaMemcpyAsync(d_sraw.p, seds_in, seds_bytes, in_kind, s));
            sraw = d_sraw.as<uint8_t>();
        }
        SpaceCensus cs;
        census("census_seds", sraw, seds_bytes, cs);
        const uint8_t* stext;
        unsigned long long ns;
        if (cs.count == 0 || cs.first + cs.count == seds_bytes) {
            stext = sraw;
            ns = cs.count ? cs.first : seds_bytes;
        } else {
            d_stext.reserve(seds_bytes + 16);
            LEDS_SCAN("strip_seds", OpSum64, seds_bytes, (StripFn{sraw, d_stext.as<uint8_t>()}));
            stext = d_stext.as<uint8_t>();
            ns = total_of();
        }
        if (ns == 0) throw std::runtime_error("sEDS input is empty");
        const uint32_t pw = kMaxPathId / 32;
        d_present.reserve((size_t)pw * 4);
        d_rank.reserve((size_t)pw * 4);
        EDSB_CUDA(cudaMemsetAsync(d_present.p, 0, (size_t)std::min<uint32_t>(pw, present_dirty_words_) * 4, s));
        present_dirty_words_ = pw;  // until this call's largest id is known
        const SetFn sf{stext, ns, d_present.as<uint32_t>(), st};
        LEDS_SCAN("seds_sets", OpSum64, ns, sf);
        const unsigned long long tot = total_of();
        const uint32_t n_sets = (uint32_t)tot;  // (the high half counts the ids, modulo 2^32: not used)
        status_now();
        bool oor = false;
        if (hst.err & (kErrSedsSyntax | kErrSedsOverflow | kErrSedsBigId) || n_sets != n_str) {
            const std::string msg = explain_seds_error(strip_host(host_view(seds_in, seds_bytes), seds_bytes), n_str, oor);
            if (oor) throw std::out_of_range(msg);
            throw std::runtime_error(msg);
        }
        const uint32_t pw_used = hst.max_id / 32u + 1u;  // words of the presence bitmap that can be non-zero
        present_dirty_words_ = pw_used;
        LEDS_SCAN("id_rank", OpSum64, pw_used, (RankFn{d_present.as<uint32_t>(), d_rank.as<uint32_t>()}));
        n_paths = (uint32_t)total_of();
        Wd = (n_paths + 31) / 32;
        d_idof.reserve((size_t)(n_paths + 1) * 4);
        id_of = d_idof.as<uint32_t>();
        LEDS_LAUNCH("k_id_table", k_id_table, G, B, d_present.as<uint32_t>(), d_rank.as<uint32_t>(), pw_used, id_of);
        uint32_t w0 = 0;
        EDSB_CUDA(cudaMemcpyAsync(&w0, d_present.p, 4, cudaMemcpyDeviceToHost, s));
        EDSB_CUDA(cudaStreamSynchronize(s));
        has_zero = w0 & 1u;
        p_bits.ensure((size_t)std::max<uint32_t>(n_str, 1) * Wd, 0, s);
        EDSB_CUDA(cudaMemsetAsync(p_bits.p, 0, (size_t)n_str * Wd * 4, s));
        LEDS_SCAN("seds_bits", OpSum64, ns, (BitsFn{sf, d_rank.as<uint32_t>(), p_bits.p, Wd, n_str}));
        d_rawbits.reserve((size_t)n_str * Wd * 4 + 16);  // the sets as given, for alternatives that are never merged
        EDSB_CUDA(cudaMemcpyAsync(d_rawbits.p, p_bits.p, (size_t)n_str * Wd * 4, cudaMemcpyDeviceToDevice, s));
        LEDS_LAUNCH("k_universal", k_universal, G, B, p_bits.p, Wd, n_str, n_paths, has_zero, st);
        status_now();
        if (hst.empty_set != kNone) throw std::runtime_error("sEDS: Empty path set at string " + std::to_string(hst.empty_set));
    }

    // ---- pool leaves, first symbol table ----------------------------------------------------------------
    uint32_t pool_n = n_str;
    p_left.ensure(std::max<size_t>(pool_n, 1), 0, s);
    p_right.ensure(std::max<size_t>(pool_n, 1), 0, s);
    p_len.ensure(std::max<size_t>(pool_n, 1), 0, s);
    for (int i = 0; i < 4; ++i) d_tab[i].reserve((size_t)(n_sym + 2) * 4);
    SymTab cur{d_tab[0].as<uint32_t>(), d_tab[1].as<uint32_t>()}, nxt{d_tab[2].as<uint32_t>(), d_tab[3].as<uint32_t>()};
    Pool pool{p_left.p, p_right.p, p_len.p, p_bits.p};
    if (n_str)
        LEDS_LAUNCH("k_init", k_init, G, B, d_str_start.as<uint32_t>(), d_str_end.as<uint32_t>(), d_sym_first.as<uint32_t>(), n_str, n_sym, pool, cur);

    if (parse_only) {
        // EDS::parse / parse_sources end here: hand the index, the statistics and the source id lists to the host
        eds_parsed& o = *parse_only;
        memset(&o, 0, sizeof(o));
        DevBuf& d_stats = B_.d[30];  // (the emit's stack workspace: not in use on this path)
        d_stats.reserve(sizeof(EdsStatsDev));
        EdsStatsDev hs;
        memset(&hs, 0, sizeof(hs));
        hs.min_ctx = 0xffffffffu;
        EDSB_CUDA(cudaMemcpyAsync(d_stats.p, &hs, sizeof(hs), cudaMemcpyHostToDevice, s));
        if (n_sym) LEDS_LAUNCH("k_eds_stats", k_eds_stats, G, B, cur, pool, n_sym, d_stats.as<EdsStatsDev>());
        unsigned long long n_ids = 0;
        if (linear && n_str) {
            d_off.reserve((size_t)(n_str + 1) * 8);
            LEDS_SCAN("src_offsets", OpSum64, n_str, (SrcCountFn{d_rawbits.as<uint32_t>(), Wd, d_off.as<unsigned long long>(), d_stats.as<EdsStatsDev>()}));
            n_ids = total_of();
            d_kept.reserve((size_t)(n_ids + 1) * 4);
            LEDS_LAUNCH("k_src_ids", k_src_ids, G, B, d_rawbits.as<uint32_t>(), Wd, n_str, d_off.as<unsigned long long>(), id_of,
                        reinterpret_cast<int32_t*>(d_kept.p));
        }
        auto grab = [&](const void* dev, size_t bytes) -> void* {
            void* h = malloc(bytes ? bytes : 1);
            if (!h) throw std::bad_alloc();
            if (bytes) {
                const cudaError_t e = cudaMemcpyAsync(h, dev, bytes, cudaMemcpyDeviceToHost, s);
                if (e != cudaSuccess) {
                    free(h);
                    throw CudaError(std::string("device to host copy: ") + cudaGetErrorString(e));
                }
            }
            return h;
        };
        try {
            o.text = static_cast<uint8_t*>(grab(text, n));
            o.text_bytes = n;
            o.str_start = static_cast<uint32_t*>(grab(d_str_start.p, (size_t)n_str * 4));
            o.str_end = static_cast<uint32_t*>(grab(d_str_end.p, (size_t)n_str * 4));
            o.sym_first = static_cast<uint32_t*>(grab(n_str ? d_sym_first.p : nullptr, n_str ? (size_t)(n_sym + 1) * 4 : 0));
            if (!n_str) o.sym_first[0] = 0;
            o.n_strings = n_str;
            o.n_symbols = n_sym;
            o.has_sources = linear ? 1u : 0u;
            if (linear) {
                o.src_off = static_cast<uint64_t*>(grab(n_str ? d_off.p : nullptr, n_str ? (size_t)n_str * 8 : 0));
                o.src_off = static_cast<uint64_t*>(realloc(o.src_off, (size_t)(n_str + 1) * 8));
                if (!o.src_off) throw std::bad_alloc();
                o.src_ids = static_cast<int32_t*>(grab(d_kept.p, (size_t)n_ids * 4));
            }
            EDSB_CUDA(cudaMemcpyAsync(&hs, d_stats.p, sizeof(hs), cudaMemcpyDeviceToHost, s));
            EDSB_CUDA(cudaStreamSynchronize(s));
            EDSB_CUDA(cudaGetLastError());
        } catch (...) {
            eds_parsed_free(&o);
            throw;
        }
        if (linear) o.src_off[n_str] = n_ids;
        o.total_chars = hs.total_chars;
        o.num_degenerate_symbols = hs.num_degenerate;
        o.num_common_chars = hs.num_common_chars;
        o.total_change_size = hs.total_change_size;
        o.num_empty_strings = hs.num_empty_strings;
        o.sum_context_length = hs.sum_ctx;
        o.num_context_blocks = hs.num_ctx_blocks;
        o.min_context_length = hs.min_ctx == 0xffffffffu ? 0u : hs.min_ctx;
        o.max_context_length = hs.max_ctx;
        o.num_paths = n_paths;
        o.max_paths_per_string = hs.max_paths_per_string;
        o.total_paths = n_ids;
        clk.resolve();
        return;
    }

    // ---- merge rounds (eds_transforms.cpp:335-359) -------------------------------------------------------
    d_cand.reserve((size_t)n_sym + 1);
    d_sel.reserve((size_t)n_sym + 1);
    d_pairs_before.reserve((size_t)(n_sym + 1) * 4);
    d_pair_list.reserve((size_t)(n_sym + 1) * 4);
    d_kept.reserve((size_t)(n_sym + 1) * 8);
    d_off.reserve((size_t)(n_sym + 1) * 8);
    uint32_t cur_n = n_sym, rounds = 0;
    const uint32_t kMaxRounds = 10000;
    if (single_pair && *single_pair + 1 >= (uint64_t)cur_n)  // eds.cpp:1437-1442
        throw std::out_of_range("Position out of range: pos1=" + std::to_string(*single_pair) + ", pos2=" + std::to_string(*single_pair + 1) +
                                ", n=" + std::to_string(cur_n));
    while (cur_n >= 2) {
        if (rounds >= kMaxRounds) throw std::runtime_error("Maximum iterations reached without convergence");
        if (single_pair) {
            if (rounds == 1) break;  // EDS::merge_adjacent: exactly this pair, once
            LEDS_LAUNCH("k_select_one", k_select_one, G, B, d_sel.as<uint8_t>(), cur_n, (uint32_t)*single_pair);
        } else {
            LEDS_LAUNCH("k_cand", k_cand, G, B, cur, pool, cur_n, l, d_cand.as<uint8_t>());
            LEDS_SCAN("run_parity", OpMax64, cur_n, (RunFn{d_cand.as<uint8_t>(), d_sel.as<uint8_t>()}));
        }
        LEDS_SCAN("pair_list", OpSum64, cur_n, (PairListFn{d_sel.as<uint8_t>(), d_pairs_before.as<uint32_t>(), d_pair_list.as<uint32_t>()}));
        const uint32_t n_pairs = (uint32_t)total_of();
        if (check_only) {
            *check_only = n_pairs == 0 ? 1 : 0;
            return;
        }
        if (n_pairs == 0) break;
        LEDS_LAUNCH("k_kept", k_kept, G, B, cur, pool, d_pair_list.as<uint32_t>(), n_pairs, Wd, d_kept.as<unsigned long long>(), st);
        LEDS_SCAN("kept_offsets", OpSum64, n_pairs, (KeptFn{d_kept.as<unsigned long long>(), d_off.as<unsigned long long>()}));
        const unsigned long long added = total_of();
        status_now();
        if (hst.empty_merge != kNone)  // eds.cpp:1513-1519
            throw std::runtime_error("Merging positions " + std::to_string(hst.empty_merge) + " and " +
                                     std::to_string(hst.empty_merge + 1) + " results in empty set (no valid source intersections)");
        if ((unsigned long long)pool_n + added >= 0xfffffff0ull || (max_output_bytes && added > max_output_bytes))
            throw BudgetError("merged EDS exceeds the output budget (" + std::to_string(added) + " alternatives in one round)");
        p_left.ensure((size_t)pool_n + added, pool_n, s);
        p_right.ensure((size_t)pool_n + added, pool_n, s);
        p_len.ensure((size_t)pool_n + added, pool_n, s);
        if (Wd) p_bits.ensure(((size_t)pool_n + added) * Wd, (size_t)pool_n * Wd, s);
        pool = Pool{p_left.p, p_right.p, p_len.p, p_bits.p};
        LEDS_LAUNCH("k_merge_write", k_merge_write, G, B, cur, pool, d_pair_list.as<uint32_t>(), n_pairs, Wd, d_off.as<unsigned long long>(), pool_n);
        LEDS_LAUNCH("k_rebuild", k_rebuild, G, B, cur, nxt, d_sel.as<uint8_t>(), d_pairs_before.as<uint32_t>(), cur_n,
                    d_kept.as<unsigned long long>(), d_off.as<unsigned long long>(), pool_n);
        std::swap(cur, nxt);
        pool_n += (uint32_t)added;
        cur_n -= n_pairs;
        ++rounds;
    }
    if (rounds_out) *rounds_out = rounds;
    if (edge_unmerged) {
        // for a caller that cut the EDS inside long conserved symbols (eds_group_leds_merge_host): did the shard's first
        // and last symbol stay the leaves they were? (a pool index below n_str is an original string)
        *edge_unmerged = 0;
        if (cur_n) {
            uint32_t first[2] = {0, 0}, last[2] = {0, 0};
            EDSB_CUDA(cudaMemcpyAsync(&first[0], cur.begin, 4, cudaMemcpyDeviceToHost, s));
            EDSB_CUDA(cudaMemcpyAsync(&first[1], cur.count, 4, cudaMemcpyDeviceToHost, s));
            EDSB_CUDA(cudaMemcpyAsync(&last[0], cur.begin + (cur_n - 1), 4, cudaMemcpyDeviceToHost, s));
            EDSB_CUDA(cudaMemcpyAsync(&last[1], cur.count + (cur_n - 1), 4, cudaMemcpyDeviceToHost, s));
            EDSB_CUDA(cudaStreamSynchronize(s));
            if (first[1] == 1u && first[0] < n_str) *edge_unmerged |= 1u;
            if (last[1] == 1u && last[0] < n_str) *edge_unmerged |= 2u;
        }
    }
    if (check_only) {  // fewer than two symbols: nothing can be merged
        *check_only = 1;
        return;
    }

    // ---- emit (EDS::save / save_sources) -------------------------------------------------------------------
    unsigned long long n_falt = 0, eds_total = 0, seds_total = 0;
    if (cur_n) {
        d_falt_off.reserve((size_t)(cur_n + 1) * 8);
        LEDS_SCAN("alt_offsets", OpSum64, cur_n, (AltCountFn{cur, d_falt_off.as<unsigned long long>()}));
        n_falt = total_of();
        d_falt_pool.reserve((size_t)(n_falt + 1) * 4);
        d_falt_flags.reserve((size_t)n_falt + 1);
        d_eds_off.reserve((size_t)(n_falt + 1) * 8);
        LEDS_LAUNCH("k_expand", k_expand, G, B, cur, cur_n, d_falt_off.as<unsigned long long>(), compact ? 1u : 0u,
                    d_falt_pool.as<uint32_t>(), d_falt_flags.as<uint8_t>());
        LEDS_SCAN("eds_offsets", OpSum64, n_falt, (EdsOffFn{d_falt_pool.as<uint32_t>(), d_falt_flags.as<uint8_t>(), pool.len, d_eds_off.as<unsigned long long>()}));
        eds_total = total_of();
        if (linear) {
            d_seds_sz.reserve((size_t)(n_falt + 1) * 4);
            d_seds_off.reserve((size_t)(n_falt + 1) * 8);
            LEDS_LAUNCH("k_seds_sizes", k_seds_sizes, G, B, d_falt_pool.as<uint32_t>(), n_falt, pool, d_rawbits.as<uint32_t>(), n_str, Wd, has_zero, id_of,
                        d_seds_sz.as<uint32_t>());
            LEDS_SCAN("seds_offsets", OpSum64, n_falt, (SedsOffFn{d_seds_sz.as<uint32_t>(), d_seds_off.as<unsigned long long>()}));
            seds_total = total_of();
        }
    }
    if (max_output_bytes && eds_total + 1 + seds_total + 1 > max_output_bytes)
        throw BudgetError("l-EDS output (" + std::to_string(eds_total + seds_total + 2) + " bytes) exceeds max_output_bytes");
    d_out.reserve(eds_total + 16);
    d_sout.reserve(seds_total + 16);
    if (cur_n) {
        const uint32_t threads_total = G * B, depth = rounds + 2;
        d_stack.reserve((size_t)threads_total * depth * 4);
        const uint32_t big_cap = (uint32_t)std::min<unsigned long long>(eds_total / kBigLeaf + 1, 0x7fffffffull);
        d_big.reserve((size_t)big_cap * sizeof(BigCopy));
        LEDS_LAUNCH("k_emit_eds", k_emit_eds, G, B, d_falt_pool.as<uint32_t>(), d_falt_flags.as<uint8_t>(), n_falt, pool,
                    d_str_start.as<uint32_t>(), text, d_eds_off.as<unsigned long long>(), d_out.as<uint8_t>(), d_stack.as<uint32_t>(),
                    depth, d_big.as<BigCopy>(), big_cap, st);
        status_now();
        if (hst.n_big > big_cap) throw std::runtime_error("edsparser_b200: big-copy list overflow");
        if (hst.n_big) LEDS_LAUNCH("k_big_copy", k_big_copy, std::min<uint32_t>(hst.n_big, G), B, d_big.as<BigCopy>(), hst.n_big, text, d_out.as<uint8_t>());
        if (linear)
            LEDS_LAUNCH("k_emit_seds", k_emit_seds, G, B, d_falt_pool.as<uint32_t>(), n_falt, pool, d_rawbits.as<uint32_t>(), n_str, Wd, has_zero,
                        id_of, d_seds_off.as<unsigned long long>(), d_sout.as<uint8_t>());
    }
    LEDS_LAUNCH("k_newline", k_newline, 1, 32, d_out.as<uint8_t>(), eds_total, linear ? d_sout.as<uint8_t>() : nullptr, seds_total);
    // results straight into their final host buffers: malloc'd (the caller frees them), or wherever `sink` puts them
    // (eds_vcf_transform_host_view: pinned memory kept by the context)
    const uint64_t eds_bytes_out = eds_total + 1, seds_bytes_out = linear ? seds_total + 1 : 0;
    auto place = [&](int which, uint64_t bytes) -> uint8_t* {
        uint8_t* p = sink ? sink(which, bytes) : static_cast<uint8_t*>(malloc(bytes ? bytes : 1));
        if (!p) throw std::bad_alloc();
        return p;
    };
    uint8_t* h_eds_p = place(0, eds_bytes_out);
    uint8_t* h_seds_p = nullptr;
    try {
        h_seds_p = place(1, seds_bytes_out);
        EDSB_CUDA(cudaMemcpyAsync(h_eds_p, d_out.p, eds_bytes_out, cudaMemcpyDeviceToHost, s));
        if (linear) EDSB_CUDA(cudaMemcpyAsync(h_seds_p, d_sout.p, seds_bytes_out, cudaMemcpyDeviceToHost, s));
        EDSB_CUDA(cudaStreamSynchronize(s));
        EDSB_CUDA(cudaGetLastError());
    } catch (...) {
        if (!sink) {
            free(h_eds_p);
            free(h_seds_p);
        }
        throw;
    }
    clk.resolve();
    leds_out->data = h_eds_p;
    leds_out->bytes = eds_bytes_out;
    seds_out->data = h_seds_p;
    seds_out->bytes = seds_bytes_out;
#undef LEDS_SCAN
#undef LEDS_LAUNCH
}

// =====================================================================================================================
// genrandomeds on the device (reference tool src/cpp/tools/genrandomeds.cpp:221-352): an EDS + SEDS
// pair of the tool's SHAPE generated straight into HBM — reference of `n` random bases, a position is a variant site
// with probability ppm / 10^6, a site has 2..4 alternatives (the reference base first; the others 70 % SNP, 15 %
// insertion of the base + 1..10 random bases, 15 % the empty string), P paths, alternative k carried by path k + 1 and
// every further path by a random alternative, conserved runs carry {0}; no trailing newline. The reference tool draws
// from libstdc++'s mt19937 stream, which is sequential; here every position is keyed on (seed, position) through
// splitmix64, so any slice can be produced anywhere (edsparser_b200/synth.py::genrandomeds is the numpy statement of
// the same function, pinned to this kernel in the tests). One scan: per-position (EDS bytes | SEDS bytes << 32).
__device__ __forceinline__ unsigned long long gmix64(unsigned long long x) {
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}

struct GenSite {
    unsigned long long h;
    uint32_t base;    // index into ACGT
    uint32_t n_alts;  // 0: not a site
};

struct GenFn {
    unsigned long long n, seed;
    uint32_t ppm, P;
    uint8_t* eds;
    uint8_t* seds;
    __device__ __forceinline__ GenSite site(unsigned long long i) const {
        GenSite g;
        g.h = gmix64(seed ^ gmix64(i));
        g.base = (uint32_t)(g.h & 3u);
        const bool is = ((g.h >> 8) % 1000000ull) < ppm;
        g.n_alts = is ? min(P, 2u + (uint32_t)((g.h >> 40) % 3ull)) : 0u;
        return g;
    }
    // alternative k >= 1 of a site: length, and its characters through out(j, ch)
    template <typename Out>
    __device__ __forceinline__ uint32_t alt(const GenSite& g, uint32_t k, Out out) const {
        const unsigned long long hk = gmix64(g.h ^ gmix64(k));
        const uint32_t r = (uint32_t)(hk % 100ull);
        if (r < 70u) {  // SNP: another base
            out(0u, (uint8_t)"ACGT"[(g.base + 1u + (uint32_t)((hk >> 8) % 3ull)) & 3u]);
            return 1u;
        }
        if (r < 85u) {  // insertion: the base, then 1..10 random bases
            const uint32_t extra = 1u + (uint32_t)((hk >> 16) % 10ull);
            out(0u, (uint8_t)"ACGT"[g.base]);
            for (uint32_t j = 0; j < extra; ++j) out(1u + j, (uint8_t)"ACGT"[gmix64(hk + j) & 3u]);
            return 1u + extra;
        }
        return 0u;  // deletion: the empty string
    }
    __device__ __forceinline__ uint32_t alt_of_path(const GenSite& g, uint32_t p) const {  // p = 1..P
        return p <= g.n_alts ? p - 1u : (uint32_t)(gmix64(g.h ^ gmix64(100ull + p)) % g.n_alts);
    }
    __device__ unsigned long long value(unsigned long long i) const {
        const GenSite g = site(i);
        if (!g.n_alts) {
            const bool first = i == 0 || site(i - 1).n_alts != 0, last = i + 1 == n || site(i + 1).n_alts != 0;
            return (unsigned long long)(1u + first + last) | ((unsigned long long)(first ? 3u : 0u) << 32);
        }
        uint32_t e = 2u + 1u + (g.n_alts - 1u);  // braces, the reference base, separators
        for (uint32_t k = 1; k < g.n_alts; ++k) e += alt(g, k, [](uint32_t, uint8_t) {});
        uint32_t sb = 2u * g.n_alts;  // braces of every set
        for (uint32_t p = 1; p <= P; ++p) sb += decimal_width(p) + 1u;  // "id," (the last ',' of a set is its '}')
        sb -= g.n_alts;
        return (unsigned long long)e | ((unsigned long long)sb << 32);
    }
    __device__ void apply(unsigned long long i, unsigned long long prefix, unsigned long long) const {
        const GenSite g = site(i);
        uint8_t* e = eds + (prefix & 0xffffffffull);
        uint8_t* sd = seds + (prefix >> 32);
        if (!g.n_alts) {
            const bool first = i == 0 || site(i - 1).n_alts != 0, last = i + 1 == n || site(i + 1).n_alts != 0;
            if (first) {
                *e++ = '{';
                sd[0] = '{';
                sd[1] = '0';
                sd[2] = '}';
            }
            *e++ = (uint8_t)"ACGT"[g.base];
            if (last) *e = '}';
            return;
        }
        *e++ = '{';
        *e++ = (uint8_t)"ACGT"[g.base];
        for (uint32_t k = 1; k < g.n_alts; ++k) {
            *e++ = ',';
            e += alt(g, k, [e](uint32_t j, uint8_t ch) { e[j] = ch; });
        }
        *e = '}';
        for (uint32_t k = 0; k < g.n_alts; ++k) {
            *sd++ = '{';
            for (uint32_t p = 1; p <= P; ++p)
                if (alt_of_path(g, p) == k) {
                    const uint32_t w = decimal_width(p);
                    write_decimal(sd, p, w);
                    sd[w] = ',';
                    sd += w + 1u;
                }
            sd[-1] = '}';
        }
    }
};

void LedsPipeline::genrandomeds(uint64_t n, uint32_t ppm, uint32_t paths, uint64_t seed, eds_buffer* eds_out, eds_buffer* seds_out) {
    if (n == 0 || n >= 0xfffffff0ull / 3) throw std::invalid_argument("eds_genrandomeds_device: reference size out of range");
    if (paths < 2 || paths > 1000000) throw std::invalid_argument("eds_genrandomeds_device: need 2 .. 10^6 paths");
    if (ppm > 1000000) throw std::invalid_argument("eds_genrandomeds_device: variability above 1");
    cudaStream_t s = ctx_->stream;
    const uint32_t sms = (uint32_t)ctx_->sm_count;
    const unsigned P = std::max(1u, std::min(ctx_->partitions ? ctx_->partitions : sms * 8u, 4096u));
    Bufs& B_ = *bufs_;
    DevBuf &d_part = B_.d[3], &d_eds = B_.d[36], &d_seds = B_.d[37];
    d_part.reserve((size_t)(P + 1) * 16);
    unsigned long long* part = d_part.as<unsigned long long>();
    // sizes first (the reduce pass of the scan), then the text
    GenFn fn{n, seed, ppm, paths, nullptr, nullptr};
    auto reduce = k_part_reduce<OpSum64, GenFn>;
    EDSB_LAUNCH(reduce, P, kScanBlock, 0, s, (unsigned long long)n, fn, part);
    std::vector<unsigned long long> h(P);
    EDSB_CUDA(cudaMemcpyAsync(h.data(), part, (size_t)P * 8, cudaMemcpyDeviceToHost, s));
    EDSB_CUDA(cudaStreamSynchronize(s));
    unsigned long long eb = 0, sb = 0;
    for (unsigned long long v : h) {
        eb += v & 0xffffffffull;
        sb += v >> 32;
    }
    if (eb >= 0xfffffff0ull || sb >= 0xfffffff0ull) throw std::invalid_argument("eds_genrandomeds_device: output above 4 GiB");
    d_eds.reserve(eb + 16);
    d_seds.reserve(sb + 16);
    fn.eds = d_eds.as<uint8_t>();
    fn.seds = d_seds.as<uint8_t>();
    auto apply = k_part_apply<OpSum64, GenFn>;
    EDSB_LAUNCH(apply, P, kScanBlock, 0, s, (unsigned long long)n, fn, part);
    EDSB_CUDA(cudaStreamSynchronize(s));
    EDSB_CUDA(cudaGetLastError());
    eds_out->data = d_eds.as<uint8_t>();
    eds_out->bytes = eb;
    seds_out->data = d_seds.as<uint8_t>();
    seds_out->bytes = sb;
}

}  // namespace edsb
