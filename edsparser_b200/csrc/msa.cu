// MSA -> EDS / l-EDS on the GPU: the kernels behind eds_msa_transform_device.
//
// Reference behaviour (draessld/EDSParser, src/cpp/lib/transforms/msa_transforms.cpp):
//   pass 1  parse_msa_and_build_variant_bv :36-90   -> k_scan, k_colbits
//   pass 2  build_eds_boundaries :101-115 / build_leds_boundaries :133-190
//                                                   -> k_compact, k_sym_count, k_sym_scatter, k_finalize
//   pass 3  generate_output :200-324                -> k_stash, k_group, k_size_*, k_emit_*
// Nothing here is derived from the reference's code: the reference walks the file with getline /
// seekg / std::map; this is a column-parallel formulation whose output bytes are identical.
//
// Data flow (all in HBM, one stream, no host round trip until the final status copy):
//   text (FASTA rows, any alignment) --k_scan--> mismatch bits in row-0-aligned byte space
//   --k_colbits--> variable-column bits V, run-start bits T, row 0 in column space
//   --k_compact--> variable-column list + rank directory, run list
//   --k_stash--> stash[k][r]: the R residues of every variable column, column-major (the only
//                re-read of the alignment: R bytes per variable column)
//   --k_sym_*--> symbol list (runs that open a symbol), --k_finalize--> owned range / shard edges
//   --k_group--> per variable symbol: hash rows, exact dedup, alternative ids, sizes
//   --k_size_*--> output offsets, --k_emit_*--> EDS and SEDS text.
#include "msa.h"

#include "idlist.cuh"

#include <string.h>

#include <algorithm>

namespace edsb {

// ---------------------------------------------------------------------------------------------
// k_scan: column-conservation scan (msa_transforms.cpp:69-84).
// One thread owns one 16-byte chunk of p-space and walks all rows: acc |= row ^ row0. A row whose
// start is not congruent to row 0's mod 16 is read as two aligned vectors and funnel-shifted.
// Algorithmic bytes: R * row_bytes read once; writes 2 bytes per chunk.
// ---------------------------------------------------------------------------------------------
#ifdef EDSB_EMU
constexpr int kScanThreads = 64;
#else
constexpr int kScanThreads = 256;
#endif
#ifndef EDSB_SCAN_MINB
#define EDSB_SCAN_MINB 3
#endif
#ifndef EDSB_SCAN_UNROLL
#define EDSB_SCAN_UNROLL 8
#endif
constexpr int kScanUnroll = EDSB_SCAN_UNROLL;
constexpr int kRowCache = 2048;

// row_pack: for every row r >= 1, the address of the aligned 16-byte vector that holds p-space byte 0 of
// the row, with the row's byte shift (0..15) in the low 4 bits; sorted on the host by word shift
// (shift >> 2) so that each class runs with a compile-time word selection and only the bit shift is a
// run-time operand of the funnel shifts (g.cls[c]..g.cls[c+1] = rows of class c). Entry 0 is row 0.
template <int WS>
__device__ __forceinline__ void xor_acc(const uint4& lo, const uint4& hi, uint32_t bs, const uint4& ref, uint4& acc) {
    uint32_t w0, w1, w2, w3, w4;
    if (WS == 0) { w0 = lo.x; w1 = lo.y; w2 = lo.z; w3 = lo.w; w4 = hi.x; }
    else if (WS == 1) { w0 = lo.y; w1 = lo.z; w2 = lo.w; w3 = hi.x; w4 = hi.y; }
    else if (WS == 2) { w0 = lo.z; w1 = lo.w; w2 = hi.x; w3 = hi.y; w4 = hi.z; }
    else { w0 = lo.w; w1 = hi.x; w2 = hi.y; w3 = hi.z; w4 = hi.w; }
    acc.x |= __funnelshift_r(w0, w1, bs) ^ ref.x;
    acc.y |= __funnelshift_r(w1, w2, bs) ^ ref.y;
    acc.z |= __funnelshift_r(w2, w3, bs) ^ ref.z;
    acc.w |= __funnelshift_r(w3, w4, bs) ^ ref.w;
}

template <int WS>
__device__ __forceinline__ void scan_class(const unsigned long long* pk, uint32_t n, long long jj, const uint4& ref,
                                           uint4& acc) {
    uint32_t r = 0;
    for (; r + kScanUnroll <= n; r += kScanUnroll) {
        uint4 lo[kScanUnroll], hi[kScanUnroll];
        uint32_t bs[kScanUnroll];
#pragma unroll
        for (int u = 0; u < kScanUnroll; ++u) {
            const unsigned long long v = pk[r + u];
            const uint4* p = reinterpret_cast<const uint4*>(v & ~15ull) + jj;
            bs[u] = ((uint32_t)v & 3u) * 8u;
            lo[u] = ldg_nc(p);
            hi[u] = ldg_nc(p + 1);
        }
#pragma unroll
        for (int u = 0; u < kScanUnroll; ++u) xor_acc<WS>(lo[u], hi[u], bs[u], ref, acc);
    }
    for (; r < n; ++r) {
        const unsigned long long v = pk[r];
        const uint4* p = reinterpret_cast<const uint4*>(v & ~15ull) + jj;
        const uint4 lo = ldg_nc(p), hi = ldg_nc(p + 1);
        xor_acc<WS>(lo, hi, ((uint32_t)v & 3u) * 8u, ref, acc);
    }
}

// every row starts on a 16-byte boundary relative to row 0 (e.g. a pitched copy): one load per row
__device__ __forceinline__ void scan_aligned(const unsigned long long* pk, uint32_t n, long long jj, const uint4& ref,
                                             uint4& acc) {
    uint32_t r = 0;
    for (; r + kScanUnroll <= n; r += kScanUnroll) {
        uint4 lo[kScanUnroll];
#pragma unroll
        for (int u = 0; u < kScanUnroll; ++u) lo[u] = ldg_nc(reinterpret_cast<const uint4*>(pk[r + u]) + jj);
#pragma unroll
        for (int u = 0; u < kScanUnroll; ++u) {
            acc.x |= lo[u].x ^ ref.x;
            acc.y |= lo[u].y ^ ref.y;
            acc.z |= lo[u].z ^ ref.z;
            acc.w |= lo[u].w ^ ref.w;
        }
    }
    for (; r < n; ++r) {
        const uint4 lo = ldg_nc(reinterpret_cast<const uint4*>(pk[r]) + jj);
        acc.x |= lo.x ^ ref.x;
        acc.y |= lo.y ^ ref.y;
        acc.z |= lo.z ^ ref.z;
        acc.w |= lo.w ^ ref.w;
    }
}

// gridDim.y > 1 (few columns per device, many rows: a column shard of a deep alignment): the rows are split into
// gridDim.y slices, every block scans one slice and ORs its 16 mismatch bits into the (pre-zeroed) word.
__device__ __forceinline__ void row_slice(uint32_t lo, uint32_t hi, uint32_t& a, uint32_t& b) {
    const uint32_t n = hi - lo;
    a = lo + (uint32_t)((unsigned long long)n * blockIdx.y / gridDim.y);
    b = lo + (uint32_t)((unsigned long long)n * (blockIdx.y + 1) / gridDim.y);
}

template <bool CACHED>
__global__ void __launch_bounds__(kScanThreads, EDSB_SCAN_MINB) k_scan(MsaGeom g, const unsigned long long* row_pack, uint16_t* mism16,
                                                       MsaStatus* st) {
    __shared__ unsigned long long s_pk[CACHED ? kRowCache : 1];
    if (CACHED) {
        for (uint32_t i = threadIdx.x; i < g.R; i += blockDim.x) s_pk[i] = row_pack[i];
        __syncthreads();
    }
    const unsigned long long* pk = CACHED ? s_pk : row_pack;

    const uint4* vec = reinterpret_cast<const uint4*>(g.text);
    const long long vmax = (long long)g.n_vec - 1;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t wpb = blockDim.x >> 5;
    const uint32_t n_tiles = (g.n_chunks + 31) / 32;
    uint32_t bad = 0;

    for (uint32_t tile = blockIdx.x * wpb + (threadIdx.x >> 5); tile < n_tiles; tile += gridDim.x * wpb) {
        const uint32_t j = tile * 32 + lane;
        const long long j0 = (long long)tile * 32;
        // interior tile: every vector any lane touches (and its right neighbour) lies inside the buffer
        const bool interior = j0 + 31 < (long long)g.n_chunks && j0 + g.d_min_vec >= 0 && j0 + 32 + g.d_max_vec <= vmax;
        if (j >= g.n_chunks) continue;
        const long long jj = (long long)j;
        uint4 ref, acc = make_uint4(0, 0, 0, 0);
        if (interior) {
            ref = ldg_nc(reinterpret_cast<const uint4*>(pk[0]) + jj);
            uint32_t a, b;
            if (g.all_aligned) {
                row_slice(1, g.R, a, b);
                scan_aligned(pk + a, b - a, jj, ref, acc);
            } else {
                row_slice(g.cls[0], g.cls[1], a, b);
                scan_class<0>(pk + a, b - a, jj, ref, acc);
                row_slice(g.cls[1], g.cls[2], a, b);
                scan_class<1>(pk + a, b - a, jj, ref, acc);
                row_slice(g.cls[2], g.cls[3], a, b);
                scan_class<2>(pk + a, b - a, jj, ref, acc);
                row_slice(g.cls[3], g.cls[4], a, b);
                scan_class<3>(pk + a, b - a, jj, ref, acc);
            }
        } else {
            // edge tile: clamp every vector index into the buffer (bytes outside the row are masked below)
            const long long d0 = (long long)g.row_off[0] - (long long)g.a0;
            long long v0 = jj + (d0 >> 4);
            v0 = v0 < 0 ? 0 : (v0 > vmax ? vmax : v0);
            ref = ldg_nc(vec + v0);
            uint32_t r_lo, r_hi;
            row_slice(1, g.R, r_lo, r_hi);
            for (uint32_t r = r_lo; r < r_hi; ++r) {
                const long long d = (long long)g.row_off[r] - (long long)g.a0;
                const long long vi = jj + (d >> 4);
                const uint32_t sh = (uint32_t)(d & 15);
                const long long va = vi < 0 ? 0 : (vi > vmax ? vmax : vi);
                long long vb = vi + 1;
                vb = vb < 0 ? 0 : (vb > vmax ? vmax : vb);
                const uint4 x = realign16(ldg_nc(vec + va), ldg_nc(vec + vb), sh);
                acc.x |= x.x ^ ref.x;
                acc.y |= x.y ^ ref.y;
                acc.z |= x.z ^ ref.z;
                acc.w |= x.w ^ ref.w;
            }
        }

        // bytes of this chunk that belong to the row segment
        const uint64_t p0 = (uint64_t)j * 16u;
        const uint64_t pend = (uint64_t)g.a0 + g.row_bytes;
        const uint32_t vlo = p0 >= g.a0 ? 0u : (uint32_t)(g.a0 - p0);
        const uint32_t vhi = pend >= p0 + 16u ? 16u : (pend > p0 ? (uint32_t)(pend - p0) : 0u);
        const uint32_t valid = low_bits(vhi) & ~low_bits(vlo);
        // where line breaks must be: u % (lw + 1) == lw
        uint32_t expect = 0;
        {
            const uint64_t u_first = g.u_begin + (p0 + vlo - g.a0);
            uint32_t rem = (uint32_t)(u_first % (uint64_t)(g.lw + 1u));
            for (uint32_t i = vlo; i < vhi; ++i) {
                if (rem == g.lw) expect |= 1u << i;
                rem = (rem == g.lw) ? 0u : rem + 1u;
            }
        }
        uint32_t mism = nonzero_bytes16(acc) | eq_bytes16(ref, 0x2d2d2d2du);  // differs from row 0, or row 0 is '-'
        const uint32_t nl = eq_bytes16(ref, 0x0a0a0a0au);
        if (((nl ^ expect) | (nonzero_bytes16(acc) & expect)) & valid) bad = 1;
        mism &= valid & ~expect;
        if (gridDim.y == 1)
            mism16[j] = (uint16_t)mism;
        else if (mism)
            atomicOr(reinterpret_cast<uint32_t*>(mism16) + (j >> 1), mism << ((j & 1u) * 16u));
    }
    if (bad) atomicOr(&st->bad_msa, (uint32_t)kBadNewlineLayout);
}

}  // namespace edsb
#include "scan_fused.cuh"
#include "scan_l2.cuh"
namespace edsb {

// ---------------------------------------------------------------------------------------------
// k_colbits: p-space mismatch bits -> column-space V (variable) and T (run start) words, row 0 in
// column space, per-partition counts. Thread per 32-column word.
// ---------------------------------------------------------------------------------------------
#ifdef EDSB_EMU
constexpr int kPartThreads = 64;  // fewer OS threads per emulated block; results do not depend on it
#else
constexpr int kPartThreads = 256;
#endif

__device__ __forceinline__ uint32_t get_bits(const uint32_t* m, uint64_t p, uint32_t n) {
    const uint64_t wi = p >> 5;
    const uint32_t v = __funnelshift_r(m[wi], m[wi + 1], (uint32_t)(p & 31u));
    return v & low_bits(n);
}

__global__ void __launch_bounds__(kPartThreads) k_colbits(MsaGeom g, const uint32_t* mism, uint32_t* vbits,
                                                          uint32_t* tbits, uint8_t* refc, uint2* part_cnt) {
    __shared__ unsigned long long s_red[33];
    const uint32_t P = gridDim.x;
    const uint32_t wpp = (g.n_words + P - 1) / P;
    const uint32_t w_begin = min(g.n_words, blockIdx.x * wpp);
    const uint32_t w_end = min(g.n_words, w_begin + wpp);
    const uint8_t* row0 = g.text + g.row_off[0];
    unsigned long long cnt = 0;
    for (uint32_t w = w_begin + threadIdx.x; w < w_end; w += blockDim.x) {
        const uint32_t c0 = w * 32u;
        const uint32_t nvalid = min(32u, g.ncols - c0);
        const uint64_t gc = g.col_begin + c0;
        const uint64_t q = gc / g.lw;
        uint32_t rem = (uint32_t)(gc - q * g.lw);
        uint64_t u = gc + q - g.u_begin;  // byte offset inside the row segment
        uint32_t V = 0, filled = 0;
        uint32_t chars[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        while (filled < nvalid) {
            const uint32_t n = min(nvalid - filled, g.lw - rem);
            V |= get_bits(mism, u + g.a0, n) << filled;
            for (uint32_t i = 0; i < n; ++i) {
                const uint32_t at = filled + i;
                chars[at >> 2] |= (uint32_t)row0[u + i] << ((at & 3u) * 8u);
            }
            filled += n;
            u += n + 1u;  // skip the line break
            rem = 0;
        }
        uint32_t prev;
        if (c0 == 0) {
            prev = (~V) & 1u;  // a run always starts at the first held column
        } else {
            const uint64_t gp = gc - 1;
            const uint64_t up = gp + gp / g.lw - g.u_begin;
            prev = get_bits(mism, up + g.a0, 1);
        }
        const uint32_t T = (V ^ ((V << 1) | prev)) & low_bits(nvalid);
        vbits[w] = V;
        tbits[w] = T;
        uint4* dst = reinterpret_cast<uint4*>(refc + (size_t)c0);
        dst[0] = make_uint4(chars[0], chars[1], chars[2], chars[3]);
        dst[1] = make_uint4(chars[4], chars[5], chars[6], chars[7]);
        cnt += (unsigned long long)__popc(V) | ((unsigned long long)__popc(T) << 32);
    }
    const unsigned long long tot = block_sum(cnt, s_red);
    if (threadIdx.x == 0) part_cnt[blockIdx.x] = uint2{(uint32_t)tot, (uint32_t)(tot >> 32)};
}

// ---------------------------------------------------------------------------------------------
// k_compact: stream compaction of V into the variable-column list (+ rank directory) and of T into
// the run list. Each block first sums the counts of the partitions before it.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPartThreads) k_compact(MsaGeom g, MsaBufs b) {
    __shared__ unsigned long long s_scan[33];
    const uint32_t P = gridDim.x;
    const uint32_t wpp = (g.n_words + P - 1) / P;
    const uint32_t w_begin = min(g.n_words, blockIdx.x * wpp);
    const uint32_t w_end = min(g.n_words, w_begin + wpp);
    unsigned long long mine = 0;
    for (uint32_t q = threadIdx.x; q < blockIdx.x; q += blockDim.x) {
        const uint2 c = b.part_cnt[q];
        mine += (unsigned long long)c.x | ((unsigned long long)c.y << 32);
    }
    const unsigned long long base = block_sum(mine, s_scan);
    uint32_t base_var = (uint32_t)base, base_run = (uint32_t)(base >> 32);
    for (uint32_t w0 = w_begin; w0 < w_end; w0 += blockDim.x) {
        const uint32_t w = w0 + threadIdx.x;
        uint32_t V = 0, T = 0;
        if (w < w_end) {
            V = b.vbits[w];
            T = b.tbits[w];
        }
        const unsigned long long cnt = (unsigned long long)__popc(V) | ((unsigned long long)__popc(T) << 32);
        unsigned long long total;
        const unsigned long long ex = block_exclusive_scan(cnt, s_scan, total);
        if (w < w_end) {
            uint32_t iv = base_var + (uint32_t)ex, ir = base_run + (uint32_t)(ex >> 32);
            b.rankdir[w] = iv;
            for (uint32_t bits = V; bits; bits &= bits - 1) {
                const uint32_t bit = (uint32_t)__ffs((int)bits) - 1u;
                if (iv < b.cap_var) b.varcol[iv] = w * 32u + bit;
                ++iv;
            }
            for (uint32_t bits = T; bits; bits &= bits - 1) {
                const uint32_t bit = (uint32_t)__ffs((int)bits) - 1u;
                if (ir < b.cap_runs) b.runs[ir] = (w * 32u + bit) | (((V >> bit) & 1u) ? 0u : kCommonFlag);
                ++ir;
            }
        }
        base_var += (uint32_t)total;
        base_run += (uint32_t)(total >> 32);
    }
    if (blockIdx.x == P - 1 && threadIdx.x == 0) {
        MsaStatus* st = b.status;
        st->n_var = base_var;
        st->n_runs = base_run;
        st->need_var = base_var;
        st->need_runs = base_run;
        if (base_run <= b.cap_runs) b.runs[base_run] = g.ncols;  // sentinel (cap_runs + 1 entries allocated)
        if (base_var > b.cap_var)
            st->abort = kAbortVarCap;
        else if (base_run > b.cap_runs)
            st->abort = kAbortRunsCap;
    }
}

// ---------------------------------------------------------------------------------------------
// k_stash: gather the R residues of every variable column into stash[k * Rp + r] (lanes over rows,
// coalesced writes; the reads are one sector per (column, row) and are the pipeline's only re-read).
// ---------------------------------------------------------------------------------------------
#ifdef EDSB_EMU
constexpr int kStashThreads = 64;
#else
constexpr int kStashThreads = 256;
#endif

__global__ void __launch_bounds__(kStashThreads) k_stash(MsaGeom g, MsaBufs b) {
    __shared__ uint8_t tile[32][36];
    MsaStatus* st = b.status;
    if (st->abort) return;
    const uint32_t n_var = st->n_var;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const uint32_t tiles_r = (g.R + 31) / 32, tiles_c = (n_var + 31) / 32;
    const uint64_t n_tiles = (uint64_t)tiles_r * tiles_c;
    uint32_t bad = 0;
    for (uint64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const uint32_t tc = (uint32_t)(t / tiles_r), tr = (uint32_t)(t % tiles_r);
        // read: a warp instruction reads 32 variable columns of ONE row (one page, nearby sectors)
        const uint32_t k = tc * 32 + lane;
        uint64_t uoff = 0;
        if (k < n_var) {
            const uint64_t gc = g.col_begin + b.varcol[k];
            uoff = gc + gc / g.lw - g.u_begin;
        }
        for (uint32_t rl = warp; rl < 32; rl += nw) {
            const uint32_t r = tr * 32 + rl;
            if (r < g.R && k < n_var) {
                const uint8_t ch = g.text[g.row_off[r] + uoff];
                tile[rl][lane] = ch;
                bad |= (ch == (uint8_t)'\n');
            }
        }
        __syncthreads();
        // write: lanes over rows, 32 consecutive bytes of one stash column
        for (uint32_t cl = warp; cl < 32; cl += nw) {
            const uint32_t k2 = tc * 32 + cl, r = tr * 32 + lane;
            if (k2 < n_var && r < g.R) b.stash[(size_t)k2 * g.Rp + r] = tile[lane][cl];
        }
        __syncthreads();
    }
    if (bad) atomicOr(&st->bad_msa, (uint32_t)kBadResidueByte);
}

// ---------------------------------------------------------------------------------------------
// Symbol boundaries. Plain EDS (msa_transforms.cpp:101-115): every run opens a symbol.
// l-EDS (msa_transforms.cpp:133-190): a conserved run is standalone when it is at least l long or
// touches an end of the alignment; run k opens a symbol iff run k or run k-1 is standalone (or k is
// the first run of the alignment). At a window edge that is not an alignment end a short conserved
// run is of unknown length: it is treated as not standalone and k_finalize decides whether that
// matters for the owned range.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool run_standalone(const MsaGeom& g, const uint32_t* runs, uint32_t k) {
    const uint32_t e = runs[k];
    if (!(e & kCommonFlag)) return false;
    const uint32_t s = e & kColMask, en = runs[k + 1] & kColMask;
    if (en - s >= g.l) return true;
    return (g.col_begin + s == 0) || (g.col_begin + en == g.total_cols);
}

__device__ __forceinline__ bool run_opens(const MsaGeom& g, const uint32_t* runs, uint32_t k) {
    if (!g.leds || g.l == 0) return true;
    if (k == 0) return g.col_begin == 0 ? true : run_standalone(g, runs, 0);
    return run_standalone(g, runs, k) || run_standalone(g, runs, k - 1);
}

// The two kernels also list the variable symbols (varsym[v] = symbol index), the work list of k_group /
// k_emit_var: counts are packed as opens | variable opens << 32.
__global__ void __launch_bounds__(kPartThreads) k_sym_count(MsaGeom g, MsaBufs b) {
    __shared__ unsigned long long s_red[33];
    const MsaStatus* st = b.status;
    if (st->abort) return;
    const uint32_t n = st->n_runs, P = gridDim.x;
    const uint32_t per = (n + P - 1) / P;
    const uint32_t k_begin = min(n, blockIdx.x * per), k_end = min(n, k_begin + per);
    unsigned long long cnt = 0;
    for (uint32_t k = k_begin + threadIdx.x; k < k_end; k += blockDim.x)
        if (run_opens(g, b.runs, k)) cnt += (b.runs[k] & kCommonFlag) ? 1ull : (1ull | (1ull << 32));
    const unsigned long long tot = block_sum(cnt, s_red);
    if (threadIdx.x == 0) b.part_sym[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(kPartThreads) k_sym_scatter(MsaGeom g, MsaBufs b) {
    __shared__ unsigned long long s_scan[33];
    MsaStatus* st = b.status;
    if (st->abort) return;
    const uint32_t n = st->n_runs, P = gridDim.x;
    const uint32_t per = (n + P - 1) / P;
    const uint32_t k_begin = min(n, blockIdx.x * per), k_end = min(n, k_begin + per);
    unsigned long long mine = 0;
    for (uint32_t q = threadIdx.x; q < blockIdx.x; q += blockDim.x) mine += b.part_sym[q];
    unsigned long long base = block_sum(mine, s_scan);
    for (uint32_t k0 = k_begin; k0 < k_end; k0 += blockDim.x) {
        const uint32_t k = k0 + threadIdx.x;
        const bool open = k < k_end && run_opens(g, b.runs, k);
        const uint32_t e = open ? b.runs[k] : kCommonFlag;
        const bool var = open && !(e & kCommonFlag);
        unsigned long long total;
        const unsigned long long ex = block_exclusive_scan((open ? 1ull : 0ull) | (var ? 1ull << 32 : 0ull), s_scan, total);
        const uint32_t ks = (uint32_t)(base + ex);
        if (open) b.sym[ks] = e;  // an opening conserved run is a whole common symbol
        if (var) b.varsym[(uint32_t)((base + ex) >> 32)] = ks;
        base += total;
    }
    if (blockIdx.x == P - 1 && threadIdx.x == 0) {
        st->n_syms = (uint32_t)base;
        st->n_varsyms_window = (uint32_t)(base >> 32);
        b.sym[(uint32_t)base] = g.ncols;
    }
}

__device__ __forceinline__ uint32_t sym_lower_bound(const uint32_t* sym, uint32_t n, uint32_t col) {
    uint32_t lo = 0, hi = n;  // first k with start(sym[k]) >= col
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if ((sym[mid] & kColMask) < col)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

// first index in [0, n) whose (masked) entry is >= key, by the whole warp: 32 probes per step instead of one
__device__ __forceinline__ uint32_t warp_lower_bound(const uint32_t* a, uint32_t n, uint32_t key, uint32_t mask) {
    const uint32_t lane = threadIdx.x & 31;
    uint32_t lo = 0, hi = n;  // answer in [lo, hi]
    while (hi - lo > 32u) {
        const uint32_t step = (hi - lo + 31u) / 32u;
        const uint32_t at = lo + lane * step;  // lane 0 probes lo
        const bool below = at < hi && (a[at] & mask) < key;
        const uint32_t m = __ballot_sync(0xffffffffu, below);
        // entries are sorted: the lanes that answer "below" form a prefix
        const uint32_t cnt = (uint32_t)__popc(m);
        if (cnt == 0) {
            hi = lo;  // a[lo] >= key
        } else {
            const uint32_t new_lo = lo + (cnt - 1u) * step + 1u;
            const uint32_t new_hi = cnt < 32u ? min(hi, lo + cnt * step) : hi;
            lo = new_lo;
            hi = new_hi;
        }
    }
    const uint32_t at = lo + lane;
    const bool below = at < hi && (a[at] & mask) < key;
    return lo + (uint32_t)__popc(__ballot_sync(0xffffffffu, below));
}

// k_finalize: owned symbol range and what happens at the shard's two edges (one warp; the four searches are 32-ary).
__global__ void k_finalize(MsaGeom g, MsaBufs b) {
    MsaStatus* st = b.status;
    if (st->abort || blockIdx.x != 0 || threadIdx.x >= 32) return;
    const uint32_t n_runs = st->n_runs, n_syms = st->n_syms;
    const uint32_t* sym = b.sym;
    const uint32_t* runs = b.runs;
    const bool leds = g.leds && g.l > 0;
    const bool right_open = g.col_begin + g.ncols < g.total_cols;  // the alignment continues past the window
    uint32_t fail = 0;
    const uint32_t k_lo = warp_lower_bound(sym, n_syms, g.own_lo, kColMask);
    const uint32_t k_hi = warp_lower_bound(sym, n_syms, g.own_hi, kColMask);
    const uint32_t nv = st->n_varsyms_window;
    const uint32_t v_lo = warp_lower_bound(b.varsym, nv, k_lo, 0xffffffffu);
    const uint32_t v_hi = warp_lower_bound(b.varsym, nv, k_hi, 0xffffffffu);
    if (threadIdx.x != 0) return;
    uint32_t lead_lo = 0, lead_hi = 0, lead_close = 0, tail_open = 0;

    if (leds && g.col_begin > 0 && n_runs > 0) {
        // run 0 is cut by the window: conserved and short => its true length is unknown
        const uint32_t e0 = runs[0], end0 = runs[1] & kColMask;
        const bool uncertain = (e0 & kCommonFlag) && end0 < g.l && !(g.col_begin + end0 == g.total_cols);
        if (uncertain && g.own_lo <= end0) fail = 1;
    }
    if (g.own_lo < g.own_hi) {
        const bool at_start = k_lo < n_syms && (sym[k_lo] & kColMask) == g.own_lo;
        if (!at_start && k_lo > 0 && (sym[k_lo - 1] & kCommonFlag)) {
            const uint32_t end = sym[k_lo] & kColMask;  // sentinel = ncols
            lead_lo = g.own_lo;
            lead_hi = min(end, g.own_hi);
            lead_close = end <= g.own_hi ? 1u : 0u;
        }
        if (k_hi > k_lo) {
            const uint32_t last = sym[k_hi - 1], end = sym[k_hi] & kColMask;
            if (last & kCommonFlag) {
                tail_open = end > g.own_hi ? 1u : 0u;
            } else if (k_hi == n_syms && right_open) {
                fail = 1;  // the last owned variable symbol does not close inside the window
            }
        }
    }
    st->v_lo = v_lo;
    st->v_hi = v_hi;
    st->k_lo = k_lo;
    st->k_hi = k_hi;
    st->lead_lo = lead_lo;
    st->lead_hi = lead_hi;
    st->lead_close = lead_close;
    st->tail_open = tail_open;
    st->halo_fail = fail;
    st->first_open_col = k_lo < k_hi ? g.col_begin + (sym[k_lo] & kColMask) : ~0ull;
}

// ---------------------------------------------------------------------------------------------
// Walking one row's characters across the columns of a variable symbol: conserved columns read
// row 0 (column space), variable columns read the stash; '-' is dropped (msa_transforms.cpp:281-286).
// ---------------------------------------------------------------------------------------------
struct RowWalk {
    const uint32_t* vbits;
    const uint8_t* refc;
    const uint8_t* stash;
    uint32_t Rp;
    // staged mode: the symbol's columns are described in shared memory (cdesc[i]: 0x100 | index of the
    // variable column inside the staged block, or the conserved character) and its stash block is a copy
    // in shared memory; every warp lane walks without touching global memory.
    const uint16_t* cdesc;
    const uint8_t* stage;
    uint32_t s0;
    bool staged;
};

constexpr uint32_t kStageCols = 64;  // widest symbol the staged path takes

__device__ __forceinline__ int next_char(const RowWalk& w, uint32_t r, uint32_t& c, uint32_t e, uint32_t& slot) {
    if (w.staged) {
        while (c < e) {
            const uint32_t d = w.cdesc[c - w.s0];
            ++c;
            const uint8_t ch = (d & 0x100u) ? w.stage[(d & 0xffu) * w.Rp + r] : (uint8_t)d;
            if (ch != (uint8_t)'-') return (int)ch;
        }
        return -1;
    }
    while (c < e) {
        const uint32_t v = (w.vbits[c >> 5] >> (c & 31u)) & 1u;
        uint8_t ch;
        if (v) {
            ch = w.stash[(size_t)slot * w.Rp + r];
            ++slot;
        } else {
            ch = w.refc[c];
        }
        ++c;
        if (ch != (uint8_t)'-') return (int)ch;
    }
    return -1;
}

// Warp-cooperative: describe columns [s, en) in cdesc and copy the symbol's stash block (its variable
// columns are consecutive slots from slot0) into `stage`. False (warp-uniform) when the symbol is too
// wide or its block does not fit; the caller then walks global memory.
__device__ __forceinline__ bool stage_symbol(const MsaGeom& g, const MsaBufs& b, uint32_t s, uint32_t en, uint32_t slot0,
                                             uint16_t* cdesc, uint8_t* stage, uint32_t stage_bytes) {
    const uint32_t lane = threadIdx.x & 31, width = en - s;
    if (stage_bytes == 0 || width > kStageCols) return false;
    uint32_t nv = 0;
    for (uint32_t i0 = 0; i0 < width; i0 += 32) {
        const uint32_t i = i0 + lane, c = s + i;
        const bool in = i < width;
        const uint32_t v = in ? (b.vbits[c >> 5] >> (c & 31u)) & 1u : 0u;
        const uint32_t m = __ballot_sync(0xffffffffu, v);
        if (in) cdesc[i] = v ? (uint16_t)(0x100u | (nv + (uint32_t)__popc(m & lanemask_lt()))) : (uint16_t)b.refc[c];
        nv += (uint32_t)__popc(m);
    }
    if (nv > 255u || nv * g.Rp > stage_bytes) return false;
    const uint4* src = reinterpret_cast<const uint4*>(b.stash + (size_t)slot0 * g.Rp);
    uint4* dst = reinterpret_cast<uint4*>(stage);
    const uint32_t n16 = nv * g.Rp / 16u;
    for (uint32_t i = lane; i < n16; i += 32) dst[i] = src[i];
    __syncwarp();
    return true;
}

__device__ bool rows_equal(const RowWalk& w, uint32_t r1, uint32_t r2, uint32_t s, uint32_t e, uint32_t slot0) {
    uint32_t c1 = s, c2 = s, s1 = slot0, s2 = slot0;
    for (;;) {
        const int a = next_char(w, r1, c1, e, s1);
        const int bch = next_char(w, r2, c2, e, s2);
        if (a != bch) return false;
        if (a < 0) return true;
    }
}

__device__ __forceinline__ uint32_t first_slot(const MsaBufs& b, uint32_t s) {
    return b.rankdir[s >> 5] + (uint32_t)__popc(b.vbits[s >> 5] & low_bits(s & 31u));
}

__device__ __forceinline__ void store_alt(const MsaGeom& g, void* altid, size_t i, uint32_t a) {
    if (g.alt32)
        reinterpret_cast<uint32_t*>(altid)[i] = a;
    else
        reinterpret_cast<uint16_t*>(altid)[i] = (uint16_t)a;
}
__device__ __forceinline__ uint32_t load_alt(const MsaGeom& g, const void* altid, size_t i) {
    return g.alt32 ? reinterpret_cast<const uint32_t*>(altid)[i] : (uint32_t)reinterpret_cast<const uint16_t*>(altid)[i];
}

// ---------------------------------------------------------------------------------------------
// Grouping of a variable symbol's rows into alternatives (generate_output's variant branch,
// msa_transforms.cpp:259-318). Two shapes of work:
//
//  * narrow: the symbol is ONE variable column (the common case: ~80 % of the variable symbols of an
//    l = 10 run at 1 % variability, ~99 % in plain EDS). A LANE owns the symbol and walks its stash column
//    serially; the distinct residues, in order of first row, live in two registers (up to 8; more, or a
//    NUL byte, demotes the symbol to the wide path). 32 symbols per warp, ~10 instructions per row.
//  * wide: a WARP owns the symbol. Rows are keyed by the gap-stripped string — packed exactly into 64 bits
//    when the symbol is at most 7 columns wide, else FNV-1a hashed — grouped through a per-warp
//    open-addressing table that keeps the smallest row of each key; hashed keys are then verified byte for
//    byte, and a true collision switches the symbol to an exact quadratic search.
// Alternatives are numbered by first row (insertion_order, :263,290-292).
// Per-warp scratch of the wide path (shared memory, or global when R is large): h[Rq] u64, len[Rq],
// lead[Rq], tab[T], then the stage of stage_symbol.
// ---------------------------------------------------------------------------------------------
struct Seen {
    uint32_t lo, hi, n;
};

// index of byte ch (non-zero) among the seen bytes, appending it if new; n == 9 signals overflow
__device__ __forceinline__ uint32_t seen_index(Seen& sn, uint32_t ch) {
    const uint32_t sp = ch * 0x01010101u;
    uint32_t x = sn.lo ^ sp;
    uint32_t z = (x - 0x01010101u) & ~x & 0x80808080u;  // lowest flagged byte is a true zero byte
    if (z) return ((uint32_t)__ffs((int)z) - 1u) >> 3;
    x = sn.hi ^ sp;
    z = (x - 0x01010101u) & ~x & 0x80808080u;
    if (z) return 4u + (((uint32_t)__ffs((int)z) - 1u) >> 3);
    const uint32_t a = sn.n;
    if (a < 4u)
        sn.lo |= ch << (8u * a);
    else if (a < 8u)
        sn.hi |= ch << (8u * (a - 4u));
    sn.n = a + 1u;
    return a;
}

__device__ __forceinline__ uint32_t seen_byte(const Seen& sn, uint32_t a) {
    return ((a < 4u ? sn.lo >> (8u * a) : sn.hi >> (8u * (a - 4u)))) & 0xffu;
}

// Lane-serial pass over one stash column: distinct residues in first-row order. False on overflow.
__device__ __forceinline__ bool narrow_scan(const uint8_t* col, uint32_t R, Seen& sn) {
    sn.lo = sn.hi = sn.n = 0;
    for (uint32_t r0 = 0; r0 < R; r0 += 4) {
        uint32_t word = *reinterpret_cast<const uint32_t*>(col + r0);
        const uint32_t lim = min(4u, R - r0);
        for (uint32_t i = 0; i < lim; ++i) {
            const uint32_t ch = word & 0xffu;
            word >>= 8;
            if (ch == 0u) return false;
            seen_index(sn, ch);
            if (sn.n > 8u) return false;
        }
    }
    return true;
}

struct GroupScratch {
    unsigned long long* hrow;
    uint32_t* len;
    uint32_t* lead;
    uint32_t* tab;
    uint8_t* stage;
    uint16_t* cdesc;
    uint32_t T, stage_bytes;
};

// CTA = false: one warp groups the symbol (many symbols in flight). CTA = true: the whole block does — the symbols that
// reach this path at R = 1000 are a few hundred wide ones, fewer than there are blocks, and a warp alone needs ~0.5 ms
// for the widest (every lane walks R / 32 strings of up to kStageCols characters three times).
template <bool CTA>
__device__ void group_wide(const MsaGeom& g, const MsaBufs& b, const GroupScratch& sc, uint32_t k) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t tid = CTA ? threadIdx.x : lane, nthr = CTA ? blockDim.x : 32u;
    auto sync = [&]() {
        if (CTA) __syncthreads();
        else __syncwarp();
    };
    const uint32_t T = sc.T;
    unsigned long long* hrow = sc.hrow;
    uint32_t *len = sc.len, *lead = sc.lead, *tab = sc.tab;
    const uint32_t e = b.sym[k];
    const uint32_t s = e & kColMask, en = b.sym[k + 1] & kColMask;
    const uint32_t slot0 = first_slot(b, s);
    RowWalk w{b.vbits, b.refc, b.stash, g.Rp, sc.cdesc, sc.stage, s, false};
    if (CTA) {
        __shared__ uint32_t s_staged;
        __syncthreads();  // the previous symbol's readers of the scratch are done
        if (threadIdx.x < 32u) {
            const bool st = stage_symbol(g, b, s, en, slot0, sc.cdesc, sc.stage, sc.stage_bytes);
            if (lane == 0) s_staged = st ? 1u : 0u;
        }
        __syncthreads();
        w.staged = s_staged != 0u;
    } else {
        w.staged = stage_symbol(g, b, s, en, slot0, sc.cdesc, sc.stage, sc.stage_bytes);
    }
    const bool exact = (en - s) <= 7u && g.hash_mask == ~0ull;  // the key IS the string: no verification needed

    for (uint32_t i = tid; i < T; i += nthr) tab[i] = kEmptySlot;
    for (uint32_t r = tid; r < g.R; r += nthr) {
        uint32_t c = s, slot = slot0, n = 0;
        unsigned long long h;
        int ch;
        if (exact) {
            h = 0;
            while ((ch = next_char(w, r, c, en, slot)) >= 0) {
                h |= (unsigned long long)ch << (8u * n);
                ++n;
            }
            h |= (unsigned long long)n << 56;
        } else {
            h = 14695981039346656037ull;
            while ((ch = next_char(w, r, c, en, slot)) >= 0) {
                h = (h ^ (unsigned long long)ch) * 1099511628211ull;
                ++n;
            }
            h = (h ^ ((unsigned long long)n * 0x9e3779b97f4a7c15ull)) & g.hash_mask;
        }
        hrow[r] = h;
        len[r] = n;
    }
    sync();
    for (uint32_t r = tid; r < g.R; r += nthr) {
        const unsigned long long h = hrow[r];
        uint32_t slot = (uint32_t)((h ^ (h >> 32)) * 0x9e3779b1u >> 7) & (T - 1u);
        for (;;) {
            const uint32_t cur = atomicCAS(&tab[slot], kEmptySlot, r);
            if (cur == kEmptySlot) break;
            if (hrow[cur] == h) {
                atomicMin(&tab[slot], r);
                break;
            }
            slot = (slot + 1u) & (T - 1u);
        }
        lead[r] = slot;
    }
    sync();
    uint32_t collided = 0;
    for (uint32_t r = tid; r < g.R; r += nthr) {
        const uint32_t m = tab[lead[r]];
        if (!exact && m != r && (len[m] != len[r] || !rows_equal(w, m, r, s, en, slot0))) collided = 1;
        lead[r] = m;
    }
    if (CTA) collided = (uint32_t)__syncthreads_or((int)collided);
    else collided = __any_sync(0xffffffffu, collided);
    sync();
    if (collided) {
        // exact fallback: first earlier row with the same string
        for (uint32_t r = tid; r < g.R; r += nthr) {
            uint32_t m = r;
            for (uint32_t r2 = 0; r2 < r; ++r2) {
                if (len[r2] == len[r] && hrow[r2] == hrow[r] && rows_equal(w, r2, r, s, en, slot0)) {
                    m = r2;
                    break;
                }
            }
            lead[r] = m;
        }
        sync();
    }
    // number the alternatives by first row; tab[0..R) is reused as "alternative of leader row"
    uint32_t nalts = 0;
    unsigned long long lensum = 0;
    if (!CTA) {
        for (uint32_t r0 = 0; r0 < g.R; r0 += 32) {
            const uint32_t r = r0 + lane;
            const bool isl = r < g.R && lead[r] == r;
            const uint32_t mask = __ballot_sync(0xffffffffu, isl);
            if (isl) {
                tab[r] = nalts + (uint32_t)__popc(mask & lanemask_lt());
                lensum += len[r];
            }
            if (lane == 0) b.leadmask[(size_t)slot0 * (g.Rp >> 5) + (r0 >> 5)] = mask;
            nalts += (uint32_t)__popc(mask);
        }
        __syncwarp();
        for (uint32_t r = lane; r < g.R; r += 32) store_alt(g, b.altid, (size_t)slot0 * g.Rp + r, tab[lead[r]]);
        lensum = warp_sum(lensum);
    } else {
        // chunks of 32 rows over the warps: leaders per chunk -> exclusive offsets (kept behind the R entries of tab that
        // the numbering writes: T = 2 Rq >= R + R / 32) -> numbers
        __shared__ unsigned long long s_lensum;
        __shared__ uint32_t s_nalts;
        const uint32_t warp = threadIdx.x >> 5, nw = blockDim.x >> 5, nchunks = (g.R + 31u) >> 5;
        uint32_t* cnt = tab + (T >> 1);
        if (threadIdx.x == 0) s_lensum = 0ull;
        for (uint32_t ck = warp; ck < nchunks; ck += nw) {
            const uint32_t r = ck * 32u + lane;
            const uint32_t mask = __ballot_sync(0xffffffffu, r < g.R && lead[r] == r);
            if (lane == 0) {
                cnt[ck] = (uint32_t)__popc(mask);
                b.leadmask[(size_t)slot0 * (g.Rp >> 5) + ck] = mask;
            }
        }
        __syncthreads();
        if (warp == 0) {
            uint32_t carry = 0;
            for (uint32_t c0 = 0; c0 < nchunks; c0 += 32) {
                const uint32_t v = c0 + lane < nchunks ? cnt[c0 + lane] : 0u;
                const uint32_t inc = warp_inclusive_scan(v);
                if (c0 + lane < nchunks) cnt[c0 + lane] = carry + inc - v;
                carry += __shfl_sync(0xffffffffu, inc, 31);
            }
            if (lane == 0) s_nalts = carry;
        }
        __syncthreads();
        for (uint32_t ck = warp; ck < nchunks; ck += nw) {
            const uint32_t r = ck * 32u + lane;
            const bool isl = r < g.R && lead[r] == r;
            const uint32_t mask = __ballot_sync(0xffffffffu, isl);
            if (isl) {
                tab[r] = cnt[ck] + (uint32_t)__popc(mask & lanemask_lt());
                lensum += len[r];
            }
        }
        lensum = warp_sum(lensum);
        if (lane == 0 && lensum) atomicAdd(&s_lensum, lensum);
        __syncthreads();
        for (uint32_t r = threadIdx.x; r < g.R; r += blockDim.x) store_alt(g, b.altid, (size_t)slot0 * g.Rp + r, tab[lead[r]]);
        nalts = s_nalts;
        lensum = s_lensum;
    }
    if (tid == 0) {
        b.sym_nalts[k] = nalts;
        b.sym_edsz[k] = 2ull + lensum + (unsigned long long)(nalts - 1u);
    }
    sync();
}

__device__ __forceinline__ void list_append(uint32_t* list, uint32_t* counter, bool mine, uint32_t k) {
    const uint32_t m = __ballot_sync(0xffffffffu, mine);
    if (!m) return;
    uint32_t base = 0;
    if ((threadIdx.x & 31) == 0) base = atomicAdd(counter, (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (mine) list[base + (uint32_t)__popc(m & lanemask_lt())] = k;
}

// ---- single-column symbols when there are too many rows for the lane-per-symbol path: a WARP owns the
// symbol, lanes run over rows, the residue list is warp-uniform. A chunk's unseen residues are appended in
// lane (= row) order, which is the order of first rows.
__device__ __forceinline__ uint32_t seen_lookup(const Seen& sn, uint32_t ch) {  // 8 = not in the list
    const uint32_t sp = ch * 0x01010101u;
    uint32_t x = sn.lo ^ sp;
    uint32_t z = (x - 0x01010101u) & ~x & 0x80808080u;
    if (z) return ((uint32_t)__ffs((int)z) - 1u) >> 3;
    x = sn.hi ^ sp;
    z = (x - 0x01010101u) & ~x & 0x80808080u;
    if (z) return 4u + (((uint32_t)__ffs((int)z) - 1u) >> 3);
    return 8u;
}

// class of this lane's residue (valid lanes), growing the warp-uniform list; false on overflow / NUL byte
__device__ __forceinline__ bool warp_classify(Seen& sn, uint32_t ch, bool valid, uint32_t& cls) {
    cls = valid ? seen_lookup(sn, ch) : 0u;
    if (__any_sync(0xffffffffu, valid && ch == 0u)) return false;
    uint32_t pending = __ballot_sync(0xffffffffu, valid && cls == 8u);
    while (pending) {
        const uint32_t leader = (uint32_t)__ffs((int)pending) - 1u;
        const uint32_t nch = __shfl_sync(0xffffffffu, ch, (int)leader);
        if (sn.n >= 8u) return false;
        const uint32_t a = seen_index(sn, nch);  // appends: same in every lane
        const bool mine = valid && cls == 8u && ch == nch;
        if (mine) cls = a;
        pending &= ~__ballot_sync(0xffffffffu, mine);
    }
    return true;
}

// byte i (0..31) of the 32 bytes held in v[0..8)
__device__ __forceinline__ uint32_t byte_of32(const uint32_t (&v)[8], uint32_t i) {
    uint32_t w = v[0];
#pragma unroll
    for (uint32_t q = 1; q < 8u; ++q) w = (i >> 2) == q ? v[q] : w;
    return (w >> ((i & 3u) * 8u)) & 0xffu;
}

// bit i = (byte i of the 32 bytes in v == ch)
__device__ __forceinline__ uint32_t eq_mask32(const uint32_t (&v)[8], uint32_t ch) {
    const uint32_t sp = ch * 0x01010101u;
    uint32_t m = 0;
#pragma unroll
    for (uint32_t q = 0; q < 8u; ++q) m |= eq_bytes4(v[q], sp) << (4u * q);
    return m;
}

// Single-column symbols, rows across lanes: a LANE owns 32 consecutive rows — one word of every alternative's row
// bitset — and a GROUP of GL lanes (GL = R/32 rounded up to a power of two, at most 32) owns a symbol, so a warp
// handles 32 / GL symbols at once (100 rows: 8 symbols per warp; 1000 rows: one). A lane peels its distinct residues
// off in order of first row (one SIMD byte compare over its 32 bytes per distinct residue: the compare IS the bitset
// word), the group merges its lanes' lists in lane order (= row order) into the symbol's residue list, and every lane
// stores its words under the merged numbering (coalesced). Every collective is executed by all 32 lanes (groups never
// diverge around one). ok = false: more than 8 residues or a NUL byte -> the symbol goes to the hashed row path.
template <uint32_t GL>
__device__ void group_single_lanes(const MsaGeom& g, const MsaBufs& b, bool active, uint32_t k, uint32_t s, uint32_t& nalts, bool& ok_out) {
    const uint32_t lane = threadIdx.x & 31, gl = lane & (GL - 1u), gbase = lane & ~(GL - 1u);
    const uint32_t gmask = GL >= 32u ? 0xffffffffu : (((1u << GL) - 1u) << gbase);
    const uint32_t Rw = g.Rp >> 5;
    const uint32_t slot0 = active ? first_slot(b, s) : 0u;
    const uint8_t* col = b.stash + (size_t)slot0 * g.Rp;
    uint32_t* rowbits = b.rowbits + (size_t)slot0 * 8u * Rw;
    Seen G;
    G.lo = G.hi = G.n = 0;
    bool ok = active;
    if (Rw > GL) {
        // deeper than 1024 rows: a residue first seen in a later chunk has no words for the earlier chunks
        if (active)
            for (uint32_t i = gl; i < 8u * Rw; i += GL) rowbits[i] = 0u;
        __syncwarp();
    }
    for (uint32_t w0 = 0; w0 < Rw; w0 += GL) {
        const uint32_t w = w0 + gl;
        const bool have = ok && w < Rw;
        uint32_t v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (have) {
            const uint4 a = *reinterpret_cast<const uint4*>(col + (size_t)w * 32u), c = *reinterpret_cast<const uint4*>(col + (size_t)w * 32u + 16u);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
            v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
        }
        const uint32_t nrows = have ? min(32u, g.R - w * 32u) : 0u;
        uint32_t remaining = low_bits(nrows);
        Seen L;
        L.lo = L.hi = L.n = 0;
        uint32_t m[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        bool lane_ok = true;
#pragma unroll
        for (uint32_t j = 0; j < 8u; ++j) {
            if (remaining) {
                const uint32_t ch = byte_of32(v, (uint32_t)__ffs((int)remaining) - 1u);
                if (ch == 0u) lane_ok = false;
                m[j] = eq_mask32(v, ch) & remaining;
                remaining &= ~m[j];
                if (j < 4u) L.lo |= ch << (8u * j); else L.hi |= ch << (8u * (j - 4u));
                L.n = j + 1u;
            }
        }
        if (remaining) lane_ok = false;  // a ninth residue
        if (__ballot_sync(0xffffffffu, !lane_ok) & gmask) ok = false;
        // merge in lane order: the lowest lane of the group that holds a residue the list does not know appends its news
        for (;;) {
            bool news = false;
            if (ok && have)
                for (uint32_t j = 0; j < L.n; ++j) news = news || seen_lookup(G, seen_byte(L, j)) == 8u;
            const uint32_t pending_all = __ballot_sync(0xffffffffu, news);
            if (!pending_all) break;
            const uint32_t pending = pending_all & gmask;
            const int src = pending ? __ffs((int)pending) - 1 : (int)lane;
            Seen src_list;
            src_list.lo = __shfl_sync(0xffffffffu, L.lo, src);
            src_list.hi = __shfl_sync(0xffffffffu, L.hi, src);
            src_list.n = __shfl_sync(0xffffffffu, L.n, src);
            if (pending)
                for (uint32_t j = 0; j < src_list.n; ++j) {
                    const uint32_t ch = seen_byte(src_list, j);
                    if (seen_lookup(G, ch) == 8u) {
                        if (G.n >= 8u) ok = false; else seen_index(G, ch);
                    }
                }
        }
        if (have && ok) {
            // words of this lane under the merged numbering
#pragma unroll
            for (uint32_t a = 0; a < 8u; ++a) {
                if (a < G.n) {
                    const uint32_t j = seen_lookup(L, seen_byte(G, a));
                    uint32_t word = 0;
#pragma unroll
                    for (uint32_t q = 0; q < 8u; ++q) word = j == q ? m[q] : word;
                    rowbits[a * Rw + w] = word;
                }
            }
        }
    }
    if (active && gl == 0u) {
        if (!ok) {
            b.seen[(size_t)slot0 * 3u + 2u] = 0xffu;  // on the hashed row path: k_emit_var skips it
        } else {
            b.seen[(size_t)slot0 * 3u] = G.lo;
            b.seen[(size_t)slot0 * 3u + 1u] = G.hi;
            b.seen[(size_t)slot0 * 3u + 2u] = G.n;
            uint32_t chars = 0;
            for (uint32_t a = 0; a < G.n; ++a) chars += seen_byte(G, a) != (uint32_t)'-';
            b.sym_nalts[k] = G.n;
            b.sym_edsz[k] = 2ull + chars + (G.n - 1u);
        }
    }
    nalts = ok ? G.n : 0u;
    ok_out = ok;
}

// k_group: single-column symbols — lane per symbol (registers only) when narrow_ok, else warp per symbol with
// rows across lanes; everything else is queued for k_group2 (two variable columns already give up to 25
// alternatives, more than the 8-entry residue list holds).
__global__ void k_group(MsaGeom g, MsaBufs b, uint32_t narrow_ok, uint32_t group_lanes) {
    MsaStatus* st = b.status;
    if (st->abort || st->halo_fail) return;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const uint32_t v_lo = st->v_lo, v_hi = st->v_hi;
    unsigned long long alts_here = 0;
    if (narrow_ok != 1u) {
        // rows across lanes: a group of GL lanes per symbol, 32 / GL symbols per warp, batches strided over the grid
        const uint32_t GL = group_lanes, NS = 32u / GL, sub = lane / GL;
        for (uint32_t v0 = v_lo + (blockIdx.x * wpb + warp) * NS; v0 < v_hi; v0 += gridDim.x * wpb * NS) {
            const uint32_t v = v0 + sub;
            bool active = v < v_hi && narrow_ok != 2u;  // narrow_ok 2: tests send every symbol down the hashed path
            uint32_t k = 0, s = 0;
            if (active) {
                k = b.varsym[v];
                s = b.sym[k] & kColMask;
                active = (b.sym[k + 1] & kColMask) - s == 1u;  // multi-column symbols: k_group3 takes them
            }
            uint32_t na = 0;
            bool ok = false;
            switch (GL) {  // (warp-uniform)
                case 1: group_single_lanes<1>(g, b, active, k, s, na, ok); break;
                case 2: group_single_lanes<2>(g, b, active, k, s, na, ok); break;
                case 4: group_single_lanes<4>(g, b, active, k, s, na, ok); break;
                case 8: group_single_lanes<8>(g, b, active, k, s, na, ok); break;
                case 16: group_single_lanes<16>(g, b, active, k, s, na, ok); break;
                default: group_single_lanes<32>(g, b, active, k, s, na, ok); break;
            }
            if (active && (lane & (GL - 1u)) == 0u) {
                if (ok) {
                    alts_here += na;
                } else {  // more than 8 residues, or a NUL byte: hashed row path
                    b.hardlist[atomicAdd(&st->n_hard, 1u)] = k;
                    b.emitlist[atomicAdd(&st->n_emit2, 1u)] = k;
                }
            }
        }
    } else
    for (uint32_t v0 = v_lo + (blockIdx.x * wpb + warp) * 32u; v0 < v_hi; v0 += gridDim.x * wpb * 32u) {
        const uint32_t v = v0 + lane;
        const bool have = v < v_hi;
        const uint32_t k = have ? b.varsym[v] : 0u;
        uint32_t s = 0, en = 0;
        if (have) {
            s = b.sym[k] & kColMask;
            en = b.sym[k + 1] & kColMask;
        }
        uint32_t cls = !have ? 0u : (en - s == 1u ? 1u : 0u);  // multi-column symbols: k_group3 takes them
        if (cls == 1u) {
            Seen sn;
            if (narrow_scan(b.stash + (size_t)first_slot(b, s) * g.Rp, g.R, sn)) {
                // alternative = the residue, or the empty string for '-'
                uint32_t chars = 0;
                for (uint32_t a = 0; a < sn.n; ++a) chars += seen_byte(sn, a) != (uint32_t)'-';
                b.sym_nalts[k] = sn.n;
                b.sym_edsz[k] = 2ull + chars + (sn.n - 1u);
                alts_here += sn.n;
            } else {
                cls = 3u;
            }
        }
        list_append(b.hardlist, &st->n_hard, cls == 3u, k);
        list_append(b.emitlist, &st->n_emit2, cls == 3u, k);
    }
    alts_here = warp_sum(alts_here);
    if (lane == 0 && alts_here) atomicAdd(&st->n_alts, alts_here);
    if (blockIdx.x == 0 && threadIdx.x == 0) st->n_var_syms = v_hi - v_lo;
}

// k_group2: the queued symbols, warp per symbol, evenly strided over the list.
__global__ void k_group2(MsaGeom g, MsaBufs b, uint32_t Rq, uint32_t T, uint32_t use_global, uint32_t stage_bytes,
                         uint32_t per_warp_smem, uint32_t cta_ok) {
    MsaStatus* st = b.status;
    if (st->abort || st->halo_fail) return;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    unsigned char* smem = EDSB_DYN_SMEM() + (size_t)warp * per_warp_smem;
    unsigned char* base = use_global ? b.group_ws + ((size_t)blockIdx.x * wpb + warp) * ((size_t)Rq * 16u + (size_t)T * 4u) : smem;
    GroupScratch sc;
    sc.hrow = reinterpret_cast<unsigned long long*>(base);
    sc.len = reinterpret_cast<uint32_t*>(sc.hrow + Rq);
    sc.lead = sc.len + Rq;
    sc.tab = sc.lead + Rq;
    sc.stage = reinterpret_cast<uint8_t*>(sc.tab + T);  // only when the scratch is in shared memory
    sc.cdesc = reinterpret_cast<uint16_t*>(sc.stage + stage_bytes);
    sc.T = T;
    sc.stage_bytes = use_global ? 0u : stage_bytes;
    const uint32_t n_wide = st->n_hard;
    unsigned long long alts_here = 0;
    if (cta_ok == 2u || (cta_ok && n_wide <= 2u * gridDim.x)) {
        // few symbols: a block each, in the first warp's scratch
        unsigned char* base0 = use_global ? b.group_ws + ((size_t)blockIdx.x * wpb) * ((size_t)Rq * 16u + (size_t)T * 4u) : EDSB_DYN_SMEM();
        sc.hrow = reinterpret_cast<unsigned long long*>(base0);
        sc.len = reinterpret_cast<uint32_t*>(sc.hrow + Rq);
        sc.lead = sc.len + Rq;
        sc.tab = sc.lead + Rq;
        sc.stage = reinterpret_cast<uint8_t*>(sc.tab + T);
        sc.cdesc = reinterpret_cast<uint16_t*>(sc.stage + stage_bytes);
        for (uint32_t item = blockIdx.x; item < n_wide; item += gridDim.x) {
            const uint32_t kw = b.hardlist[item];
            group_wide<true>(g, b, sc, kw);
            if (threadIdx.x == 0) alts_here += b.sym_nalts[kw];
        }
        if (threadIdx.x == 0 && alts_here) atomicAdd(&st->n_alts, alts_here);
        return;
    }
    for (uint32_t item = blockIdx.x * wpb + warp; item < n_wide; item += gridDim.x * wpb) {
        const uint32_t kw = b.hardlist[item];
        group_wide<false>(g, b, sc, kw);
        if (lane == 0) alts_here += b.sym_nalts[kw];
    }
    if (lane == 0 && alts_here) atomicAdd(&st->n_alts, alts_here);
}


// ---------------------------------------------------------------------------------------------
// Multi-column symbols, tuple formulation (k_group3 / k_emit3). The conserved columns inside a variable symbol are
// the same in every row, so a row's string is a function of its residues at the symbol's nv VARIABLE columns only.
//   1. key(row) = those nv bytes (nv <= 7: 56 bits), read as coalesced words from the stash: ~20 instructions a row
//      instead of a walk over every column of the symbol;
//   2. the distinct keys and the first row of each go through a 128-slot table in shared memory;
//   3. only the distinct keys (a few dozen) are turned into gap-stripped strings, compared exactly (strings of up
//      to 7 characters are their own 64-bit key; longer ones hash first and are verified) and numbered by first row;
//   4. emit: rows -> per-alternative row bitsets in shared memory (one __match_any_sync per 32 rows), each rendered
//      as "{ids}" through idlist.cuh into a staged, 16-byte aligned copy-out.
// Scratch per warp is ~3.6 KB + R bytes (the hashed row path needs 28 KB at 1000 rows), so an SM holds 32 of these
// warps instead of 6. Symbols outside the envelope (more than 7 variable or 64 total columns, more than 128
// distinct keys) go to `hardlist` and take the hashed row path (k_group2 / k_emit2).
// ---------------------------------------------------------------------------------------------
constexpr uint32_t kTS = 256;          // tuple table slots (open addressing, at most half full)
constexpr uint32_t kTD = 128;          // most distinct tuples of a symbol
constexpr uint32_t kTupleBitWords = 1024;  // k_emit3: row-bitset words per warp (alternatives x R/32 per batch)

struct TupleScratch {
    uint16_t* cdesc;            // [64]
    unsigned long long* tkey;   // [kTS] 0 = empty
    unsigned long long* thash;  // [kTS] string key of the tuple
    uint32_t* trow;             // [kTS] first row
    uint16_t* tlen;             // [kTS]
    uint16_t* tlead;            // [kTS] slot of the tuple whose first row introduces this tuple's string
    uint16_t* talt;             // [kTS]
    uint8_t* dl;                // [kTS] occupied slots
    uint32_t* ndist;            // distinct tuples so far
    uint8_t* rslot;             // [R] slot of each row's tuple
};

inline size_t tuple_group_smem(uint32_t R) { return 128 + kTS * (8 + 8 + 4 + 2 + 2 + 2 + 1) + 16 + ((R + 15u) & ~15u) + 16; }

__device__ __forceinline__ TupleScratch tuple_scratch(unsigned char* base) {
    TupleScratch t;
    t.cdesc = reinterpret_cast<uint16_t*>(base);
    t.tkey = reinterpret_cast<unsigned long long*>(base + 128);
    t.thash = t.tkey + kTS;
    t.trow = reinterpret_cast<uint32_t*>(t.thash + kTS);
    t.tlen = reinterpret_cast<uint16_t*>(t.trow + kTS);
    t.tlead = t.tlen + kTS;
    t.talt = t.tlead + kTS;
    t.dl = reinterpret_cast<uint8_t*>(t.talt + kTS);
    t.ndist = reinterpret_cast<uint32_t*>(t.dl + kTS);
    t.rslot = reinterpret_cast<uint8_t*>(t.ndist + 4);
    return t;
}

// columns [s, en) of a symbol: cdesc[i] = 0x100 | index of the variable column inside the symbol, or the conserved
// character. Returns the number of variable columns (warp-uniform).
__device__ __forceinline__ uint32_t describe_columns(const MsaBufs& b, uint32_t s, uint32_t en, uint16_t* cdesc) {
    const uint32_t lane = threadIdx.x & 31, width = en - s;
    uint32_t nv = 0;
    for (uint32_t i0 = 0; i0 < width; i0 += 32) {
        const uint32_t i = i0 + lane, c = s + i;
        const bool in = i < width;
        const uint32_t v = in ? (b.vbits[c >> 5] >> (c & 31u)) & 1u : 0u;
        const uint32_t m = __ballot_sync(0xffffffffu, v);
        if (in) cdesc[i] = v ? (uint16_t)(0x100u | (nv + (uint32_t)__popc(m & lanemask_lt()))) : (uint16_t)b.refc[c];
        nv += (uint32_t)__popc(m);
    }
    __syncwarp();
    return nv;
}

// gap-stripped string of a tuple: length, and its key (the string itself below 8 characters, else a tagged hash)
__device__ __forceinline__ void tuple_string_key(const uint16_t* cdesc, uint32_t width, unsigned long long key, uint64_t hash_mask,
                                                 unsigned long long& out, uint32_t& len) {
    unsigned long long pk = 0, h = 14695981039346656037ull;
    uint32_t n = 0;
    for (uint32_t i = 0; i < width; ++i) {
        const uint32_t d = cdesc[i];
        const uint32_t ch = (d & 0x100u) ? (uint32_t)(key >> (8u * (d & 0xffu))) & 0xffu : d;
        if (ch == (uint32_t)'-') continue;
        if (n < 7u) pk |= (unsigned long long)ch << (8u * n);
        h = (h ^ (unsigned long long)ch) * 1099511628211ull;
        ++n;
    }
    len = n;
    const bool exact = n <= 7u && hash_mask == ~0ull;
    out = exact ? (pk | ((unsigned long long)n << 56)) : ((h & hash_mask) | (1ull << 63));
}

__device__ bool tuple_strings_equal(const uint16_t* cdesc, uint32_t width, unsigned long long ka, unsigned long long kb) {
    uint32_t ia = 0, ib = 0;
    for (;;) {
        int ca = -1, cb = -1;
        while (ia < width) {
            const uint32_t d = cdesc[ia++];
            const uint32_t ch = (d & 0x100u) ? (uint32_t)(ka >> (8u * (d & 0xffu))) & 0xffu : d;
            if (ch != (uint32_t)'-') { ca = (int)ch; break; }
        }
        while (ib < width) {
            const uint32_t d = cdesc[ib++];
            const uint32_t ch = (d & 0x100u) ? (uint32_t)(kb >> (8u * (d & 0xffu))) & 0xffu : d;
            if (ch != (uint32_t)'-') { cb = (int)ch; break; }
        }
        if (ca != cb) return false;
        if (ca < 0) return true;
    }
}

// One warp, one symbol. False (warp-uniform): outside the envelope, nothing was written.
__device__ bool group_tuple(const MsaGeom& g, const MsaBufs& b, const TupleScratch& t, uint32_t k, uint32_t& nalts_out) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t s = b.sym[k] & kColMask, en = b.sym[k + 1] & kColMask, width = en - s;
    if (width > kStageCols || (g.Rp >> 5) > kTupleBitWords || g.alt32) return false;
    const uint32_t slot0 = first_slot(b, s);
    const uint32_t nv = describe_columns(b, s, en, t.cdesc);
    if (nv < 2u || nv > 7u) return false;  // (one variable column: the symbol owns a single stash / altid slot)
    for (uint32_t i = lane; i < kTS; i += 32) {
        t.tkey[i] = 0ull;
        t.trow[i] = 0xffffffffu;
    }
    if (lane == 0) *t.ndist = 0u;
    __syncwarp();
    bool over = false;
    const uint8_t* col0 = b.stash + (size_t)slot0 * g.Rp;
    for (uint32_t r0 = 4u * lane; r0 < g.R; r0 += 128u) {
        uint32_t w[7];
#pragma unroll
        for (uint32_t i = 0; i < 7u; ++i) w[i] = i < nv ? *reinterpret_cast<const uint32_t*>(col0 + (size_t)i * g.Rp + r0) : 0u;
#pragma unroll
        for (uint32_t q = 0; q < 4u; ++q) {
            const uint32_t r = r0 + q;
            if (r >= g.R) break;
            unsigned long long key = 1ull << 63;
#pragma unroll
            for (uint32_t i = 0; i < 7u; ++i) key |= (unsigned long long)((w[i] >> (8u * q)) & 0xffu) << (8u * i);
            uint32_t h = (uint32_t)((key ^ (key >> 29)) * 0x9e3779b1u >> 11) & (kTS - 1u);
            uint32_t probes = 0;
            for (;;) {
                const unsigned long long old = atomicCAS(&t.tkey[h], 0ull, key);
                if (old == key) break;
                if (old == 0ull) {  // a new tuple: the table stays at most half full (short probe chains)
                    if (atomicAdd(t.ndist, 1u) >= kTD) over = true;
                    break;
                }
                h = (h + 1u) & (kTS - 1u);
                if (++probes >= kTS) {
                    over = true;
                    break;
                }
            }
            if (over) break;
            atomicMin(&t.trow[h], r);
            t.rslot[r] = (uint8_t)h;
        }
        if (*reinterpret_cast<volatile uint32_t*>(t.ndist) > kTD) over = true;  // some lane ran out of room: stop early
    }
    if (__any_sync(0xffffffffu, over)) return false;
    __syncwarp();
    // occupied slots
    uint32_t D = 0;
    for (uint32_t i0 = 0; i0 < kTS; i0 += 32) {
        const bool occ = t.tkey[i0 + lane] != 0ull;
        const uint32_t m = __ballot_sync(0xffffffffu, occ);
        if (occ) t.dl[D + (uint32_t)__popc(m & lanemask_lt())] = (uint8_t)(i0 + lane);
        D += (uint32_t)__popc(m);
    }
    __syncwarp();
    for (uint32_t q = lane; q < D; q += 32) {
        const uint32_t slot = t.dl[q];
        unsigned long long sk;
        uint32_t n;
        tuple_string_key(t.cdesc, width, t.tkey[slot], g.hash_mask, sk, n);
        t.thash[slot] = sk;
        t.tlen[slot] = (uint16_t)n;
    }
    __syncwarp();
    // the tuple with the smallest first row among those that spell the same string introduces the alternative
    for (uint32_t q = lane; q < D; q += 32) {
        const uint32_t slot = t.dl[q];
        const unsigned long long mine = t.thash[slot], key = t.tkey[slot];
        const uint32_t len = t.tlen[slot];
        uint32_t best = t.trow[slot], best_slot = slot;
        for (uint32_t u = 0; u < D; ++u) {
            const uint32_t us = t.dl[u];
            if (us == slot || t.thash[us] != mine || t.tlen[us] != len) continue;
            if ((mine >> 63) && !tuple_strings_equal(t.cdesc, width, key, t.tkey[us])) continue;
            if (t.trow[us] < best) {
                best = t.trow[us];
                best_slot = us;
            }
        }
        t.tlead[slot] = (uint16_t)best_slot;
    }
    __syncwarp();
    uint32_t nalts = 0, lensum = 0;
    for (uint32_t q0 = 0; q0 < D; q0 += 32) {
        const uint32_t q = q0 + lane;
        const uint32_t slot = q < D ? t.dl[q] : 0u;
        const bool leader = q < D && t.tlead[slot] == slot;
        if (leader) {
            uint32_t before = 0;
            const uint32_t row = t.trow[slot];
            for (uint32_t u = 0; u < D; ++u) {
                const uint32_t us = t.dl[u];
                before += (t.tlead[us] == us && t.trow[us] < row) ? 1u : 0u;
            }
            t.talt[slot] = (uint16_t)before;
            lensum += t.tlen[slot];
            store_alt(g, b.altid, (size_t)(slot0 + 1u) * g.Rp + before, row);  // first row of alternative `before`
        }
        nalts += (uint32_t)__popc(__ballot_sync(0xffffffffu, leader));
    }
    __syncwarp();
    for (uint32_t q = lane; q < D; q += 32) {
        const uint32_t slot = t.dl[q];
        if (t.tlead[slot] != slot) t.talt[slot] = t.talt[t.tlead[slot]];
    }
    __syncwarp();
    // rows that introduce an alternative (emit_wide reads them when rows are few): a row leads iff it is the first
    // row of its tuple and that tuple leads its string
    for (uint32_t r0 = 0; r0 < g.R; r0 += 32) {
        const uint32_t r = r0 + lane;
        uint32_t a = 0;
        bool leads = false;
        if (r < g.R) {
            const uint32_t slot = t.rslot[r];
            a = t.talt[slot];
            leads = t.tlead[slot] == slot && t.trow[slot] == r;
            store_alt(g, b.altid, (size_t)slot0 * g.Rp + r, a);
        }
        const uint32_t m = __ballot_sync(0xffffffffu, leads);
        if (lane == 0) b.leadmask[(size_t)slot0 * (g.Rp >> 5) + (r0 >> 5)] = m;
    }
    lensum = warp_sum(lensum);
    if (lane == 0) {
        b.sym_nalts[k] = nalts;
        b.sym_edsz[k] = 2ull + lensum + (unsigned long long)(nalts - 1u);
    }
    nalts_out = nalts;
    __syncwarp();
    return true;
}

// k_widelist: the owned multi-column symbols (all_symbols, a test switch: every owned variable symbol) -> widelist.
__global__ void k_widelist(MsaGeom g, MsaBufs b, uint32_t all_symbols) {
    MsaStatus* st = b.status;
    if (st->abort || st->halo_fail) return;
    const uint32_t v_lo = st->v_lo, v_hi = st->v_hi;
    for (uint32_t v0 = v_lo + ((blockIdx.x * blockDim.x + threadIdx.x) & ~31u); v0 < v_hi; v0 += gridDim.x * blockDim.x) {
        const uint32_t v = v0 + (threadIdx.x & 31);
        const uint32_t k = v < v_hi ? b.varsym[v] : 0u;
        const bool wide = v < v_hi && (all_symbols || (b.sym[k + 1] & kColMask) - (b.sym[k] & kColMask) > 1u);
        list_append(b.widelist, &st->n_wide, wide, k);
    }
}

// k_group3: the multi-column symbols of widelist, warp per symbol, on its own stream beside k_group; what the tuple
// formulation takes goes to easylist (k_emit3) or, with few rows, emitlist; the rest to hardlist (k_group2 / k_emit2).
__global__ void k_group3(MsaGeom g, MsaBufs b, uint32_t per_warp_smem, uint32_t force_hard, uint32_t emit_by_rows) {
    MsaStatus* st = b.status;
    if (st->abort || st->halo_fail) return;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const TupleScratch t = tuple_scratch(EDSB_DYN_SMEM() + (size_t)warp * per_warp_smem);
    const uint32_t n_wide = st->n_wide;
    unsigned long long alts_here = 0;
    {
      for (uint32_t item = blockIdx.x * wpb + warp; item < n_wide; item += gridDim.x * wpb) {
        const uint32_t kw = b.widelist[item];
        uint32_t na = 0;
        const bool ok = !force_hard && group_tuple(g, b, t, kw, na);
        if (lane == 0) {
            if (ok) {
                // emit_by_rows (few rows): emit_wide renders it from altid / leadmask, like the symbols of hardlist
                if (emit_by_rows) b.emitlist[atomicAdd(&st->n_emit2, 1u)] = kw;
                else b.easylist[atomicAdd(&st->n_easy, 1u)] = kw;
                alts_here += na;
            } else {
                b.hardlist[atomicAdd(&st->n_hard, 1u)] = kw;
                b.emitlist[atomicAdd(&st->n_emit2, 1u)] = kw;
            }
        }
        __syncwarp();
      }
    }
    if (lane == 0 && alts_here) atomicAdd(&st->n_alts, alts_here);
}

// k_emit3: "{alt,...}" and "{ids}" per alternative for the symbols of easylist.
// Scratch per warp: cdesc[64], row bitsets [kTupleBitWords], the id-list stage.
__global__ void k_emit3(MsaGeom g, MsaBufs b, uint32_t per_warp_smem) {
    const MsaStatus* st = b.status;
    if (st->abort || st->halo_fail) return;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    unsigned char* base = EDSB_DYN_SMEM() + (size_t)warp * per_warp_smem;
    uint16_t* cdesc = reinterpret_cast<uint16_t*>(base);
    uint32_t* bits = reinterpret_cast<uint32_t*>(base + 128);
    uint8_t* stage = base + 128 + kTupleBitWords * 4u;
    const uint32_t Rw = g.Rp >> 5;
    const uint32_t n_easy = st->n_easy;
    for (uint32_t item = blockIdx.x * wpb + warp; item < n_easy; item += gridDim.x * wpb) {
        const uint32_t k = b.easylist[item];
        const uint32_t s = b.sym[k] & kColMask, en = b.sym[k + 1] & kColMask, width = en - s;
        const uint32_t slot0 = first_slot(b, s);
        const uint32_t A = b.sym_nalts[k];
        describe_columns(b, s, en, cdesc);
        const uint8_t* col0 = b.stash + (size_t)slot0 * g.Rp;
        // ---- EDS text: the alternatives in order of first row
        uint8_t* eds = b.eds_out + b.eds_off[k];
        unsigned long long run = 1;
        for (uint32_t a0 = 0; a0 < A; a0 += 32) {
            const uint32_t a = a0 + lane;
            const bool have = a < A;
            const uint32_t row = have ? load_alt(g, b.altid, (size_t)(slot0 + 1u) * g.Rp + a) : 0u;
            uint32_t n = 0;
            if (have)
                for (uint32_t i = 0; i < width; ++i) {
                    const uint32_t d = cdesc[i];
                    const uint32_t ch = (d & 0x100u) ? (uint32_t)col0[(size_t)(d & 0xffu) * g.Rp + row] : d;
                    n += ch != (uint32_t)'-';
                }
            const unsigned long long contrib = have ? (unsigned long long)n + 1ull : 0ull;
            const unsigned long long inc = warp_inclusive_scan(contrib);
            if (have) {
                unsigned long long at = run + inc - contrib;
                eds[at - 1] = a == 0 ? '{' : ',';
                for (uint32_t i = 0; i < width; ++i) {
                    const uint32_t d = cdesc[i];
                    const uint32_t ch = (d & 0x100u) ? (uint32_t)col0[(size_t)(d & 0xffu) * g.Rp + row] : d;
                    if (ch != (uint32_t)'-') eds[at++] = (uint8_t)ch;
                }
            }
            run += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) eds[run - 1] = '}';
        // ---- SEDS: row bitsets of a batch of alternatives, then one rendered list each
        unsigned long long at = b.seds_off[k];
        const uint32_t batch = max(1u, min(A, kTupleBitWords / Rw));
        for (uint32_t a_lo = 0; a_lo < A; a_lo += batch) {
            const uint32_t nb = min(batch, A - a_lo);
            for (uint32_t i = lane; i < nb * Rw; i += 32) bits[i] = 0u;
            __syncwarp();
            for (uint32_t r0 = 0; r0 < g.R; r0 += 32) {
                const uint32_t r = r0 + lane;
                const uint32_t a = r < g.R ? load_alt(g, b.altid, (size_t)slot0 * g.Rp + r) : kEmptySlot;
                const bool inb = a >= a_lo && a < a_lo + nb;
                const uint32_t peers = __match_any_sync(0xffffffffu, inb ? a : kEmptySlot);
                if (inb && lane == (uint32_t)__ffs((int)peers) - 1u) bits[(a - a_lo) * Rw + (r0 >> 5)] = peers;
            }
            __syncwarp();
            // sizes: a lane per alternative walks its words (ids, bytes); offsets by a warp scan.
            // Then the alternatives with few rows (most of them, in a multi-column symbol) are written by their lane,
            // the crowded ones one after the other by the whole warp through the staged renderer.
            for (uint32_t j0 = 0; j0 < nb; j0 += 32) {
                const uint32_t j = j0 + lane;
                uint32_t ids = 0, bytes = 0;
                if (j < nb) {
                    for (uint32_t w = 0; w < Rw; ++w) {
                        const uint32_t v = bits[j * Rw + w];
                        if (v) {
                            ids += (uint32_t)__popc(v);
                            bytes += word_id_bytes(w, v);
                        }
                    }
                    bytes += 1u;  // '{' + "id," each, the last ',' is the '}'
                }
                const uint32_t incl = warp_inclusive_scan(bytes);
                const unsigned long long mine = at + (incl - bytes);
                const bool crowded = j < nb && ids > 24u;
                if (j < nb && !crowded) {
                    uint8_t* out = b.seds_out + mine;
                    *out++ = '{';
                    for (uint32_t w = 0; w < Rw; ++w)
                        for (uint32_t v = bits[j * Rw + w]; v; v &= v - 1u) {
                            const uint32_t id = w * 32u + (uint32_t)__ffs((int)v);
                            const unsigned long long e = __ldg(b.id_text + id);  // digits, then ','; width in the top byte
                            const uint32_t dw = (uint32_t)(e >> 56);             // (ids here are below 2^16: never 0)
                            for (uint32_t q = 0; q <= dw; ++q) out[q] = (uint8_t)(e >> (8u * q));
                            out += dw + 1u;
                        }
                    out[-1] = '}';
                }
                uint32_t todo = __ballot_sync(0xffffffffu, crowded);
                while (todo) {
                    const int src = __ffs((int)todo) - 1;
                    todo &= todo - 1u;
                    const unsigned long long so = __shfl_sync(0xffffffffu, mine, src);
                    const uint32_t sb = __shfl_sync(0xffffffffu, bytes, src);
                    warp_render_id_list(stage, bits + (j0 + (uint32_t)src) * Rw, Rw, b.id_text, b.seds_out, so, sb, 0xffffffffu);
                }
                at += __shfl_sync(0xffffffffu, incl, 31);
            }
            __syncwarp();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Output sizes -> offsets. Symbol k emits (eds, seds) bytes:
//   common   : '{' + its own columns (+ '}' when it ends inside the owned range), "{0}"
//   variable : 2 + sum(len) + (nalts - 1),  nalts + sum_id_width + R
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void symbol_sizes(const MsaGeom& g, const MsaBufs& b, const MsaStatus* st, uint32_t k,
                                             unsigned long long& eds, unsigned long long& seds) {
    eds = 0;
    seds = 0;
    if (k < st->k_lo || k >= st->k_hi) return;
    const uint32_t e = b.sym[k];
    const uint32_t s = e & kColMask, en = b.sym[k + 1] & kColMask;
    if (e & kCommonFlag) {
        const uint32_t endc = min(en, g.own_hi);
        eds = 1ull + (endc - s) + (en <= g.own_hi ? 1u : 0u);
        seds = 3;
    } else {
        eds = b.sym_edsz[k];
        seds = (unsigned long long)b.sym_nalts[k] + g.sum_id_width + g.R;
    }
}

__global__ void __launch_bounds__(kPartThreads) k_size_count(MsaGeom g, MsaBufs b) {
    __shared__ unsigned long long s_red[33];
    const MsaStatus* st = b.status;
    if (st->abort || st->halo_fail) return;
    const uint32_t n = st->n_syms, P = gridDim.x;
    const uint32_t per = (n + P - 1) / P;
    const uint32_t k_begin = min(n, blockIdx.x * per), k_end = min(n, k_begin + per);
    unsigned long long se = 0, ss = 0;
    for (uint32_t k = k_begin + threadIdx.x; k < k_end; k += blockDim.x) {
        unsigned long long a, c;
        symbol_sizes(g, b, st, k, a, c);
        se += a;
        ss += c;
    }
    se = block_sum(se, s_red);
    ss = block_sum(ss, s_red);
    if (threadIdx.x == 0) {
        b.part_sz[2 * blockIdx.x] = se;
        b.part_sz[2 * blockIdx.x + 1] = ss;
    }
}

__global__ void __launch_bounds__(kPartThreads) k_size_scatter(MsaGeom g, MsaBufs b) {
    __shared__ unsigned long long s_scan[33];
    MsaStatus* st = b.status;
    if (st->abort || st->halo_fail) return;
    const uint32_t n = st->n_syms, P = gridDim.x;
    const uint32_t per = (n + P - 1) / P;
    const uint32_t k_begin = min(n, blockIdx.x * per), k_end = min(n, k_begin + per);
    unsigned long long me = 0, ms = 0;
    for (uint32_t q = threadIdx.x; q < blockIdx.x; q += blockDim.x) {
        me += b.part_sz[2 * q];
        ms += b.part_sz[2 * q + 1];
    }
    unsigned long long base_e = block_sum(me, s_scan) + (st->lead_hi - st->lead_lo) + st->lead_close;
    unsigned long long base_s = block_sum(ms, s_scan);
    for (uint32_t k0 = k_begin; k0 < k_end; k0 += blockDim.x) {
        const uint32_t k = k0 + threadIdx.x;
        unsigned long long a = 0, c = 0;
        if (k < k_end) symbol_sizes(g, b, st, k, a, c);
        unsigned long long ta, tc;
        const unsigned long long ea = block_exclusive_scan(a, s_scan, ta);
        const unsigned long long ec = block_exclusive_scan(c, s_scan, tc);
        if (k < k_end) {
            b.eds_off[k] = base_e + ea;
            b.seds_off[k] = base_s + ec;
        }
        base_e += ta;
        base_s += tc;
    }
    if (blockIdx.x == P - 1 && threadIdx.x == 0) {
        st->eds_total = base_e;
        st->seds_total = base_s;
        st->need_eds = base_e;
        st->need_seds = base_s;
        if (base_e > b.cap_eds)
            st->abort = kAbortEdsCap;
        else if (base_s > b.cap_seds)
            st->abort = kAbortSedsCap;
    }
}

// ---------------------------------------------------------------------------------------------
// k_emit_common: conserved text. A thread owns 16 consecutive window columns, finds the symbol of
// the first by binary search and walks on from there; bytes of common symbols (and of the leading
// continuation of a lower shard's symbol) go to eds_off[k] + 1 + (c - start). The same kernel
// writes each owned common symbol's braces and its "{0}".
// ---------------------------------------------------------------------------------------------
__global__ void k_emit_common(MsaGeom g, MsaBufs b) {
    const MsaStatus* st = b.status;
    if (st->abort || st->halo_fail) return;
    const uint32_t n_syms = st->n_syms, k_lo = st->k_lo, k_hi = st->k_hi;
    const uint32_t lead_lo = st->lead_lo, lead_hi = st->lead_hi;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;

    // braces and {0}
    for (uint32_t k = k_lo + tid; k < k_hi; k += nth) {
        const uint32_t e = b.sym[k];
        if (!(e & kCommonFlag)) continue;
        const uint32_t en = b.sym[k + 1] & kColMask;
        const unsigned long long off = b.eds_off[k];
        b.eds_out[off] = '{';
        if (en <= g.own_hi) b.eds_out[off + 1 + (en - (e & kColMask))] = '}';
        uint8_t* sd = b.seds_out + b.seds_off[k];
        sd[0] = '{';
        sd[1] = '0';
        sd[2] = '}';
    }
    if (tid == 0 && st->lead_close) b.eds_out[lead_hi - lead_lo] = '}';

    // text
    const uint32_t g_lo = g.own_lo >> 4, g_hi = (g.own_hi + 15u) >> 4;
    for (uint32_t grp = g_lo + tid; grp < g_hi; grp += nth) {
        const uint32_t c_first = max(grp * 16u, g.own_lo), c_last = min(grp * 16u + 16u, g.own_hi);
        if (c_first >= c_last) continue;
        // symbol containing c_first: last k with start <= c_first (may not exist)
        uint32_t k = sym_lower_bound(b.sym, n_syms, c_first + 1u);  // first start > c_first
        uint32_t next_start = b.sym[k] & kColMask;                  // sentinel at n_syms
        bool have = k > 0;
        if (have) --k;
        uint32_t cur = have ? b.sym[k] : 0u;
        unsigned long long off = (have && k >= k_lo && k < k_hi) ? b.eds_off[k] : 0ull;
        for (uint32_t c = c_first; c < c_last; ++c) {
            while (c >= next_start) {  // step into the next symbol
                k = have ? k + 1 : 0;
                have = true;
                cur = b.sym[k];
                next_start = b.sym[k + 1] & kColMask;
                off = (k >= k_lo && k < k_hi) ? b.eds_off[k] : 0ull;
            }
            if (c >= lead_lo && c < lead_hi) {
                b.eds_out[c - lead_lo] = b.refc[c];
            } else if (have && (cur & kCommonFlag) && k >= k_lo && k < k_hi) {
                b.eds_out[off + 1u + (c - (cur & kColMask))] = b.refc[c];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// k_emit_var: "{alt0,alt1,...}" and one "{ids}" per alternative for every owned variable symbol
// (msa_transforms.cpp:297-317); SEDS ids are 1-based rows in ascending order.
//  * wide (warp per symbol): a row's byte offset is base(alt) + bytes of lower rows of the same alternative,
//    obtained per 32-row chunk from __match_any_sync plus per-alternative running totals. Within a chunk
//    decimal widths take at most two values, so the in-chunk prefix is two popcounts.
//    Scratch: altw[Rq], altbase[Rq], running[Rq] (uint32) + the stage of stage_symbol.
//  * narrow (lane per single-column symbol, `nb` symbols per warp pass): the lane re-derives the residue
//    order (cheaper than storing it), counts bytes per alternative, then places every id in a per-lane
//    segment of shared memory; segments are copied out with 16-byte stores by the whole warp, so the SEDS
//    (70 % of the output bytes) leaves the SM coalesced. The row number's digits are warp-uniform.
//    Scratch (aliases the wide scratch): cnt[8][32], cur[8][32] (uint32), nb segments of seg_pitch bytes.
// ---------------------------------------------------------------------------------------------
struct EmitScratch {
    uint32_t* altw;
    uint32_t* altbase;
    uint32_t* running;
    uint8_t* stage;
    uint16_t* cdesc;
    uint32_t stage_bytes;
};

__device__ void emit_wide(const MsaGeom& g, const MsaBufs& b, const EmitScratch& sc, uint32_t k) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t lt = lanemask_lt();
    uint32_t *altw = sc.altw, *altbase = sc.altbase, *running = sc.running;
    const uint32_t e = b.sym[k];
    const uint32_t s = e & kColMask, en = b.sym[k + 1] & kColMask;
    const uint32_t slot0 = first_slot(b, s);
    const uint32_t nalts = b.sym_nalts[k];
    RowWalk w{b.vbits, b.refc, b.stash, g.Rp, sc.cdesc, sc.stage, s, false};
    w.staged = stage_symbol(g, b, s, en, slot0, sc.cdesc, sc.stage, sc.stage_bytes);
    uint8_t* eds = b.eds_out + b.eds_off[k];
    uint8_t* seds = b.seds_out + b.seds_off[k];

    // ---- EDS: '{' alt0 ',' alt1 ... '}' in leader-row order
    unsigned long long run = 1;
    uint32_t a_base = 0;
    for (uint32_t r0 = 0; r0 < g.R; r0 += 32) {
        const uint32_t r = r0 + lane;
        const uint32_t mask = b.leadmask[(size_t)slot0 * (g.Rp >> 5) + (r0 >> 5)];
        const bool isl = (mask >> lane) & 1u;
        uint32_t n = 0;
        if (isl) {
            uint32_t c = s, slot = slot0;
            while (next_char(w, r, c, en, slot) >= 0) ++n;
        }
        const unsigned long long contrib = isl ? (unsigned long long)n + 1ull : 0ull;
        const unsigned long long inc = warp_inclusive_scan(contrib);
        if (isl) {
            const unsigned long long pos = run + inc - contrib;
            const uint32_t a = a_base + (uint32_t)__popc(mask & lt);
            eds[pos - 1] = a == 0 ? '{' : ',';
            uint32_t c = s, slot = slot0;
            unsigned long long at = pos;
            int ch;
            while ((ch = next_char(w, r, c, en, slot)) >= 0) eds[at++] = (uint8_t)ch;
        }
        run += __shfl_sync(0xffffffffu, inc, 31);
        a_base += (uint32_t)__popc(mask);
    }
    if (lane == 0) eds[run - 1] = '}';

    // ---- SEDS pass 1: bytes per alternative (each id costs separator + digits)
    for (uint32_t i = lane; i < nalts; i += 32) {
        altw[i] = 0;
        running[i] = 0;
    }
    __syncwarp();
    for (uint32_t r0 = 0; r0 < g.R; r0 += 32) {
        const uint32_t r = r0 + lane;
        const bool valid = r < g.R;
        const uint32_t a = valid ? load_alt(g, b.altid, (size_t)slot0 * g.Rp + r) : kEmptySlot;
        const uint32_t wd = decimal_width(r + 1u) + 1u;
        const uint32_t wfirst = __shfl_sync(0xffffffffu, wd, 0);
        const uint32_t mA = __ballot_sync(0xffffffffu, wd == wfirst);
        const uint32_t peers = __match_any_sync(0xffffffffu, a);
        if (valid && lane == (uint32_t)__ffs((int)peers) - 1u)
            altw[a] += (uint32_t)__popc(peers & mA) * wfirst + (uint32_t)__popc(peers & ~mA) * (wfirst + 1u);
        __syncwarp();
    }
    // exclusive scan over alternatives; every finished alternative adds its '}'
    uint32_t carry = 0;
    for (uint32_t i0 = 0; i0 < nalts; i0 += 32) {
        const uint32_t i = i0 + lane;
        const uint32_t v = i < nalts ? altw[i] + 1u : 0u;
        const uint32_t inc = warp_inclusive_scan(v);
        if (i < nalts) altbase[i] = carry + inc - v;
        carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    __syncwarp();
    // ---- SEDS pass 2: place every id
    for (uint32_t r0 = 0; r0 < g.R; r0 += 32) {
        const uint32_t r = r0 + lane;
        const bool valid = r < g.R;
        const uint32_t a = valid ? load_alt(g, b.altid, (size_t)slot0 * g.Rp + r) : kEmptySlot;
        const uint32_t wd = decimal_width(r + 1u) + 1u;
        const uint32_t wfirst = __shfl_sync(0xffffffffu, wd, 0);
        const uint32_t mA = __ballot_sync(0xffffffffu, wd == wfirst);
        const uint32_t peers = __match_any_sync(0xffffffffu, a);
        const uint32_t below = peers & lt;
        const uint32_t pre = (uint32_t)__popc(below & mA) * wfirst + (uint32_t)__popc(below & ~mA) * (wfirst + 1u);
        uint32_t done = 0;
        if (valid) {
            done = running[a] + pre;
            uint8_t* dst = seds + altbase[a] + done;
            dst[0] = done == 0 ? '{' : ',';
            write_decimal(dst + 1, r + 1u, wd - 1u);
            if (done + wd == altw[a]) dst[wd] = '}';
        }
        __syncwarp();
        if (valid && lane == 31u - (uint32_t)__clz((int)peers)) running[a] = done + wd;
        __syncwarp();
    }
    __syncwarp();
}

// warp-cooperative copy of n bytes from shared to global memory; (dst - src) is a multiple of 16
__device__ __forceinline__ void warp_copy_out(uint8_t* dst, const uint8_t* src, uint32_t n) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t head = min(n, (uint32_t)((16u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u));
    if (lane < head) dst[lane] = src[lane];
    const uint32_t body = (n - head) >> 4;
    const uint4* s4 = reinterpret_cast<const uint4*>(src + head);
    uint4* d4 = reinterpret_cast<uint4*>(dst + head);
    for (uint32_t i = lane; i < body; i += 32) d4[i] = s4[i];
    const uint32_t done = head + (body << 4);
    if (done + lane < n) dst[done + lane] = src[done + lane];
}

// Narrow SEDS segment of one lane: alternative a occupies [start[a], start[a] + cnt[a]], '{' first and '}'
// last. cnt[a * 32 + lane] holds the byte counts on entry; cur[] receives the write cursors (top bit = no id
// placed yet). Returns the segment length.
__device__ __forceinline__ uint32_t seg_layout(uint8_t* seg, const uint32_t* cnt, uint32_t* cur, uint32_t nalts) {
    const uint32_t lane = threadIdx.x & 31;
    uint32_t at = 0;
    for (uint32_t a = 0; a < nalts; ++a) {
        const uint32_t c = cnt[a * 32u + lane];
        cur[a * 32u + lane] = at | 0x80000000u;
        seg[at] = '{';
        seg[at + c] = '}';
        at += c + 1u;
    }
    return at;
}

__device__ __forceinline__ void seg_place(uint8_t* seg, uint32_t* cur, uint32_t a, uint32_t id) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t wd = decimal_width(id);
    const uint32_t cv = cur[a * 32u + lane], pos = cv & 0x7fffffffu;
    if (!(cv >> 31)) seg[pos] = ',';
    write_decimal(seg + pos + 1u, id, wd);
    cur[a * 32u + lane] = pos + wd + 1u;
}

// copy the finished segments out, one symbol at a time, whole warp per symbol
__device__ __forceinline__ void segs_copy_out(const uint8_t* segs, uint32_t seg_pitch, uint32_t seg_bytes, uint32_t shift,
                                              uint8_t* seds_dst) {
    uint32_t ready = __ballot_sync(0xffffffffu, seg_bytes != 0u);
    while (ready) {
        const uint32_t bit = (uint32_t)__ffs((int)ready) - 1u;
        ready &= ready - 1u;
        const uint32_t n = __shfl_sync(0xffffffffu, seg_bytes, (int)bit);
        const uint32_t sh = __shfl_sync(0xffffffffu, shift, (int)bit);
        const unsigned long long dst =
            __shfl_sync(0xffffffffu, (unsigned long long)reinterpret_cast<uintptr_t>(seds_dst), (int)bit);
        warp_copy_out(reinterpret_cast<uint8_t*>((uintptr_t)dst), segs + (size_t)bit * seg_pitch + sh, n);
    }
}

// Single-column symbol, rows across lanes (any number of rows). Pass 1 grows the residue list and sums the
// bytes of every alternative; pass 2 places the ids: per chunk and alternative one ballot gives the lanes of
// that alternative, their in-chunk prefix is two popcounts (decimal widths take at most two values in a
// chunk). All per-alternative state is warp-uniform and lives in registers (8 alternatives at most).
// The same symbol from what k_group left behind (row bitsets + residue list): EDS text from the list, one staged
// "{ids}" per alternative (idlist.cuh). No second classification of the 1000 rows.
// bytes "id," of the ids of word w (id = 32 w + bit + 1) given the word's id widths: lo ids are wl wide, the rest wl + 1
__device__ __forceinline__ uint32_t word_bytes_fast(uint32_t v, uint32_t lowmask, uint32_t wl) {
    return (uint32_t)__popc(v) * (wl + 1u) + (uint32_t)__popc(v & ~lowmask);
}

// one id's text (digits + ',') into the stage: predicated byte stores, no branch on the width
__device__ __forceinline__ void stage_id(uint8_t* stage, uint32_t q, unsigned long long e, uint32_t dw) {
#ifdef EDSB_EMU
    for (uint32_t j = 0; j <= dw; ++j) stage[q + j] = (uint8_t)(e >> (8u * j));
#else
    const uint32_t lo4 = (uint32_t)e, hi4 = (uint32_t)(e >> 32);
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(stage) + q;
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(lo4) : "memory");
    asm volatile("st.shared.u8 [%0+1], %1;" ::"r"(a), "r"(lo4 >> 8) : "memory");
    if (dw >= 2u) asm volatile("st.shared.u8 [%0+2], %1;" ::"r"(a), "r"(lo4 >> 16) : "memory");
    if (dw >= 3u) asm volatile("st.shared.u8 [%0+3], %1;" ::"r"(a), "r"(lo4 >> 24) : "memory");
    if (dw >= 4u) asm volatile("st.shared.u8 [%0+4], %1;" ::"r"(a), "r"(hi4) : "memory");
    if (dw >= 5u) asm volatile("st.shared.u8 [%0+5], %1;" ::"r"(a), "r"(hi4 >> 8) : "memory");
    if (dw >= 6u) asm volatile("st.shared.u8 [%0+6], %1;" ::"r"(a), "r"(hi4 >> 16) : "memory");
#endif
}

template <uint32_t GL, typename T>
__device__ __forceinline__ T group_inclusive_scan(T v, uint32_t gl) {
#pragma unroll
    for (uint32_t d = 1; d < GL; d <<= 1) {
        const T o = __shfl_up_sync(0xffffffffu, v, d);
        if (gl >= d) v += o;
    }
    return v;
}

template <uint32_t GL, typename T>
__device__ __forceinline__ T group_sum(T v) {
#pragma unroll
    for (uint32_t d = GL >> 1; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, (int)d);
    return v;
}

// Single-column symbols from what k_group left behind (row bitsets + residue list), a group of GL lanes per symbol:
// EDS text from the list; the SEDS segment of the WHOLE symbol — every alternative's "{ids}" — is laid out in the
// group's stage at the output's own 16-byte phase and leaves with aligned 16-byte stores. A lane owns word w of every
// alternative, i.e. rows 32 w .. 32 w + 31: every row is in exactly one alternative, so every lane renders exactly 32
// ids — balanced, one scan per alternative, one copy-out per symbol. All collectives are executed by all 32 lanes.
// (stage: the symbol's SEDS bytes + 32; alignments deeper than 32 GL rows take one more round per 32 GL rows.)
template <uint32_t GL>
__device__ void emit_single_lanes(const MsaGeom& g, const MsaBufs& b, bool active, uint32_t k, uint32_t s, uint8_t* stage) {
    const uint32_t lane = threadIdx.x & 31, gl = lane & (GL - 1u), gbase = lane & ~(GL - 1u);
    const uint32_t Rw = g.Rp >> 5;
    const uint32_t slot0 = active ? first_slot(b, s) : 0u;
    Seen sn;
    sn.lo = sn.hi = sn.n = 0;
    if (active) {
        sn.lo = b.seen[(size_t)slot0 * 3u];
        sn.hi = b.seen[(size_t)slot0 * 3u + 1u];
        sn.n = b.seen[(size_t)slot0 * 3u + 2u];
        if (sn.n > 8u) {  // on the hashed row path
            active = false;
            sn.n = 0;
        }
    }
    const uint32_t* rowbits = b.rowbits + (size_t)slot0 * 8u * Rw;
    if (active && gl == 0u) {
        uint8_t* eds = b.eds_out + b.eds_off[k];
        *eds++ = '{';
        for (uint32_t a = 0; a < sn.n; ++a) {
            if (a) *eds++ = ',';
            const uint32_t ch = seen_byte(sn, a);
            if (ch != (uint32_t)'-') *eds++ = (uint8_t)ch;
        }
        *eds = '}';
    }
    const unsigned long long so = active ? b.seds_off[k] : 0ull;
    const uint32_t phase = (uint32_t)(so & 15u);
    uint32_t amax = sn.n;  // most alternatives of any symbol of this warp (warp-uniform loop bound)
    for (int d = 16; d > 0; d >>= 1) amax = max(amax, __shfl_xor_sync(0xffffffffu, amax, d));
    // bytes of every alternative (all words), hence where each starts in the stage
    uint32_t start[8], fill[8];  // fill[a]: bytes of alternative a placed by the rounds so far
    uint32_t run = phase;
#pragma unroll
    for (uint32_t a = 0; a < 8u; ++a) {
        start[a] = run;
        fill[a] = 0;
        if (a >= amax) continue;
        uint32_t bytes = 0;
        if (a < sn.n)
            for (uint32_t w = gl; w < Rw; w += GL) bytes += word_id_bytes(w, rowbits[a * Rw + w]);
        bytes = group_sum<GL>(bytes);
        if (a < sn.n) run += 1u + bytes;  // '{' + "id," each; the last ',' becomes '}'
    }
    for (uint32_t w0 = 0; w0 < Rw; w0 += GL) {
        const uint32_t w = w0 + gl;
        // id widths of this lane's word: ids below `pow` are wl wide, the others wl + 1 (a word spans at most one power of ten)
        const uint32_t lo_id = w * 32u + 1u, wl = decimal_width(lo_id);
        uint32_t pow = 10;
        for (uint32_t i = 1; i < wl; ++i) pow *= 10u;
        const uint32_t lowmask = pow - lo_id >= 32u ? 0xffffffffu : low_bits(pow - lo_id);
#pragma unroll
        for (uint32_t a = 0; a < 8u; ++a) {
            if (a >= amax) continue;
            const uint32_t v = (a < sn.n && w < Rw) ? rowbits[a * Rw + w] : 0u;
            const uint32_t mine = word_bytes_fast(v, lowmask, wl);
            const uint32_t incl = group_inclusive_scan<GL>(mine, gl);
            uint32_t q = start[a] + 1u + fill[a] + (incl - mine);
            for (uint32_t rem = v; rem; rem &= rem - 1u) {
                const uint32_t id = w * 32u + (uint32_t)__ffs((int)rem);
                const unsigned long long e = __ldg(b.id_text + id);
                uint32_t dw = (uint32_t)(e >> 56);
                if (dw) {
                    stage_id(stage, q, e, dw);
                } else {  // more than six digits
                    dw = decimal_width(id);
                    write_decimal(stage + q, id, dw);
                    stage[q + dw] = (uint8_t)',';
                }
                q += dw + 1u;
            }
            fill[a] += __shfl_sync(0xffffffffu, incl, (int)(gbase + GL - 1u));
        }
    }
    __syncwarp();
    // braces of alternative a by lane a of the group (groups of fewer than 8 lanes: lanes take several)
    for (uint32_t a0 = gl; a0 < 8u; a0 += GL) {
        uint32_t st0 = 0, fl = 0;
#pragma unroll
        for (uint32_t a = 0; a < 8u; ++a)
            if (a == a0) {
                st0 = start[a];
                fl = fill[a];
            }
        if (a0 < sn.n) {
            stage[st0] = (uint8_t)'{';
            stage[st0 + fl] = (uint8_t)'}';
        }
    }
    __syncwarp();
    // copy out [phase, run): aligned 16-byte stores, bytes at the two ragged ends
    if (active) {
        uint8_t* const dst = b.seds_out + (so - phase);
        for (uint32_t j = gl * 16u; j < run; j += GL * 16u) {
            if (j >= phase && j + 16u <= run) {
                *reinterpret_cast<uint4*>(dst + j) = *reinterpret_cast<const uint4*>(stage + j);
            } else {
                const uint32_t lo_b = j > phase ? j : phase, hi_b = j + 16u < run ? j + 16u : run;
                for (uint32_t i = lo_b; i < hi_b; ++i) dst[i] = stage[i];
            }
        }
    }
    __syncwarp();
}

__device__ bool emit_single_warp(const MsaGeom& g, const MsaBufs& b, uint32_t k, uint32_t s) {
    const uint32_t lane = threadIdx.x & 31, lt = lanemask_lt();
    const uint8_t* col = b.stash + (size_t)first_slot(b, s) * g.Rp;
    Seen sn;
    sn.lo = sn.hi = sn.n = 0;
    uint32_t altw[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (uint32_t r0 = 0; r0 < g.R; r0 += 32) {
        const uint32_t r = r0 + lane;
        const bool valid = r < g.R;
        uint32_t cls;
        if (!warp_classify(sn, valid ? col[r] : 0u, valid, cls)) return false;  // queued as wide by k_group
        const uint32_t wd = decimal_width(r + 1u) + 1u;
        const uint32_t wfirst = __shfl_sync(0xffffffffu, wd, 0);
        const uint32_t mA = __ballot_sync(0xffffffffu, wd == wfirst);
#pragma unroll
        for (uint32_t a = 0; a < 8u; ++a) {
            if (a < sn.n) {  // warp-uniform
                const uint32_t m = __ballot_sync(0xffffffffu, valid && cls == a);
                altw[a] += (uint32_t)__popc(m & mA) * wfirst + (uint32_t)__popc(m & ~mA) * (wfirst + 1u);
            }
        }
    }
    uint32_t base[8];
    uint32_t at = 0;
#pragma unroll
    for (uint32_t a = 0; a < 8u; ++a) {
        base[a] = at;
        at += a < sn.n ? altw[a] + 1u : 0u;
    }
    uint8_t* seds = b.seds_out + b.seds_off[k];
    if (lane == 0) {
        uint8_t* eds = b.eds_out + b.eds_off[k];
        *eds++ = '{';
        for (uint32_t a = 0; a < sn.n; ++a) {
            if (a) *eds++ = ',';
            const uint32_t ch = seen_byte(sn, a);
            if (ch != (uint32_t)'-') *eds++ = (uint8_t)ch;
        }
        *eds = '}';
    }
    if (lane < sn.n) {
        // lane a closes alternative a; its '{' is written with the alternative's first id
        uint32_t b0 = 0, w0 = 0;
#pragma unroll
        for (uint32_t a = 0; a < 8u; ++a)
            if (a == lane) {
                b0 = base[a];
                w0 = altw[a];
            }
        seds[b0 + w0] = '}';
    }
    uint32_t running[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (uint32_t r0 = 0; r0 < g.R; r0 += 32) {
        const uint32_t r = r0 + lane;
        const bool valid = r < g.R;
        const uint32_t cls = valid ? seen_lookup(sn, col[r]) : 8u;
        const uint32_t wd = decimal_width(r + 1u) + 1u;
        const uint32_t wfirst = __shfl_sync(0xffffffffu, wd, 0);
        const uint32_t mA = __ballot_sync(0xffffffffu, wd == wfirst);
        uint32_t pos = 0;
#pragma unroll
        for (uint32_t a = 0; a < 8u; ++a) {
            if (a < sn.n) {  // warp-uniform
                const uint32_t m = __ballot_sync(0xffffffffu, cls == a);
                if (cls == a) pos = base[a] + running[a] + (uint32_t)__popc(m & lt & mA) * wfirst + (uint32_t)__popc(m & lt & ~mA) * (wfirst + 1u);
                running[a] += (uint32_t)__popc(m & mA) * wfirst + (uint32_t)__popc(m & ~mA) * (wfirst + 1u);
            }
        }
        if (valid) {
            uint32_t b0 = 0;
#pragma unroll
            for (uint32_t a = 0; a < 8u; ++a)
                if (a == cls) b0 = base[a];
            seds[pos] = pos == b0 ? '{' : ',';
            write_decimal(seds + pos + 1u, r + 1u, wd - 1u);
        }
    }
    return true;
}

// k_emit_var: the single-column symbols (lane per symbol, nb symbols per warp pass).
// idtab[r] = decimal digits of r + 1 packed in 24 bits (first digit lowest) | width << 24, built once per
// block so the per-row decimal work is a table lookup (R <= 999 here: the host turns the narrow path off
// for more rows).
__global__ void k_emit_var(MsaGeom g, MsaBufs b, uint32_t per_warp_smem, uint32_t nb, uint32_t seg_pitch) {
    const MsaStatus* st = b.status;
    if (st->abort || st->halo_fail) return;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    if (nb == 0u) {
        // rows across lanes: a group of GL lanes per symbol (the bitsets k_group left in b.rowbits), 32 / GL symbols per
        // warp; per_warp_smem = the groups' stages. Without the bitsets / a stage: round 1's row-parallel form.
        uint8_t* stage = (per_warp_smem && b.rowbits) ? EDSB_DYN_SMEM() + (size_t)warp * per_warp_smem : nullptr;
        const uint32_t v_lo = st->v_lo, v_hi = st->v_hi;
        if (stage) {
            const uint32_t GL = seg_pitch, NS = 32u / GL, sub = lane / GL;  // (seg_pitch carries the group width here)
            uint8_t* my_stage = stage + (size_t)sub * (per_warp_smem / NS);
            for (uint32_t v0 = v_lo + (blockIdx.x * wpb + warp) * NS; v0 < v_hi; v0 += gridDim.x * wpb * NS) {
                const uint32_t v = v0 + sub;
                bool active = v < v_hi;
                uint32_t k = 0, s = 0;
                if (active) {
                    k = b.varsym[v];
                    s = b.sym[k] & kColMask;
                    active = (b.sym[k + 1] & kColMask) - s == 1u;
                }
                switch (GL) {  // (warp-uniform)
                    case 1: emit_single_lanes<1>(g, b, active, k, s, my_stage); break;
                    case 2: emit_single_lanes<2>(g, b, active, k, s, my_stage); break;
                    case 4: emit_single_lanes<4>(g, b, active, k, s, my_stage); break;
                    case 8: emit_single_lanes<8>(g, b, active, k, s, my_stage); break;
                    case 16: emit_single_lanes<16>(g, b, active, k, s, my_stage); break;
                    default: emit_single_lanes<32>(g, b, active, k, s, my_stage); break;
                }
            }
            return;
        }
        for (uint32_t v = v_lo + blockIdx.x * wpb + warp; v < v_hi; v += gridDim.x * wpb) {
            const uint32_t k = b.varsym[v];
            const uint32_t s = b.sym[k] & kColMask, en = b.sym[k + 1] & kColMask;
            if (en - s == 1u) {
                emit_single_warp(g, b, k, s);
                __syncwarp();
            }
        }
        return;
    }
    uint32_t* idtab = reinterpret_cast<uint32_t*>(EDSB_DYN_SMEM());
    for (uint32_t r = threadIdx.x; r < g.R; r += blockDim.x) {
        const uint32_t id = r + 1u, wd = decimal_width(id);
        uint8_t d[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        write_decimal(d, id, wd);
        idtab[r] = (uint32_t)d[0] | ((uint32_t)d[1] << 8) | ((uint32_t)d[2] << 16) | (wd << 24);
    }
    __syncthreads();
    unsigned char* smem = EDSB_DYN_SMEM() + (((size_t)g.R * 4u + 15u) & ~(size_t)15u) + (size_t)warp * per_warp_smem;
    uint32_t* cnt = reinterpret_cast<uint32_t*>(smem);  // [8][32]
    uint32_t* cur = cnt + 256;                           // [8][32]
    uint8_t* segs = reinterpret_cast<uint8_t*>(cur + 256);
    const uint32_t v_lo = st->v_lo, v_hi = st->v_hi;

    for (uint32_t v0 = v_lo + (blockIdx.x * wpb + warp) * nb; v0 < v_hi; v0 += gridDim.x * wpb * nb) {
        const uint32_t v = v0 + lane;
        const bool have = lane < nb && v < v_hi;
        const uint32_t k = have ? b.varsym[v] : 0u;
        uint32_t s = 0, en = 0;
        if (have) {
            s = b.sym[k] & kColMask;
            en = b.sym[k + 1] & kColMask;
        }
        uint32_t seg_bytes = 0, shift = 0;
        uint8_t* seds_dst = nullptr;
        if (have && en - s == 1u) {
            const uint8_t* col = b.stash + (size_t)first_slot(b, s) * g.Rp;
            // pass A: residue order and rows per alternative, counted per id-width band (ids 1-9, 10-99, 100-999
            // cost 2, 3, 4 bytes with their separator): eight 8-bit counters per band in one 64-bit register
            // (R <= 160 here, so a band holds at most 90 rows). No shared memory in this loop.
            Seen sn;
            sn.lo = sn.hi = sn.n = 0;
            unsigned long long rows1 = 0, rows2 = 0, rows3 = 0;
            bool ok = true;
            const uint32_t full = g.R & ~3u;
            auto count_row = [&](uint32_t ch, uint32_t r) {
                const uint32_t a = ch ? seen_index(sn, ch) : 8u;
                if (a >= 8u) {
                    ok = false;  // queued as wide by k_group
                    return;
                }
                const unsigned long long one = 1ull << (8u * a);
                if (r < 9u) rows1 += one; else if (r < 99u) rows2 += one; else rows3 += one;  // r is warp-uniform
            };
            for (uint32_t r0 = 0; r0 < full && ok; r0 += 4) {
                const uint32_t word = *reinterpret_cast<const uint32_t*>(col + r0);
                count_row(word & 0xffu, r0);
                count_row((word >> 8) & 0xffu, r0 + 1u);
                count_row((word >> 16) & 0xffu, r0 + 2u);
                count_row(word >> 24, r0 + 3u);
            }
            for (uint32_t r = full; r < g.R && ok; ++r) count_row(col[r], r);
            if (ok) {
                // EDS text straight from the lane (a dozen bytes)
                uint8_t* eds = b.eds_out + b.eds_off[k];
                *eds++ = '{';
                for (uint32_t a = 0; a < sn.n; ++a) {
                    if (a) *eds++ = ',';
                    const uint32_t ch = seen_byte(sn, a);
                    if (ch != (uint32_t)'-') *eds++ = (uint8_t)ch;
                    cnt[a * 32u + lane] = 2u * (uint32_t)((rows1 >> (8u * a)) & 0xffull) + 3u * (uint32_t)((rows2 >> (8u * a)) & 0xffull) +
                                          4u * (uint32_t)((rows3 >> (8u * a)) & 0xffull);
                }
                *eds = '}';
                seds_dst = b.seds_out + b.seds_off[k];
                shift = ((uint32_t)(reinterpret_cast<uintptr_t>(seds_dst) & 15u) - ((lane * seg_pitch) & 15u)) & 15u;
                uint8_t* seg = segs + (size_t)lane * seg_pitch + shift;
                seg_bytes = seg_layout(seg, cnt, cur, sn.n);
                // pass B: place every id (digits and width of the row number come from the per-block table)
                auto place_row = [&](uint32_t ch, uint32_t r) {
                    const uint32_t a = seen_lookup(sn, ch);
                    const uint32_t t = idtab[r], wd = t >> 24;
                    uint32_t* cp = cur + a * 32u + lane;
                    const uint32_t cv = *cp, pos = cv & 0x7fffffffu;
                    if (!(cv >> 31)) seg[pos] = ',';
                    seg[pos + 1u] = (uint8_t)t;
                    if (wd > 1u) seg[pos + 2u] = (uint8_t)(t >> 8);  // wd is warp-uniform
                    if (wd > 2u) seg[pos + 3u] = (uint8_t)(t >> 16);
                    *cp = pos + wd + 1u;
                };
                for (uint32_t r0 = 0; r0 < full; r0 += 4) {
                    const uint32_t word = *reinterpret_cast<const uint32_t*>(col + r0);
                    place_row(word & 0xffu, r0);
                    place_row((word >> 8) & 0xffu, r0 + 1u);
                    place_row((word >> 16) & 0xffu, r0 + 2u);
                    place_row(word >> 24, r0 + 3u);
                }
                for (uint32_t r = full; r < g.R; ++r) place_row(col[r], r);
            }
        }
        __syncwarp();
        segs_copy_out(segs, seg_pitch, seg_bytes, shift, seds_dst);
        __syncwarp();
    }
}

// k_emit2: the queued symbols, warp per symbol.
__global__ void k_emit2(MsaGeom g, MsaBufs b, uint32_t Rq, uint32_t use_global, uint32_t stage_bytes,
                        uint32_t per_warp_smem) {
    const MsaStatus* st = b.status;
    if (st->abort || st->halo_fail) return;
    const uint32_t warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    unsigned char* smem = EDSB_DYN_SMEM() + (size_t)warp * per_warp_smem;
    unsigned char* base = use_global ? b.group_ws + ((size_t)blockIdx.x * wpb + warp) * ((size_t)Rq * 12u) : smem;
    EmitScratch sc;
    sc.altw = reinterpret_cast<uint32_t*>(base);
    sc.altbase = sc.altw + Rq;
    sc.running = sc.altbase + Rq;
    sc.stage = reinterpret_cast<uint8_t*>(sc.running + Rq);
    sc.cdesc = reinterpret_cast<uint16_t*>(sc.stage + stage_bytes);
    sc.stage_bytes = use_global ? 0u : stage_bytes;
    const uint32_t n_wide = st->n_emit2;
    for (uint32_t item = blockIdx.x * wpb + warp; item < n_wide; item += gridDim.x * wpb) emit_wide(g, b, sc, b.emitlist[item]);
}

// ---------------------------------------------------------------------------------------------
// k_synth: the synthetic alignment of BASELINE.json configs 2 / 4, written as FASTA text straight
// into HBM. Keyed hashing (splitmix64 of seed, row, column) makes any column window of the same
// alignment reproducible on any rank.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}

__device__ __forceinline__ uint8_t synth_residue(uint64_t seed, uint32_t r, uint64_t col, uint32_t variable_ppm) {
    const uint64_t hc = mix64(seed ^ mix64(col));
    const uint8_t base0 = (uint8_t)"ACGT"[hc & 3u];
    if (r == 0) return base0;
    const bool variable = ((hc >> 8) % 1000000ull) < variable_ppm;
    if (!variable) return base0;
    const uint64_t hr = mix64(hc ^ mix64(((uint64_t)r << 1) | 1ull));
    const uint32_t roll = (uint32_t)((hr >> 8) % 100ull);
    if (roll < 30) return (uint8_t)"ACGT"[hr & 3u];  // substitution (may equal row 0)
    if (roll < 40) return (uint8_t)'-';
    return base0;
}

__global__ void k_synth(uint8_t* text, const uint64_t* row_off, uint32_t R, uint32_t lw, uint64_t u_begin,
                        uint64_t row_bytes, uint64_t seed, uint32_t variable_ppm) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint32_t r = blockIdx.y; r < R; r += gridDim.y) {
        uint8_t* row = text + row_off[r];
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < row_bytes; i += stride) {
            const uint64_t u = u_begin + i;
            const uint64_t line = u / (lw + 1u);
            const uint32_t rem = (uint32_t)(u - line * (lw + 1u));
            row[i] = rem == lw ? (uint8_t)'\n' : synth_residue(seed, r, line * lw + rem, variable_ppm);
        }
    }
}

// =============================================================================================
// Host side
// =============================================================================================

#ifdef EDSB_EMU
constexpr uint32_t kSymWarps = 2;  // fewer OS threads per emulated block
#else
constexpr uint32_t kSymWarps = 8;
#endif

static uint32_t pow2_ceil(uint32_t v) {
    uint32_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

static uint64_t sum_decimal_widths(uint32_t R) {
    uint64_t total = 0, lo = 1, width = 1;
    while (lo <= R) {
        const uint64_t hi = std::min<uint64_t>(R, lo * 10 - 1);
        total += (hi - lo + 1) * width;
        lo *= 10;
        ++width;
    }
    return total;
}

MsaPipeline::MsaPipeline(eds_ctx* ctx) : ctx_(ctx) {
    EDSB_CUDA(cudaMallocHost(&h_status_, sizeof(MsaStatus)));
    d_status_.reserve(sizeof(MsaStatus));
}

MsaPipeline::~MsaPipeline() {
    DevBuf* all[] = {&d_rows_, &d_mism_, &d_vbits_, &d_tbits_, &d_rank_, &d_refc_, &d_part_, &d_varcol_, &d_runs_,
                     &d_sym_, &d_stash_, &d_altid_, &d_leadmask_, &d_symmeta_, &d_eds_, &d_seds_, &d_ws_, &d_status_,
                     &d_rowbits_, &d_seen_, &d_id_text_, &d_fz_rows_, &d_fz_tmp_, &d_fz_col_, &d_fz_cnt_};
    for (DevBuf* d : all) d->release();
    if (h_status_) cudaFreeHost(h_status_);
}

uint32_t MsaPipeline::partitions() const {
    uint32_t p = ctx_->partitions ? ctx_->partitions : 4u * (uint32_t)ctx_->sm_count;  // 4 blocks per SM: 592 on a B200
    return std::max(1u, std::min(p, 4096u));
}

void MsaPipeline::prepare(const eds_msa_view& v, uint32_t l, int leds) {
    if (!v.text || !v.row_start) throw std::invalid_argument("eds_msa_view: null text or row_start");
    if (reinterpret_cast<uintptr_t>(v.text) & 15u) throw std::invalid_argument("eds_msa_view: text must be 16-byte aligned");
    if (v.n_rows < 2) throw BadMsa("alignment needs at least 2 rows (undefined in the reference, msa_transforms.cpp:53-57)");
    if (v.line_width == 0 || v.total_cols == 0 || v.col_count == 0) throw BadMsa("empty alignment");
    if (v.col_begin + v.col_count > v.total_cols) throw std::invalid_argument("eds_msa_view: window exceeds the alignment");
    if (v.col_count >= 0x7fffffffull - 64) throw std::invalid_argument("eds_msa_view: col_count must be < 2^31 - 64");
    const uint64_t win_end = v.col_begin + v.col_count;
    if (v.own_begin < v.col_begin || v.own_end > win_end || v.own_begin > v.own_end)
        throw std::invalid_argument("eds_msa_view: owned range outside the window");
    if (v.own_begin < v.own_end) {
        if (v.col_begin > 0 && v.own_begin == v.col_begin)
            throw std::invalid_argument("eds_msa_view: a shard that does not start the alignment needs a left halo");
        if (win_end < v.total_cols && v.own_end == win_end)
            throw std::invalid_argument("eds_msa_view: a shard that does not end the alignment needs a right halo");
    }
    MsaGeom& g = geom_;
    memset(&g, 0, sizeof(g));
    g.text = v.text;
    g.n_vec = (v.text_bytes + 15) / 16;
    g.total_cols = v.total_cols;
    g.col_begin = v.col_begin;
    g.lw = v.line_width;
    g.u_begin = v.col_begin + v.col_begin / v.line_width;
    const uint64_t last = win_end - 1;
    g.row_bytes = last + last / v.line_width - g.u_begin + 1;
    g.R = v.n_rows;
    g.Rp = (v.n_rows + 31u) & ~31u;
    g.ncols = (uint32_t)v.col_count;
    g.own_lo = (uint32_t)(v.own_begin - v.col_begin);
    g.own_hi = (uint32_t)(v.own_end - v.col_begin);
    g.a0 = (uint32_t)(v.row_start[0] & 15u);
    const uint64_t chunks = (g.a0 + g.row_bytes + 15) / 16;
    if (chunks >= 0xffffffffull) throw std::invalid_argument("eds_msa_view: window too large");
    g.n_chunks = (uint32_t)chunks;
    g.n_words = (g.ncols + 31u) / 32u;
    g.l = l;
    g.leds = leds ? 1u : 0u;
    g.alt32 = v.n_rows > 65535u ? 1u : 0u;
    g.sum_id_width = sum_decimal_widths(v.n_rows);
    g.hash_mask = ctx_->hash_mask;
    for (uint32_t r = 0; r < v.n_rows; ++r)
        if (v.row_start[r] + g.row_bytes > v.text_bytes) throw BadMsa("a row runs past the end of the buffer (rows of unequal length?)");

    cudaStream_t s = ctx_->stream;
    // device copy of the row offsets, followed by the packed (vector address | byte shift) of the rows
    // sorted by word shift: [row 0][class 0 rows][class 1]...[class 3]
    h_rows_.resize((size_t)v.n_rows * 2);
    long long dmin = 0, dmax = 0;
    uint32_t count[4] = {0, 0, 0, 0};
    bool all_aligned = true;
    for (uint32_t r = 0; r < v.n_rows; ++r) {
        const long long d = (long long)v.row_start[r] - (long long)g.a0;
        const long long dv = d >> 4;
        if (r == 0 || dv < dmin) dmin = dv;
        if (r == 0 || dv > dmax) dmax = dv;
        h_rows_[r] = v.row_start[r];
        if (r > 0) {
            ++count[(d & 15) >> 2];
            all_aligned = all_aligned && (d & 15) == 0;
        }
    }
    g.cls[0] = 1;
    for (int c = 0; c < 4; ++c) g.cls[c + 1] = g.cls[c] + count[c];
    g.all_aligned = all_aligned ? 1u : 0u;
    {
        uint32_t at[4] = {g.cls[0], g.cls[1], g.cls[2], g.cls[3]};
        uint64_t* pack = h_rows_.data() + v.n_rows;
        for (uint32_t r = 0; r < v.n_rows; ++r) {
            const long long d = (long long)v.row_start[r] - (long long)g.a0;
            const uint64_t e =
                (uint64_t)(reinterpret_cast<uintptr_t>(v.text) + (unsigned long long)((d >> 4) * 16)) | (uint64_t)(d & 15);
            pack[r == 0 ? 0 : at[(d & 15) >> 2]++] = e;
        }
    }
    g.d_min_vec = dmin;
    g.d_max_vec = dmax;
    d_rows_.reserve((size_t)v.n_rows * 16);
    EDSB_CUDA(cudaMemcpyAsync(d_rows_.p, h_rows_.data(), (size_t)v.n_rows * 16, cudaMemcpyHostToDevice, s));
    g.row_off = d_rows_.as<uint64_t>();

    const uint32_t P = partitions();
    d_mism_.reserve(((size_t)g.n_chunks / 2 + 4) * 4);
    d_vbits_.reserve((size_t)(g.n_words + 1) * 4);
    d_tbits_.reserve((size_t)(g.n_words + 1) * 4);
    d_rank_.reserve((size_t)(g.n_words + 1) * 4);
    d_refc_.reserve((size_t)g.n_words * 32 + 32);
    d_part_.reserve((size_t)P * (8 + 8 + 16));
    plan_fused();
}

// k_scan_fused geometry for the alignment of prepare(): cluster size (rows per CTA <= 128), stages that fit in shared
// memory, the per-CTA row tables (slot 0 = row 0, then the CTA's rows sorted by word shift), co-resident clusters.
// the instance of k_scan_fused the plan asks for: 16-chunk tiles, 32-chunk tiles, 32-chunk tiles fetched in pairs
#define FZ_KERNEL(EXPR)                                   \
    do {                                                  \
        if (fz_.T == 16u) {                               \
            auto kern = k_scan_fused<16, 1, false>;       \
            EXPR;                                         \
        } else if (fz_.H == 2u) {                         \
            auto kern = k_scan_fused<32, 2, false>;       \
            EXPR;                                         \
        } else if (fz_.direct) {                          \
            auto kern = k_scan_fused<32, 1, true>;        \
            EXPR;                                         \
        } else {                                          \
            auto kern = k_scan_fused<32, 1, false>;       \
            EXPR;                                         \
        }                                                 \
    } while (0)

void MsaPipeline::plan_fused() {
    const MsaGeom& g = geom_;
    fz_ = FzPlan();
    if (!ctx_->fused || g.R < ctx_->fused_min_rows) return;
    if (ctx_->fused_l2 && g.R <= (uint32_t)kRowCache) {
        // k_scan_l2: plain loads, the caches as the stage; a persistent grid, every duty warp of every CTA owns a region
        fz_.l2 = true;
        fz_.T = 32;
        fz_.DW = std::max(1u, std::min(ctx_->fused_dw ? ctx_->fused_dw : (uint32_t)kL2MaxDW, (uint32_t)kL2MaxDW));
        fz_.CW = std::max(1u, std::min(ctx_->fused_cw ? ctx_->fused_cw : (uint32_t)kL2MaxCW, (uint32_t)kL2MaxCW));
        const uint32_t n_tiles = (g.n_chunks + 31) / 32;
#ifdef EDSB_EMU
        uint32_t regions = 3;
#else
        const uint32_t key = 0x80000000u ^ (fz_.CW << 8) ^ fz_.DW;
        if (fz_occ_key_ != key) {
            int per_sm = 0;
            EDSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_scan_l2<true>, (int)((fz_.CW + fz_.DW) * 32), 0));
            fz_occ_regions_ = (uint32_t)std::max(0, per_sm) * (uint32_t)ctx_->sm_count;
            fz_occ_key_ = key;
        }
        uint32_t regions = fz_occ_regions_;
#endif
        regions = std::min(regions, n_tiles);
        if (regions == 0) {
            fz_ = FzPlan();
            return;
        }
        fz_.regions = regions;
        d_fz_cnt_.reserve((size_t)regions * fz_.DW * 4);
        FzParams& f = fzp_;
        memset(&f, 0, sizeof(f));
        f.region_count = d_fz_cnt_.as<uint32_t>();
        f.n_tiles = n_tiles;
        f.all_aligned = g.all_aligned;
        f.DW = fz_.DW;
        f.T = 32;
        f.NC = 1;
        f.probe = ctx_->fused_probe;
        fz_.on = true;
        return;
    }
    uint32_t NC = 1;
    while (NC < kFzMaxNC && (g.R + NC - 1) / NC > kFzGroupRows) NC <<= 1;
    if (ctx_->fused_nc) NC = ctx_->fused_nc;
    if ((g.R + NC - 1) / NC > kFzGroupRows || NC > kFzMaxNC) return;  // too deep for one cluster: k_scan + k_stash
#ifdef EDSB_EMU
    if (NC > 1) return;  // clusters are not emulated
#endif
    const uint32_t RG = (((g.R + NC - 1) / NC) + 31u) & ~31u;
    const uint32_t T = ctx_->fused_t == 16u ? 16u : 32u;  // 16-chunk tiles lose: twice the bulk copies per byte (profiles/r02_a)
    // pairs of adjacent tiles per bulk copy (scan_fused.cuh): bulk mode, 32-chunk tiles; S and DW even
    const uint32_t H = (T == 32u && ctx_->fused_mode == 0u && ctx_->fused_pair) ? 2u : 1u;  // off by default: see profiles/r02_a
    // rows that bypass the ring (scan_fused.cuh: the last slots of every CTA, loaded straight into registers)
    const uint32_t n_direct = (T == 32u && H == 1u) ? std::min(ctx_->fused_direct, kFzMaxDirect) : 0u;
    uint32_t slot_pitch = 1;  // row 0 + the most rows any CTA of the cluster holds
    uint32_t stage_slots = 1; // ... of which a stage holds
    for (uint32_t c = 0; c < NC; ++c) {
        const uint32_t lo = std::max(c * RG, 1u), hi = std::min(g.R, (c + 1) * RG);
        const uint32_t nslots = hi > lo ? 1u + hi - lo : 1u;
        slot_pitch = std::max(slot_pitch, nslots);
        stage_slots = std::max(stage_slots, nslots - std::min(n_direct, (nslots - 1u) / 2u));
    }
    uint32_t DW = std::max(1u, std::min(ctx_->fused_dw ? ctx_->fused_dw : 4u, kFzMaxDW));
#ifdef EDSB_EMU
    DW = std::min(DW, 2u);
#endif
    uint32_t S = 0;
    for (uint32_t s = 12; s >= 2; --s)
        if (fz_smem_bytes(T, s, NC, RG, slot_pitch, DW, stage_slots) + 1024 <= ctx_->smem_optin) {
            S = s;
            break;
        }
    // a communicator posts its all-gather beside the next scan: its kernel needs room on the SMs this persistent grid
    // fills, or one of our CTAs starts only when the slowest peer has arrived (8 GPUs: 0.67 -> 0.61 ms per config-2 step).
    // Not at the price of a ring shallower than three stages.
    if (ctx_->peer_headroom)
        for (uint32_t s = S; s >= 3; --s)
            if (fz_smem_bytes(T, s, NC, RG, slot_pitch, DW, stage_slots) + 1024 + ctx_->peer_headroom <= ctx_->smem_optin) {
                S = s;
                break;
            }
    if (ctx_->fused_stages && ctx_->fused_stages <= S) S = ctx_->fused_stages;
    if (S < 2) return;
    if (H == 2u) {
        S &= ~1u;
        DW = std::max(2u, DW & ~1u);
    }
    // a duty warp waits for phase q of a stage's barrier by parity, which is only sound while phase q - 1 is known
    // to be complete: true when its previous tile (DW tiles back) is not older than the stage's previous use (S back)
    DW = std::min(DW, S);
    const uint32_t n_tiles = (g.n_chunks + T - 1) / T;
    fz_.PW = std::max(1u, std::min(ctx_->fused_pw ? ctx_->fused_pw : 4u, kFzMaxPW));
#ifdef EDSB_EMU
    fz_.PW = std::min(fz_.PW, 2u);
#endif
    fz_.T = T;
    fz_.H = H;
    fz_.direct = n_direct > 0;
    fz_.S = S;
    fz_.DW = DW;
    fz_.NC = NC;
    fz_.RG = RG;
    fz_.slot_pitch = slot_pitch;
    fz_.smem = fz_smem_bytes(T, S, NC, RG, slot_pitch, DW, stage_slots);

    // per-CTA tables, one upload: pack[NC][slot_pitch] u64 | meta[NC][8] u32 | info[NC][RG] u16
    const size_t off_meta = (size_t)NC * slot_pitch * 8, off_info = off_meta + (size_t)NC * 32;
    h_fz_.assign(off_info + (((size_t)NC * RG * 2 + 15) & ~(size_t)15), 0);
    uint64_t* pack = reinterpret_cast<uint64_t*>(h_fz_.data());
    uint32_t* meta = reinterpret_cast<uint32_t*>(h_fz_.data() + off_meta);
    uint16_t* info = reinterpret_cast<uint16_t*>(h_fz_.data() + off_info);
    const uint64_t* row_start = h_rows_.data();
    auto packed = [&](uint32_t r) {
        const long long d = (long long)row_start[r] - (long long)g.a0;
        return (uint64_t)(reinterpret_cast<uintptr_t>(g.text) + (unsigned long long)((d >> 4) * 16)) | (uint64_t)(d & 15);
    };
    for (uint32_t c = 0; c < NC; ++c) {
        const uint32_t lo = c * RG, hi = std::min(g.R, lo + RG);
        uint32_t count[4] = {0, 0, 0, 0};
        for (uint32_t r = std::max(lo, 1u); r < hi; ++r) ++count[(packed(r) & 15) >> 2];
        uint32_t* m = meta + (size_t)c * 8;
        m[1] = 1;
        for (int k = 0; k < 4; ++k) m[2 + k] = m[1 + k] + count[k];
        m[0] = m[5];
        m[6] = lo;
        m[7] = hi > lo ? hi - lo : 0;
        uint32_t at[4] = {m[1], m[2], m[3], m[4]};
        uint64_t* pk = pack + (size_t)c * slot_pitch;
        pk[0] = packed(0);
        for (uint32_t rl = 0; rl < RG; ++rl) {
            const uint32_t r = lo + rl;
            uint16_t v = 0xffffu;
            if (r == 0) {
                v = 0;
            } else if (r < hi) {
                const uint64_t e = packed(r);
                const uint32_t slot = at[(e & 15) >> 2]++;
                pk[slot] = e;
                v = (uint16_t)((slot << 4) | (uint32_t)(e & 15));
            }
            info[(size_t)c * RG + rl] = v;
        }
    }
    cudaStream_t s = ctx_->stream;
    d_fz_rows_.reserve(h_fz_.size());
    EDSB_CUDA(cudaMemcpyAsync(d_fz_rows_.p, h_fz_.data(), h_fz_.size(), cudaMemcpyHostToDevice, s));

    // co-resident clusters (one CTA per SM at these shared-memory sizes): the grid is persistent
    uint32_t regions = 0;
#ifdef EDSB_EMU
    regions = 3;
#else
    const uint32_t key = (NC << 24) ^ (S << 16) ^ slot_pitch ^ (fz_.PW << 28) ^ (fz_.DW << 12) ^ (T << 20) ^ (H << 9) ^ (stage_slots * 2654435761u) ^ (n_direct ? 0x55u : 0u);
    if (fz_attr_smem_ < fz_.smem) {
        EDSB_CUDA(cudaFuncSetAttribute(k_scan_fused<16, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fz_.smem));
        EDSB_CUDA(cudaFuncSetAttribute(k_scan_fused<32, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fz_.smem));
        EDSB_CUDA(cudaFuncSetAttribute(k_scan_fused<32, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fz_.smem));
        EDSB_CUDA(cudaFuncSetAttribute(k_scan_fused<32, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fz_.smem));
        fz_attr_smem_ = fz_.smem;
    }
    if (fz_occ_key_ != key) {
        if (NC == 1) {
            int per_sm = 0;
            FZ_KERNEL(EDSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, (kFzCW + fz_.PW + fz_.DW) * 32, fz_.smem)));
            fz_occ_regions_ = (uint32_t)std::max(0, per_sm) * (uint32_t)ctx_->sm_count;
        } else {
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof(cfg));
            cfg.gridDim = dim3(NC * (uint32_t)ctx_->sm_count, 1, 1);
            cfg.blockDim = dim3((kFzCW + fz_.PW + fz_.DW) * 32, 1, 1);
            cfg.dynamicSmemBytes = fz_.smem;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = NC;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            int n = 0;
            FZ_KERNEL(EDSB_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &cfg)));
            fz_occ_regions_ = (uint32_t)std::max(0, n);
        }
        fz_occ_key_ = key;
    }
    regions = fz_occ_regions_;
#endif
    regions = std::min(regions, n_tiles);
    if (regions == 0) return;
    fz_.regions = regions;
    d_fz_cnt_.reserve((size_t)regions * DW * 4);

    FzParams& f = fzp_;
    memset(&f, 0, sizeof(f));
    f.pack = reinterpret_cast<const unsigned long long*>(d_fz_rows_.as<unsigned char>());
    f.meta = reinterpret_cast<const uint32_t*>(d_fz_rows_.as<unsigned char>() + off_meta);
    f.info = reinterpret_cast<const uint16_t*>(d_fz_rows_.as<unsigned char>() + off_info);
    f.region_count = d_fz_cnt_.as<uint32_t>();
    f.S = S;
    f.NC = NC;
    f.RG = RG;
    f.slot_pitch = slot_pitch;
    f.stage_slots = stage_slots;
    f.n_direct = n_direct;
    f.split = ctx_->fused_split;
    f.n_tiles = n_tiles;
    f.all_aligned = g.all_aligned;
    f.PW = fz_.PW;
    f.DW = fz_.DW;
    f.mode = ctx_->fused_mode;
    f.probe = ctx_->fused_probe;
    f.PWB = 0;
    f.n_bulk = 0;
    if (f.mode == 2u) {
        // both copy engines: the first warps issue bulk copies for a share of the rows, the others cp.async for the rest
        if (fz_.PW < 2u || fz_.T != 32u) {
            f.mode = 0;
        } else {
            f.PWB = std::max(1u, std::min(fz_.PW - 1u, ctx_->fused_pwb ? ctx_->fused_pwb : fz_.PW / 2u));
            f.n_bulk = (uint32_t)((uint64_t)slot_pitch * std::min(ctx_->fused_bulk_pct ? ctx_->fused_bulk_pct : 50u, 100u) / 100u);
        }
    }
    // tile t is fetched with bulk copies when vectors [32 t + dmin, 32 t + 32 + dmax] all lie inside the buffer
    const long long vmax = (long long)g.n_vec - 1;
    const long long Tl = (long long)fz_.T;
    f.tile_lo_ok = g.d_min_vec >= 0 ? 0 : (-g.d_min_vec + Tl - 1) / Tl;
    const long long top = vmax - Tl - g.d_max_vec;
    f.tile_hi_ok = top < 0 ? 0 : top / Tl + 1;
    f.T = fz_.T;
    fz_.on = true;
}

void MsaPipeline::bind(MsaBufs& b) {
    const uint32_t P = partitions();
    const MsaGeom& g = geom_;
    memset(&b, 0, sizeof(b));
    b.mism = d_mism_.as<uint32_t>();
    b.vbits = d_vbits_.as<uint32_t>();
    b.tbits = d_tbits_.as<uint32_t>();
    b.rankdir = d_rank_.as<uint32_t>();
    b.refc = d_refc_.as<uint8_t>();
    unsigned char* part = d_part_.as<unsigned char>();
    b.part_sz = reinterpret_cast<unsigned long long*>(part);
    b.part_cnt = reinterpret_cast<uint2*>(part + (size_t)P * 16);
    b.part_sym = reinterpret_cast<unsigned long long*>(part + (size_t)P * 24);

    d_varcol_.reserve((size_t)cap_var_ * 4);
    d_runs_.reserve((size_t)(cap_runs_ + 2) * 4);
    d_sym_.reserve((size_t)(cap_runs_ + 2) * 24);  // sym[], varsym[], hardlist[], widelist[], easylist[], emitlist[]
    d_stash_.reserve((size_t)cap_var_ * g.Rp);
    d_altid_.reserve((size_t)cap_var_ * g.Rp * (g.alt32 ? 4 : 2));
    d_leadmask_.reserve((size_t)cap_var_ * (g.Rp / 32) * 4);
    const bool rows_across_lanes = ctx_->narrow_off != 0 || g.R > 160u;  // mirrors the choice of nb in enqueue
    if (rows_across_lanes) {
        d_rowbits_.reserve((size_t)cap_var_ * 8 * (g.Rp / 32) * 4);
        d_seen_.reserve((size_t)cap_var_ * 3 * 4);
    }
    d_symmeta_.reserve((size_t)(cap_runs_ + 2) * (4 + 8 + 8 + 8));
    d_eds_.reserve(cap_eds_);
    d_seds_.reserve(cap_seds_);
    if (fz_.on) {
        // every cluster fills its own region of the temporary stash; 25 % + 64 slots of slack over an even split
        // every duty warp of every cluster fills its own region of the temporary stash (tiles reach the warps in
        // rotation, so the regions fill evenly); 25 % + 64 slots of slack over an even split
        const uint64_t n_regions = (uint64_t)fz_.regions * fz_.DW;
        fz_.capc = (uint32_t)(((uint64_t)cap_var_ + cap_var_ / 4) / n_regions + 64);
        d_fz_tmp_.reserve((size_t)n_regions * fz_.capc * g.Rp);
        d_fz_col_.reserve((size_t)n_regions * fz_.capc * 8);
        fzp_.capc = fz_.capc;
        fzp_.tmp_stash = d_fz_tmp_.as<uint8_t>();
        fzp_.tmp_col = d_fz_col_.as<unsigned long long>();
        fzp_.mism16 = reinterpret_cast<uint16_t*>(d_mism_.p);
    }
    b.varcol = d_varcol_.as<uint32_t>();
    b.runs = d_runs_.as<uint32_t>();
    b.sym = d_sym_.as<uint32_t>();
    b.varsym = b.sym + (size_t)cap_runs_ + 2;
    b.hardlist = b.varsym + (size_t)cap_runs_ + 2;
    b.widelist = b.hardlist + (size_t)cap_runs_ + 2;
    b.easylist = b.widelist + (size_t)cap_runs_ + 2;
    b.emitlist = b.easylist + (size_t)cap_runs_ + 2;
    b.stash = d_stash_.as<uint8_t>();
    b.altid = d_altid_.p;
    b.leadmask = d_leadmask_.as<uint32_t>();
    b.rowbits = rows_across_lanes ? d_rowbits_.as<uint32_t>() : nullptr;
    b.seen = rows_across_lanes ? d_seen_.as<uint32_t>() : nullptr;
    b.id_text = d_id_text_.as<unsigned long long>();
    unsigned char* meta = d_symmeta_.as<unsigned char>();
    const size_t n = (size_t)cap_runs_ + 2;
    b.sym_edsz = reinterpret_cast<unsigned long long*>(meta);
    b.eds_off = reinterpret_cast<unsigned long long*>(meta + n * 8);
    b.seds_off = reinterpret_cast<unsigned long long*>(meta + n * 16);
    b.sym_nalts = reinterpret_cast<uint32_t*>(meta + n * 24);
    b.eds_out = d_eds_.as<uint8_t>();
    b.seds_out = d_seds_.as<uint8_t>();
    b.group_ws = d_ws_.as<uint8_t>();
    b.status = d_status_.as<MsaStatus>();
    b.cap_var = cap_var_;
    b.cap_runs = cap_runs_;
    b.cap_eds = cap_eds_;
    b.cap_seds = cap_seds_;
}

void MsaPipeline::launch_scan(const MsaBufs& b, bool allow_fused) {
    const MsaGeom& g = geom_;
    cudaStream_t s = ctx_->stream;
    if (allow_fused && fz_.on && fz_.l2) {
        ctx_->clock.begin("k_scan_l2");
        const unsigned long long* pack = reinterpret_cast<const unsigned long long*>(g.row_off + g.R);
        EDSB_LAUNCH(k_scan_l2<true>, fz_.regions, (fz_.CW + fz_.DW) * 32, 0, s, g, pack, fzp_, b.status);
        ctx_->clock.end();
        ctx_->clock.begin("k_colbits");
        EDSB_LAUNCH(k_colbits, partitions(), kPartThreads, 0, s, g, b.mism, b.vbits, b.tbits, b.refc, b.part_cnt);
        ctx_->clock.end();
        return;
    }
    if (allow_fused && fz_.on) {
        ctx_->clock.begin("k_scan_fused");
        const uint32_t threads = (kFzCW + fz_.PW + fz_.DW) * 32, blocks = fz_.regions * fz_.NC;
#ifdef EDSB_EMU
        FZ_KERNEL(EDSB_LAUNCH(kern, blocks, threads, fz_.smem, s, g, fzp_, b.status));
#else
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(blocks, 1, 1);
        cfg.blockDim = dim3(threads, 1, 1);
        cfg.dynamicSmemBytes = fz_.smem;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = fz_.NC;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = fz_.NC > 1 ? 1 : 0;
        FZ_KERNEL(EDSB_CUDA(cudaLaunchKernelEx(&cfg, kern, g, fzp_, b.status)));
#endif
        ctx_->clock.end();
        ctx_->clock.begin("k_colbits");
        EDSB_LAUNCH(k_colbits, partitions(), kPartThreads, 0, s, g, b.mism, b.vbits, b.tbits, b.refc, b.part_cnt);
        ctx_->clock.end();
        return;
    }
    const uint32_t per_sm = ctx_->scan_blocks_per_sm ? ctx_->scan_blocks_per_sm : 16u;
    const uint32_t tiles = (g.n_chunks + 31) / 32;
    const uint32_t blocks = std::max(1u, std::min((tiles + 7) / 8, (uint32_t)ctx_->sm_count * per_sm));
    // too few column blocks to fill the device and rows to spare: split the rows too (blocks x slices)
    uint32_t slices = 1;
    const uint32_t want = (uint32_t)ctx_->sm_count * per_sm * 2u;
    if (blocks < want && g.R >= 128u) slices = std::min(std::min((want + blocks - 1) / blocks, g.R / 64u), 32u);
    if (ctx_->scan_row_slices) slices = std::max(1u, std::min(ctx_->scan_row_slices, g.R > 1 ? g.R - 1 : 1u));
    if (slices > 1) EDSB_CUDA(cudaMemsetAsync(b.mism, 0, ((size_t)g.n_chunks / 2 + 1) * 4, s));
    ctx_->clock.begin("k_scan");
    const unsigned long long* pack = reinterpret_cast<const unsigned long long*>(g.row_off + g.R);
    const dim3 grid(blocks, slices);
    if (g.R <= (uint32_t)kRowCache) {
        EDSB_LAUNCH(k_scan<true>, grid, kScanThreads, 0, s, g, pack, reinterpret_cast<uint16_t*>(b.mism), b.status);
    } else {
        EDSB_LAUNCH(k_scan<false>, grid, kScanThreads, 0, s, g, pack, reinterpret_cast<uint16_t*>(b.mism), b.status);
    }
    ctx_->clock.end();
    ctx_->clock.begin("k_colbits");
    EDSB_LAUNCH(k_colbits, partitions(), kPartThreads, 0, s, g, b.mism, b.vbits, b.tbits, b.refc, b.part_cnt);
    ctx_->clock.end();
}

void MsaPipeline::run_once(MsaBufs& b) {
    const MsaGeom& g = geom_;
    cudaStream_t s = ctx_->stream;
    const uint32_t P = partitions();
    const uint32_t sms = (uint32_t)ctx_->sm_count;
    EDSB_CUDA(cudaMemsetAsync(b.status, 0, sizeof(MsaStatus), s));
    launch_scan(b, true);

    ctx_->clock.begin("k_compact");
    EDSB_LAUNCH(k_compact, P, kPartThreads, 0, s, g, b);
    ctx_->clock.end();

    // fork: the stash gather (needs the variable-column list) runs beside the symbol-boundary kernels
    // (need the run list); they join before the grouping kernels.
    cudaStream_t s1 = ctx_->serial ? s : ctx_->aux[0], s2 = ctx_->serial ? s : ctx_->aux[1];
    auto after = [&](cudaStream_t waiter, cudaStream_t producer, cudaEvent_t ev) {
        if (waiter == producer) return;
        EDSB_CUDA(cudaEventRecord(ev, producer));
        EDSB_CUDA(cudaStreamWaitEvent(waiter, ev, 0));
    };
    after(s1, s, ctx_->ev[0]);
    if (fz_.on) {
        ctx_->clock.begin("k_restash", s1);
        EDSB_LAUNCH(k_restash, sms * 8u, 256, 0, s1, g, b, fzp_, fz_.regions * fz_.DW);
        ctx_->clock.end();
    } else {
        ctx_->clock.begin("k_stash", s1);
        EDSB_LAUNCH(k_stash, sms * 8u, kStashThreads, 0, s1, g, b);
        ctx_->clock.end();
    }

    ctx_->clock.begin("k_sym_count");
    EDSB_LAUNCH(k_sym_count, P, kPartThreads, 0, s, g, b);
    ctx_->clock.end();
    ctx_->clock.begin("k_sym_scatter");
    EDSB_LAUNCH(k_sym_scatter, P, kPartThreads, 0, s, g, b);
    ctx_->clock.end();
    ctx_->clock.begin("k_finalize");
    EDSB_LAUNCH(k_finalize, 1, 32, 0, s, g, b);
    ctx_->clock.end();

    // ---- variable symbols. Single-column ones (lane per symbol) are handled by k_group / k_emit_var; the
    // rest is queued for k_group2 / k_emit2 (warp per symbol). Warp-path scratch lives in shared memory
    // while it fits, else in a global workspace, plus a stage for the symbol's stash block (>= 2 variable
    // columns; 2 KB when rows are few). k_emit_var needs nb segments of seg_pitch bytes per warp; nb == 0
    // (too many rows for shared memory) turns the lane-per-symbol path off in k_group as well.
    const uint32_t Rq = std::max(32u, pow2_ceil(g.R));
    const uint32_t T = 2u * Rq;
    const size_t group_per_warp = (size_t)Rq * 16 + (size_t)T * 4;
    const size_t emit_per_warp = (size_t)Rq * 12;
    const uint32_t stage_bytes = std::max<uint32_t>(2048u, 4u * g.Rp);
    const size_t stage_per_warp = (size_t)stage_bytes + kStageCols * 2;
    const size_t smem_budget = std::min<size_t>(ctx_->smem_optin, 200 * 1024);
    // segments: 8 + sum of id widths + R bytes at most, pitch/4 odd (bank spread), room for the 16-byte shift
    uint32_t seg_pitch = (uint32_t)((8 + g.sum_id_width + g.R + 16 + 3) & ~3ull);
    if (((seg_pitch >> 2) & 1u) == 0) seg_pitch += 4;
    uint32_t nb = 0, evw = kSymWarps;
    if (ctx_->narrow_off == 0 && g.R <= 160u) {  // beyond: rows-across-lanes path (per-lane segments would crowd shared memory)
        // warps per block that put the most warps on an SM: a pass of the kernel is lane-serial, so what counts
        // is covering all batches of 32 symbols in as few rounds as possible
        const size_t per_warp = 2048 + 32 * (size_t)seg_pitch, fixed = 4 * (size_t)g.R + 16, room = ctx_->smem_optin;
        uint32_t best = 0;
        for (uint32_t w = kSymWarps; w >= 1; --w) {
            const size_t blk = w * per_warp + fixed;
            if (blk > room) continue;
            const uint32_t resident = w * (uint32_t)std::min<size_t>(room / (blk + 1024), 32 / w ? 32 / w : 1);
            if (resident > best) {
                best = resident;
                evw = w;
            }
        }
        if (best) nb = 32;
    }
    const uint32_t narrow_ok = ctx_->narrow_off == 2 ? 2u : (nb ? 1u : 0u);  // 1 lane per symbol, 0 rows across lanes, 2 all wide
    // rows across lanes: lanes per symbol = words of a row bitset, rounded up to a power of two (at most a warp)
    const uint32_t group_lanes = std::min(32u, pow2_ceil(g.Rp >> 5));
    size_t ev_warp_smem = (2048 + (size_t)nb * seg_pitch + 15) & ~(size_t)15;
    size_t ev_smem = narrow_ok ? evw * ev_warp_smem + (((size_t)g.R * 4 + 15) & ~(size_t)15) : 0;
    if (narrow_ok == 0u && b.rowbits) {
        // rows across lanes: one id-list stage per warp; the row bitsets come from k_group through global memory
        // a symbol's whole SEDS segment + phase, for each of the 32 / GL symbols a warp renders at once
        const size_t per_warp = (size_t)((8 + g.sum_id_width + g.R + 32 + 15) & ~15ull) * (32u / group_lanes);
        evw = kSymWarps;
        while (evw > 1 && evw * per_warp > smem_budget / 2) evw >>= 1;
        ev_warp_smem = evw * per_warp <= smem_budget ? per_warp : 0;
        ev_smem = evw * ev_warp_smem;
    }
    if (id_text_n_ < (uint64_t)g.R + 1) {
        id_text_n_ = (uint64_t)g.R + 1;
        d_id_text_.reserve((size_t)id_text_n_ * 8);
        b.id_text = d_id_text_.as<unsigned long long>();
        EDSB_LAUNCH(k_id_text, 8, kPartThreads, 0, s, d_id_text_.as<unsigned long long>(), (unsigned long long)id_text_n_);
    }

    // warps per block: the choice that keeps the most warps resident on an SM (the warp-per-symbol kernels are
    // latency-bound: at R = 1000 a warp's scratch is 29 KB and 2-warp blocks fit 3 times = 6 warps, 1-warp blocks 7)
    auto best_warps = [&](size_t per_warp) {
        uint32_t best_w = 1, best_resident = 0;
        for (uint32_t w = kSymWarps; w >= 1; w >>= 1) {
            const size_t blk = w * per_warp + 1024;  // + the per-block reservation
            if (blk > smem_budget) continue;
            const uint32_t resident = w * (uint32_t)std::min<size_t>(ctx_->smem_optin / blk, 32);
            if (resident > best_resident) {
                best_resident = resident;
                best_w = w;
            }
        }
        return best_w;
    };
    uint32_t gw = best_warps(group_per_warp + stage_per_warp);
    const bool group_global = gw * group_per_warp > smem_budget;
    const uint32_t group_stage = (!group_global && gw * (group_per_warp + stage_per_warp) <= smem_budget) ? stage_bytes : 0u;
    const size_t group_warp_smem = group_global ? 0 : ((group_per_warp + (group_stage ? stage_per_warp : 0) + 15) & ~(size_t)15);
    uint32_t ew = best_warps(emit_per_warp + stage_per_warp);
    const bool emit_global = ew * emit_per_warp > smem_budget;
    const uint32_t emit_stage = (!emit_global && ew * (emit_per_warp + stage_per_warp) <= smem_budget) ? stage_bytes : 0u;
    const size_t emit_warp_smem = emit_global ? 0 : ((emit_per_warp + (emit_stage ? stage_per_warp : 0) + 15) & ~(size_t)15);
    const size_t group_smem = gw * group_warp_smem, emit_smem = ew * emit_warp_smem;
    // tuple formulation of the multi-column symbols (k_group3 / k_emit3)
    const size_t g3_warp_smem = (tuple_group_smem(g.R) + 15) & ~(size_t)15;
    const size_t e3_warp_smem = (128 + (size_t)kTupleBitWords * 4 + id_list_stage_bytes(g.R) + 15) & ~(size_t)15;
    const uint32_t g3w = kSymWarps, e3w = kSymWarps;
    const size_t g3_smem = g3w * g3_warp_smem, e3_smem = e3w * e3_warp_smem;
    // occupancy and shared-memory attributes depend on the row count and the test switches only: looked up once
    const uint64_t plan_key = ((uint64_t)g.R << 8) ^ ((uint64_t)ctx_->narrow_off << 4) ^ (b.rowbits ? 1u : 0u) ^ ((uint64_t)nb << 40);
    if (sym_plan_key_ != plan_key) {
        sym_plan_key_ = plan_key;
        int g1_occ = 4, g2_occ = 2, ev_occ = 2, e2_occ = 2, g3_occ = 2, e3_occ = 2;
#ifndef EDSB_EMU
        if (group_smem > 48 * 1024)
            EDSB_CUDA(cudaFuncSetAttribute(k_group2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)group_smem));
        if (emit_smem > 48 * 1024)
            EDSB_CUDA(cudaFuncSetAttribute(k_emit2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)emit_smem));
        if (ev_smem > 48 * 1024)
            EDSB_CUDA(cudaFuncSetAttribute(k_emit_var, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ev_smem));
        if (g3_smem > 48 * 1024)
            EDSB_CUDA(cudaFuncSetAttribute(k_group3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g3_smem));
        if (e3_smem > 48 * 1024)
            EDSB_CUDA(cudaFuncSetAttribute(k_emit3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e3_smem));
        EDSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g1_occ, k_group, kPartThreads, 0));
        EDSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g2_occ, k_group2, (int)(gw * 32u), group_smem));
        EDSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ev_occ, k_emit_var, (int)(evw * 32u), ev_smem));
        EDSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&e2_occ, k_emit2, (int)(ew * 32u), emit_smem));
        EDSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g3_occ, k_group3, (int)(g3w * 32u), g3_smem));
        EDSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&e3_occ, k_emit3, (int)(e3w * 32u), e3_smem));
#endif
        sym_occ_[0] = std::max(1, g1_occ);
        sym_occ_[1] = std::max(1, g2_occ);
        sym_occ_[2] = std::max(1, ev_occ);
        sym_occ_[3] = std::max(1, e2_occ);
        sym_occ_[4] = std::max(1, g3_occ);
        sym_occ_[5] = std::max(1, e3_occ);
    }
    const int g1_occ = sym_occ_[0], g2_occ = sym_occ_[1], ev_occ = sym_occ_[2], e2_occ = sym_occ_[3], g3_occ = sym_occ_[4],
              e3_occ = sym_occ_[5];
    const uint32_t g2_blocks = sms * (uint32_t)std::max(1, g2_occ);
    const uint32_t e2_blocks = sms * (uint32_t)std::max(1, e2_occ);
    if (group_global || emit_global) {
        const size_t need = std::max(group_global ? (size_t)g2_blocks * gw * group_per_warp : 0,
                                     emit_global ? (size_t)e2_blocks * ew * emit_per_warp : 0);
        d_ws_.reserve(need);
        b.group_ws = d_ws_.as<uint8_t>();
    }
    // the stash is ready on s1 (k_restash / k_stash ran there): the multi-column symbols start right behind it on s1,
    // the single-column ones on the main stream once it has seen the stash; they join before the hashed row path
    after(s, s1, ctx_->ev[1]);   // main stream: the stash is ready
    after(s1, s, ctx_->ev[7]);   // side stream: the symbol list and the owned range are ready (k_finalize)
    ctx_->clock.begin("k_widelist", s1);
    EDSB_LAUNCH(k_widelist, sms * 2u, kPartThreads, 0, s1, g, b, ctx_->narrow_off == 2 ? 1u : 0u);
    ctx_->clock.end();
    ctx_->clock.begin("k_group3", s1);
    EDSB_LAUNCH(k_group3, sms * (uint32_t)g3_occ, g3w * 32u, g3_smem, s1, g, b, (uint32_t)g3_warp_smem,
                (ctx_->narrow_off == 2 || ctx_->tuple_off) ? 1u : 0u, g.R <= 160u ? 1u : 0u);
    ctx_->clock.end();
    ctx_->clock.begin("k_group");
    EDSB_LAUNCH(k_group, sms * (uint32_t)std::max(1, g1_occ), kPartThreads, 0, s, g, b, narrow_ok, group_lanes);
    ctx_->clock.end();
    after(s, s1, ctx_->ev[6]);   // join: k_group3 done
    ctx_->clock.begin("k_group2");
    EDSB_LAUNCH(k_group2, g2_blocks, gw * 32u, group_smem, s, g, b, Rq, T, group_global ? 1u : 0u, group_stage,
                (uint32_t)group_warp_smem, ctx_->group_cta);
    ctx_->clock.end();

    ctx_->clock.begin("k_size_count");
    EDSB_LAUNCH(k_size_count, P, kPartThreads, 0, s, g, b);
    ctx_->clock.end();
    ctx_->clock.begin("k_size_scatter");
    EDSB_LAUNCH(k_size_scatter, P, kPartThreads, 0, s, g, b);
    ctx_->clock.end();

    // fork: the three emit kernels write disjoint bytes of the outputs
    after(s1, s, ctx_->ev[2]);
    after(s2, s, ctx_->ev[3]);
    ctx_->clock.begin("k_emit_common", s1);
    EDSB_LAUNCH(k_emit_common, sms * 8u, kPartThreads, 0, s1, g, b);
    ctx_->clock.end();
    if (narrow_ok != 2u) {
        ctx_->clock.begin("k_emit_var");
        EDSB_LAUNCH(k_emit_var, sms * (uint32_t)std::max(1, ev_occ), evw * 32u, ev_smem, s, g, b, (uint32_t)ev_warp_smem, nb,
                    nb ? seg_pitch : group_lanes);
        ctx_->clock.end();
    }
    if (g.R > 160u) {  // with few rows k_group3 hands its symbols to k_emit2 (easylist stays empty)
        ctx_->clock.begin("k_emit3", s2);
        EDSB_LAUNCH(k_emit3, sms * (uint32_t)e3_occ, e3w * 32u, e3_smem, s2, g, b, (uint32_t)e3_warp_smem);
        ctx_->clock.end();
    }
    ctx_->clock.begin("k_emit2", s2);
    EDSB_LAUNCH(k_emit2, e2_blocks, ew * 32u, emit_smem, s2, g, b, Rq, emit_global ? 1u : 0u, emit_stage,
                (uint32_t)emit_warp_smem);
    ctx_->clock.end();
    after(s, s1, ctx_->ev[4]);  // join
    after(s, s2, ctx_->ev[5]);

    EDSB_CUDA(cudaMemcpyAsync(h_status_, b.status, sizeof(MsaStatus), cudaMemcpyDeviceToHost, s));
    EDSB_CUDA(cudaStreamSynchronize(s));
    EDSB_CUDA(cudaGetLastError());
}

void MsaPipeline::transform(const eds_msa_view& view, uint32_t l, int leds, eds_buffer* eds_out, eds_buffer* seds_out,
                            eds_msa_stats* stats) {
    ctx_->clock.reset();
    prepare(view, l, leds);
    const MsaGeom& g = geom_;
    // first guesses; a stage that runs out of room reports what it needs and the pipeline is re-run
    cap_var_ = std::max<uint32_t>(cap_var_, std::max<uint32_t>(1024u, g.ncols / 64u));
    cap_runs_ = std::max<uint32_t>(cap_runs_, 2u * cap_var_ + 2u);
    cap_eds_ = std::max<uint64_t>(cap_eds_, (uint64_t)g.ncols + g.ncols / 4 + 4096);
    cap_seds_ = std::max<uint64_t>(cap_seds_, 1u << 20);
    uint32_t retries = 0;
    MsaStatus st;
    for (;;) {
        MsaBufs b;
        bind(b);
        run_once(b);
        st = *h_status_;
        if (st.bad_msa & kBadNewlineLayout)
            throw BadMsa("line breaks are not at the same positions in every row (unequal row lengths or wrap widths)");
        if (st.bad_msa & kBadResidueByte) throw BadMsa("a row is shorter than the first row (line break inside a residue column)");
        if (st.abort == kAbortNone) break;
        if (++retries > 8) throw std::runtime_error("edsparser_b200: buffer sizing did not converge");
        if (st.abort == kAbortVarCap || st.abort == kAbortRunsCap) {
            cap_var_ = std::max<uint32_t>(cap_var_, (uint32_t)std::min<uint64_t>(st.need_var + st.need_var / 16 + 64, 0x7fffffffu));
            if (fz_.on && st.fz_need)  // a cluster's region of the temporary stash overflowed: size for the fullest one
                cap_var_ = std::max<uint32_t>(cap_var_, (uint32_t)std::min<uint64_t>((uint64_t)st.fz_need * fz_.regions * fz_.DW, 0x7fffffffu));
            cap_runs_ = std::max<uint32_t>(cap_runs_, (uint32_t)std::min<uint64_t>(st.need_runs + st.need_runs / 16 + 64, 0xfffffff0u));
        } else {
            cap_eds_ = std::max<uint64_t>(cap_eds_, st.need_eds + st.need_eds / 16 + 4096);
            cap_seds_ = std::max<uint64_t>(cap_seds_, st.need_seds + st.need_seds / 16 + 4096);
        }
    }
    if (st.halo_fail) throw HaloError("a symbol of the owned range does not resolve inside the window; widen the halo");
    ctx_->clock.resolve();
    if (eds_out) {
        eds_out->data = d_eds_.as<uint8_t>();
        eds_out->bytes = st.eds_total;
    }
    if (seds_out) {
        seds_out->data = d_seds_.as<uint8_t>();
        seds_out->bytes = st.seds_total;
    }
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->n_variable_cols = st.n_var;
        stats->n_runs = st.n_runs;
        stats->n_symbols = st.k_hi - st.k_lo;
        stats->n_variable = st.n_var_syms;
        stats->n_alternatives = st.n_alts;
        stats->first_open_col = st.first_open_col;
        stats->eds_bytes = st.eds_total;
        stats->seds_bytes = st.seds_total;
        stats->eds_lead_bytes = (uint64_t)(st.lead_hi - st.lead_lo) + st.lead_close;
        stats->tail_open_common = st.tail_open;
        stats->gpu_launches = ctx_->clock.launches;
        stats->retries = retries;
        stats->n_hashed_symbols = st.n_hard;
    }
}

void MsaPipeline::conserved_bits(const eds_msa_view& view, uint8_t* out_bits, uint64_t out_bytes) {
    ctx_->clock.reset();
    prepare(view, 0, 0);
    const MsaGeom& g = geom_;
    const uint64_t need = ((uint64_t)g.ncols + 7) / 8;
    if (out_bytes < need) throw std::invalid_argument("eds_msa_conserved_bits: output too small");
    MsaBufs b;
    bind(b);
    cudaStream_t s = ctx_->stream;
    EDSB_CUDA(cudaMemsetAsync(b.status, 0, sizeof(MsaStatus), s));
    launch_scan(b, false);
    std::vector<uint32_t> words(g.n_words);
    EDSB_CUDA(cudaMemcpyAsync(words.data(), b.vbits, (size_t)g.n_words * 4, cudaMemcpyDeviceToHost, s));
    EDSB_CUDA(cudaMemcpyAsync(h_status_, b.status, sizeof(MsaStatus), cudaMemcpyDeviceToHost, s));
    EDSB_CUDA(cudaStreamSynchronize(s));
    EDSB_CUDA(cudaGetLastError());
    if (h_status_->bad_msa) throw BadMsa("line breaks are not at the same positions in every row");
    for (uint64_t i = 0; i < need; ++i) {
        const uint32_t w = words[i / 4];
        uint8_t v = (uint8_t)~(w >> ((i & 3u) * 8u));
        const uint64_t first = i * 8;
        if (first + 8 > g.ncols) v &= (uint8_t)((1u << (g.ncols - first)) - 1u);
        out_bits[i] = v;
    }
}

// Synthetic alignment text for a column window, laid out like a FASTA file of that window:
// ">seq<r+1>\n" + row segment + "\n" per row (so rows start at arbitrary byte offsets).
void msa_synth(eds_ctx* ctx, uint32_t n_rows, uint64_t total_cols, uint32_t lw, uint64_t col_begin, uint64_t col_count,
               uint64_t seed, uint32_t variable_ppm, eds_msa_view* view) {
    if (n_rows < 2 || lw == 0 || col_count == 0 || col_begin + col_count > total_cols)
        throw std::invalid_argument("eds_msa_synth_device: bad shape");
    const uint64_t u_begin = col_begin + col_begin / lw;
    const uint64_t last = col_begin + col_count - 1;
    const uint64_t row_bytes = last + last / lw - u_begin + 1;
    std::vector<uint64_t>& rows = ctx->synth_rows;
    rows.assign(n_rows, 0);
    std::string headers;
    uint64_t at = 0;
    std::vector<uint64_t> header_at(n_rows);
    for (uint32_t r = 0; r < n_rows; ++r) {
        const std::string h = ">seq" + std::to_string(r + 1) + "\n";
        header_at[r] = at;
        at += h.size();
        rows[r] = at;
        at += row_bytes + 1;
        headers += h;
    }
    const uint64_t total = at;
    ctx->synth_text.reserve(total + 64);
    uint8_t* text = ctx->synth_text.as<uint8_t>();
    cudaStream_t s = ctx->stream;
    // headers and the line break after each row: small strided copies from one host staging string
    {
        size_t hpos = 0;
        static const uint8_t nl = '\n';
        for (uint32_t r = 0; r < n_rows; ++r) {
            const size_t hl = (size_t)(rows[r] - header_at[r]);
            EDSB_CUDA(cudaMemcpyAsync(text + header_at[r], headers.data() + hpos, hl, cudaMemcpyHostToDevice, s));
            EDSB_CUDA(cudaMemcpyAsync(text + rows[r] + row_bytes, &nl, 1, cudaMemcpyHostToDevice, s));
            hpos += hl;
        }
        EDSB_CUDA(cudaStreamSynchronize(s));  // `headers` goes out of scope
    }
    DevBuf d_rows;
    d_rows.reserve((size_t)n_rows * 8);
    EDSB_CUDA(cudaMemcpyAsync(d_rows.p, rows.data(), (size_t)n_rows * 8, cudaMemcpyHostToDevice, s));
    const uint32_t bx = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>((row_bytes + 1023) / 1024, 4096));
    EDSB_LAUNCH(k_synth, dim3(bx, std::min(n_rows, 65535u), 1), dim3(256, 1, 1), 0, s, text, d_rows.as<uint64_t>(), n_rows, lw, u_begin,
                row_bytes, seed, variable_ppm);
    EDSB_CUDA(cudaStreamSynchronize(s));
    EDSB_CUDA(cudaGetLastError());
    d_rows.release();
    memset(view, 0, sizeof(*view));
    view->text = text;
    view->text_bytes = total;
    view->row_start = rows.data();
    view->n_rows = n_rows;
    view->line_width = lw;
    view->total_cols = total_cols;
    view->col_begin = col_begin;
    view->col_count = col_count;
    view->own_begin = col_begin;
    view->own_end = col_begin + col_count;
}

}  // namespace edsb
