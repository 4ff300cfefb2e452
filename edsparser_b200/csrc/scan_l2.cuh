// k_scan_l2: the column-conservation scan (msa_transforms.cpp:69-84) and the gather of the variable columns
// (the seek + read of msa_transforms.cpp:266-286) in one pass, with the L1 / L2 caches as the staging buffer.
//
// k_scan_fused stages every tile in shared memory through the TMA unit, and the TMA unit takes one bulk copy per ~28
// cycles per SM whatever its size: with one 528-byte copy per row and tile the fetch alone tops out at 5.3 TB/s
// (EDSB_FUSED_PROBE=3: the ring with no compare and no gather at all runs at 0.82 of the copy peak). Plain 16-byte loads
// into registers do not have that limit (k_scan: 0.88). So here the rows of a tile go straight from global memory into
// the registers of CW consumer warps (lane = 16-byte chunk, the rows split over the warps, unrolled eight deep), the
// warps OR their mismatch bits together in shared memory, and a duty warp — DW of them take the tiles in rotation — turns
// the bits into the tile's final mask and reads the residues of the few variable columns AGAIN from global memory: a CTA
// finishes a tile of 100 rows in about two microseconds, so those lines are still in L1 or L2 (126 MB against the
// ~16 MB all CTAs touch in that time) and DRAM sees every byte once. Nothing holds a stage, so nothing waits for the
// gather: the consumers only need the duty warp's 128-byte OR buffer back, DW tiles later.
//
// Temporary stash, regions, k_restash: as for k_scan_fused (every duty warp of every CTA fills its own region).
#pragma once
#include "scan_fused.h"

namespace edsb {

#ifdef EDSB_EMU
constexpr int kL2MaxCW = 2, kL2MaxDW = 2;
#else
constexpr int kL2MaxCW = 8, kL2MaxDW = 2;
#endif

// rows [0, n) of pk (any word-shift class: the shift is a run-time operand here), U rows in flight at a time. A warp only
// has R / CW rows of a tile, so batches must not break at class boundaries or leave a serial tail: the last batch
// repeats the last row instead (the same OR again, served by L1).
template <int U>
__device__ __forceinline__ void scan_rows_any(const unsigned long long* pk, uint32_t n, long long jj, const uint4& ref, uint4& acc) {
    for (uint32_t r = 0; r < n; r += (uint32_t)U) {
        uint4 lo[U], hi[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned long long v = pk[min(r + (uint32_t)u, n - 1u)];
            const uint4* p = reinterpret_cast<const uint4*>(v & ~15ull) + jj;
            lo[u] = ldg_nc(p);
            hi[u] = ldg_nc(p + 1);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            // (the shift is read again rather than kept: eight registers the loads need more)
            const uint4 x = realign16(lo[u], hi[u], (uint32_t)pk[min(r + (uint32_t)u, n - 1u)] & 15u);
            acc.x |= x.x ^ ref.x;
            acc.y |= x.y ^ ref.y;
            acc.z |= x.z ^ ref.z;
            acc.w |= x.w ^ ref.w;
        }
    }
}

template <bool CACHED>
__global__ void __launch_bounds__((kL2MaxCW + kL2MaxDW) * 32, 2)
    k_scan_l2(MsaGeom g, const unsigned long long* row_pack, FzParams f, MsaStatus* st) {
    __shared__ unsigned long long s_pk[CACHED ? kRowCache : 1];
    __shared__ uint32_t red16[kL2MaxDW * 32];
    __shared__ uint16_t s_vpos[kL2MaxDW * 512];
    __shared__ Mbar full[kL2MaxDW], spare[kL2MaxDW];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t DW = f.DW, CW = (blockDim.x >> 5) - DW;
    if (CACHED)
        for (uint32_t i = threadIdx.x; i < g.R; i += blockDim.x) s_pk[i] = row_pack[i];
    for (uint32_t i = threadIdx.x; i < DW * 32u; i += blockDim.x) red16[i] = 0u;
    if (threadIdx.x == 0) {
        for (uint32_t d = 0; d < DW; ++d) {
            mbar_init(&full[d], CW);
            mbar_init(&spare[d], 1);
        }
        mbar_fence_init();
    }
    __syncthreads();
    const unsigned long long* pk = CACHED ? s_pk : row_pack;
    const uint4* vec = reinterpret_cast<const uint4*>(g.text);
    const long long vmax = (long long)g.n_vec - 1;
    const uint32_t n_tiles = f.n_tiles;

    if (warp < CW) {
        // ---------------------------------------------------------------- consumer warps: a contiguous share of the sorted rows
        const uint32_t n_rows = g.R - 1u;
        const uint32_t my_lo = 1u + (uint32_t)((unsigned long long)n_rows * warp / CW);
        const uint32_t my_hi = 1u + (uint32_t)((unsigned long long)n_rows * (warp + 1u) / CW);
        uint32_t it = 0;
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const uint32_t d = it % DW, q = it / DW;
            const uint32_t j = tile * 32u + lane;
            const long long j0 = (long long)tile * 32;
            const bool interior = j0 + 31 < (long long)g.n_chunks && j0 + g.d_min_vec >= 0 && j0 + 32 + g.d_max_vec <= vmax;
            uint4 acc = make_uint4(0, 0, 0, 0);
            if (j < g.n_chunks && !(f.probe & 1u)) {
                const long long jj = (long long)j;
                if (interior) {
                    const uint4 ref = ldg_nc(reinterpret_cast<const uint4*>(pk[0]) + jj);
                    if (my_hi > my_lo) scan_rows_any<kScanUnroll>(pk + my_lo, my_hi - my_lo, jj, ref, acc);
                } else {
                    // edge tile: rows in file order, every vector index clamped into the buffer (bytes outside the row are
                    // masked by the duty warp)
                    const long long d0 = (long long)g.row_off[0] - (long long)g.a0;
                    long long v0 = jj + (d0 >> 4);
                    v0 = v0 < 0 ? 0 : (v0 > vmax ? vmax : v0);
                    const uint4 ref = ldg_nc(vec + v0);
                    for (uint32_t r = my_lo; r < my_hi; ++r) {
                        const long long dd = (long long)g.row_off[r] - (long long)g.a0;
                        const long long vi = jj + (dd >> 4);
                        const uint32_t sh = (uint32_t)(dd & 15);
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
            }
            // the duty warp cleared red16[d] before it gave the buffer back (tile it - DW)
            if (q >= 1u) mbar_wait(&spare[d], (q - 1u) & 1u);
            const uint32_t nz = nonzero_bytes16(acc);
            if (nz) atomicOr(&red16[d * 32u + lane], nz);
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[d]);
        }
    } else {
        // ---------------------------------------------------------------- duty warps: lane = 16-byte chunk of the tile
        const uint32_t d = warp - CW;
        const uint32_t line = g.lw + 1u;
        const uint64_t pend = (uint64_t)g.a0 + g.row_bytes;
        const uint32_t region = blockIdx.x * DW + d;
        uint16_t* vp = s_vpos + d * 512u;
        uint32_t local_cnt = 0, bad = 0, overflow = 0;
        // this lane's rows of the gather, 128 rows a round: bases of the first round stay in registers
        const uint8_t* base[4];
        uint32_t rmask0 = 0;
#pragma unroll
        for (uint32_t k = 0; k < 4u; ++k) {
            const uint32_t r = 4u * lane + k;
            base[k] = g.text + (r < g.R ? (long long)g.row_off[r] - (long long)g.a0 : 0ll);
            if (r < g.R) rmask0 |= 0xffu << (8u * k);
        }
        for (uint32_t it = d;; it += DW) {
            const unsigned long long t64 = (unsigned long long)blockIdx.x + (unsigned long long)it * gridDim.x;
            if (t64 >= n_tiles) break;
            const uint32_t tile = (uint32_t)t64, q = it / DW;
            mbar_wait(&full[d], q & 1u);
            const uint32_t nz = red16[d * 32u + lane];
            red16[d * 32u + lane] = 0u;
            __syncwarp();
            if (lane == 0) mbar_arrive(&spare[d]);  // the consumers may OR into the buffer again
            const uint32_t j = tile * 32u + lane;
            const long long jj = (long long)j;
            const uint64_t tp = (uint64_t)tile * 512u;
            if (f.probe & 2u) {
                if (j < g.n_chunks) f.mism16[j] = 0;
                continue;
            }
            const bool live = j < g.n_chunks;
            // row 0's 16 bytes of this chunk (clamped like the consumers' loads; L1 hit)
            uint4 ref = make_uint4(0, 0, 0, 0);
            if (live) {
                const long long d0 = (long long)g.row_off[0] - (long long)g.a0;
                long long v0 = jj + (d0 >> 4);
                v0 = v0 < 0 ? 0 : (v0 > vmax ? vmax : v0);
                ref = ldg_nc(vec + v0);
            }
            // bytes of this chunk that belong to the row segment; where line breaks must be: u % (lw + 1) == lw
            const uint64_t p0 = (uint64_t)j * 16u;
            const uint32_t vlo = p0 >= g.a0 ? 0u : (uint32_t)(g.a0 - p0);
            const uint32_t vhi = pend >= p0 + 16u ? 16u : (pend > p0 ? (uint32_t)(pend - p0) : 0u);
            uint32_t valid = live ? (low_bits(vhi) & ~low_bits(vlo)) : 0u;
            uint32_t expect = 0;
            if (live && vlo < vhi) {
                const uint64_t u_first = g.u_begin + (p0 + vlo - g.a0);
                uint32_t rem = (uint32_t)(u_first % (uint64_t)line);
                if (line > 16u && vlo == 0u && vhi == 16u) {  // at most one line break in 16 bytes
                    const uint32_t i0 = g.lw - rem;
                    if (i0 < 16u) expect = 1u << i0;
                } else {
                    for (uint32_t i = vlo; i < vhi; ++i) {
                        if (rem == g.lw) expect |= 1u << i;
                        rem = (rem == g.lw) ? 0u : rem + 1u;
                    }
                }
            }
            const uint32_t nl = eq_bytes16(ref, 0x0a0a0a0au);
            if (((nl ^ expect) | (nz & expect)) & valid) bad |= (uint32_t)kBadNewlineLayout;
            const uint32_t mism = (nz | eq_bytes16(ref, 0x2d2d2d2du)) & valid & ~expect;  // differs from row 0, or row 0 is '-'
            if (live) f.mism16[j] = (uint16_t)mism;
            const uint32_t cnt = (uint32_t)__popc(mism);
            const uint32_t incl = warp_inclusive_scan(cnt);
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
            if (total) {
                uint32_t at = incl - cnt;
                for (uint32_t bits = mism; bits; bits &= bits - 1u) vp[at++] = (uint16_t)(16u * lane + (uint32_t)__ffs((int)bits) - 1u);
                __syncwarp();
                if (f.probe & 4u) {
                } else if (local_cnt + total <= f.capc) {
                    // the residues of the tile's variable columns, read again (L1 / L2: the consumers have just been here);
                    // two columns per round, a lane covers four rows of every 128
                    uint8_t* out = f.tmp_stash + ((size_t)region * f.capc + local_cnt) * g.Rp;
                    for (uint32_t i = 0; i < total; i += 2u) {
                        const uint32_t i1 = min(i + 1u, total - 1u);
                        const uint64_t c0 = tp + vp[i], c1 = tp + vp[i1];
                        for (uint32_t rb = 0; rb < g.Rp; rb += 128u) {
                            const uint32_t r0 = rb + 4u * lane;
                            if (r0 >= g.Rp) break;
                            uint32_t w0 = 0, w1 = 0, rmask = rmask0;
                            if (rb == 0u) {
#pragma unroll
                                for (uint32_t k = 0; k < 4u; ++k) {
                                    w0 |= (uint32_t)base[k][c0] << (8u * k);
                                    w1 |= (uint32_t)base[k][c1] << (8u * k);
                                }
                            } else {
                                rmask = 0;
#pragma unroll
                                for (uint32_t k = 0; k < 4u; ++k) {
                                    const uint32_t r = r0 + k;
                                    if (r < g.R) {
                                        const uint8_t* bp = g.text + ((long long)g.row_off[r] - (long long)g.a0);
                                        w0 |= (uint32_t)bp[c0] << (8u * k);
                                        w1 |= (uint32_t)bp[c1] << (8u * k);
                                        rmask |= 0xffu << (8u * k);
                                    }
                                }
                            }
                            w0 &= rmask;
                            w1 &= rmask;
                            if (eq_bytes4(w0, 0x0a0a0a0au) | eq_bytes4(w1, 0x0a0a0a0au)) bad |= (uint32_t)kBadResidueByte;
                            *reinterpret_cast<uint32_t*>(out + (size_t)i * g.Rp + r0) = w0;
                            if (i1 != i) *reinterpret_cast<uint32_t*>(out + (size_t)i1 * g.Rp + r0) = w1;
                        }
                    }
                    for (uint32_t i = lane; i < total; i += 32) f.tmp_col[(size_t)region * f.capc + local_cnt + i] = tp + vp[i];
                } else {
                    overflow = 1;
                }
                local_cnt += total;
            }
            __syncwarp();
        }
        if (bad) atomicOr(&st->bad_msa, bad);
        if (lane == 0) {
            f.region_count[region] = overflow ? 0u : local_cnt;
            if (overflow) {
                atomicMax(&st->fz_need, local_cnt);
                st->abort = kAbortVarCap;
            }
        }
    }
}

}  // namespace edsb
