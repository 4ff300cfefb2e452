// k_scan_fused: the column-conservation scan (msa_transforms.cpp:69-84) and the gather of the variable columns
// (the seek + read of msa_transforms.cpp:266-286) in ONE pass over the alignment.
//
// k_scan + k_stash read the alignment 1.7 times: once to compare, and again — one DRAM burst per residue — to
// fetch the R bytes of every variable column. Here a CTA streams a tile of [rows x 512 bytes of p-space] into
// shared memory with one 1-D bulk copy per row (cp.async.bulk: the TMA engine, completion on an mbarrier),
// compares from shared memory, and while the tile is still there copies the columns it found variable into the
// stash. Nothing is read twice.
//
//   * ring of S stages, one producer warp (waits for an empty stage, arms the stage's barrier with the byte count,
//     issues one bulk copy per row: the aligned 528-byte superset of the row's 512 bytes) and CW consumer warps
//     (lane = 16-byte chunk, warps over rows; rows are sorted by word shift so the funnel shifts select words at
//     compile time, as in k_scan);
//   * deep alignments: the rows of a tile are split over the CTAs of a thread-block cluster (NC <= 8, 128 rows
//     each). Every CTA ORs its rows, the 32 x 16 partial mismatch bits go to every CTA of the cluster through
//     distributed shared memory (st.async + complete_tx on the receiver's mbarrier, double-buffered), and each CTA
//     gathers ITS rows of the variable columns out of its own stage: the stash column is assembled by the cluster;
//   * the slot of a variable column is not known yet (it is its global rank): every cluster owns a region of a
//     temporary stash and fills it in tile order — all CTAs of a cluster count identically, no atomics — and
//     k_restash moves the columns to their ranks once k_compact has them (2 x 1 % of the alignment).
// Tiles whose 16-byte vectors would fall outside the buffer (the first / last of a window) are staged by the
// producer warp with clamped plain loads instead of bulk copies.
#pragma once
#include "scan_fused.h"

namespace edsb {

template <int WS>
__device__ __forceinline__ void fz_rows(const uint8_t* stg, const unsigned long long* s_pack, uint32_t first, uint32_t end,
                                        uint32_t lane, const uint4& ref, uint4& acc) {
    uint32_t slot = first;
    for (; slot + 3u * kFzCW < end; slot += 4u * kFzCW) {
        uint4 lo[4], hi[4];
        uint32_t bs[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint8_t* row = stg + (size_t)(slot + u * kFzCW) * kFzPitch + 16u * lane;
            lo[u] = *reinterpret_cast<const uint4*>(row);
            hi[u] = *reinterpret_cast<const uint4*>(row + 16);
            bs[u] = ((uint32_t)s_pack[slot + u * kFzCW] & 3u) * 8u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) xor_acc<WS>(lo[u], hi[u], bs[u], ref, acc);
    }
    for (; slot < end; slot += kFzCW) {
        const uint8_t* row = stg + (size_t)slot * kFzPitch + 16u * lane;
        const uint4 lo = *reinterpret_cast<const uint4*>(row), hi = *reinterpret_cast<const uint4*>(row + 16);
        xor_acc<WS>(lo, hi, ((uint32_t)s_pack[slot] & 3u) * 8u, ref, acc);
    }
}

__device__ __forceinline__ void fz_rows_aligned(const uint8_t* stg, uint32_t first, uint32_t end, uint32_t lane, const uint4& ref,
                                                uint4& acc) {
    for (uint32_t slot = first; slot < end; slot += kFzCW) {
        const uint4 lo = *reinterpret_cast<const uint4*>(stg + (size_t)slot * kFzPitch + 16u * lane);
        acc.x |= lo.x ^ ref.x;
        acc.y |= lo.y ^ ref.y;
        acc.z |= lo.z ^ ref.z;
        acc.w |= lo.w ^ ref.w;
    }
}

__global__ void __launch_bounds__((kFzCW + 1) * 32, 1) k_scan_fused(MsaGeom g, FzParams f, MsaStatus* st) {
    unsigned char* smem = EDSB_DYN_SMEM();
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t S = f.S, NC = f.NC, RG = f.RG;
    const uint32_t rank = NC > 1 ? cluster_rank() : 0u;
    const uint32_t cid = NC > 1 ? cluster_id_x() : blockIdx.x, ncl = NC > 1 ? cluster_count_x() : gridDim.x;
    const uint32_t stage_bytes = f.slot_pitch * kFzPitch;

    uint8_t* stages = smem;
    uint4* red = reinterpret_cast<uint4*>(stages + (size_t)S * stage_bytes);
    uint32_t* mask_in = reinterpret_cast<uint32_t*>(red + kFzCW * 32);
    unsigned long long* s_pack = reinterpret_cast<unsigned long long*>(mask_in + 2u * NC * 32u);
    uint16_t* s_info = reinterpret_cast<uint16_t*>(s_pack + f.slot_pitch);
    uint16_t* s_vpos = reinterpret_cast<uint16_t*>(reinterpret_cast<unsigned char*>(s_info) + ((RG * 2u + 15u) & ~15u));
    uint32_t* s_misc = reinterpret_cast<uint32_t*>(s_vpos + 16u * kFzT);
    Mbar* full = reinterpret_cast<Mbar*>(s_misc + 4);
    Mbar* empty = full + S;
    Mbar* maskbar = empty + S;  // 2
    Mbar* cbar = maskbar + 2;   // emulator: barrier of the consumer threads

    const uint32_t* meta = f.meta + rank * 8u;
    const uint32_t nslots = meta[0];
    const uint32_t cls0 = meta[1], cls1 = meta[2], cls2 = meta[3], cls3 = meta[4], cls4 = meta[5];
    for (uint32_t i = threadIdx.x; i < nslots; i += blockDim.x) s_pack[i] = f.pack[(size_t)rank * f.slot_pitch + i];
    for (uint32_t i = threadIdx.x; i < RG; i += blockDim.x) s_info[i] = f.info[(size_t)rank * RG + i];
    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kFzCW);
        }
        mbar_init(&maskbar[0], 1);
        mbar_init(&maskbar[1], 1);
        mbar_init(cbar, kFzCW * 32);
        mbar_fence_init();
    }
    __syncthreads();
    if (NC > 1) cluster_sync_all();  // peers complete transactions on our barriers: they must exist first

    if (warp == (uint32_t)kFzCW) {
        // ---------------------------------------------------------------- producer warp
        const uint4* vec = reinterpret_cast<const uint4*>(g.text);
        const long long vmax = (long long)g.n_vec - 1;
        uint32_t it = 0;
        for (uint32_t tile = cid; tile < f.n_tiles; tile += ncl, ++it) {
            const uint32_t s = it % S;
            if (it >= S) mbar_wait(&empty[s], ((it / S) - 1u) & 1u);
            uint8_t* dst = stages + (size_t)s * stage_bytes;
            if ((long long)tile >= f.tile_lo_ok && (long long)tile < f.tile_hi_ok) {
                if (lane == 0) mbar_arrive_expect_tx(&full[s], nslots * kFzPitch);
                __syncwarp();
                for (uint32_t slot = lane; slot < nslots; slot += 32)
                    bulk_g2s(dst + (size_t)slot * kFzPitch,
                             reinterpret_cast<const uint8_t*>((uintptr_t)(s_pack[slot] & ~15ull)) + (size_t)tile * (16u * kFzT), kFzPitch,
                             &full[s]);
            } else {
                for (uint32_t idx = lane; idx < nslots * (kFzT + 1u); idx += 32) {
                    const uint32_t slot = idx / (kFzT + 1u), k = idx % (kFzT + 1u);
                    const long long d16 = (long long)((s_pack[slot] & ~15ull) - (unsigned long long)(uintptr_t)g.text) >> 4;
                    long long vi = d16 + (long long)tile * kFzT + k;
                    vi = vi < 0 ? 0 : (vi > vmax ? vmax : vi);
                    reinterpret_cast<uint4*>(dst + (size_t)slot * kFzPitch)[k] = ldg_nc(vec + vi);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[s]);
            }
        }
    } else {
        // ---------------------------------------------------------------- consumer warps
        uint32_t it = 0, local_cnt = 0, bad = 0, overflow = 0, cphase = 0;
        auto consumer_sync = [&]() {
#ifdef EDSB_EMU
            mbar_arrive(cbar);
            mbar_wait(cbar, cphase);
            cphase ^= 1u;
#else
            asm volatile("bar.sync 1, %0;" ::"n"(kFzCW * 32) : "memory");
#endif
        };
        (void)cphase;
        for (uint32_t tile = cid; tile < f.n_tiles; tile += ncl, ++it) {
            const uint32_t s = it % S;
            mbar_wait(&full[s], (it / S) & 1u);
            const uint8_t* stg = stages + (size_t)s * stage_bytes;
            const uint4 ref = *reinterpret_cast<const uint4*>(stg + 16u * lane);  // slot 0 = row 0, shift 0
            uint4 acc = make_uint4(0, 0, 0, 0);
            if (f.all_aligned) {
                fz_rows_aligned(stg, cls0 + warp, cls4, lane, ref, acc);
            } else {
                fz_rows<0>(stg, s_pack, cls0 + warp, cls1, lane, ref, acc);
                fz_rows<1>(stg, s_pack, cls1 + warp, cls2, lane, ref, acc);
                fz_rows<2>(stg, s_pack, cls2 + warp, cls3, lane, ref, acc);
                fz_rows<3>(stg, s_pack, cls3 + warp, cls4, lane, ref, acc);
            }
            red[warp * 32u + lane] = acc;
            consumer_sync();
            if (warp == 0) {
                uint4 a = red[lane];
                for (uint32_t w = 1; w < (uint32_t)kFzCW; ++w) {
                    const uint4 o = red[w * 32u + lane];
                    a.x |= o.x;
                    a.y |= o.y;
                    a.z |= o.z;
                    a.w |= o.w;
                }
                uint32_t nz = nonzero_bytes16(a);
                if (NC > 1) {
                    // partial mismatch bits of this CTA's rows -> every CTA of the cluster (distributed shared memory)
                    const uint32_t par = it & 1u;
                    uint32_t* mine = mask_in + (par * NC + rank) * 32u + lane;
                    if (lane == 0) mbar_arrive_expect_tx(&maskbar[par], NC * 128u);
                    for (uint32_t peer = 0; peer < NC; ++peer) st_async_u32(mine, nz, &maskbar[par], peer);
                    mbar_wait(&maskbar[par], (it >> 1) & 1u);
                    nz = 0;
                    for (uint32_t p = 0; p < NC; ++p) nz |= mask_in[(par * NC + p) * 32u + lane];
                }
                // bytes of this chunk that belong to the row segment; where line breaks must be: u % (lw + 1) == lw
                const uint64_t j = (uint64_t)tile * kFzT + lane;
                const uint64_t p0 = j * 16u;
                const uint64_t pend = (uint64_t)g.a0 + g.row_bytes;
                const uint32_t vlo = p0 >= g.a0 ? 0u : (uint32_t)(g.a0 - p0);
                const uint32_t vhi = pend >= p0 + 16u ? 16u : (pend > p0 ? (uint32_t)(pend - p0) : 0u);
                const uint32_t valid = low_bits(vhi) & ~low_bits(vlo);
                uint32_t expect = 0;
                if (vlo < vhi) {
                    const uint64_t u_first = g.u_begin + (p0 + vlo - g.a0);
                    uint32_t rem = (uint32_t)(u_first % (uint64_t)(g.lw + 1u));
                    for (uint32_t i = vlo; i < vhi; ++i) {
                        if (rem == g.lw) expect |= 1u << i;
                        rem = (rem == g.lw) ? 0u : rem + 1u;
                    }
                }
                const uint32_t nl = eq_bytes16(ref, 0x0a0a0a0au);
                if (((nl ^ expect) | (nz & expect)) & valid) bad |= (uint32_t)kBadNewlineLayout;
                const uint32_t mism = (nz | eq_bytes16(ref, 0x2d2d2d2du)) & valid & ~expect;  // differs from row 0, or row 0 is '-'
                if (rank == 0 && j < g.n_chunks) f.mism16[j] = (uint16_t)mism;
                const uint32_t cnt = (uint32_t)__popc(mism);
                const uint32_t incl = warp_inclusive_scan(cnt);
                uint32_t at = incl - cnt;
                for (uint32_t bits = mism; bits; bits &= bits - 1u) s_vpos[at++] = (uint16_t)(16u * lane + (uint32_t)__ffs((int)bits) - 1u);
                if (lane == 31) s_misc[0] = incl;
            }
            consumer_sync();
            const uint32_t total = s_misc[0];
            if (total) {
                if (local_cnt + total <= f.capc) {
                    for (uint32_t i = warp; i < total; i += (uint32_t)kFzCW) {
                        const uint32_t vp = s_vpos[i];
                        const size_t slotg = (size_t)cid * f.capc + local_cnt + i;
                        for (uint32_t q = lane; q < RG / 4u; q += 32) {
                            if (rank * RG + 4u * q >= g.Rp) break;
                            uint32_t word = 0;
#pragma unroll
                            for (uint32_t k = 0; k < 4u; ++k) {
                                const uint32_t inf = s_info[4u * q + k];
                                if (inf != 0xffffu) {
                                    const uint32_t ch = stg[(size_t)(inf >> 4) * kFzPitch + (inf & 15u) + vp];
                                    word |= ch << (8u * k);
                                    if (ch == (uint32_t)'\n') bad |= (uint32_t)kBadResidueByte;
                                }
                            }
                            *reinterpret_cast<uint32_t*>(f.tmp_stash + slotg * g.Rp + rank * RG + 4u * q) = word;
                        }
                        if (rank == 0 && lane == 0) {
                            const uint64_t u = g.u_begin + ((uint64_t)tile * (16u * kFzT) + vp - g.a0);
                            f.tmp_col[slotg] = (uint32_t)(u - u / (uint64_t)(g.lw + 1u) - g.col_begin);
                        }
                    }
                } else {
                    overflow = 1;
                }
                local_cnt += total;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);  // this warp is done with the stage
        }
        if (bad) atomicOr(&st->bad_msa, bad);
        if (warp == 0 && lane == 0 && rank == 0) {
            f.region_count[cid] = overflow ? 0u : local_cnt;
            if (overflow) {
                atomicMax(&st->fz_need, local_cnt);
                st->abort = kAbortVarCap;
            }
        }
    }
    if (NC > 1) cluster_sync_all();  // nobody leaves while a peer may still write into its shared memory
}

// k_restash: temporary slots -> stash[rank of the column] once the ranks exist (k_compact). Warp per column.
__global__ void __launch_bounds__(256) k_restash(MsaGeom g, MsaBufs b, FzParams f, uint32_t regions) {
    const MsaStatus* st = b.status;
    if (st->abort) return;
    const uint32_t lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const uint64_t n = (uint64_t)regions * f.capc;
    const uint32_t v16 = g.Rp / 16u;
    for (uint64_t idx = (uint64_t)blockIdx.x * wpb + (threadIdx.x >> 5); idx < n; idx += (uint64_t)gridDim.x * wpb) {
        const uint32_t k = (uint32_t)(idx / f.capc), i = (uint32_t)(idx % f.capc);
        if (i >= f.region_count[k]) continue;
        const uint32_t c = f.tmp_col[idx];
        const uint32_t kg = b.rankdir[c >> 5] + (uint32_t)__popc(b.vbits[c >> 5] & low_bits(c & 31u));
        if (kg >= b.cap_var) continue;
        const uint4* src = reinterpret_cast<const uint4*>(f.tmp_stash + idx * g.Rp);
        uint4* dst = reinterpret_cast<uint4*>(b.stash + (size_t)kg * g.Rp);
        for (uint32_t v = lane; v < v16; v += 32) dst[v] = src[v];
    }
}

}  // namespace edsb
