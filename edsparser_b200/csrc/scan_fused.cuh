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

// rows [first, end) of one word-shift class; a warp instruction covers 32 / T rows (lane = chunk + T * sub)
template <int WS, int T>
__device__ __forceinline__ void fz_rows(const uint8_t* stg, const unsigned long long* s_pack, uint32_t first, uint32_t end,
                                        uint32_t chunk, uint32_t sub, const uint4& ref, uint4& acc) {
    constexpr uint32_t kPitch = 16u * (uint32_t)T + 16u, kRows = 32u / (uint32_t)T;
    uint32_t slot = first + sub;
    for (; slot + 3u * kRows < end; slot += 4u * kRows) {
        uint4 lo[4], hi[4];
        uint32_t bs[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint8_t* row = stg + (size_t)(slot + u * kRows) * kPitch + 16u * chunk;
            lo[u] = *reinterpret_cast<const uint4*>(row);
            hi[u] = *reinterpret_cast<const uint4*>(row + 16);
            bs[u] = (reinterpret_cast<const uint32_t*>(s_pack)[2u * (slot + u * kRows)] & 3u) * 8u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) xor_acc<WS>(lo[u], hi[u], bs[u], ref, acc);
    }
    for (; slot < end; slot += kRows) {
        const uint8_t* row = stg + (size_t)slot * kPitch + 16u * chunk;
        const uint4 lo = *reinterpret_cast<const uint4*>(row), hi = *reinterpret_cast<const uint4*>(row + 16);
        xor_acc<WS>(lo, hi, (reinterpret_cast<const uint32_t*>(s_pack)[2u * slot] & 3u) * 8u, ref, acc);
    }
}

template <int T>
__device__ __forceinline__ void fz_rows_aligned(const uint8_t* stg, uint32_t first, uint32_t end, uint32_t chunk, uint32_t sub,
                                                const uint4& ref, uint4& acc) {
    constexpr uint32_t kPitch = 16u * (uint32_t)T + 16u, kRows = 32u / (uint32_t)T;
    for (uint32_t slot = first + sub; slot < end; slot += kRows) {
        const uint4 lo = *reinterpret_cast<const uint4*>(stg + (size_t)slot * kPitch + 16u * chunk);
        acc.x |= lo.x ^ ref.x;
        acc.y |= lo.y ^ ref.y;
        acc.z |= lo.z ^ ref.z;
        acc.w |= lo.w ^ ref.w;
    }
}

// Warp roles of a CTA: [0, CW) consumers, [CW, CW + PW) producers, [CW + PW, CW + PW + DW) duty warps.
//   producers -> full[s] -> consumers -> red_full[s] -> duty warp (tile % DW) -> empty[s]
// Every arrow is an mbarrier. A duty warp turns the OR of the consumers' mismatch bits into the tile's final mask
// (gaps of row 0, line breaks, window edges), stores it, lists the variable columns and copies them out of the
// stage into its own region of the temporary stash. That is a few hundred dependent instructions per tile — more
// than a tile's time budget — so DW warps take the tiles in rotation while the consumers stream on.
template <int T>
__global__ void __launch_bounds__((kFzCW + kFzMaxPW + kFzMaxDW) * 32, 1) k_scan_fused(MsaGeom g, FzParams f, MsaStatus* st) {
    constexpr uint32_t kFzT = (uint32_t)T, kFzPitch = 16u * kFzT + 16u;
    const uint32_t chunk = (threadIdx.x & 31) % kFzT, sub = (threadIdx.x & 31) / kFzT;
    unsigned char* smem = EDSB_DYN_SMEM();
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t S = f.S, NC = f.NC, RG = f.RG, PW = f.PW, DW = f.DW;
    const uint32_t rank = NC > 1 ? cluster_rank() : 0u;
    const uint32_t cid = NC > 1 ? cluster_id_x() : blockIdx.x, ncl = NC > 1 ? cluster_count_x() : gridDim.x;
    const uint32_t stage_bytes = f.slot_pitch * kFzPitch;

    uint8_t* stages = smem;
    uint32_t* red16 = reinterpret_cast<uint32_t*>(stages + (size_t)S * stage_bytes);  // [S][32]
    uint32_t* mask_in = red16 + S * 32u;                                              // [2S][NC][32] (NC > 1)
    unsigned long long* s_pack = reinterpret_cast<unsigned long long*>(mask_in + (NC > 1 ? 2u * S * NC * 32u : 0u));
    uint16_t* s_info = reinterpret_cast<uint16_t*>(s_pack + f.slot_pitch);
    uint32_t* s_off16 = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(s_info) + ((RG * 2u + 15u) & ~15u));  // [slot_pitch]
    uint16_t* s_vpos = reinterpret_cast<uint16_t*>(s_off16 + ((f.slot_pitch + 3u) & ~3u));  // [DW][512]
    Mbar* full = reinterpret_cast<Mbar*>(s_vpos + DW * 16u * kFzT);
    Mbar* empty = full + S;
    Mbar* red_full = empty + S;    // [S]
    Mbar* maskbar = red_full + S;  // [2S] cluster exchange of the partial masks

    const uint32_t* meta = f.meta + rank * 8u;
    const uint32_t nslots = meta[0];
    const uint32_t cls0 = meta[1], cls1 = meta[2], cls2 = meta[3], cls3 = meta[4], cls4 = meta[5];
    for (uint32_t i = threadIdx.x; i < nslots; i += blockDim.x) {
        const unsigned long long e = f.pack[(size_t)rank * f.slot_pitch + i];
        s_pack[i] = e;
        s_off16[i] = (uint32_t)(((long long)((e & ~15ull) - (unsigned long long)(uintptr_t)g.text) >> 4) - g.d_min_vec);
    }
    for (uint32_t i = threadIdx.x; i < RG; i += blockDim.x) s_info[i] = f.info[(size_t)rank * RG + i];
    for (uint32_t i = threadIdx.x; i < S * 32u; i += blockDim.x) red16[i] = 0u;
    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < S; ++s) {
            mbar_init(&full[s], (f.mode == 2u && kFzT == 32u) ? f.PWB + (PW - f.PWB) * 32u : ((f.mode == 0 || kFzT != 32u) ? PW : PW * 32u));
            mbar_init(&empty[s], kFzCW + 1);
            mbar_init(&red_full[s], kFzCW);
            mbar_init(&maskbar[2u * s], 1);
            mbar_init(&maskbar[2u * s + 1u], 1);
        }
        mbar_fence_init();
    }
    __syncthreads();
    if (NC > 1) cluster_sync_all();  // peers complete transactions on our barriers: they must exist first

    if (warp < (uint32_t)kFzCW) {
        // ---------------------------------------------------------------- consumer warps: a contiguous share of the slots
        const uint32_t n_rows = nslots - 1u;
        const uint32_t my_lo = 1u + n_rows * warp / (uint32_t)kFzCW, my_hi = 1u + n_rows * (warp + 1u) / (uint32_t)kFzCW;
        uint32_t it = 0;
        for (uint32_t tile = cid; tile < f.n_tiles; tile += ncl, ++it) {
            const uint32_t s = it % S;
            mbar_wait(&full[s], (it / S) & 1u);
            const uint8_t* stg = stages + (size_t)s * stage_bytes;
            const uint4 ref = *reinterpret_cast<const uint4*>(stg + 16u * chunk);  // slot 0 = row 0, shift 0
            uint4 acc = make_uint4(0, 0, 0, 0);
            if (f.all_aligned) {
                fz_rows_aligned<T>(stg, my_lo, my_hi, chunk, sub, ref, acc);
            } else {
                fz_rows<0, T>(stg, s_pack, max(my_lo, cls0), min(my_hi, cls1), chunk, sub, ref, acc);
                fz_rows<1, T>(stg, s_pack, max(my_lo, cls1), min(my_hi, cls2), chunk, sub, ref, acc);
                fz_rows<2, T>(stg, s_pack, max(my_lo, cls2), min(my_hi, cls3), chunk, sub, ref, acc);
                fz_rows<3, T>(stg, s_pack, max(my_lo, cls3), min(my_hi, cls4), chunk, sub, ref, acc);
            }
            // red16[s] is zero again by now: the duty warp of the tile that last used stage s cleared it before it let
            // the stage go, and full[s] completed after that
            const uint32_t nz = nonzero_bytes16(acc);
            if (nz) atomicOr(&red16[s * 32u + chunk], nz);
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&red_full[s]);
                mbar_arrive(&empty[s]);  // this warp is done with the stage
            }
        }
    } else if (warp < (uint32_t)kFzCW + PW) {
        // ---------------------------------------------------------------- producer warps
        // mode 0: a lane issues one bulk copy per row it owns (UBLKCP takes uniform operands: the warp issues its
        // lanes' copies one after the other, which is why the rows are spread over PW warps);
        // mode 1: a warp copies a row with one 16-byte cp.async per lane, the 33rd vectors lane-per-row.
        // mode 2 (T = 32): both engines at once — the first PWB warps feed slots [0, n_bulk) through bulk copies (the TMA
        // unit takes one 528-byte copy per ~28 cycles per SM, which alone caps the kernel near 5 TB/s), the other warps
        // feed the rest with cp.async through the load/store path.
        const uint32_t pw_all = warp - (uint32_t)kFzCW;
        const bool hybrid = f.mode == 2u && kFzT == 32u;
        const bool bulk = hybrid ? pw_all < f.PWB : (f.mode == 0u || kFzT != 32u);
        const uint32_t cut = min(f.n_bulk, nslots);
        const uint32_t lo_slot = hybrid ? (bulk ? 0u : cut) : 0u, hi_slot = hybrid ? (bulk ? cut : nslots) : nslots;
        const uint32_t pw = hybrid ? (bulk ? pw_all : pw_all - f.PWB) : pw_all, PWg = hybrid ? (bulk ? f.PWB : PW - f.PWB) : PW;
        const uint4* vec = reinterpret_cast<const uint4*>(g.text);
        const long long vmax = (long long)g.n_vec - 1;
        const uint8_t* base0 = g.text + g.d_min_vec * 16;  // s_off16[slot] counts 16-byte vectors from here
        uint32_t my_slots = 0;  // slots lo + pw * 32 + lane + 32 * PWg * m of all lanes together
        for (uint32_t base = lo_slot + pw * 32u; base < hi_slot; base += 32u * PWg) my_slots += min(32u, hi_slot - base);
        uint32_t it = 0;
        for (uint32_t tile = cid; tile < f.n_tiles; tile += ncl, ++it) {
            const uint32_t s = it % S;
            if (it >= S) mbar_wait(&empty[s], ((it / S) - 1u) & 1u);
            uint8_t* dst = stages + (size_t)s * stage_bytes;
            const size_t toff = (size_t)tile * (16u * kFzT);
            if ((long long)tile >= f.tile_lo_ok && (long long)tile < f.tile_hi_ok) {
                if (bulk) {
                    if (lane == 0) mbar_arrive_expect_tx(&full[s], my_slots * kFzPitch);
                    __syncwarp();
                    for (uint32_t slot = lo_slot + pw * 32u + lane; slot < hi_slot; slot += 32u * PWg)
                        bulk_g2s(dst + (size_t)slot * kFzPitch, reinterpret_cast<const uint8_t*>((uintptr_t)(s_pack[slot] & ~15ull)) + toff,
                                 kFzPitch, &full[s]);
                } else {
                    // lean issue loop: one shared-memory read (the row's offset in 16-byte units), one 64-bit multiply-add
                    // and the copy per row; the destination steps by a constant
                    const uint8_t* src0 = base0 + toff + 16u * lane;
                    auto d = smem_addr(dst + (size_t)(lo_slot + pw) * kFzPitch + 16u * lane);
#pragma unroll 4
                    for (uint32_t slot = lo_slot + pw; slot < hi_slot; slot += PWg) {
                        cp_async16_at(d, src0 + (size_t)s_off16[slot] * 16u);
                        d += PWg * kFzPitch;
                    }
                    for (uint32_t slot = lo_slot + pw * 32u + lane; slot < hi_slot; slot += 32u * PWg)
                        cp_async16(dst + (size_t)slot * kFzPitch + 16u * kFzT,
                                   reinterpret_cast<const uint8_t*>((uintptr_t)(s_pack[slot] & ~15ull)) + toff + 16u * kFzT);
                    cp_async_arrive_noinc(&full[s]);
                }
            } else {
                for (uint32_t base = lo_slot + pw * 32u; base < hi_slot; base += 32u * PWg) {
                    const uint32_t top = min(hi_slot, base + 32u);
                    for (uint32_t idx = lane; idx < (top - base) * (kFzT + 1u); idx += 32) {
                        const uint32_t slot = base + idx / (kFzT + 1u), k = idx % (kFzT + 1u);
                        const long long d16 = (long long)((s_pack[slot] & ~15ull) - (unsigned long long)(uintptr_t)g.text) >> 4;
                        long long vi = d16 + (long long)tile * kFzT + k;
                        vi = vi < 0 ? 0 : (vi > vmax ? vmax : vi);
                        reinterpret_cast<uint4*>(dst + (size_t)slot * kFzPitch)[k] = ldg_nc(vec + vi);
                    }
                }
                __syncwarp();
                if (!bulk || lane == 0) mbar_arrive(&full[s]);
            }
        }
    } else {
        // ---------------------------------------------------------------- duty warps: lane = 16-byte chunk of the tile
        const uint32_t dw = warp - (uint32_t)kFzCW - PW;
        const uint32_t line = g.lw + 1u;
        const uint64_t pend = (uint64_t)g.a0 + g.row_bytes;
        const uint32_t region = cid * DW + dw;
        uint16_t* vp = s_vpos + dw * (16u * kFzT);
        uint32_t local_cnt = 0, bad = 0, overflow = 0;
        // this lane's four rows of the gather: byte offset of the row inside a stage, and which of the four exist
        uint32_t roff[4], rmask = 0;
        const bool have_rows = 4u * lane < RG && rank * RG + 4u * lane < g.Rp;
#pragma unroll
        for (uint32_t k = 0; k < 4u; ++k) {
            const uint32_t inf = have_rows ? s_info[4u * lane + k] : 0xffffu;
            roff[k] = inf == 0xffffu ? 0u : (inf >> 4) * kFzPitch + (inf & 15u);
            if (inf != 0xffffu) rmask |= 0xffu << (8u * k);
        }
        // position of this lane's chunk inside a text line, kept up to date by addition (no division per tile)
        const uint32_t step = (uint32_t)(((uint64_t)DW * ncl * (16u * kFzT)) % line);
        uint32_t rb;
        {
            const uint64_t t0 = (uint64_t)cid + (uint64_t)dw * ncl;
            const uint64_t v = g.u_begin + t0 * (16u * kFzT) + 16u * lane + 16ull * line - g.a0;  // + 16 lines: never negative
            rb = (uint32_t)(v % line);
        }
        uint32_t it = 0;
        for (uint32_t tile = cid; tile < f.n_tiles; tile += ncl, ++it) {
            if (it % DW != dw) continue;
            const uint32_t s = it % S;
            mbar_wait(&red_full[s], (it / S) & 1u);
            const bool live = lane < kFzT;  // lanes beyond the tile's chunks idle (16-chunk tiles)
            uint32_t nz = red16[s * 32u + lane];
            red16[s * 32u + lane] = 0u;
            const uint8_t* stg = stages + (size_t)s * stage_bytes;  // the stage is held until this warp lets go
            const uint4 ref = *reinterpret_cast<const uint4*>(stg + 16u * (live ? lane : 0u));
            if (NC > 1) {
                // partial mismatch bits of this CTA's rows -> every CTA of the cluster (distributed shared memory).
                // 2S buffers: a peer's bits for tile it + 2S can only arrive after its duty for it + S, hence after OUR
                // bits for it + S, hence after our stage of tile it was released, i.e. after this read.
                const uint32_t mb = it % (2u * S);
                uint32_t* mine = mask_in + (mb * NC + rank) * 32u + lane;
                if (lane == 0) mbar_arrive_expect_tx(&maskbar[mb], NC * 128u);
                for (uint32_t peer = 0; peer < NC; ++peer) st_async_u32(mine, nz, &maskbar[mb], peer);
                mbar_wait(&maskbar[mb], (it / (2u * S)) & 1u);
                nz = 0;
                for (uint32_t p = 0; p < NC; ++p) nz |= mask_in[(mb * NC + p) * 32u + lane];
            }
            // bytes of this chunk that belong to the row segment; where line breaks must be: u % (lw + 1) == lw
            const uint64_t j = (uint64_t)tile * kFzT + lane;
            const uint64_t tp = (uint64_t)tile * (16u * kFzT);
            uint32_t valid = 0xffffu, expect = 0;
            if (tp >= g.a0 && tp + 16u * kFzT <= pend) {
                // the whole tile lies inside the row segment (all but the first and the last tile of a window)
                if (line > 16u) {  // at most one line break in 16 bytes
                    const uint32_t i0 = g.lw - rb;
                    if (i0 < 16u) expect = 1u << i0;
                } else {
                    uint32_t rem = rb;
                    for (uint32_t i = 0; i < 16u; ++i) {
                        if (rem == g.lw) expect |= 1u << i;
                        rem = (rem == g.lw) ? 0u : rem + 1u;
                    }
                }
            } else {
                const uint64_t p0 = j * 16u;
                const uint32_t vlo = p0 >= g.a0 ? 0u : (uint32_t)(g.a0 - p0);
                const uint32_t vhi = pend >= p0 + 16u ? 16u : (pend > p0 ? (uint32_t)(pend - p0) : 0u);
                valid = low_bits(vhi) & ~low_bits(vlo);
                if (vlo < vhi) {
                    const uint64_t u_first = g.u_begin + (p0 + vlo - g.a0);
                    uint32_t rem = (uint32_t)(u_first % (uint64_t)line);
                    for (uint32_t i = vlo; i < vhi; ++i) {
                        if (rem == g.lw) expect |= 1u << i;
                        rem = (rem == g.lw) ? 0u : rem + 1u;
                    }
                }
            }
            rb += step;
            if (rb >= line) rb -= line;
            const uint32_t nl = eq_bytes16(ref, 0x0a0a0a0au);
            if (live && (((nl ^ expect) | (nz & expect)) & valid)) bad |= (uint32_t)kBadNewlineLayout;
            if (!live) valid = 0u;
            const uint32_t mism = (nz | eq_bytes16(ref, 0x2d2d2d2du)) & valid & ~expect;  // differs from row 0, or row 0 is '-'
            if (live && rank == 0 && j < g.n_chunks) f.mism16[j] = (uint16_t)mism;
            const uint32_t cnt = (uint32_t)__popc(mism);
            const uint32_t incl = warp_inclusive_scan(cnt);
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
            if (total) {
                uint32_t at = incl - cnt;
                for (uint32_t bits = mism; bits; bits &= bits - 1u) vp[at++] = (uint16_t)(16u * lane + (uint32_t)__ffs((int)bits) - 1u);
                __syncwarp();
                if (local_cnt + total <= f.capc) {
                    // two columns per round: eight independent byte reads in flight per lane
                    uint8_t* out = f.tmp_stash + ((size_t)region * f.capc + local_cnt) * g.Rp + rank * RG + 4u * lane;
                    for (uint32_t i = 0; i < total; i += 2u) {
                        const uint32_t v0 = vp[i], v1 = vp[min(i + 1u, total - 1u)];
                        const uint8_t* c0 = stg + v0;
                        const uint8_t* c1 = stg + v1;
                        uint32_t w0 = (uint32_t)c0[roff[0]] | ((uint32_t)c0[roff[1]] << 8) | ((uint32_t)c0[roff[2]] << 16) | ((uint32_t)c0[roff[3]] << 24);
                        uint32_t w1 = (uint32_t)c1[roff[0]] | ((uint32_t)c1[roff[1]] << 8) | ((uint32_t)c1[roff[2]] << 16) | ((uint32_t)c1[roff[3]] << 24);
                        w0 &= rmask;
                        w1 &= rmask;
                        if (have_rows) {
                            if (eq_bytes4(w0, 0x0a0a0a0au) | eq_bytes4(w1, 0x0a0a0a0au)) bad |= (uint32_t)kBadResidueByte;
                            *reinterpret_cast<uint32_t*>(out + (size_t)i * g.Rp) = w0;
                            if (i + 1u < total) *reinterpret_cast<uint32_t*>(out + (size_t)(i + 1u) * g.Rp) = w1;
                        }
                    }
                    if (rank == 0)
                        for (uint32_t i = lane; i < total; i += 32)
                            f.tmp_col[(size_t)region * f.capc + local_cnt + i] = tp + vp[i];  // p-space position
                } else {
                    overflow = 1;
                }
                local_cnt += total;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        if (bad) atomicOr(&st->bad_msa, bad);
        if (lane == 0 && rank == 0) {
            f.region_count[region] = overflow ? 0u : local_cnt;
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
        const uint64_t u = g.u_begin + (f.tmp_col[idx] - g.a0);  // position inside the wrapped row -> window column
        const uint32_t c = (uint32_t)(u - u / (uint64_t)(g.lw + 1u) - g.col_begin);
        const uint32_t kg = b.rankdir[c >> 5] + (uint32_t)__popc(b.vbits[c >> 5] & low_bits(c & 31u));
        if (kg >= b.cap_var) continue;
        const uint4* src = reinterpret_cast<const uint4*>(f.tmp_stash + idx * g.Rp);
        uint4* dst = reinterpret_cast<uint4*>(b.stash + (size_t)kg * g.Rp);
        for (uint32_t v = lane; v < v16; v += 32) dst[v] = src[v];
    }
}

}  // namespace edsb
