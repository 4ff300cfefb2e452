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
template <int WS, int T, int H>
__device__ __forceinline__ void fz_rows(const uint8_t* stg, const unsigned long long* s_pack, uint32_t first, uint32_t end,
                                        uint32_t chunk, uint32_t sub, const uint4& ref, uint4& acc) {
    constexpr uint32_t kPitch = 16u * (uint32_t)T * (uint32_t)H + 16u, kRows = 32u / (uint32_t)T;
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

template <int T, int H>
__device__ __forceinline__ void fz_rows_aligned(const uint8_t* stg, uint32_t first, uint32_t end, uint32_t chunk, uint32_t sub,
                                                const uint4& ref, uint4& acc) {
    constexpr uint32_t kPitch = 16u * (uint32_t)T * (uint32_t)H + 16u, kRows = 32u / (uint32_t)T;
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
//
// H = 2 (pairs): the TMA unit takes one bulk copy per ~28 cycles per SM whatever its size, so 528-byte copies cap the
// kernel near 5 TB/s. A CTA therefore takes its tiles in adjacent PAIRS — iteration `it` is tile
// 2 (cid + (it / 2) ncl) + (it & 1) — and the producers fetch both tiles of a row with ONE 1040-byte copy into two
// stages that interleave row by row (row pitch 1040: stage s at + 512 (s & 1)); everything downstream still works tile
// by tile on its own stage index, barriers and all. Only full[even stage] is armed; the odd tile waits on it too (the
// phase cannot advance before the odd stage has been released). S and DW are even in this form.
// Measured (profiles/r02_a_fused_scan.md): slower at the shapes of the benchmark — a pair is refilled only when BOTH its
// stages are free, so a ring that holds 4 stages keeps half of shared memory in flight instead of three quarters
// (R = 100: 0.230 ms against 0.208; R = 1000, where only one pair fits: 11.5 ms against 6.7). With the rows split over
// two CTAs (four pairs fit) pairs do win, 0.28 ms against 0.38, but the cluster exchange costs more than they give.
// Kept behind EDSB_FUSED_PAIR=1.
template <int T, int H, bool DIRECT>
__global__ void __launch_bounds__((kFzCW + kFzMaxPW + kFzMaxDW) * 32, 1) k_scan_fused(MsaGeom g, FzParams f, MsaStatus* st) {
    constexpr uint32_t kFzT = (uint32_t)T, kFzH = (uint32_t)H, kFzPitch = 16u * kFzT * kFzH + 16u, kTileBytes = 16u * kFzT;
    static_assert(H == 1 || (H == 2 && T == 32), "pairs are built for 32-chunk tiles");
    const uint32_t chunk = (threadIdx.x & 31) % kFzT, sub = (threadIdx.x & 31) / kFzT;
    unsigned char* smem = EDSB_DYN_SMEM();
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t S = f.S, NC = f.NC, RG = f.RG, PW = f.PW, DW = f.DW;
    const uint32_t rank = NC > 1 ? cluster_rank() : 0u;
    const uint32_t cid = NC > 1 ? cluster_id_x() : blockIdx.x, ncl = NC > 1 ? cluster_count_x() : gridDim.x;
    const uint32_t stage_bytes = f.stage_slots * (kTileBytes + 16u);  // per stage; a pair of stages interleaves inside 2 of them
    auto stage_ptr = [&](uint32_t s) -> uint8_t* {
        return kFzH == 1u ? smem + (size_t)s * stage_bytes : smem + (size_t)(s >> 1) * (2u * stage_bytes) + (s & 1u) * kTileBytes;
    };
    auto tile_of = [&](uint32_t it) -> uint32_t {
        const unsigned long long t = kFzH == 1u ? (unsigned long long)cid + (unsigned long long)it * ncl
                                                 : 2ull * ((unsigned long long)cid + (unsigned long long)(it >> 1) * ncl) + (it & 1u);
        return t > 0xffffffffull ? 0xffffffffu : (uint32_t)t;
    };

    uint8_t* stages = smem;
    uint32_t* red16 = reinterpret_cast<uint32_t*>(stages + (size_t)S * stage_bytes);  // [S][32]
    uint32_t* mask_in = red16 + S * 32u;                                              // [2S][NC][32] (NC > 1)
    unsigned long long* s_pack = reinterpret_cast<unsigned long long*>(mask_in + (NC > 1 ? 2u * S * NC * 32u : 0u));
    uint16_t* s_info = reinterpret_cast<uint16_t*>(s_pack + f.slot_pitch);
    uint32_t* s_off16 = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(s_info) + ((RG * 2u + 15u) & ~15u));  // [slot_pitch]
    uint16_t* s_vpos = reinterpret_cast<uint16_t*>(s_off16 + ((f.slot_pitch + 3u) & ~3u));  // [DW][512]
    // full: one barrier per stage, or (f.split) one per stage and producer warp — the completions of a stage's ~100 bulk
    // copies then update PW barriers instead of queueing on one
    Mbar* full_base = reinterpret_cast<Mbar*>(s_vpos + DW * 16u * kFzT);  // [S][kFzMaxPW]
    const bool split = f.split && f.mode == 0u && kFzH == 1u && kFzT == 32u;
    const uint32_t fstep = split ? kFzMaxPW : 1u;
    Mbar* full = full_base;  // stage s, part p: full[s * fstep + p]
    Mbar* empty = full_base + S * kFzMaxPW;
    Mbar* red_full = empty + S;    // [S]
    Mbar* maskbar = red_full + S;  // [2S] cluster exchange of the partial masks

    const uint32_t* meta = f.meta + rank * 8u;
    const uint32_t nslots = meta[0];
    // Direct rows: the TMA unit takes one bulk copy per ~28 cycles whatever its size, which caps a ring fed with 528-byte
    // copies at 5.3 TB/s. The last nd slots therefore never enter the ring: the consumer warps load their 16-byte vectors
    // of those rows straight into registers BEFORE they wait for the stage (the loads fly while the stage fills), and
    // the duty warp reads their residues of the variable columns from global memory (L2: just loaded).
    // Measured (profiles/r02_a_fused_scan.md): slower, 0.27 - 0.30 ms against 0.21 at R = 100 — the loads of tile `it` are only
    // issued when the warp reaches tile `it`, so their DRAM latency is paid per tile instead of hidden by the ring; the
    // registers to issue them a tile ahead are not there. Kept behind EDSB_FUSED_DIRECT=n (its own instantiation).
    const uint32_t nd = (DIRECT && kFzT == 32u && kFzH == 1u) ? min(f.n_direct, (nslots - 1u) / 2u) : 0u;
    const uint32_t ns = nslots - nd;  // staged slots
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
            if (split)
                for (uint32_t p = 0; p < PW; ++p) mbar_init(&full[s * fstep + p], 1);
            else
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
        const uint32_t n_rows = ns - 1u;
        const uint32_t my_lo = 1u + n_rows * warp / (uint32_t)kFzCW, my_hi = 1u + n_rows * (warp + 1u) / (uint32_t)kFzCW;
        constexpr uint32_t kDirectPerWarp = DIRECT ? (kFzMaxDirect + (uint32_t)kFzCW - 1u) / (uint32_t)kFzCW : 1u;
        const uint4* gvec = reinterpret_cast<const uint4*>(g.text);
        const long long gvmax = (long long)g.n_vec - 1;
        for (uint32_t it = 0; tile_of(it) < f.n_tiles; ++it) {
            const uint32_t s = it % S;
            uint4 dlo[kDirectPerWarp], dhi[kDirectPerWarp];
            if (DIRECT && nd) {
                const uint32_t tile = tile_of(it);
                const bool inside = (long long)tile >= f.tile_lo_ok && (long long)tile < f.tile_hi_ok;
#pragma unroll
                for (uint32_t k = 0; k < kDirectPerWarp; ++k) {
                    const uint32_t slot = ns + warp + k * (uint32_t)kFzCW;
                    const unsigned long long e = s_pack[slot < nslots ? slot : 0u];
                    long long vi = ((long long)((e & ~15ull) - (unsigned long long)(uintptr_t)g.text) >> 4) + (long long)tile * kFzT + lane;
                    if (inside) {
                        dlo[k] = ldg_nc(gvec + vi);
                        dhi[k] = ldg_nc(gvec + vi + 1);
                    } else {
                        const long long va = vi < 0 ? 0 : (vi > gvmax ? gvmax : vi);
                        long long vb = vi + 1;
                        vb = vb < 0 ? 0 : (vb > gvmax ? gvmax : vb);
                        dlo[k] = ldg_nc(gvec + va);
                        dhi[k] = ldg_nc(gvec + vb);
                    }
                }
            }
            if (split)
                for (uint32_t p = 0; p < PW; ++p) mbar_wait(&full[s * fstep + p], (it / S) & 1u);
            else
                mbar_wait(&full[kFzH == 1u ? s : (s & ~1u)], (it / S) & 1u);
            const uint8_t* stg = stage_ptr(s);
            const uint4 ref = *reinterpret_cast<const uint4*>(stg + 16u * chunk);  // slot 0 = row 0, shift 0
            uint4 acc = make_uint4(0, 0, 0, 0);
            if (DIRECT && nd && !(f.probe & 1u)) {
#pragma unroll
                for (uint32_t k = 0; k < kDirectPerWarp; ++k) {
                    const uint32_t slot = ns + warp + k * (uint32_t)kFzCW;
                    if (slot < nslots) {
                        const uint4 x = realign16(dlo[k], dhi[k], (uint32_t)s_pack[slot] & 15u);
                        acc.x |= x.x ^ ref.x;
                        acc.y |= x.y ^ ref.y;
                        acc.z |= x.z ^ ref.z;
                        acc.w |= x.w ^ ref.w;
                    }
                }
            }
            if (f.probe & 1u) {
            } else if (f.all_aligned) {
                fz_rows_aligned<T, H>(stg, my_lo, my_hi, chunk, sub, ref, acc);
            } else {
                fz_rows<0, T, H>(stg, s_pack, max(my_lo, cls0), min(my_hi, cls1), chunk, sub, ref, acc);
                fz_rows<1, T, H>(stg, s_pack, max(my_lo, cls1), min(my_hi, cls2), chunk, sub, ref, acc);
                fz_rows<2, T, H>(stg, s_pack, max(my_lo, cls2), min(my_hi, cls3), chunk, sub, ref, acc);
                fz_rows<3, T, H>(stg, s_pack, max(my_lo, cls3), min(my_hi, cls4), chunk, sub, ref, acc);
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
        const uint32_t cut = min(f.n_bulk, ns);
        const uint32_t lo_slot = hybrid ? (bulk ? 0u : cut) : 0u, hi_slot = hybrid ? (bulk ? cut : ns) : ns;
        const uint32_t pw = hybrid ? (bulk ? pw_all : pw_all - f.PWB) : pw_all, PWg = hybrid ? (bulk ? f.PWB : PW - f.PWB) : PW;
        const uint4* vec = reinterpret_cast<const uint4*>(g.text);
        const long long vmax = (long long)g.n_vec - 1;
        const uint8_t* base0 = g.text + g.d_min_vec * 16;  // s_off16[slot] counts 16-byte vectors from here
        uint32_t my_slots = 0;  // slots lo + pw * 32 + lane + 32 * PWg * m of all lanes together
        for (uint32_t base = lo_slot + pw * 32u; base < hi_slot; base += 32u * PWg) my_slots += min(32u, hi_slot - base);
        // one tile (H = 1) or a pair of adjacent tiles (H = 2) per step; copy_tiles = how many of them exist
        for (uint32_t it = 0; tile_of(it) < f.n_tiles; it += kFzH) {
            const uint32_t tile = tile_of(it);
            const uint32_t s = it % S;
            const uint32_t n_here = (kFzH == 2u && tile + 1u < f.n_tiles) ? 2u : 1u;
            if (it >= S) {
                mbar_wait(&empty[s], ((it / S) - 1u) & 1u);
                if (kFzH == 2u) mbar_wait(&empty[s + 1u], ((it / S) - 1u) & 1u);
            }
            uint8_t* dst = stage_ptr(s);
            const size_t toff = (size_t)tile * kTileBytes;
            if ((long long)tile >= f.tile_lo_ok && (long long)(tile + n_here - 1u) < f.tile_hi_ok) {
                if (bulk) {
                    const uint32_t copy_bytes = n_here * kTileBytes + 16u;
                    Mbar* fb = &full[split ? s * fstep + pw_all : s];
                    if (lane == 0) mbar_arrive_expect_tx(fb, my_slots * copy_bytes);
                    __syncwarp();
                    for (uint32_t slot = lo_slot + pw * 32u + lane; slot < hi_slot; slot += 32u * PWg)
                        bulk_g2s(dst + (size_t)slot * kFzPitch, reinterpret_cast<const uint8_t*>((uintptr_t)(s_pack[slot] & ~15ull)) + toff,
                                 copy_bytes, fb);
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
                // the first / last tiles of a window: clamped plain loads, vector by vector
                const uint32_t n_vec_row = n_here * kFzT + 1u;
                for (uint32_t base = lo_slot + pw * 32u; base < hi_slot; base += 32u * PWg) {
                    const uint32_t top = min(hi_slot, base + 32u);
                    for (uint32_t idx = lane; idx < (top - base) * n_vec_row; idx += 32) {
                        const uint32_t slot = base + idx / n_vec_row, k = idx % n_vec_row;
                        const long long d16 = (long long)((s_pack[slot] & ~15ull) - (unsigned long long)(uintptr_t)g.text) >> 4;
                        long long vi = d16 + (long long)tile * kFzT + k;
                        vi = vi < 0 ? 0 : (vi > vmax ? vmax : vi);
                        reinterpret_cast<uint4*>(dst + (size_t)slot * kFzPitch)[k] = ldg_nc(vec + vi);
                    }
                }
                __syncwarp();
                if (!bulk || lane == 0) mbar_arrive(&full[split ? s * fstep + pw_all : s]);
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
        const uint8_t* gsrc[4];  // direct rows: p-space byte 0 of the row in global memory (else null: the row is in the stage)
        const bool have_rows = 4u * lane < RG && rank * RG + 4u * lane < g.Rp;
#pragma unroll
        for (uint32_t k = 0; k < 4u; ++k) {
            const uint32_t inf = have_rows ? s_info[4u * lane + k] : 0xffffu;
            roff[k] = inf == 0xffffu ? 0u : (inf >> 4) * kFzPitch + (inf & 15u);
            if (inf != 0xffffu) rmask |= 0xffu << (8u * k);
            gsrc[k] = nullptr;
            if (inf != 0xffffu && (inf >> 4) >= ns) {
                roff[k] = 0u;
                gsrc[k] = reinterpret_cast<const uint8_t*>((uintptr_t)(s_pack[inf >> 4] & ~15ull)) + (inf & 15u);
            }
        }
        // position of this lane's chunk inside a text line, kept up to date by addition (no division per tile)
        // (pairs: DW is even, so a duty warp keeps its parity inside the pairs and still advances DW * ncl tiles a turn)
        const uint32_t step = (uint32_t)(((uint64_t)DW * ncl * (16u * kFzT)) % line);
        uint32_t rb;
        {
            const uint64_t t0 = kFzH == 1u ? (uint64_t)cid + (uint64_t)dw * ncl : 2ull * ((uint64_t)cid + (uint64_t)(dw >> 1) * ncl) + (dw & 1u);
            const uint64_t v = g.u_begin + t0 * (16u * kFzT) + 16u * lane + 16ull * line - g.a0;  // + 16 lines: never negative
            rb = (uint32_t)(v % line);
        }
        for (uint32_t it = 0; tile_of(it) < f.n_tiles; ++it) {
            if (it % DW != dw) continue;
            const uint32_t tile = tile_of(it);
            const uint32_t s = it % S;
            mbar_wait(&red_full[s], (it / S) & 1u);
            const bool live = lane < kFzT;  // lanes beyond the tile's chunks idle (16-chunk tiles)
            uint32_t nz = red16[s * 32u + lane];
            red16[s * 32u + lane] = 0u;
            if (f.probe & 2u) {
                const uint64_t jj = (uint64_t)tile * kFzT + lane;
                if (live && rank == 0 && jj < g.n_chunks) f.mism16[jj] = 0;
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);
                continue;
            }
            const uint8_t* stg = stage_ptr(s);  // the stage is held until this warp lets go
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
                if (f.probe & 4u) {
                } else if (local_cnt + total <= f.capc) {
                    // two columns per round: eight independent byte reads in flight per lane
                    uint8_t* out = f.tmp_stash + ((size_t)region * f.capc + local_cnt) * g.Rp + rank * RG + 4u * lane;
                    for (uint32_t i = 0; i < total; i += 2u) {
                        const uint32_t v0 = vp[i], v1 = vp[min(i + 1u, total - 1u)];
                        const uint8_t* c0 = stg + v0;
                        const uint8_t* c1 = stg + v1;
                        uint32_t w0 = 0, w1 = 0;
#pragma unroll
                        for (uint32_t k = 0; k < 4u; ++k) {
                            const uint8_t* p0 = (DIRECT && gsrc[k]) ? gsrc[k] + tp + v0 : c0 + roff[k];
                            const uint8_t* p1 = (DIRECT && gsrc[k]) ? gsrc[k] + tp + v1 : c1 + roff[k];
                            w0 |= (uint32_t)*p0 << (8u * k);
                            w1 |= (uint32_t)*p1 << (8u * k);
                        }
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
