// "{id,id,...}" lists from bitsets: the source-set text of the SEDS files (msa2eds with many rows, vcf2eds).
// A warp renders 32 bitset words per round into a shared-memory stage laid out at the output's own 16-byte phase
// and writes it out with aligned 16-byte stores; ids are bit index + 1, their digits come from an id -> text table
// that stays L1-resident (k_id_text builds it once per context).
#pragma once
#include "common.cuh"

namespace edsb {

// bytes "id," for the ids of one bitset word (id = 32 w + bit + 1)
__device__ __forceinline__ uint32_t word_id_bytes(uint32_t w, uint32_t bits) {
    if (!bits) return 0;
    const uint32_t lo_id = w * 32u + 1u, wl = decimal_width(lo_id), wh = decimal_width(lo_id + 31u);
    if (wl == wh) return (uint32_t)__popc(bits) * (wl + 1u);
    uint32_t pow = 10;
    for (uint32_t i = 1; i < wl; ++i) pow *= 10u;
    const uint32_t low = bits & low_bits(pow - lo_id);  // ids below 10^wl
    return (uint32_t)__popc(low) * (wl + 1u) + (uint32_t)__popc(bits & ~low) * (wh + 1u);
}

// DW digits and the ',' of one id-text table entry into the stage at byte offset q: DW + 1 byte stores, no branches
template <uint32_t DW>
__device__ __forceinline__ void put_id(uint8_t* stage, uint32_t q, unsigned long long e) {
    const uint32_t lo4 = (uint32_t)e, hi4 = (uint32_t)(e >> 32);
#ifdef EDSB_EMU
    uint8_t* const d = stage + q;
#pragma unroll
    for (uint32_t j = 0; j <= DW; ++j) d[j] = (uint8_t)((j < 4u ? lo4 >> (8u * j) : hi4 >> (8u * (j - 4u))));
#else
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(stage) + q;
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(lo4) : "memory");
    asm volatile("st.shared.u8 [%0+1], %1;" ::"r"(a), "r"(lo4 >> 8) : "memory");
    if (DW >= 2u) asm volatile("st.shared.u8 [%0+2], %1;" ::"r"(a), "r"(lo4 >> 16) : "memory");
    if (DW >= 3u) asm volatile("st.shared.u8 [%0+3], %1;" ::"r"(a), "r"(lo4 >> 24) : "memory");
    if (DW >= 4u) asm volatile("st.shared.u8 [%0+4], %1;" ::"r"(a), "r"(hi4) : "memory");
    if (DW >= 5u) asm volatile("st.shared.u8 [%0+5], %1;" ::"r"(a), "r"(hi4 >> 8) : "memory");
    if (DW >= 6u) asm volatile("st.shared.u8 [%0+6], %1;" ::"r"(a), "r"(hi4 >> 16) : "memory");
#endif
}

// id_text[id]: the decimal digits of id, first digit in the low byte, then ','; width in the top byte (0: more than six digits)
static __global__ void k_id_text(unsigned long long* id_text, unsigned long long n) {
    for (unsigned long long id = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x); id < n; id += ((unsigned long long)gridDim.x * blockDim.x)) {
        const uint32_t dw = id < 1000000ull ? decimal_width((uint32_t)id) : 0u;
        unsigned long long e = ((unsigned long long)dw << 56) | ((unsigned long long)(uint32_t)',' << (8 * dw));
        uint32_t v = (uint32_t)id;
        for (int j = (int)dw - 1; j >= 0; --j) {
            e |= (unsigned long long)((uint32_t)'0' + v % 10u) << (8 * j);
            v /= 10u;
        }
        id_text[id] = e;
    }
}

// bytes of the stage a warp needs for lists over ids 1..n_ids (one round = 1024 ids, the phase, the '{')
inline uint32_t id_list_stage_bytes(unsigned long long n_ids) {
    uint32_t id_width = 1;
    for (unsigned long long top = n_ids; top >= 10; top /= 10) ++id_width;
    return ((1024u * (id_width + 1u) + 32u + 15u) / 16u) * 16u;
}

// One warp: the list of the set bits of bits_in[0..W) as "{id,id,...}" = sb bytes at sout[so..so+sb) (sb = 1 + sum of
// word_id_bytes; the caller has it from its size pass). sout must be 16-byte aligned; stage: id_list_stage_bytes().
// lane_per_word_below: rounds of fewer bytes than this are rendered lane-per-word (each lane its own word's ids, one
// after the other); larger ones lane-per-bit, word by word. vcf2eds (four-digit ids, thousands of samples) measured
// the word-by-word form faster for crowded rounds; msa2eds renders every round lane-per-word (~100 warp
// instructions per crowded word against ~10 per id).
__device__ __forceinline__ void warp_render_id_list(uint8_t* stage, const uint32_t* bits_in, uint32_t W,
                                                    const unsigned long long* id_text, uint8_t* sout,
                                                    unsigned long long so, uint32_t sb, uint32_t lane_per_word_below = 1536u) {
    const unsigned lane = threadIdx.x & 31;
    // "{id,id,...}": 32 bitset words per round are rendered into the warp's shared-memory stage at the
    // output's own 16-byte phase, then copied out with aligned 16-byte stores (bytes at the two ragged ends)
    const unsigned long long end = so + sb;
    unsigned long long gpos = so;
    for (uint32_t w0 = 0; w0 < W; w0 += 32) {
        const uint32_t w = w0 + lane;
        uint32_t v = w < W ? bits_in[w] : 0u;
        const uint32_t mine = word_id_bytes(w, v);
        const uint32_t incl = warp_inclusive_scan(mine);
        const uint32_t phase = (uint32_t)(gpos & 15u), open = w0 == 0 ? 1u : 0u;
        const uint32_t tile = open + __shfl_sync(0xffffffffu, incl, 31);
        if (open && lane == 0) stage[phase] = (uint8_t)'{';
        // lane = bit: the ids of one word are rendered side by side (neighbouring shared-memory banks);
        // digits come from the id table (L1-resident), offsets from a popcount when the word's ids share a width
        const uint32_t my_off = phase + open + (incl - mine);
        const uint32_t lo_id = w * 32u + 1u, wl = decimal_width(lo_id);
        const uint32_t my_dw = (wl == decimal_width(lo_id + 31u) && wl <= 6u) ? wl : 0u;
        const uint32_t end_q = (uint32_t)(end - gpos) + phase;
        const uint32_t live = __ballot_sync(0xffffffffu, v != 0);
        const uint32_t id0 = w0 * 32u + lane + 1u, lt = lanemask_lt();
        if (tile < lane_per_word_below) {
            // sparse round (rare-variant carriers): every lane renders the few ids of its own word
            uint32_t q = my_off;
            for (uint32_t rem = v; rem; rem &= rem - 1) {
                const uint32_t id = w * 32u + (uint32_t)__ffs((int)rem);
                const unsigned long long e = __ldg(id_text + id);
                uint32_t dw = (uint32_t)(e >> 56);
                if (dw) {
                    uint8_t* const d = stage + q;
                    for (uint32_t j = 0; j <= dw; ++j) d[j] = (uint8_t)(e >> (8u * j));
                } else {
                    dw = decimal_width(id);
                    write_decimal(stage + q, id, dw);
                    stage[q + dw] = (uint8_t)',';
                }
                q += dw + 1u;
            }
        } else
        for (uint32_t rest = live; rest; rest &= rest - 1) {
            const int k = __ffs((int)rest) - 1;
            const uint32_t wk = __shfl_sync(0xffffffffu, v, k);
            const uint32_t base_k = __shfl_sync(0xffffffffu, my_off, k);
            const uint32_t dwk = __shfl_sync(0xffffffffu, my_dw, k);
            if ((wk >> lane) & 1u) {
                const uint32_t id = id0 + (uint32_t)k * 32u;
                const uint32_t below = (uint32_t)__popc(wk & lt);
                // dwk is warp-uniform; most ids of a wide matrix have four digits
                if (dwk == 4u) put_id<4>(stage, base_k + below * 5u, __ldg(id_text + id));
                else if (dwk == 3u) put_id<3>(stage, base_k + below * 4u, __ldg(id_text + id));
                else if (dwk == 5u) put_id<5>(stage, base_k + below * 6u, __ldg(id_text + id));
                else if (dwk == 2u) put_id<2>(stage, base_k + below * 3u, __ldg(id_text + id));
                else if (dwk == 6u) put_id<6>(stage, base_k + below * 7u, __ldg(id_text + id));
                else if (dwk == 1u) put_id<1>(stage, base_k + below * 2u, __ldg(id_text + id));
                else {  // the word straddles a power of ten, or more than six digits
                    uint8_t* const d = stage + base_k + word_id_bytes(w0 + (uint32_t)k, wk & lt);
                    const uint32_t w7 = decimal_width(id);
                    write_decimal(d, id, w7);
                    d[w7] = (uint8_t)',';
                }
            }
        }
        if (tile && end_q <= phase + tile) {  // this round holds the end of the list: its last ',' is the '}'
            __syncwarp();
            if (lane == 0) stage[end_q - 1] = (uint8_t)'}';
        }
        __syncwarp();
        uint8_t* const dst = sout + (gpos - phase);  // 16-byte aligned
        const uint32_t hi = phase + tile;
        for (uint32_t j = lane * 16u; j < hi; j += 512u) {
            if (j >= phase && j + 16u <= hi) {
                *reinterpret_cast<uint4*>(dst + j) = *reinterpret_cast<const uint4*>(stage + j);
            } else {
                const uint32_t lo_b = j > phase ? j : phase, hi_b = j + 16u < hi ? j + 16u : hi;
                for (uint32_t i = lo_b; i < hi_b; ++i) dst[i] = stage[i];
            }
        }
        __syncwarp();
        gpos += tile;
    }
}

}  // namespace edsb
