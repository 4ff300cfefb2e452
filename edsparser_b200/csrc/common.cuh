// Shared device/host helpers for the edsparser_b200 kernels (sm_100a only).
//
// The same sources also compile under tests/emu/cuda_emu.h (g++, -DEDSB_EMU) so that the CPU-only
// test tier can step through kernel logic. That build is test infrastructure; the product library is
// nvcc-only and has no CPU code path.
#pragma once
#ifdef EDSB_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include <stdexcept>
#include <string>

namespace edsb {

struct CudaError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

#define EDSB_CUDA(expr)                                                                                   \
    do {                                                                                                  \
        cudaError_t _e = (expr);                                                                          \
        if (_e != cudaSuccess)                                                                            \
            throw ::edsb::CudaError(std::string(#expr) + ": " + cudaGetErrorString(_e));                  \
    } while (0)

#ifdef EDSB_EMU
#define EDSB_LAUNCH(kernel, grid, block, smem, stream, ...) \
    emu::launch(dim3(grid), dim3(block), (smem), [=]() { kernel(__VA_ARGS__); })
#define EDSB_DYN_SMEM() (reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(emu::state().dyn_smem) + 15) & ~uintptr_t(15)))
#else
#define EDSB_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define EDSB_DYN_SMEM() (::edsb::edsb_dyn_smem_)
#endif

// Growable device allocation (never shrinks; contents are NOT preserved on growth).
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    void reserve(size_t bytes) {
        if (bytes <= cap) return;
        if (p) EDSB_CUDA(cudaFree(p));
        p = nullptr;
        cap = 0;
        size_t want = ((bytes + bytes / 8 + 511) / 256) * 256;
        EDSB_CUDA(cudaMalloc(&p, want));
        cap = want;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T* as() const {
        return static_cast<T*>(p);
    }
};

#if defined(__CUDACC__) || defined(EDSB_EMU)

#if defined(__CUDACC__) && !defined(EDSB_EMU)
extern __shared__ __align__(16) unsigned char edsb_dyn_smem_[];
#endif

// 16-byte read of input that is streamed once per kernel: read-only path. L1 allocation is kept on
// purpose: a row that is not 16-byte aligned is read as two overlapping aligned vectors and the
// second one is an L1 hit.
__device__ __forceinline__ uint4 ldg_nc(const uint4* p) {
    // __ldg = ld.global.nc; as an intrinsic (not opaque inline asm) the compiler knows it is a long-latency
    // load and keeps an unrolled batch of them in flight instead of interleaving each with its use.
    return __ldg(p);
}

// bit k of the result = (byte k of w != 0), k = 0..3
__device__ __forceinline__ uint32_t nonzero_bytes4(uint32_t w) {
    uint32_t t = ((w & 0x7f7f7f7fu) + 0x7f7f7f7fu) | w;  // bit 7 of each byte set iff byte != 0
    t &= 0x80808080u;
    return (t * 0x00204081u) >> 28;  // gather bits 7,15,23,31 into bits 0..3
}

__device__ __forceinline__ uint32_t eq_bytes4(uint32_t w, uint32_t splat) { return nonzero_bytes4(w ^ splat) ^ 0xfu; }

__device__ __forceinline__ uint32_t nonzero_bytes16(uint4 v) {
    return nonzero_bytes4(v.x) | (nonzero_bytes4(v.y) << 4) | (nonzero_bytes4(v.z) << 8) | (nonzero_bytes4(v.w) << 12);
}

__device__ __forceinline__ uint32_t eq_bytes16(uint4 v, uint32_t splat) {
    return eq_bytes4(v.x, splat) | (eq_bytes4(v.y, splat) << 4) | (eq_bytes4(v.z, splat) << 8) |
           (eq_bytes4(v.w, splat) << 12);
}

// bytes [s, s+16) of the 32-byte string lo:hi, 0 <= s < 16
__device__ __forceinline__ uint4 realign16(uint4 lo, uint4 hi, uint32_t s) {
    const uint32_t bs = (s & 3u) * 8u;
    uint32_t w0, w1, w2, w3, w4;
    switch (s >> 2) {
        case 0: w0 = lo.x; w1 = lo.y; w2 = lo.z; w3 = lo.w; w4 = hi.x; break;
        case 1: w0 = lo.y; w1 = lo.z; w2 = lo.w; w3 = hi.x; w4 = hi.y; break;
        case 2: w0 = lo.z; w1 = lo.w; w2 = hi.x; w3 = hi.y; w4 = hi.z; break;
        default: w0 = lo.w; w1 = hi.x; w2 = hi.y; w3 = hi.z; w4 = hi.w; break;
    }
    uint4 o;
    o.x = __funnelshift_r(w0, w1, bs);
    o.y = __funnelshift_r(w1, w2, bs);
    o.z = __funnelshift_r(w2, w3, bs);
    o.w = __funnelshift_r(w3, w4, bs);
    return o;
}

__device__ __forceinline__ uint32_t lanemask_lt() { return (1u << (threadIdx.x & 31)) - 1u; }

// mask with bits [0, n) set, 0 <= n <= 32
__device__ __forceinline__ uint32_t low_bits(uint32_t n) { return n >= 32u ? 0xffffffffu : ((1u << n) - 1u); }

template <typename T>
__device__ __forceinline__ T warp_inclusive_scan(T v) {
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= (unsigned)d) v += o;
    }
    return v;
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// Exclusive scan over the block (blockDim.x a multiple of 32, <= 1024). `total` = block sum.
// smem: at least 33 elements of T. Contains __syncthreads(): call from all threads.
template <typename T>
__device__ __forceinline__ T block_exclusive_scan(T v, T* smem, T& total) {
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    T inc = warp_inclusive_scan(v);
    __syncthreads();  // protect smem reuse across back-to-back calls
    if (lane == 31) smem[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        T w = lane < nw ? smem[lane] : T(0);
        T winc = warp_inclusive_scan(w);
        smem[lane] = winc - w;
        if (lane == 31) smem[32] = winc;
    }
    __syncthreads();
    total = smem[32];
    return smem[wid] + inc - v;
}

// Block-wide sum, result valid in every thread. smem: at least 32 elements of T.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* smem) {
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) smem[wid] = v;
    __syncthreads();
    T r = T(0);
    for (unsigned i = 0; i < nw; ++i) r += smem[i];
    return r;
}

__device__ __forceinline__ uint32_t decimal_width(uint32_t v) {
    return 1u + (v >= 10u) + (v >= 100u) + (v >= 1000u) + (v >= 10000u) + (v >= 100000u) + (v >= 1000000u) +
           (v >= 10000000u) + (v >= 100000000u) + (v >= 1000000000u);
}

// writes the decimal digits of v (width w = decimal_width(v)) to dst[0..w)
__device__ __forceinline__ void write_decimal(uint8_t* dst, uint32_t v, uint32_t w) {
    for (int i = (int)w - 1; i >= 0; --i) {
        dst[i] = (uint8_t)('0' + v % 10u);
        v /= 10u;
    }
}

#endif  // __CUDACC__ || EDSB_EMU

}  // namespace edsb
