// Device-wide "transform -> exclusive scan -> apply" in two launches over P partitions, the pattern all
// of the l-EDS merge kernels are built from (stream compaction, offsets, segmented run starts).
//
//   Fn::value(i)              -> V   the element's contribution (computed on the fly from the inputs)
//   Fn::apply(i, prefix, v)          called once per element with its EXCLUSIVE prefix and its value
//   Op::identity(), Op::combine(a, b)   associative
// part[] holds one V per partition; part[P] receives the grand total.
#pragma once
#include "common.cuh"

namespace edsb {

struct OpSum64 {
    typedef unsigned long long V;
    static __device__ __forceinline__ V identity() { return 0ull; }
    static __device__ __forceinline__ V combine(V a, V b) { return a + b; }
};
// two packed sums scanned together (eds2leds: strings | symbols << 32 and opens | closes << 32 in ONE pass over the text)
struct U64x2 {
    unsigned long long a, b;
};
struct OpSum64x2 {
    typedef U64x2 V;
    static __device__ __forceinline__ V identity() { return U64x2{0ull, 0ull}; }
    static __device__ __forceinline__ V combine(V x, V y) { return U64x2{x.a + y.a, x.b + y.b}; }
};
struct OpMax64 {
    typedef unsigned long long V;
    static __device__ __forceinline__ V identity() { return 0ull; }
    static __device__ __forceinline__ V combine(V a, V b) { return a > b ? a : b; }
};

#ifdef EDSB_EMU
constexpr int kScanBlock = 64;
constexpr int kScanItems = 2;
#else
#ifndef EDSB_SCAN_BLOCK
#define EDSB_SCAN_BLOCK 256
#endif
#ifndef EDSB_SCAN_ITEMS
#define EDSB_SCAN_ITEMS 8
#endif
constexpr int kScanBlock = EDSB_SCAN_BLOCK;
constexpr int kScanItems = EDSB_SCAN_ITEMS;
#endif

__device__ __forceinline__ unsigned long long shfl_up_v(unsigned long long v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
__device__ __forceinline__ U64x2 shfl_up_v(U64x2 v, int d) {
    return U64x2{__shfl_up_sync(0xffffffffu, v.a, d), __shfl_up_sync(0xffffffffu, v.b, d)};
}

template <typename Op>
__device__ __forceinline__ typename Op::V warp_scan_incl(typename Op::V v) {
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        typename Op::V o = shfl_up_v(v, d);
        if (lane >= (unsigned)d) v = Op::combine(o, v);
    }
    return v;
}

// exclusive scan across the block; total = combination of all. smem: 33 entries.
template <typename Op>
__device__ __forceinline__ typename Op::V block_scan_excl(typename Op::V v, typename Op::V* smem, typename Op::V& total) {
    typedef typename Op::V V;
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    const V inc = warp_scan_incl<Op>(v);
    V excl = shfl_up_v(inc, 1);
    if (lane == 0) excl = Op::identity();
    __syncthreads();
    if (lane == 31) smem[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        const V w = lane < nw ? smem[lane] : Op::identity();
        const V winc = warp_scan_incl<Op>(w);
        V wex = shfl_up_v(winc, 1);
        if (lane == 0) wex = Op::identity();
        smem[lane] = wex;
        if (lane == 31) smem[32] = winc;
    }
    __syncthreads();
    total = smem[32];
    return Op::combine(smem[wid], excl);
}

template <typename Op>
__device__ __forceinline__ typename Op::V block_reduce(typename Op::V v, typename Op::V* smem) {
    typename Op::V total;
    block_scan_excl<Op>(v, smem, total);
    return total;
}

template <typename Op, typename Fn>
__global__ void __launch_bounds__(kScanBlock) k_part_reduce(unsigned long long n, Fn fn, typename Op::V* part) {
    typedef typename Op::V V;
    __shared__ V s_red[33];
    const unsigned long long P = gridDim.x, per = (n + P - 1) / P;
    const unsigned long long lo = min(n, (unsigned long long)blockIdx.x * per), hi = min(n, lo + per);
    V acc = Op::identity();
    for (unsigned long long i = lo + threadIdx.x; i < hi; i += blockDim.x) acc = Op::combine(acc, fn.value(i));
    // order inside a partition does not matter for the partition total only when Op is commutative: both are
    const V tot = block_reduce<Op>(acc, s_red);
    if (threadIdx.x == 0) part[blockIdx.x] = tot;
}

// Each thread owns kScanItems consecutive elements of a batch: one block scan per kScanItems * blockDim.x elements.
template <typename Op, typename Fn>
__global__ void __launch_bounds__(kScanBlock) k_part_apply(unsigned long long n, Fn fn, typename Op::V* part) {
    typedef typename Op::V V;
    __shared__ V s_scan[33];
    const unsigned long long P = gridDim.x, per = (n + P - 1) / P;
    const unsigned long long lo = min(n, (unsigned long long)blockIdx.x * per), hi = min(n, lo + per);
    V mine = Op::identity();
    for (unsigned q = threadIdx.x; q < blockIdx.x; q += blockDim.x) mine = Op::combine(mine, part[q]);
    V base = block_reduce<Op>(mine, s_scan);
    const unsigned long long batch = (unsigned long long)blockDim.x * kScanItems;
    for (unsigned long long i0 = lo; i0 < hi; i0 += batch) {
        const unsigned long long first = i0 + (unsigned long long)threadIdx.x * kScanItems;
        V vals[kScanItems];
        V sum = Op::identity();
#pragma unroll
        for (int j = 0; j < kScanItems; ++j) {
            vals[j] = first + j < hi ? fn.value(first + j) : Op::identity();
            sum = Op::combine(sum, vals[j]);
        }
        V total;
        V run = Op::combine(base, block_scan_excl<Op>(sum, s_scan, total));
#pragma unroll
        for (int j = 0; j < kScanItems; ++j) {
            if (first + j < hi) fn.apply(first + j, run, vals[j]);
            run = Op::combine(run, vals[j]);
        }
        base = Op::combine(base, total);
    }
    if (blockIdx.x == P - 1 && threadIdx.x == 0) part[P] = base;
}

// Host helper: both launches on `stream`; part must hold P + 1 entries. The grand total lands in part[P].
template <typename Op, typename Fn>
inline void device_scan(cudaStream_t stream, unsigned P, unsigned long long n, const Fn& fn, typename Op::V* part) {
    auto reduce = k_part_reduce<Op, Fn>;
    auto apply = k_part_apply<Op, Fn>;
    EDSB_LAUNCH(reduce, P, kScanBlock, 0, stream, n, fn, part);
    EDSB_LAUNCH(apply, P, kScanBlock, 0, stream, n, fn, part);
}

}  // namespace edsb
