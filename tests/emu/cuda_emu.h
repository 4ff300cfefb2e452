// ============================================================================
// TEST INFRASTRUCTURE ONLY — a minimal CUDA execution-model shim so that the
// kernel SOURCES under edsparser_b200/csrc/ can also be compiled with g++ and
// stepped through on a CPU-only CI box (tests/emu/libedsparser_emu.so).
//
// This is NOT a product path and NOT a fallback: libedsparser_b200.so (the
// product) is built by nvcc only, has no CPU code path, and never links or
// loads anything under tests/. The emulator exists so the `-m "not gpu"` test
// tier can check kernel LOGIC (bit tricks, scans, offsets) against the oracle
// before GPU time is spent; parity claims are made on the GPU tier only.
//
// Model: one kernel at a time, blocks run sequentially, every CUDA thread of a
// block is an OS thread from a pool. __syncthreads is a block barrier, warp
// collectives exchange through a per-warp mailbox + warp barrier, so all 32
// lanes of a warp must reach every collective (the kernels are written that
// way: no early return before a collective, full masks only).
// ============================================================================
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#define EDSB_EMU 1

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint4 {
    unsigned x, y, z, w;
};
struct uint2 {
    unsigned x, y;
};
using std::max;
using std::min;
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

namespace emu {

// ---- sense-reversing barrier usable by a varying number of participants ----
class Barrier {
   public:
    void reset(unsigned n) {
        n_ = n;
        count_.store(0);
        gen_.store(0);
    }
    void wait() {
        unsigned g = gen_.load(std::memory_order_acquire);
        if (count_.fetch_add(1, std::memory_order_acq_rel) + 1 == n_) {
            count_.store(0, std::memory_order_relaxed);
            gen_.fetch_add(1, std::memory_order_acq_rel);
            std::lock_guard<std::mutex> lk(m_);
            cv_.notify_all();
        } else {
            int spins = 0;
            while (gen_.load(std::memory_order_acquire) == g) {
                if (++spins < 200) {
                    std::this_thread::yield();
                } else {
                    std::unique_lock<std::mutex> lk(m_);
                    cv_.wait_for(lk, std::chrono::microseconds(200),
                                 [&] { return gen_.load(std::memory_order_acquire) != g; });
                }
            }
        }
    }

   private:
    unsigned n_ = 1;
    std::atomic<unsigned> count_{0};
    std::atomic<unsigned> gen_{0};
    std::mutex m_;
    std::condition_variable cv_;
};

struct WarpBox {
    Barrier bar;
    unsigned long long slot[32];
    unsigned lanes = 32;
};

struct State {
    dim3 grid, block;
    Barrier block_bar;
    std::vector<WarpBox*> warps;
    unsigned char* dyn_smem = nullptr;
};
State& state();

struct Tls {
    dim3 tid, bid;
    unsigned lin = 0;
};
extern thread_local Tls tls;

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);

inline WarpBox& my_warp() { return *state().warps[tls.lin >> 5]; }
inline unsigned my_lane() { return tls.lin & 31; }

template <typename T>
inline T exchange(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    WarpBox& w = my_warp();
    unsigned long long raw = 0;
    memcpy(&raw, &v, sizeof(T));
    w.slot[my_lane()] = raw;
    w.bar.wait();
    T out = v;
    if (src_lane >= 0 && (unsigned)src_lane < w.lanes) {
        unsigned long long r = w.slot[src_lane];
        memcpy(&out, &r, sizeof(T));
    }
    w.bar.wait();
    return out;
}

}  // namespace emu

#define threadIdx (emu::tls.tid)
#define blockIdx (emu::tls.bid)
#define blockDim (emu::state().block)
#define gridDim (emu::state().grid)

inline void __syncthreads() { emu::state().block_bar.wait(); }
// block-wide OR of a predicate: accumulate, barrier, read, barrier, (one thread) reset, barrier
inline int __syncthreads_or(int pred) {
    static std::atomic<int> acc{0};
    if (pred) acc.store(1, std::memory_order_relaxed);
    emu::state().block_bar.wait();
    const int v = acc.load(std::memory_order_relaxed);
    emu::state().block_bar.wait();
    if (emu::tls.lin == 0) acc.store(0, std::memory_order_relaxed);
    emu::state().block_bar.wait();
    return v;
}
inline void __syncwarp(unsigned = 0xffffffffu) { emu::my_warp().bar.wait(); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }

template <typename T>
inline T __shfl_sync(unsigned, T v, int src, int = 32) {
    return emu::exchange(v, src & 31);
}
template <typename T>
inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) {
    int src = (int)emu::my_lane() - (int)d;
    return emu::exchange(v, src < 0 ? (int)emu::my_lane() : src);
}
template <typename T>
inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) {
    int src = (int)emu::my_lane() + (int)d;
    return emu::exchange(v, src > 31 ? (int)emu::my_lane() : src);
}
template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) {
    return emu::exchange(v, (int)(emu::my_lane() ^ (unsigned)m));
}
inline unsigned __ballot_sync(unsigned, int pred) {
    emu::WarpBox& w = emu::my_warp();
    w.slot[emu::my_lane()] = pred ? 1ull : 0ull;
    w.bar.wait();
    unsigned m = 0;
    for (unsigned i = 0; i < w.lanes; ++i)
        if (w.slot[i]) m |= 1u << i;
    w.bar.wait();
    return m;
}
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
inline int __all_sync(unsigned m, int pred) {
    unsigned b = __ballot_sync(m, pred);
    unsigned lanes = emu::my_warp().lanes;
    unsigned full = lanes >= 32 ? 0xffffffffu : ((1u << lanes) - 1u);
    return b == full;
}
template <typename T>
inline unsigned __match_any_sync(unsigned, T v) {
    emu::WarpBox& w = emu::my_warp();
    unsigned long long raw = 0;
    memcpy(&raw, &v, sizeof(T));
    w.slot[emu::my_lane()] = raw;
    w.bar.wait();
    unsigned m = 0;
    for (unsigned i = 0; i < w.lanes; ++i)
        if (w.slot[i] == raw) m |= 1u << i;
    w.bar.wait();
    return m;
}

// or-reduction over the lanes named by `mask` (every lane of the warp calls it, each with the mask of its own group)
inline unsigned emu_reduce_or(unsigned mask, unsigned v) {
    emu::WarpBox& w = emu::my_warp();
    w.slot[emu::my_lane()] = v;
    w.bar.wait();
    unsigned r = 0;
    for (unsigned i = 0; i < w.lanes; ++i)
        if (mask & (1u << i)) r |= (unsigned)w.slot[i];
    w.bar.wait();
    return r;
}

inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline unsigned __brev(unsigned v) {
    unsigned r = 0;
    for (int i = 0; i < 32; ++i)
        if (v & (1u << i)) r |= 1u << (31 - i);
    return r;
}
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) {
    unsigned long long v = ((unsigned long long)hi << 32) | lo;
    return (unsigned)(v >> (sh & 31));
}
inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned sh) {
    unsigned long long v = ((unsigned long long)hi << 32) | lo;
    return (unsigned)((v << (sh & 31)) >> 32);
}
template <typename T>
inline T __ldg(const T* p) {
    return *p;
}

// ---- atomics (relaxed device-scope semantics are enough for the kernels) ----
#define EMU_ATOMIC(name, builtin)                                                                  \
    template <typename T>                                                                          \
    inline T name(T* p, T v) {                                                                     \
        return builtin(p, v, __ATOMIC_SEQ_CST);                                                    \
    }
EMU_ATOMIC(atomicAdd, __atomic_fetch_add)
EMU_ATOMIC(atomicOr, __atomic_fetch_or)
EMU_ATOMIC(atomicAnd, __atomic_fetch_and)
EMU_ATOMIC(atomicExch, __atomic_exchange_n)
template <typename T>
inline T atomicMin(T* p, T v) {
    T cur = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (v < cur && !__atomic_compare_exchange_n(p, &cur, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {
    }
    return cur;
}
template <typename T>
inline T atomicMax(T* p, T v) {
    T cur = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (v > cur && !__atomic_compare_exchange_n(p, &cur, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {
    }
    return cur;
}
template <typename T>
inline T atomicCAS(T* p, T expected, T desired) {
    __atomic_compare_exchange_n(p, &expected, desired, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return expected;
}

// ---- the slice of the CUDA runtime API the host launcher uses ----
typedef int cudaError_t;
typedef void* cudaStream_t;
struct EmuEvent {
    std::chrono::steady_clock::time_point t;
};
typedef EmuEvent* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaStreamNonBlocking = 1, cudaHostAllocDefault = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp {
    int multiProcessorCount;
    size_t sharedMemPerBlockOptin;
    int major, minor;
    char name[64];
};

inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) {
    *n = 1;
    return cudaSuccess;
}
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDevice(int* d) {
    *d = 0;
    return cudaSuccess;
}
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
    memset(p, 0, sizeof(*p));
    p->multiProcessorCount = 1;
    p->sharedMemPerBlockOptin = 227 * 1024;
    p->major = 10;
    strcpy(p->name, "emulated");
    return cudaSuccess;
}
inline cudaError_t cudaMalloc(void** p, size_t n) {
    // page-pad so that 16-byte over-reads the kernels are allowed to make stay in bounds
    *p = calloc(1, n + 64);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
template <typename T>
inline cudaError_t cudaMalloc(T** p, size_t n) {
    return cudaMalloc(reinterpret_cast<void**>(p), n);
}
inline cudaError_t cudaFree(void* p) {
    free(p);
    return cudaSuccess;
}
inline cudaError_t cudaMallocHost(void** p, size_t n) {
    *p = malloc(n ? n : 1);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
template <typename T>
inline cudaError_t cudaMallocHost(T** p, size_t n) {
    return cudaMallocHost(reinterpret_cast<void**>(p), n);
}
inline cudaError_t cudaFreeHost(void* p) {
    free(p);
    return cudaSuccess;
}
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) {
    if (n) memmove(d, s, n);
    return cudaSuccess;
}
inline cudaError_t cudaMemcpyPeerAsync(void* d, int, const void* s, int, size_t n, cudaStream_t = nullptr) {
    memcpy(d, s, n);
    return cudaSuccess;
}
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) {
    if (n) memmove(d, s, n);
    return cudaSuccess;
}
inline cudaError_t cudaMemset(void* d, int v, size_t n) {
    if (n) memset(d, v, n);
    return cudaSuccess;
}
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) {
    if (n) memset(d, v, n);
    return cudaSuccess;
}
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) {
    *s = nullptr;
    return cudaSuccess;
}
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) {
    *e = new EmuEvent();
    return cudaSuccess;
}
enum { cudaEventDisableTiming = 2 };
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) {
    *e = new EmuEvent();
    return cudaSuccess;
}
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) {
    delete e;
    return cudaSuccess;
}
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) {
    e->t = std::chrono::steady_clock::now();
    return cudaSuccess;
}
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
    return cudaSuccess;
}
template <typename F>
inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) {
    return cudaSuccess;
}
