// sm_100a asynchronous-copy plumbing for the streaming kernels: mbarrier objects in shared memory, 1-D bulk copies
// global -> shared (cp.async.bulk, the TMA engine: SASS UBLKCP) that complete on an mbarrier, and the thread-block
// cluster pieces (rank, distributed-shared-memory address mapping, st.async with remote complete_tx, cluster barrier).
//
// Under EDSB_EMU (tests/emu, g++: kernel LOGIC on a CPU-only box) the same calls are played by a small software
// mbarrier; clusters are not emulated (the launcher uses cluster size 1 there).
#pragma once
#include "common.cuh"

namespace edsb {

#ifdef EDSB_EMU

struct Mbar {
    int init, pending;
    long long tx;
    unsigned phase;
};
inline std::mutex& mbar_mutex() {
    static std::mutex m;
    return m;
}
inline void mbar_check_(Mbar* b) {
    if (b->pending == 0 && b->tx == 0) {
        ++b->phase;
        b->pending = b->init;
    }
}
inline void mbar_init(Mbar* b, uint32_t count) {
    std::lock_guard<std::mutex> lk(mbar_mutex());
    b->init = b->pending = (int)count;
    b->tx = 0;
    b->phase = 0;
}
inline void mbar_fence_init() {}
inline void mbar_arrive(Mbar* b) {
    std::lock_guard<std::mutex> lk(mbar_mutex());
    --b->pending;
    mbar_check_(b);
}
inline void mbar_arrive_expect_tx(Mbar* b, uint32_t bytes) {
    std::lock_guard<std::mutex> lk(mbar_mutex());
    b->tx += bytes;
    --b->pending;
    mbar_check_(b);
}
inline void mbar_wait(Mbar* b, uint32_t parity) {
    for (;;) {
        {
            std::lock_guard<std::mutex> lk(mbar_mutex());
            if ((b->phase & 1u) != parity) return;
        }
        std::this_thread::yield();
    }
}
inline void bulk_g2s(void* dst, const void* src, uint32_t bytes, Mbar* b) {
    memcpy(dst, src, bytes);
    std::lock_guard<std::mutex> lk(mbar_mutex());
    b->tx -= bytes;
    mbar_check_(b);
}
inline void cp_async16(void* dst, const void* src) { memcpy(dst, src, 16); }
inline void cp_async16_at(uintptr_t dst, const void* src) { memcpy(reinterpret_cast<void*>(dst), src, 16); }
inline uintptr_t smem_addr(void* p) { return reinterpret_cast<uintptr_t>(p); }
inline void cp_async_arrive_noinc(Mbar* b) { mbar_arrive(b); }
inline uint32_t cluster_rank() { return 0; }
inline uint32_t cluster_id_x() { return blockIdx.x; }
inline uint32_t cluster_count_x() { return gridDim.x; }
inline void cluster_sync_all() {}
inline void st_async_u32(void*, uint32_t, Mbar*, uint32_t) { abort(); }  // no clusters under the emulator

#else  // ---------------------------------------------------------------------------------------------- sm_100a

struct __align__(8) Mbar {
    unsigned long long v;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(Mbar* b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
// make the initialised barriers visible to the async proxy and to the other CTAs of the cluster
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(Mbar* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(Mbar* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(Mbar* b, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(b)),
        "r"(parity)
        : "memory");
}
// 1-D bulk copy global -> this CTA's shared memory; dst, src and bytes are multiples of 16
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, Mbar* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(b))
                 : "memory");
}
// 16 bytes global -> shared without a register round trip (LDGSTS, L1 bypassed); completion is reported to an mbarrier
// by cp_async_arrive_noinc: "arrive once every cp.async this thread has issued so far has landed"
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
// the same with the destination as a shared-window address (kept in a register and stepped by the caller)
__device__ __forceinline__ void cp_async16_at(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ uint32_t smem_addr(void* p) { return smem_u32(p); }
__device__ __forceinline__ void cp_async_arrive_noinc(Mbar* b) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_count_x() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 4 bytes into the shared memory of CTA `rank` of this cluster at the address that `dst` has here; the store completes
// 4 transaction bytes on that CTA's copy of barrier `b` (distributed shared memory: no global round trip, no fence)
__device__ __forceinline__ void st_async_u32(void* dst, uint32_t value, Mbar* b, uint32_t rank) {
    uint32_t rd, rb;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rd) : "r"(smem_u32(dst)), "r"(rank));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(smem_u32(b)), "r"(rank));
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(rd), "r"(value), "r"(rb)
                 : "memory");
}

#endif

}  // namespace edsb
