// Multi-GPU msa2eds behind the C ABI: the alignment's columns shard across the GPUs of one node.
//
//   eds_group  — ONE process drives N devices: a context and a host thread per device, an in-process NCCL
//                communicator (ncclCommInitAll). eds_group_msa_transform_host / _fd: every device receives its
//                column window of the .msa bytes (+ halo, widened and retried on EDS_ERR_HALO), transforms it,
//                the devices all-gather their (eds, seds) byte counts over NCCL, and each copies its slice to its
//                offset of the one output pair (host buffers, or pwrite into the two files).
//   eds_comm   — one process PER GPU (torchrun / mpirun): the caller ships the 128-byte NCCL id from rank 0,
//                every rank builds the communicator here (ncclCommInitRank) and posts its counts behind each
//                transform; offsets are read when the caller is about to write.
// The data path itself has no collective: 16 bytes per rank and step are all that crosses NVLink.
// NCCL is loaded with dlopen("libnccl.so.2") on first use, so single-GPU users do not need it installed.
// Under EDSB_EMU (CPU test tier) the devices are played one after the other and the counts are summed on the host.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#ifndef EDSB_EMU
#include <dlfcn.h>
#endif

#include "ctx.h"
#include "leds.h"
#include "msa.h"
#include "vcf.h"

namespace edsb {
void set_last_error(const std::string& msg);  // capi.cu
}

namespace {

// ---- the slice of NCCL this file uses, resolved at run time ---------------------------------------------------------
typedef struct ncclComm* ncclComm_t;
typedef struct {
    char internal[128];
} ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclUint64 = 5 };

struct Nccl {
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string why;
    bool ok = false;
};

Nccl& nccl() {
    static Nccl n;
    static std::once_flag once;
    std::call_once(once, [] {
#ifdef EDSB_EMU
        n.why = "NCCL is not part of the emulated build";
#else
        void* h = nullptr;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (h) break;
        }
        if (!h) {
            n.why = std::string("NCCL not found (dlopen libnccl.so.2): ") + dlerror();
            return;
        }
        auto sym = [&](const char* s) { return dlsym(h, s); };
        n.GetUniqueId = reinterpret_cast<decltype(n.GetUniqueId)>(sym("ncclGetUniqueId"));
        n.CommInitRank = reinterpret_cast<decltype(n.CommInitRank)>(sym("ncclCommInitRank"));
        n.CommInitAll = reinterpret_cast<decltype(n.CommInitAll)>(sym("ncclCommInitAll"));
        n.CommDestroy = reinterpret_cast<decltype(n.CommDestroy)>(sym("ncclCommDestroy"));
        n.AllGather = reinterpret_cast<decltype(n.AllGather)>(sym("ncclAllGather"));
        n.GroupStart = reinterpret_cast<decltype(n.GroupStart)>(sym("ncclGroupStart"));
        n.GroupEnd = reinterpret_cast<decltype(n.GroupEnd)>(sym("ncclGroupEnd"));
        n.GetErrorString = reinterpret_cast<decltype(n.GetErrorString)>(sym("ncclGetErrorString"));
        n.ok = n.GetUniqueId && n.CommInitRank && n.CommInitAll && n.CommDestroy && n.AllGather && n.GroupStart && n.GroupEnd;
        if (!n.ok) n.why = "libnccl.so.2 lacks an expected symbol";
#endif
    });
    return n;
}

void nccl_check(int rc, const char* what) {
    if (rc == ncclSuccess) return;
    Nccl& n = nccl();
    throw edsb::CudaError(std::string(what) + ": " + (n.GetErrorString ? n.GetErrorString(rc) : "NCCL error"));
}

struct Barrier {
    std::mutex m;
    std::condition_variable cv;
    unsigned n = 1, count = 0, gen = 0;
    void wait() {
        std::unique_lock<std::mutex> lk(m);
        const unsigned g = gen;
        if (++count == n) {
            count = 0;
            ++gen;
            cv.notify_all();
        } else {
            cv.wait(lk, [&] { return gen != g; });
        }
    }
};

// columns [lo, hi) owned by shard k of n, held window [wb, we)
void shard_plan(uint64_t total, uint32_t n, uint32_t k, uint64_t halo, uint64_t& lo, uint64_t& hi, uint64_t& wb, uint64_t& we) {
    lo = (uint64_t)((unsigned __int128)total * k / n);
    hi = (uint64_t)((unsigned __int128)total * (k + 1) / n);
    wb = lo > halo ? lo - halo : 0;
    we = std::min(total, hi + halo);
}

template <typename F>
eds_status guarded_shard(F&& body) {
    try {
        body();
        return EDS_OK;
    } catch (const edsb::BadMsa& e) {
        edsb::set_last_error(e.what());
        return EDS_ERR_BAD_MSA;
    } catch (const edsb::HaloError& e) {
        edsb::set_last_error(e.what());
        return EDS_ERR_HALO;
    } catch (const edsb::CudaError& e) {
        edsb::set_last_error(e.what());
        return EDS_ERR_CUDA;
    } catch (const edsb::BudgetError& e) {
        edsb::set_last_error(e.what());
        return EDS_ERR_BUDGET;
    } catch (const std::invalid_argument& e) {
        edsb::set_last_error(e.what());
        return EDS_ERR_INVALID_ARGUMENT;
    } catch (const std::out_of_range& e) {
        edsb::set_last_error(e.what());
        return EDS_ERR_OUT_OF_RANGE;
    } catch (const std::bad_alloc&) {
        edsb::set_last_error("out of host memory");
        return EDS_ERR_RUNTIME;
    } catch (const std::exception& e) {
        edsb::set_last_error(e.what());
        return EDS_ERR_RUNTIME;
    }
}

}  // namespace

// =====================================================================================================================
struct eds_comm {
    eds_ctx* ctx = nullptr;
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    static constexpr int kRing = 4;
    uint64_t* h_mine[kRing] = {};  // pinned: this rank's (eds, seds) bytes
    uint64_t* h_all[kRing] = {};   // pinned: everyone's, as gathered
    uint64_t* d_mine = nullptr;    // device: kRing x 2
    uint64_t* d_all = nullptr;     // device: kRing x 2 x world
    cudaEvent_t done[kRing] = {};
    // The counts are host values (the transform has returned them), so the exchange depends on nothing the context's
    // stream still holds: it runs on a stream of its own and the next transform does not queue behind the all-gather
    // (which would make every step wait for the slowest rank).
    cudaStream_t side = nullptr;
    uint64_t posted = 0;
};

struct eds_group {
    std::vector<eds_ctx*> ctx;
    std::vector<int> devices;
    std::vector<ncclComm_t> comm;       // empty: single device, or the emulated build
    std::vector<edsb::DevBuf> window;   // per device: the window's rows, 16-byte aligned pitch
    std::vector<uint64_t*> d_counts;    // per device: 2 + 2 n
    std::vector<uint64_t*> h_counts;    // per device, pinned: 2 + 2 n
    edsb::DevBuf gather[2];             // first device: the slices' EDS / SEDS text joined over NVLink (vcf2eds, l > 0)
    void* host_out[2] = {nullptr, nullptr};  // pinned results of the *_view calls, grow-only
    size_t host_out_cap[2] = {0, 0};
    uint8_t* view_slot(int which, uint64_t bytes) {
        if (host_out_cap[which] < bytes + 1) {
            if (host_out[which]) cudaFreeHost(host_out[which]);
            host_out[which] = nullptr;
            host_out_cap[which] = 0;
            const size_t want = ((bytes + bytes / 8 + 4096) / 4096) * 4096;
            EDSB_CUDA(cudaMallocHost(&host_out[which], want));  // unified addressing: every device of the group can write it
            host_out_cap[which] = want;
        }
        return static_cast<uint8_t*>(host_out[which]);
    }
};

extern "C" {

// ---- one process per GPU ---------------------------------------------------------------------------------------------
eds_status eds_nccl_unique_id(uint8_t out[128]) {
    return guarded_shard([&] {
        if (!out) throw std::invalid_argument("eds_nccl_unique_id: null argument");
        Nccl& n = nccl();
        if (!n.ok) throw edsb::CudaError(n.why);
        ncclUniqueId id;
        nccl_check(n.GetUniqueId(&id), "ncclGetUniqueId");
        memcpy(out, id.internal, 128);
    });
}

eds_status eds_comm_create(eds_ctx* ctx, const uint8_t id[128], int rank, int world, eds_comm** out) {
    return guarded_shard([&] {
        if (!ctx || !out || world < 1 || rank < 0 || rank >= world) throw std::invalid_argument("eds_comm_create: bad argument");
        *out = nullptr;
        EDSB_CUDA(cudaSetDevice(ctx->device));
        eds_comm* c = new eds_comm();
        c->ctx = ctx;
        c->rank = rank;
        c->world = world;
        try {
            if (world > 1) {
                if (!id) throw std::invalid_argument("eds_comm_create: null id");
                Nccl& n = nccl();
                if (!n.ok) throw edsb::CudaError(n.why);
                ncclUniqueId uid;
                memcpy(uid.internal, id, 128);
                nccl_check(n.CommInitRank(&c->comm, world, uid, rank), "ncclCommInitRank");
                // msa.cu plan_fused: the scan leaves room for the all-gather's kernel (EDSB_PEER_HEADROOM=bytes, 0 = none)
                const char* ph = getenv("EDSB_PEER_HEADROOM");
                ctx->peer_headroom = ph ? (uint32_t)atoi(ph) : 56u * 1024u;
            }
            for (int i = 0; i < eds_comm::kRing; ++i) {
                EDSB_CUDA(cudaMallocHost(&c->h_mine[i], 16));
                EDSB_CUDA(cudaMallocHost(&c->h_all[i], (size_t)world * 16));
                EDSB_CUDA(cudaEventCreateWithFlags(&c->done[i], cudaEventDisableTiming));
            }
            EDSB_CUDA(cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking));
            EDSB_CUDA(cudaMalloc(&c->d_mine, eds_comm::kRing * 16));
            EDSB_CUDA(cudaMalloc(&c->d_all, (size_t)eds_comm::kRing * world * 16));
        } catch (...) {
            eds_comm_destroy(c);
            throw;
        }
        *out = c;
    });
}

void eds_comm_destroy(eds_comm* c) {
    if (!c) return;
    cudaSetDevice(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream);
    if (c->side) cudaStreamSynchronize(c->side);
    if (c->comm) nccl().CommDestroy(c->comm);
    if (c->side) cudaStreamDestroy(c->side);
    for (int i = 0; i < eds_comm::kRing; ++i) {
        if (c->h_mine[i]) cudaFreeHost(c->h_mine[i]);
        if (c->h_all[i]) cudaFreeHost(c->h_all[i]);
        if (c->done[i]) cudaEventDestroy(c->done[i]);
    }
    if (c->d_mine) cudaFree(c->d_mine);
    if (c->d_all) cudaFree(c->d_all);
    delete c;
}

// Enqueue, behind whatever the context's stream holds: counts -> device, all-gather, gathered counts -> pinned host.
// No host synchronisation; two posts may be in flight.
eds_status eds_comm_post(eds_comm* c, uint64_t eds_bytes, uint64_t seds_bytes) {
    return guarded_shard([&] {
        if (!c) throw std::invalid_argument("eds_comm_post: null comm");
        EDSB_CUDA(cudaSetDevice(c->ctx->device));
        const int k = (int)(c->posted % eds_comm::kRing);
        cudaStream_t s = c->side;
        if (c->posted >= (uint64_t)eds_comm::kRing) EDSB_CUDA(cudaEventSynchronize(c->done[k]));  // the post that used this slot
        c->h_mine[k][0] = eds_bytes;
        c->h_mine[k][1] = seds_bytes;
        if (c->world == 1) {
            c->h_all[k][0] = eds_bytes;
            c->h_all[k][1] = seds_bytes;
        } else {
            uint64_t* dm = c->d_mine + (size_t)k * 2;
            uint64_t* da = c->d_all + (size_t)k * 2 * c->world;
            EDSB_CUDA(cudaMemcpyAsync(dm, c->h_mine[k], 16, cudaMemcpyHostToDevice, s));
            nccl_check(nccl().AllGather(dm, da, 2, ncclUint64, c->comm, s), "ncclAllGather");
            EDSB_CUDA(cudaMemcpyAsync(c->h_all[k], da, (size_t)c->world * 16, cudaMemcpyDeviceToHost, s));
        }
        EDSB_CUDA(cudaEventRecord(c->done[k], s));
        ++c->posted;
    });
}

// Offsets of this rank's slices from the LAST post: {eds offset, seds offset, eds total, seds total}. Waits for it.
eds_status eds_comm_offsets(eds_comm* c, uint64_t out[4]) {
    return guarded_shard([&] {
        if (!c || !out) throw std::invalid_argument("eds_comm_offsets: null argument");
        if (c->posted == 0) throw std::invalid_argument("eds_comm_offsets: nothing was posted");
        EDSB_CUDA(cudaSetDevice(c->ctx->device));
        const int k = (int)((c->posted - 1) % eds_comm::kRing);
        EDSB_CUDA(cudaEventSynchronize(c->done[k]));
        out[0] = out[1] = out[2] = out[3] = 0;
        for (int r = 0; r < c->world; ++r) {
            if (r < c->rank) {
                out[0] += c->h_all[k][2 * r];
                out[1] += c->h_all[k][2 * r + 1];
            }
            out[2] += c->h_all[k][2 * r];
            out[3] += c->h_all[k][2 * r + 1];
        }
    });
}

// Make the context's stream wait for every post still in flight (inside a timed region: no host synchronisation).
eds_status eds_comm_flush(eds_comm* c) {
    return guarded_shard([&] {
        if (!c) throw std::invalid_argument("eds_comm_flush: null comm");
        if (c->posted == 0) return;
        EDSB_CUDA(cudaSetDevice(c->ctx->device));
        // the posts run on the comm's own stream: the context's stream waits for the last one (they complete in order)
        EDSB_CUDA(cudaStreamWaitEvent(c->ctx->stream, c->done[(c->posted - 1) % eds_comm::kRing], 0));
    });
}

// ---- one process, N devices ------------------------------------------------------------------------------------------
eds_status eds_group_create(const int* devices, int n_devices, eds_group** out) {
    return guarded_shard([&] {
        if (!out || n_devices < 1 || n_devices > 64) throw std::invalid_argument("eds_group_create: bad argument");
        *out = nullptr;
        eds_group* g = new eds_group();
        try {
            for (int i = 0; i < n_devices; ++i) {
                const int dev = devices ? devices[i] : i;
                eds_ctx* c = nullptr;
#ifdef EDSB_EMU
                const eds_status rc = eds_ctx_create(0, nullptr, &c);  // the emulator has one device: played in turn
#else
                const eds_status rc = eds_ctx_create(dev, nullptr, &c);
#endif
                if (rc != EDS_OK) throw edsb::CudaError(eds_last_error());
                g->ctx.push_back(c);
                g->devices.push_back(dev);
            }
            g->window.resize(n_devices);
            g->d_counts.assign(n_devices, nullptr);
            g->h_counts.assign(n_devices, nullptr);
            for (int i = 0; i < n_devices; ++i) {
                EDSB_CUDA(cudaSetDevice(g->ctx[i]->device));
                EDSB_CUDA(cudaMalloc(&g->d_counts[i], (size_t)(2 + 2 * n_devices) * 8));
                EDSB_CUDA(cudaMallocHost(&g->h_counts[i], (size_t)(2 + 2 * n_devices) * 8));
            }
#ifndef EDSB_EMU
            // repeated device ids (several contexts sharing a device: tests on a one-GPU box) -> no communicator, the
            // byte counts are exchanged on the host
            std::vector<int> uniq(g->devices);
            std::sort(uniq.begin(), uniq.end());
            const bool distinct = std::adjacent_find(uniq.begin(), uniq.end()) == uniq.end();
            if (n_devices > 1 && distinct) {
                Nccl& n = nccl();
                if (!n.ok) throw edsb::CudaError(n.why);
                g->comm.assign(n_devices, nullptr);
                nccl_check(n.CommInitAll(g->comm.data(), n_devices, g->devices.data()), "ncclCommInitAll");
            }
#endif
        } catch (...) {
            eds_group_destroy(g);
            throw;
        }
        *out = g;
    });
}

void eds_group_destroy(eds_group* g) {
    if (!g) return;
    if (!g->ctx.empty() && g->ctx[0]) {  // buffers of the first device, while its context is still there
        cudaSetDevice(g->ctx[0]->device);
        g->gather[0].release();
        g->gather[1].release();
    }
    for (size_t i = 0; i < g->ctx.size(); ++i) {
        cudaSetDevice(g->ctx[i]->device);
        if (i < g->comm.size() && g->comm[i]) nccl().CommDestroy(g->comm[i]);
        if (i < g->window.size()) g->window[i].release();
        if (i < g->d_counts.size() && g->d_counts[i]) cudaFree(g->d_counts[i]);
        if (i < g->h_counts.size() && g->h_counts[i]) cudaFreeHost(g->h_counts[i]);
        eds_ctx_destroy(g->ctx[i]);
    }
    for (int which = 0; which < 2; ++which)
        if (g->host_out[which]) cudaFreeHost(g->host_out[which]);
    delete g;
}

int eds_group_size(const eds_group* g) { return g ? (int)g->ctx.size() : 0; }
eds_ctx* eds_group_ctx(eds_group* g, int i) { return (g && i >= 0 && i < (int)g->ctx.size()) ? g->ctx[i] : nullptr; }

}  // extern "C"

namespace {

// What the device threads share during one sharded transform.
struct ShardRun {
    eds_group* g;
    const uint8_t* file;
    uint64_t file_bytes;
    uint32_t l;
    int leds;
    uint64_t halo;
    eds_msa_index idx;
    int eds_fd = -1, seds_fd = -1;  // >= 0: pwrite the slices; else into the host buffers below
    uint8_t* h_eds = nullptr;
    uint8_t* h_seds = nullptr;
    uint64_t eds_total = 0, seds_total = 0;
    eds_msa_stats* stats = nullptr;
    Barrier bar;
    std::vector<std::string> error;  // per device
    std::vector<eds_status> status;
    std::vector<uint64_t> counts;    // emulated build: 2 per device
};

void pwrite_all(int fd, const uint8_t* p, uint64_t n, uint64_t off) {
    while (n) {
        const ssize_t w = pwrite(fd, p, (size_t)std::min<uint64_t>(n, 1u << 30), (off_t)off);
        if (w <= 0) throw std::runtime_error("pwrite failed while writing an output slice");
        p += w;
        n -= (uint64_t)w;
        off += (uint64_t)w;
    }
}

// One device's part. Every path reaches the barriers the same number of times (an error is recorded and the thread
// keeps stepping), so no thread is left waiting.
void shard_worker(ShardRun& run, uint32_t k) {
    eds_group* g = run.g;
    const uint32_t n = (uint32_t)g->ctx.size();
    eds_ctx* ctx = g->ctx[k];
    const eds_msa_index& idx = run.idx;
    eds_buffer de{nullptr, 0}, ds{nullptr, 0};
    eds_msa_stats st;
    memset(&st, 0, sizeof(st));
    auto fail = [&](eds_status rc, const std::string& msg) {
        if (run.status[k] == EDS_OK) {
            run.status[k] = rc;
            run.error[k] = msg;
        }
    };
    try {
        EDSB_CUDA(cudaSetDevice(ctx->device));
        cudaStream_t s = ctx->stream;
        uint64_t halo = std::max<uint64_t>(run.halo, (uint64_t)run.l + 2);
        for (int attempt = 0;; ++attempt) {
            uint64_t lo, hi, wb, we;
            shard_plan(idx.n_cols, n, k, halo, lo, hi, wb, we);
            // the window's bytes of every row, rows at a 16-byte aligned pitch (k_scan's one-load-per-row path)
            const uint64_t lw = idx.line_width;
            const uint64_t u_begin = wb + wb / lw, last = we - 1;
            const uint64_t row_bytes = last + last / lw - u_begin + 1;
            const uint64_t pitch = ((row_bytes + 15) & ~15ull) + 16;
            g->window[k].reserve((size_t)(pitch * idx.n_rows + 64));
            uint8_t* dwin = g->window[k].as<uint8_t>();
            std::vector<uint64_t> rows(idx.n_rows);
            for (uint32_t r = 0; r < idx.n_rows; ++r) {
                rows[r] = (uint64_t)r * pitch;
                EDSB_CUDA(cudaMemcpyAsync(dwin + rows[r], run.file + idx.row_start[r] + u_begin, (size_t)row_bytes, cudaMemcpyHostToDevice, s));
            }
            eds_msa_view v;
            memset(&v, 0, sizeof(v));
            v.text = dwin;
            v.text_bytes = pitch * idx.n_rows;
            v.row_start = rows.data();
            v.n_rows = idx.n_rows;
            v.line_width = idx.line_width;
            v.total_cols = idx.n_cols;
            v.col_begin = wb;
            v.col_count = we - wb;
            v.own_begin = lo;
            v.own_end = hi;
            try {
                ctx->msa->transform(v, run.l, run.leds, &de, &ds, &st);
                break;
            } catch (const edsb::HaloError&) {
                // a symbol of the owned range does not close inside the window: widen and go again
                if (we - wb >= idx.n_cols || attempt > 24) throw;
                halo *= 4;
            }
        }
    } catch (const edsb::BadMsa& e) {
        fail(EDS_ERR_BAD_MSA, e.what());
    } catch (const edsb::HaloError& e) {
        fail(EDS_ERR_HALO, e.what());
    } catch (const edsb::CudaError& e) {
        fail(EDS_ERR_CUDA, e.what());
    } catch (const std::invalid_argument& e) {
        fail(EDS_ERR_INVALID_ARGUMENT, e.what());
    } catch (const std::exception& e) {
        fail(EDS_ERR_RUNTIME, e.what());
    }
    const bool bad = run.status[k] != EDS_OK;
    if (run.stats) run.stats[k] = st;

    // ---- the exchange: every device learns every device's byte counts (a failed device reports ~0: all abort)
    uint64_t mine[2] = {bad ? ~0ull : de.bytes, bad ? ~0ull : ds.bytes};
    std::vector<uint64_t> all(2 * (size_t)n, 0);
    try {
        if (g->comm.empty()) {
            run.counts[2 * k] = mine[0];
            run.counts[2 * k + 1] = mine[1];
            run.bar.wait();
            all = run.counts;
        } else {
            cudaStream_t s = ctx->stream;
            uint64_t* h = g->h_counts[k];
            uint64_t* d = g->d_counts[k];
            h[0] = mine[0];
            h[1] = mine[1];
            EDSB_CUDA(cudaMemcpyAsync(d, h, 16, cudaMemcpyHostToDevice, s));
            nccl_check(nccl().AllGather(d, d + 2, 2, ncclUint64, g->comm[k], s), "ncclAllGather");
            EDSB_CUDA(cudaMemcpyAsync(h + 2, d + 2, (size_t)n * 16, cudaMemcpyDeviceToHost, s));
            EDSB_CUDA(cudaStreamSynchronize(s));
            for (uint32_t i = 0; i < 2 * n; ++i) all[i] = h[2 + i];
            run.bar.wait();
        }
    } catch (const std::exception& e) {
        fail(EDS_ERR_CUDA, e.what());
        run.bar.wait();
        for (uint32_t i = 0; i < 2 * n; ++i) all[i] = ~0ull;
    }
    bool any_bad = false;
    uint64_t eoff = 0, soff = 0, etot = 0, stot = 0;
    for (uint32_t i = 0; i < n; ++i) {
        if (all[2 * i] == ~0ull) any_bad = true;
        if (i < k) {
            eoff += all[2 * i];
            soff += all[2 * i + 1];
        }
        etot += all[2 * i];
        stot += all[2 * i + 1];
    }
    if (k == 0 && !any_bad) {
        run.eds_total = etot;
        run.seds_total = stot;
        if (run.eds_fd < 0) {
            run.h_eds = static_cast<uint8_t*>(malloc(etot ? etot : 1));
            run.h_seds = static_cast<uint8_t*>(malloc(stot ? stot : 1));
            if (!run.h_eds || !run.h_seds) fail(EDS_ERR_RUNTIME, "out of host memory");
        } else {
            if (ftruncate(run.eds_fd, (off_t)etot) != 0 || ftruncate(run.seds_fd, (off_t)stot) != 0)
                fail(EDS_ERR_RUNTIME, "ftruncate failed on an output file");
        }
    }
    run.bar.wait();  // the destination exists
    if (any_bad || run.status[0] != EDS_OK) return;
    try {
        cudaStream_t s = ctx->stream;
        if (run.eds_fd < 0) {
            if (de.bytes) EDSB_CUDA(cudaMemcpyAsync(run.h_eds + eoff, de.data, de.bytes, cudaMemcpyDeviceToHost, s));
            if (ds.bytes) EDSB_CUDA(cudaMemcpyAsync(run.h_seds + soff, ds.data, ds.bytes, cudaMemcpyDeviceToHost, s));
            EDSB_CUDA(cudaStreamSynchronize(s));
        } else {
            // through the context's pinned result buffers, then pwrite at this device's offsets
            std::vector<uint8_t> tmp;
            for (int which = 0; which < 2; ++which) {
                const eds_buffer& d = which ? ds : de;
                tmp.resize(d.bytes ? d.bytes : 1);
                if (d.bytes) {
                    EDSB_CUDA(cudaMemcpyAsync(tmp.data(), d.data, d.bytes, cudaMemcpyDeviceToHost, s));
                    EDSB_CUDA(cudaStreamSynchronize(s));
                }
                pwrite_all(which ? run.seds_fd : run.eds_fd, tmp.data(), d.bytes, which ? soff : eoff);
            }
        }
    } catch (const std::exception& e) {
        fail(EDS_ERR_RUNTIME, e.what());
    }
}

eds_status group_transform(eds_group* g, const uint8_t* file, uint64_t file_bytes, uint32_t l, int leds, uint64_t halo, int eds_fd,
                           int seds_fd, eds_buffer* eds_out, eds_buffer* seds_out, uint64_t* totals, eds_msa_stats* stats) {
    if (eds_out) *eds_out = eds_buffer{nullptr, 0};
    if (seds_out) *seds_out = eds_buffer{nullptr, 0};
    return guarded_shard([&] {
        if (!g || !file) throw std::invalid_argument("eds_group_msa_transform: null argument");
        if (eds_fd < 0 && (!eds_out || !seds_out)) throw std::invalid_argument("eds_group_msa_transform: no destination");
        const uint32_t n = (uint32_t)g->ctx.size();
        ShardRun run;
        run.g = g;
        run.file = file;
        run.file_bytes = file_bytes;
        run.l = l;
        run.leds = leds;
        run.halo = halo ? halo : 4096;
        run.eds_fd = eds_fd;
        run.seds_fd = seds_fd;
        run.stats = stats;
        run.bar.n = n;
        run.error.assign(n, "");
        run.status.assign(n, EDS_OK);
        run.counts.assign(2 * (size_t)n, 0);
        memset(&run.idx, 0, sizeof(run.idx));
        if (eds_msa_index_host(file, file_bytes, &run.idx) != EDS_OK) throw edsb::BadMsa(eds_last_error());
        if (run.idx.n_cols < n) {
            eds_msa_index_free(&run.idx);
            throw std::invalid_argument("eds_group_msa_transform: fewer columns than devices");
        }
#ifdef EDSB_EMU
        // the emulator runs one kernel at a time: play the devices on threads all the same (the barrier needs them)
#endif
        std::vector<std::thread> th;
        for (uint32_t k = 0; k < n; ++k) th.emplace_back([&run, k] { shard_worker(run, k); });
        for (auto& t : th) t.join();
        eds_msa_index_free(&run.idx);
        for (uint32_t k = 0; k < n; ++k)
            if (run.status[k] != EDS_OK) {
                free(run.h_eds);
                free(run.h_seds);
                const std::string msg = run.error[k];
                switch (run.status[k]) {
                    case EDS_ERR_BAD_MSA: throw edsb::BadMsa(msg);
                    case EDS_ERR_HALO: throw edsb::HaloError(msg);
                    case EDS_ERR_CUDA: throw edsb::CudaError(msg);
                    case EDS_ERR_INVALID_ARGUMENT: throw std::invalid_argument(msg);
                    default: throw std::runtime_error(msg);
                }
            }
        if (eds_fd < 0) {
            eds_out->data = run.h_eds;
            eds_out->bytes = run.eds_total;
            seds_out->data = run.h_seds;
            seds_out->bytes = run.seds_total;
        }
        if (totals) {
            totals[0] = run.eds_total;
            totals[1] = run.seds_total;
        }
    });
}

}  // namespace

extern "C" {

eds_status eds_group_msa_transform_host(eds_group* g, const uint8_t* file, uint64_t file_bytes, uint32_t l, int leds, uint64_t halo,
                                        eds_buffer* eds_out, eds_buffer* seds_out, eds_msa_stats* per_device_stats) {
    return group_transform(g, file, file_bytes, l, leds, halo, -1, -1, eds_out, seds_out, nullptr, per_device_stats);
}

eds_status eds_group_msa_transform_fd(eds_group* g, const uint8_t* file, uint64_t file_bytes, uint32_t l, int leds, uint64_t halo,
                                      int eds_fd, int seds_fd, uint64_t totals[2], eds_msa_stats* per_device_stats) {
    if (eds_fd < 0 || seds_fd < 0) {
        edsb::set_last_error("eds_group_msa_transform_fd: bad file descriptor");
        return EDS_ERR_INVALID_ARGUMENT;
    }
    return group_transform(g, file, file_bytes, l, leds, halo, eds_fd, seds_fd, nullptr, nullptr, totals, per_device_stats);
}

}  // extern "C"

// =====================================================================================================================
// eds2leds over the devices of a group (symbol ranges). A merge never crosses a long conserved
// symbol that no candidate pair ever touches, so the EDS is cut INSIDE such symbols — `{AAAAAAAA}` becomes `...{AAAA}`
// and `{AAAA}...` — every shard is merged on its own device, and the results are joined by gluing the two halves back
// together (their source set is kept once). Whether a cut was really inert is CHECKED, not assumed: every shard
// reports whether its first / last symbol is still the unmerged half it was given; one touched seam and the call falls
// back to a single-device merge of the whole text. Output bytes are those of the single-device merge either way.
namespace {

struct LedsCut {
    uint64_t eds_at;    // byte offset inside the conserved symbol's characters where the text is cut
    uint64_t n_strings; // strings that start before the cut (the split symbol's own string included)
    bool braced;        // the symbol is written {…} (else bare text of the compact dialect)
};

inline bool is_space_c(uint8_t c) { return c == ' ' || (c >= 9 && c <= 13); }

// cuts near the byte positions k * n / parts; false when the text has no usable symbol near one of them
bool plan_leds_cuts(const uint8_t* t, uint64_t n, uint32_t parts, uint32_t l, std::vector<LedsCut>& cuts) {
    cuts.clear();
    uint64_t strings = 0, counted = 0;  // strings that start in t[0, counted)
    auto count_to = [&](uint64_t upto) {
        // a string starts after '{' or ',' and at the first character of a bare run (after '}' or at the start)
        for (uint64_t i = counted; i < upto; ++i) {
            const uint8_t c = t[i];
            if (c == '{' || c == ',') ++strings;
            else if (c != '}' && (i == 0 || t[i - 1] == '}')) ++strings;
        }
        counted = upto;
    };
    uint64_t prev_cut = 0;
    for (uint32_t k = 1; k < parts; ++k) {
        uint64_t p = std::max<uint64_t>(n / parts * k, prev_cut + 1);
        bool found = false;
        // walk forward over symbols until a conserved one of at least 2 l + 2 characters turns up (bounded search)
        for (uint64_t scanned = 0; p < n && scanned < (1u << 22) && !found;) {
            // start of the next symbol at or after p
            while (p < n && t[p] != '{' && !(p > 0 && t[p - 1] == '}' && t[p] != '{')) {
                ++p;
                ++scanned;
            }
            if (p >= n) break;
            const bool braced = t[p] == '{';
            const uint64_t body = braced ? p + 1 : p;
            uint64_t e = body;
            bool solid = true;
            while (e < n && t[e] != '}' && t[e] != '{') {
                if (t[e] == ',') solid = false;
                ++e;
            }
            if (braced && (e >= n || t[e] != '}')) return false;  // malformed: let the single-device path report it
            const uint64_t len = e - body;
            if (solid && len >= 2ull * l + 2 && body > prev_cut) {
                const uint64_t at = body + len / 2;
                count_to(body + 1);  // the symbol's own string is counted (it starts at `body`, or with the '{' before it)
                cuts.push_back(LedsCut{at, strings, braced});
                prev_cut = at;
                found = true;
            } else {
                scanned += e - p + 1;
                p = braced ? e + 1 : e;
            }
        }
        if (!found) return false;
    }
    return true;
}

// byte offset of the set with index `j` (its '{') in the SEDS text, scanning on from a known (offset, index) pair
bool seds_seek(const uint8_t* s, uint64_t n, uint64_t j, uint64_t& at, uint64_t& idx) {
    for (; at < n; ++at)
        if (s[at] == '{') {
            if (idx == j) return true;
            ++idx;
        }
    return idx == j;
}

}  // namespace

extern "C" eds_status eds_group_leds_merge_host(eds_group* g, const uint8_t* eds, uint64_t eds_bytes, const uint8_t* seds,
                                                uint64_t seds_bytes, uint32_t l, int compact, eds_buffer* leds_out, eds_buffer* seds_out,
                                                uint32_t* rounds_out, uint32_t* shards_used) {
    if (leds_out) *leds_out = eds_buffer{nullptr, 0};
    if (seds_out) *seds_out = eds_buffer{nullptr, 0};
    if (shards_used) *shards_used = 1;
    return guarded_shard([&] {
        if (!g || !eds || !leds_out || !seds_out) throw std::invalid_argument("eds_group_leds_merge_host: null argument");
        if (l == 0) throw std::invalid_argument("context_length must be > 0 for l-EDS transformation");
        const uint32_t n = (uint32_t)g->ctx.size();
        const bool linear = seds != nullptr;
        auto single = [&] {
            eds_ctx* c = g->ctx[0];
            EDSB_CUDA(cudaSetDevice(c->device));
            c->leds->merge_host(eds, eds_bytes, seds, seds_bytes, l, compact != 0, 0, leds_out, seds_out, rounds_out);
        };
        // the planner reads plain text: white space inside the text (legal, stripped by the parser) -> one device
        uint64_t ne = eds_bytes, ns = seds_bytes;
        while (ne && is_space_c(eds[ne - 1])) --ne;
        while (linear && ns && is_space_c(seds[ns - 1])) --ns;
        bool plain = n > 1 && ne >= (uint64_t)n * (4ull * l + 64);
        for (uint64_t i = 0; plain && i < ne; ++i) plain = !is_space_c(eds[i]);
        for (uint64_t i = 0; plain && linear && i < ns; ++i) plain = !is_space_c(seds[i]);
        std::vector<LedsCut> cuts;
        if (!plain || !plan_leds_cuts(eds, ne, n, l, cuts)) {
            single();
            return;
        }
        // shard texts: [cut k-1, cut k) with the split symbol closed on the left and reopened on the right
        std::vector<std::string> se(n), ss(n);
        std::vector<uint64_t> set_at(n + 1, 0);
        if (linear) {
            uint64_t at = 0, idx = 0;
            for (uint32_t k = 1; k < n; ++k) {
                // set index of the split symbol's string = n_strings - 1: the right shard starts with a copy of that set
                if (!seds_seek(seds, ns, cuts[k - 1].n_strings - 1, at, idx)) {
                    single();
                    return;
                }
                set_at[k] = at;
            }
            set_at[n] = ns;
        }
        for (uint32_t k = 0; k < n; ++k) {
            const uint64_t lo = k ? cuts[k - 1].eds_at : 0, hi = k + 1 < n ? cuts[k].eds_at : ne;
            std::string& e = se[k];
            if (k && cuts[k - 1].braced) e.push_back('{');
            e.append(reinterpret_cast<const char*>(eds + lo), hi - lo);
            if (k + 1 < n && cuts[k].braced) e.push_back('}');
            if (linear) {
                // sets of the strings that start in this shard; the split symbol's set opens the right shard as well
                const uint64_t a = set_at[k];
                uint64_t b = set_at[k + 1];
                if (k + 1 < n) {  // include the split symbol's own set: up to and including its '}'
                    while (b < ns && seds[b] != '}') ++b;
                    ++b;
                }
                ss[k].assign(reinterpret_cast<const char*>(seds + a), b - a);
            }
        }
        // one thread per device
        std::vector<eds_buffer> oe(n, eds_buffer{nullptr, 0}), os(n, eds_buffer{nullptr, 0});
        std::vector<uint32_t> rounds(n, 0), edges(n, 0);
        std::vector<std::string> err(n);
        std::vector<int> status(n, 0);
        std::vector<std::thread> th;
        for (uint32_t k = 0; k < n; ++k)
            th.emplace_back([&, k] {
                try {
                    eds_ctx* c = g->ctx[k];
                    EDSB_CUDA(cudaSetDevice(c->device));
                    c->leds->merge_host(reinterpret_cast<const uint8_t*>(se[k].data()), se[k].size(),
                                        linear ? reinterpret_cast<const uint8_t*>(ss[k].data()) : nullptr, ss[k].size(), l, compact != 0, 0,
                                        &oe[k], &os[k], &rounds[k], nullptr, false, {}, nullptr, nullptr, &edges[k]);
                } catch (const std::exception& e) {
                    status[k] = 1;
                    err[k] = e.what();
                }
            });
        for (auto& t : th) t.join();
        bool inert = true;
        for (uint32_t k = 0; k < n; ++k) {
            if (status[k]) inert = false;  // (an error inside a shard carries shard-local positions: redo it whole)
            if (k > 0 && !(edges[k] & 1u)) inert = false;
            if (k + 1 < n && !(edges[k] & 2u)) inert = false;
        }
        auto drop = [&] {
            for (uint32_t k = 0; k < n; ++k) {
                eds_buffer_free_host(&oe[k]);
                eds_buffer_free_host(&os[k]);
            }
        };
        if (!inert) {
            drop();
            single();
            return;
        }
        // join: glue the halves (drop "}{" in the braced dialect), keep the split symbol's set once, one final newline
        std::string je, js;
        for (uint32_t k = 0; k < n; ++k) {
            std::string part(reinterpret_cast<const char*>(oe[k].data), oe[k].bytes);
            while (!part.empty() && part.back() == '\n') part.pop_back();
            if (k > 0 && !je.empty() && je.back() == '}' && !part.empty() && part.front() == '{') {
                je.pop_back();
                part.erase(0, 1);
            }
            je += part;
            if (linear) {
                std::string sp(reinterpret_cast<const char*>(os[k].data), os[k].bytes);
                while (!sp.empty() && sp.back() == '\n') sp.pop_back();
                if (k > 0) sp.erase(0, sp.find('}') + 1);  // the copy of the split symbol's set
                js += sp;
            }
        }
        drop();
        je.push_back('\n');
        if (linear) js.push_back('\n');
        uint8_t* he = static_cast<uint8_t*>(malloc(je.size() ? je.size() : 1));
        uint8_t* hs = static_cast<uint8_t*>(malloc(js.size() ? js.size() : 1));
        if (!he || !hs) {
            free(he);
            free(hs);
            throw std::bad_alloc();
        }
        memcpy(he, je.data(), je.size());
        memcpy(hs, js.data(), js.size());
        leds_out->data = he;
        leds_out->bytes = je.size();
        seds_out->data = hs;
        seds_out->bytes = linear ? js.size() : 0;
        if (rounds_out) *rounds_out = *std::max_element(rounds.begin(), rounds.end());
        if (shards_used) *shards_used = n;
    });
}

// =====================================================================================================================
// vcf2eds over the devices of a group (ranges of record lines). Every device gets a slice of the
// VCF's lines and the whole FASTA record; the slices meet once in the middle of the transform (VcfShardHook):
//   * the order of records that share a position comes from the reference's ONE unstable std::sort over all records
//     (vcf_transforms.cpp:715-718), so that very call is made once over every slice's positions and each slice takes
//     its part of the permutation;
//   * no group of overlapping records may span a cut (largest record end of the slices before <= first POS - 1), and
//     the sorted order must keep every record inside its own slice;
//   * slice k renders the reference bases from its first group to the next slice's first group.
// Anything else — a slice without records, an error in a slice (its positions are slice-local), a spanning group —
// and the whole input runs on the first device, which also produces the reference's error texts. l > 0: the joined
// text goes through eds_group_leds_merge_host (parse_vcf_to_leds_streaming :750-752 = transform, then LINEAR merge).
namespace {

struct VcfMeet {
    std::mutex m;
    std::condition_variable cv;
    uint32_t n = 1, arrived = 0;
    bool done = false, ok = true;
    std::vector<char> here;
    std::vector<const std::vector<uint64_t>*> pos;
    std::vector<uint64_t> max_end, n_bases, lo, hi;
    std::vector<std::vector<uint32_t>> perm;

    explicit VcfMeet(uint32_t n_) : n(n_), here(n_, 0), pos(n_, nullptr), max_end(n_, 0), n_bases(n_, 0), lo(n_, 0), hi(n_, 0), perm(n_) {}

    void solve() {
        uint64_t total = 0;
        for (uint32_t k = 0; k < n; ++k) {
            if (!pos[k] || pos[k]->empty() || n_bases[k] != n_bases[0]) {
                ok = false;
                return;
            }
            total += pos[k]->size();
        }
        if (total >= 0xfffffff0ull) {
            ok = false;
            return;
        }
        std::vector<std::pair<uint64_t, uint32_t>> keyed(total);
        bool unsorted = false;
        uint64_t at = 0, prev = 0;
        for (uint32_t k = 0; k < n; ++k)
            for (uint64_t p : *pos[k]) {
                if (at && p <= prev) unsorted = true;
                prev = p;
                keyed[at] = {p, (uint32_t)at};
                ++at;
            }
        // the same call, comparator and sequence as the single-device transform (vcf.cu) and the reference make
        if (unsorted)
            std::sort(keyed.begin(), keyed.end(),
                      [](const std::pair<uint64_t, uint32_t>& a, const std::pair<uint64_t, uint32_t>& b) { return a.first < b.first; });
        uint64_t base = 0, end_before = 0;
        std::vector<uint64_t> first_pos(n, 0);
        for (uint32_t k = 0; k < n; ++k) {
            const uint64_t cnt = pos[k]->size();
            first_pos[k] = keyed[base].first;
            if (k && (first_pos[k] == 0 || end_before > first_pos[k] - 1)) {
                ok = false;  // a group of overlapping records spans the cut
                return;
            }
            if (unsorted) {
                perm[k].resize(cnt);
                for (uint64_t j = 0; j < cnt; ++j) {
                    const uint64_t idx = keyed[base + j].second;
                    if (idx < base || idx >= base + cnt) {
                        ok = false;  // the sort moves a record into another slice
                        return;
                    }
                    perm[k][j] = (uint32_t)(idx - base);
                }
            }
            end_before = std::max(end_before, max_end[k]);
            base += cnt;
        }
        for (uint32_t k = 0; k < n; ++k) {
            lo[k] = k ? first_pos[k] - 1 : 0;
            hi[k] = k + 1 < n ? first_pos[k + 1] - 1 : n_bases[0];
        }
    }

    // returns false when the slices cannot be joined (every caller sees the same answer)
    bool arrive(uint32_t k, const std::vector<uint64_t>* p, uint64_t me, uint64_t nb, bool failed) {
        std::unique_lock<std::mutex> lk(m);
        if (here[k]) return ok;
        here[k] = 1;
        pos[k] = p;
        max_end[k] = me;
        n_bases[k] = nb;
        if (failed) ok = false;
        if (++arrived == n) {
            if (ok) solve();
            done = true;
            cv.notify_all();
        } else {
            cv.wait(lk, [&] { return done; });
        }
        return ok;
    }
};

struct VcfCannotShard : std::runtime_error {
    using std::runtime_error::runtime_error;
};

struct VcfSliceHook : edsb::VcfShardHook {
    VcfMeet* meet;
    uint32_t k;
    VcfSliceHook(VcfMeet* m, uint32_t k_) : meet(m), k(k_) {}
    void exchange(const std::vector<uint64_t>& pos, uint64_t max_end, uint64_t n_bases, std::vector<uint32_t>* perm, uint64_t* ref_lo,
                  uint64_t* ref_hi) override {
        if (!meet->arrive(k, &pos, max_end, n_bases, false)) throw VcfCannotShard("the VCF slices cannot be joined");
        *perm = std::move(meet->perm[k]);
        *ref_lo = meet->lo[k];
        *ref_hi = meet->hi[k];
    }
};

// POS and the 0-based end of the record line at `at` (CHROM \t POS \t ID \t REF \t ...); false: not a plain record line
bool vcf_line_span(const uint8_t* t, uint64_t n, uint64_t at, uint64_t& pos, uint64_t& end) {
    if (at >= n || t[at] == '#' || t[at] == '\n') return false;
    uint64_t i = at;
    auto skip_field = [&]() {
        while (i < n && t[i] != '\t' && t[i] != '\n') ++i;
        if (i >= n || t[i] != '\t') return false;
        ++i;
        return true;
    };
    if (!skip_field()) return false;
    pos = 0;
    uint64_t digits = 0;
    while (i < n && t[i] >= '0' && t[i] <= '9' && digits < 18) {
        pos = pos * 10 + (t[i] - '0');
        ++i;
        ++digits;
    }
    if (!digits || pos == 0 || i >= n || t[i] != '\t') return false;
    ++i;
    if (!skip_field()) return false;
    const uint64_t ref0 = i;
    while (i < n && t[i] != '\t' && t[i] != '\n') ++i;
    if (i == ref0) return false;
    end = pos - 1 + (i - ref0);
    return true;
}

}  // namespace

namespace {
eds_status group_vcf_transform(eds_group* g, const uint8_t* vcf, uint64_t vcf_bytes, const uint8_t* fasta, uint64_t fasta_bytes,
                               uint32_t l, eds_buffer* eds_out, eds_buffer* seds_out, eds_vcf_stats* stats, uint64_t** sv_lines,
                               uint64_t* n_sv_lines, uint32_t* shards_used, bool view) {
    if (eds_out) *eds_out = eds_buffer{nullptr, 0};
    if (seds_out) *seds_out = eds_buffer{nullptr, 0};
    if (sv_lines) *sv_lines = nullptr;
    if (n_sv_lines) *n_sv_lines = 0;
    if (shards_used) *shards_used = 1;
    if (!g || g->ctx.empty() || !eds_out || !seds_out || (!vcf && vcf_bytes) || (!fasta && fasta_bytes)) {
        edsb::set_last_error("eds_group_vcf_transform_host: null argument");
        return EDS_ERR_INVALID_ARGUMENT;
    }
    const uint32_t n = (uint32_t)g->ctx.size();
    auto single = [&]() {
        return view ? eds_vcf_transform_host_view(g->ctx[0], vcf, vcf_bytes, fasta, fasta_bytes, l, eds_out, seds_out, stats, sv_lines, n_sv_lines)
                    : eds_vcf_transform_host(g->ctx[0], vcf, vcf_bytes, fasta, fasta_bytes, l, eds_out, seds_out, stats, sv_lines, n_sv_lines);
    };
    auto drop = [&](uint8_t* p) {
        if (!view && l == 0) free(p);  // (l > 0: the group's gather buffers; view: its pinned slots)
    };
    if (n < 2 || vcf_bytes < (uint64_t)n * 64) return single();

    // ---- cuts: the first line start at or after k / n of the bytes where the record does not touch the one before it
    std::vector<uint64_t> cut(n + 1, 0);
    cut[n] = vcf_bytes;
    for (uint32_t k = 1; k < n; ++k) {
        uint64_t at = (uint64_t)((unsigned __int128)vcf_bytes * k / n);
        at = std::max(at, cut[k - 1] + 1);
        bool found = false;
        for (int tries = 0; tries < 256 && at < vcf_bytes; ++tries) {
            const void* nl = memchr(vcf + at, '\n', (size_t)(vcf_bytes - at));
            if (!nl) break;
            const uint64_t line = (uint64_t)(static_cast<const uint8_t*>(nl) - vcf) + 1;  // a line starts here
            if (line >= vcf_bytes) break;
            // the line before it
            uint64_t before = line - 1;
            while (before > 0 && vcf[before - 1] != '\n') --before;
            uint64_t p0, e0, p1, e1;
            if (vcf_line_span(vcf, vcf_bytes, before, p0, e0) && vcf_line_span(vcf, vcf_bytes, line, p1, e1) && p1 > p0 && e0 <= p1 - 1) {
                cut[k] = line;
                found = true;
                break;
            }
            at = line;
        }
        if (!found) return single();
    }

    struct Slice {
        eds_buffer de{nullptr, 0}, ds{nullptr, 0};
        eds_vcf_stats st;
        std::vector<uint64_t> sv;
        int status = 0;
    };
    std::vector<Slice> slice(n);
    VcfMeet meet(n);
    Barrier bar;
    bar.n = n;
    uint8_t *h_eds = nullptr, *h_seds = nullptr;
    uint64_t etot = 0, stot = 0;
    bool all_ok = false;
    std::vector<std::thread> th;
    for (uint32_t k = 0; k < n; ++k)
        th.emplace_back([&, k] {
            Slice& me = slice[k];
            memset(&me.st, 0, sizeof(me.st));
            eds_ctx* c = g->ctx[k];
            try {
                EDSB_CUDA(cudaSetDevice(c->device));
                const uint64_t bytes = cut[k + 1] - cut[k];
                c->vcf_in[0].reserve(bytes + 16);
                c->vcf_in[1].reserve(fasta_bytes + 16);
                EDSB_CUDA(cudaMemcpyAsync(c->vcf_in[0].p, vcf + cut[k], bytes, cudaMemcpyHostToDevice, c->stream));
                if (fasta_bytes) EDSB_CUDA(cudaMemcpyAsync(c->vcf_in[1].p, fasta, fasta_bytes, cudaMemcpyHostToDevice, c->stream));
                VcfSliceHook hook(&meet, k);
                c->vcf->transform_device(c->vcf_in[0].as<uint8_t>(), bytes, c->vcf_in[1].as<uint8_t>(), fasta_bytes, &me.de, &me.ds, &me.st,
                                         &me.sv, &hook);
            } catch (const std::exception&) {
                me.status = 1;
            }
            meet.arrive(k, nullptr, 0, 0, true);  // no-op when the slice has been to the meeting point
            bar.wait();
            if (k == 0) {
                all_ok = meet.ok;
                for (uint32_t i = 0; i < n; ++i) all_ok = all_ok && slice[i].status == 0;
                if (all_ok) {
                    for (uint32_t i = 0; i < n; ++i) {
                        etot += slice[i].de.bytes;
                        stot += slice[i].ds.bytes;
                    }
                    if (l > 0) {
                        // the joined text stays in HBM: every device pushes its slice to the first one (peer copies,
                        // NVLink where the devices have it), which merges it there
                        try {
                            g->gather[0].reserve(etot + 16);
                            g->gather[1].reserve(stot + 16);
                            h_eds = g->gather[0].as<uint8_t>();
                            h_seds = g->gather[1].as<uint8_t>();
                        } catch (const std::exception&) {
                            all_ok = false;
                        }
                    } else if (view) {
                        try {
                            h_eds = g->view_slot(0, etot);
                            h_seds = g->view_slot(1, stot);
                        } catch (const std::exception&) {
                            all_ok = false;
                        }
                    } else {
                        h_eds = static_cast<uint8_t*>(malloc(etot ? etot : 1));
                        h_seds = static_cast<uint8_t*>(malloc(stot ? stot : 1));
                    }
                    if (!h_eds || !h_seds) all_ok = false;
                }
            }
            bar.wait();
            if (!all_ok) return;
            uint64_t eo = 0, so = 0;
            for (uint32_t i = 0; i < k; ++i) {
                eo += slice[i].de.bytes;
                so += slice[i].ds.bytes;
            }
            cudaError_t e = cudaSuccess;
            if (l > 0) {
                const int d0 = g->ctx[0]->device;
                if (me.de.bytes) e = cudaMemcpyPeerAsync(h_eds + eo, d0, me.de.data, c->device, me.de.bytes, c->stream);
                if (e == cudaSuccess && me.ds.bytes) e = cudaMemcpyPeerAsync(h_seds + so, d0, me.ds.data, c->device, me.ds.bytes, c->stream);
            } else {
                if (me.de.bytes) e = cudaMemcpyAsync(h_eds + eo, me.de.data, me.de.bytes, cudaMemcpyDeviceToHost, c->stream);
                if (e == cudaSuccess && me.ds.bytes) e = cudaMemcpyAsync(h_seds + so, me.ds.data, me.ds.bytes, cudaMemcpyDeviceToHost, c->stream);
            }
            if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
            if (e != cudaSuccess) me.status = 2;
        });
    for (auto& t : th) t.join();
    for (uint32_t i = 0; i < n; ++i) all_ok = all_ok && slice[i].status == 0;
    if (!all_ok) {
        drop(h_eds);
        drop(h_seds);
        for (uint32_t i = 0; i < n; ++i) cudaGetLastError();
        return single();
    }
    eds_vcf_stats tot;
    memset(&tot, 0, sizeof(tot));
    std::vector<uint64_t> sv_all;
    for (uint32_t i = 0; i < n; ++i) {
        const eds_vcf_stats& s = slice[i].st;
        tot.total_variants += s.total_variants;
        tot.processed_variants += s.processed_variants;
        tot.skipped_malformed += s.skipped_malformed;
        tot.skipped_unsupported_sv += s.skipped_unsupported_sv;
        tot.variant_groups += s.variant_groups;
        tot.n_lines += s.n_lines;
        tot.n_alleles += s.n_alleles;
        tot.n_haplotype_slots += s.n_haplotype_slots;
        tot.n_samples_max = std::max(tot.n_samples_max, s.n_samples_max);
        tot.gpu_launches += s.gpu_launches;
        tot.retries += s.retries;
        tot.host_sorted = std::max(tot.host_sorted, s.host_sorted);
        tot.n_bases = s.n_bases;
        for (uint64_t at : slice[i].sv) sv_all.push_back(at + cut[i]);
    }
    tot.eds_bytes = etot;
    tot.seds_bytes = stot;
    if (sv_lines && n_sv_lines && !sv_all.empty()) {
        *sv_lines = static_cast<uint64_t*>(malloc(sv_all.size() * sizeof(uint64_t)));
        if (!*sv_lines) {
            drop(h_eds);
            drop(h_seds);
            edsb::set_last_error("out of host memory");
            return EDS_ERR_RUNTIME;
        }
        memcpy(*sv_lines, sv_all.data(), sv_all.size() * sizeof(uint64_t));
        *n_sv_lines = sv_all.size();
    }
    if (shards_used) *shards_used = n;
    if (l == 0) {
        eds_out->data = h_eds;
        eds_out->bytes = etot;
        seds_out->data = h_seds;
        seds_out->bytes = stot;
        if (stats) *stats = tot;
        return EDS_OK;
    }
    // parse_vcf_to_leds_streaming :750-752: LINEAR merge (compact) of the text just joined, still in HBM, on the first device
    if (stats) *stats = tot;  // the reference fills the counters before the merge can throw
    uint32_t rounds = 0;
    const eds_status rc = guarded_shard([&] {
        eds_ctx* c0 = g->ctx[0];
        EDSB_CUDA(cudaSetDevice(c0->device));
        std::function<uint8_t*(int, uint64_t)> sink;
        if (view) sink = [g](int which, uint64_t bytes) -> uint8_t* { return g->view_slot(which, bytes); };
        c0->leds->merge_host(h_eds, etot, h_seds, stot, l, true, 0, eds_out, seds_out, &rounds, nullptr, true, sink);
    });
    if (rc != EDS_OK) {
        if (view) {
            *eds_out = eds_buffer{nullptr, 0};
            *seds_out = eds_buffer{nullptr, 0};
        } else {
            eds_buffer_free_host(eds_out);
            eds_buffer_free_host(seds_out);
        }
        if (sv_lines && *sv_lines) {
            free(*sv_lines);
            *sv_lines = nullptr;
            *n_sv_lines = 0;
        }
        return rc;
    }
    tot.leds_rounds = rounds;
    if (stats) *stats = tot;
    return EDS_OK;
}
}  // namespace

extern "C" eds_status eds_group_vcf_transform_host(eds_group* g, const uint8_t* vcf, uint64_t vcf_bytes, const uint8_t* fasta,
                                                   uint64_t fasta_bytes, uint32_t l, eds_buffer* eds_out, eds_buffer* seds_out,
                                                   eds_vcf_stats* stats, uint64_t** sv_lines, uint64_t* n_sv_lines,
                                                   uint32_t* shards_used) {
    return group_vcf_transform(g, vcf, vcf_bytes, fasta, fasta_bytes, l, eds_out, seds_out, stats, sv_lines, n_sv_lines, shards_used, false);
}

extern "C" eds_status eds_group_vcf_transform_host_view(eds_group* g, const uint8_t* vcf, uint64_t vcf_bytes, const uint8_t* fasta,
                                                        uint64_t fasta_bytes, uint32_t l, eds_buffer* eds_out, eds_buffer* seds_out,
                                                        eds_vcf_stats* stats, uint64_t** sv_lines, uint64_t* n_sv_lines,
                                                        uint32_t* shards_used) {
    return group_vcf_transform(g, vcf, vcf_bytes, fasta, fasta_bytes, l, eds_out, seds_out, stats, sv_lines, n_sv_lines, shards_used, true);
}
