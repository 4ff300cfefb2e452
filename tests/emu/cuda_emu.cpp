// TEST INFRASTRUCTURE ONLY — see cuda_emu.h. Thread pool that plays the threads of one block.
#include "cuda_emu.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>

namespace emu {

thread_local Tls tls;

State& state() {
    static State s;
    return s;
}

namespace {

struct Pool {
    std::mutex m;
    std::condition_variable cv_go, cv_done;
    std::vector<std::thread> workers;
    const std::function<void()>* body = nullptr;
    unsigned long long epoch = 0;
    unsigned active = 0;   // workers taking part in the current block
    unsigned pending = 0;  // workers that have not finished the current block
    dim3 bid;

    void ensure(unsigned n) {
        while (workers.size() < n) {
            unsigned id = (unsigned)workers.size();
            workers.emplace_back([this, id] { run(id); });
        }
    }
    void run(unsigned id) {
        unsigned long long seen = 0;
        for (;;) {
            const std::function<void()>* fn;
            {
                std::unique_lock<std::mutex> lk(m);
                cv_go.wait(lk, [&] { return epoch != seen; });
                seen = epoch;
                if (id >= active) continue;
                fn = body;
                tls.bid = bid;
            }
            const dim3& b = state().block;
            tls.lin = id;
            tls.tid = dim3(id % b.x, (id / b.x) % b.y, id / (b.x * b.y));
            (*fn)();
            {
                std::lock_guard<std::mutex> lk(m);
                if (--pending == 0) cv_done.notify_all();
            }
        }
    }
    void run_block(const std::function<void()>& fn, dim3 b, unsigned n) {
        std::unique_lock<std::mutex> lk(m);
        body = &fn;
        bid = b;
        active = n;
        pending = n;
        ++epoch;
        cv_go.notify_all();
        cv_done.wait(lk, [&] { return pending == 0; });
    }
};

Pool& pool() {
    static Pool* p = new Pool();  // leaked on purpose: workers are detached-for-life
    return *p;
}

std::mutex g_launch_mutex;

}  // namespace

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    std::lock_guard<std::mutex> guard(g_launch_mutex);
    State& s = state();
    unsigned n = block.x * block.y * block.z;
    s.grid = grid;
    s.block = block;
    s.block_bar.reset(n);
    unsigned nwarps = (n + 31) / 32;
    while (s.warps.size() < nwarps) s.warps.push_back(new WarpBox());
    for (unsigned w = 0; w < nwarps; ++w) {
        unsigned lanes = std::min(32u, n - w * 32);
        s.warps[w]->lanes = lanes;
        s.warps[w]->bar.reset(lanes);
    }
    std::vector<unsigned char> dyn(smem + 16);
    s.dyn_smem = dyn.data();
    Pool& p = pool();
    p.ensure(n);
    static const bool trace = getenv("EMU_TRACE") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    for (unsigned z = 0; z < grid.z; ++z)
        for (unsigned y = 0; y < grid.y; ++y)
            for (unsigned x = 0; x < grid.x; ++x) p.run_block(body, dim3(x, y, z), n);
    s.dyn_smem = nullptr;
    if (trace)
        fprintf(stderr, "emu launch grid %u block %u smem %zu: %.1f ms\n", grid.x * grid.y * grid.z, n, smem,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
}

}  // namespace emu
