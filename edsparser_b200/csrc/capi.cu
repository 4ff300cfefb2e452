// C ABI of libedsparser_b200.so (include/edsparser_b200.h): context management, the host-side MSA
// loader, and exception -> status mapping around the pipelines. No CPU compute path lives here.
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <functional>
#include <new>
#include <stdio.h>
#include <string>
#include <vector>

#include "ctx.h"
#include "leds.h"
#include "msa.h"
#include "vcf.h"

namespace {

thread_local std::string g_last_error;

template <typename F>
eds_status guarded(F&& body) {
    try {
        body();
        return EDS_OK;
    } catch (const edsb::BadMsa& e) {
        g_last_error = e.what();
        return EDS_ERR_BAD_MSA;
    } catch (const edsb::BadVcf& e) {
        g_last_error = e.what();
        return EDS_ERR_BAD_VCF;
    } catch (const edsb::HaloError& e) {
        g_last_error = e.what();
        return EDS_ERR_HALO;
    } catch (const edsb::BudgetError& e) {
        g_last_error = e.what();
        return EDS_ERR_BUDGET;
    } catch (const edsb::CudaError& e) {
        g_last_error = e.what();
        return EDS_ERR_CUDA;
    } catch (const std::invalid_argument& e) {
        g_last_error = e.what();
        return EDS_ERR_INVALID_ARGUMENT;
    } catch (const std::out_of_range& e) {
        g_last_error = e.what();
        return EDS_ERR_OUT_OF_RANGE;
    } catch (const std::bad_alloc&) {
        g_last_error = "out of host memory";
        return EDS_ERR_RUNTIME;
    } catch (const std::exception& e) {
        g_last_error = e.what();
        return EDS_ERR_RUNTIME;
    }
}

void use_device(eds_ctx* ctx) {
    if (!ctx) throw std::invalid_argument("null eds_ctx");
    EDSB_CUDA(cudaSetDevice(ctx->device));
    (void)cudaGetLastError();  // a non-sticky error some earlier call of this thread left behind is not this call's
}

uint8_t* to_host(eds_ctx* ctx, const eds_buffer& dev) {
    uint8_t* h = static_cast<uint8_t*>(malloc(dev.bytes ? dev.bytes : 1));
    if (!h) throw std::bad_alloc();
    if (dev.bytes) {
        cudaError_t e = cudaMemcpyAsync(h, dev.data, dev.bytes, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            free(h);
            throw edsb::CudaError(std::string("device to host copy: ") + cudaGetErrorString(e));
        }
    }
    return h;
}

// where a result goes under the *_view contract: pinned memory kept by the context, grown when needed
uint8_t* view_slot(eds_ctx* ctx, int which, uint64_t bytes) {
    if (ctx->host_out_cap[which] < bytes + 1) {
        if (ctx->host_out[which]) cudaFreeHost(ctx->host_out[which]);
        ctx->host_out[which] = nullptr;
        ctx->host_out_cap[which] = 0;
        const size_t want = ((bytes + bytes / 8 + 4096) / 4096) * 4096;
        EDSB_CUDA(cudaMallocHost(&ctx->host_out[which], want));
        ctx->host_out_cap[which] = want;
    }
    return static_cast<uint8_t*>(ctx->host_out[which]);
}

// pinned, ctx-owned destination of a device buffer (grow-only)
uint8_t* to_host_view(eds_ctx* ctx, int which, const eds_buffer& dev) {
    view_slot(ctx, which, dev.bytes);
    if (dev.bytes) {
        EDSB_CUDA(cudaMemcpyAsync(ctx->host_out[which], dev.data, dev.bytes, cudaMemcpyDeviceToHost, ctx->stream));
        EDSB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return static_cast<uint8_t*>(ctx->host_out[which]);
}

}  // namespace

namespace edsb {
void set_last_error(const std::string& msg) { g_last_error = msg; }
}  // namespace edsb

extern "C" {

const char* eds_last_error(void) { return g_last_error.c_str(); }

const char* eds_version(void) {
#ifdef EDSB_EMU
    return "edsparser_b200 0.1.0 emulated (test build, not a product path)";
#else
    return "edsparser_b200 0.1.0 sm_100a";
#endif
}

eds_status eds_ctx_create(int device, void* stream, eds_ctx** out) {
    return guarded([&] {
        if (!out) throw std::invalid_argument("eds_ctx_create: null out");
        *out = nullptr;
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n == 0)
            throw edsb::CudaError(std::string("no usable CUDA device (this library has no CPU fallback): ") +
                                  cudaGetErrorString(e));
        if (device < 0 || device >= n) throw std::invalid_argument("eds_ctx_create: no such device");
        EDSB_CUDA(cudaSetDevice(device));
        eds_ctx* ctx = new eds_ctx();
        struct Guard {  // a throw below must not leak the streams / events / pipelines created so far
            eds_ctx* c;
            ~Guard() { if (c) eds_ctx_destroy(c); }
        } guard{ctx};
        ctx->device = device;
        cudaDeviceProp prop;
        EDSB_CUDA(cudaGetDeviceProperties(&prop, device));
        ctx->sm_count = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : 148;
        ctx->smem_optin = prop.sharedMemPerBlockOptin ? prop.sharedMemPerBlockOptin : 48 * 1024;
        if (stream) {
            ctx->stream = static_cast<cudaStream_t>(stream);
        } else {
            EDSB_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
            ctx->own_stream = true;
        }
        ctx->clock.stream = ctx->stream;
        for (auto& a : ctx->aux) EDSB_CUDA(cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking));
        for (auto& e : ctx->ev) EDSB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        if (const char* se = getenv("EDSB_DEBUG_SERIAL")) ctx->serial = atoi(se) != 0;
#ifndef EDSB_EMU
        if (const char* fg = getenv("EDSB_L2_FETCH")) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(fg));
#endif
        if (const char* fz = getenv("EDSB_FUSED")) ctx->fused = atoi(fz) != 0;
        if (const char* fz = getenv("EDSB_FUSED_MIN_ROWS")) ctx->fused_min_rows = (uint32_t)atoi(fz);
        if (const char* fz = getenv("EDSB_FUSED_NC")) ctx->fused_nc = (uint32_t)atoi(fz);
        if (const char* fz = getenv("EDSB_FUSED_STAGES")) ctx->fused_stages = (uint32_t)atoi(fz);
        if (const char* fz = getenv("EDSB_FUSED_PW")) ctx->fused_pw = (uint32_t)atoi(fz);
        if (const char* fz = getenv("EDSB_FUSED_PWB")) ctx->fused_pwb = (uint32_t)atoi(fz);
        if (const char* fz = getenv("EDSB_FUSED_BULK_PCT")) ctx->fused_bulk_pct = (uint32_t)atoi(fz);
        if (const char* fz = getenv("EDSB_FUSED_T")) ctx->fused_t = (uint32_t)atoi(fz);
        if (const char* fz = getenv("EDSB_FUSED_DW")) ctx->fused_dw = (uint32_t)atoi(fz);
        if (const char* fz = getenv("EDSB_FUSED_MODE")) ctx->fused_mode = (uint32_t)atoi(fz);
        if (const char* fz = getenv("EDSB_FUSED_PROBE")) ctx->fused_probe = (uint32_t)atoi(fz);
        if (const char* fz = getenv("EDSB_FUSED_L2")) ctx->fused_l2 = (uint32_t)atoi(fz);
        if (const char* fz = getenv("EDSB_FUSED_DIRECT")) ctx->fused_direct = (uint32_t)atoi(fz);
        if (const char* fz = getenv("EDSB_FUSED_SPLIT")) ctx->fused_split = (uint32_t)atoi(fz);
        if (const char* fz = getenv("EDSB_FUSED_CW")) ctx->fused_cw = (uint32_t)atoi(fz);
        if (const char* fz = getenv("EDSB_FUSED_PAIR")) ctx->fused_pair = atoi(fz) != 0 ? 1u : 0u;
        if (const char* hm = getenv("EDSB_DEBUG_HASH_MASK")) ctx->hash_mask = strtoull(hm, nullptr, 0);
        if (const char* no = getenv("EDSB_DEBUG_NARROW_OFF")) ctx->narrow_off = atoi(no);
        if (const char* no = getenv("EDSB_DEBUG_TUPLE_OFF")) ctx->tuple_off = atoi(no);
        if (const char* no = getenv("EDSB_DEBUG_GROUP_CTA")) ctx->group_cta = (uint32_t)atoi(no);
        if (const char* rs = getenv("EDSB_DEBUG_ROW_SLICES")) ctx->scan_row_slices = (uint32_t)atoi(rs);
        ctx->msa = new edsb::MsaPipeline(ctx);
        ctx->leds = new edsb::LedsPipeline(ctx);
        ctx->vcf = new edsb::VcfPipeline(ctx);
        guard.c = nullptr;
        *out = ctx;
    });
}

void eds_ctx_destroy(eds_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    delete ctx->msa;
    delete ctx->leds;
    delete ctx->vcf;
    for (auto& b : ctx->vcf_in) b.release();
    for (auto& b : ctx->vcf_out) b.release();
    for (void* h : ctx->host_out)
        if (h) cudaFreeHost(h);
    ctx->synth_text.release();
    ctx->file_buf.release();
    ctx->clock.release();
    for (auto& a : ctx->aux)
        if (a) cudaStreamDestroy(a);
    for (auto& e : ctx->ev)
        if (e) cudaEventDestroy(e);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

eds_status eds_ctx_synchronize(eds_ctx* ctx) {
    return guarded([&] {
        use_device(ctx);
        EDSB_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}

eds_status eds_ctx_set_tuning(eds_ctx* ctx, uint32_t partitions, uint32_t scan_blocks_per_sm) {
    return guarded([&] {
        if (!ctx) throw std::invalid_argument("null eds_ctx");
        ctx->partitions = partitions;
        ctx->scan_blocks_per_sm = scan_blocks_per_sm;
    });
}

eds_status eds_ctx_set_profiling(eds_ctx* ctx, int on) {
    return guarded([&] {
        if (!ctx) throw std::invalid_argument("null eds_ctx");
        ctx->clock.on = on != 0;
    });
}

uint32_t eds_ctx_kernel_times(eds_ctx* ctx, const char** names, float* ms, uint32_t cap) {
    if (!ctx) return 0;
    const uint32_t n = (uint32_t)ctx->clock.names.size();
    for (uint32_t i = 0; i < n && i < cap; ++i) {
        if (names) names[i] = ctx->clock.names[i];
        if (ms) ms[i] = i < ctx->clock.ms.size() ? ctx->clock.ms[i] : 0.f;
    }
    return ctx->clock.on ? n : ctx->clock.launches;
}

// MSAMetadata of msa_transforms.cpp:18-24 without touching the residues of rows 1..R-1:
// row 0 fixes C and the wrap width, every other row is located from its header.
eds_status eds_msa_index_host(const uint8_t* text, uint64_t n, eds_msa_index* out) {
    return guarded([&] {
        if (!text || !out) throw std::invalid_argument("eds_msa_index_host: null argument");
        memset(out, 0, sizeof(*out));
        std::vector<uint64_t> rows;
        uint64_t p = 0;
        while (p < n && text[p] == '\n') ++p;
        if (p >= n || text[p] != '>') throw edsb::BadMsa("not a FASTA alignment: expected '>'");
        const uint8_t* nl = static_cast<const uint8_t*>(memchr(text + p, '\n', n - p));
        if (!nl) throw edsb::BadMsa("header without sequence data");
        p = (uint64_t)(nl - text) + 1;
        const uint64_t start0 = p;
        uint64_t C = 0, lw = 0, row_end = p;
        bool short_seen = false;
        while (p < n && text[p] != '>' && text[p] != '\n') {
            const uint8_t* e = static_cast<const uint8_t*>(memchr(text + p, '\n', n - p));
            const uint64_t len = e ? (uint64_t)(e - (text + p)) : n - p;
            if (lw == 0) lw = len;
            if (short_seen || len > lw) throw edsb::BadMsa("first row is not wrapped at a uniform width");
            if (len < lw) short_seen = true;
            C += len;
            row_end = p + len;
            p = e ? (uint64_t)(e - text) + 1 : n;
        }
        if (C == 0) throw edsb::BadMsa("first row is empty");
        if (lw > 0xffffffffull) throw edsb::BadMsa("line too long");
        const uint64_t row_bytes = row_end - start0;
        rows.push_back(start0);
        p = row_end;
        for (;;) {
            while (p < n && text[p] == '\n') ++p;
            if (p >= n) break;
            if (text[p] != '>') throw edsb::BadMsa("row longer than the first row, or blank line inside a record");
            const uint8_t* he = static_cast<const uint8_t*>(memchr(text + p, '\n', n - p));
            if (!he) throw edsb::BadMsa("header without sequence data");
            const uint64_t start = (uint64_t)(he - text) + 1;
            if (start + row_bytes > n) throw edsb::BadMsa("row shorter than the first row");
            if (start + row_bytes < n && text[start + row_bytes] != '\n')
                throw edsb::BadMsa("row longer than the first row (or differently wrapped)");
            if (rows.size() >= 0xffffffffull) throw edsb::BadMsa("too many rows");
            rows.push_back(start);
            p = start + row_bytes;
        }
        if (rows.size() < 2)
            throw edsb::BadMsa("alignment needs at least 2 rows (undefined in the reference, msa_transforms.cpp:53-57)");
        out->row_start = static_cast<uint64_t*>(malloc(rows.size() * sizeof(uint64_t)));
        if (!out->row_start) throw std::bad_alloc();
        memcpy(out->row_start, rows.data(), rows.size() * sizeof(uint64_t));
        out->n_rows = (uint32_t)rows.size();
        out->n_cols = C;
        out->line_width = (uint32_t)lw;
        out->row_bytes = row_bytes;
    });
}

void eds_msa_index_free(eds_msa_index* idx) {
    if (!idx) return;
    free(idx->row_start);
    memset(idx, 0, sizeof(*idx));
}

eds_status eds_msa_transform_device(eds_ctx* ctx, const eds_msa_view* view, uint32_t l, int leds, eds_buffer* eds_out,
                                    eds_buffer* seds_out, eds_msa_stats* stats) {
    return guarded([&] {
        use_device(ctx);
        if (!view) throw std::invalid_argument("eds_msa_transform_device: null view");
        ctx->msa->transform(*view, l, leds, eds_out, seds_out, stats);
    });
}

static eds_status msa_transform_host_impl(eds_ctx* ctx, const uint8_t* file, uint64_t file_bytes, uint32_t l, int leds,
                                          eds_buffer* eds_out, eds_buffer* seds_out, eds_msa_stats* stats, bool view) {
    eds_msa_index idx;
    memset(&idx, 0, sizeof(idx));
    // the caller's structs may be uninitialised: clear them before anything can fail (the error path frees them)
    if (eds_out) *eds_out = eds_buffer{nullptr, 0};
    if (seds_out) *seds_out = eds_buffer{nullptr, 0};
    eds_status rc = guarded([&] {
        use_device(ctx);
        if (!file || !eds_out || !seds_out) throw std::invalid_argument("eds_msa_transform_host: null argument");
        const bool trace = getenv("EDSB_TRACE") != nullptr;
        auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
        // the copy is started first; rows are located on the host while the DMA engine runs
        const double t0 = now();
        ctx->file_buf.reserve(file_bytes + 64);
        EDSB_CUDA(cudaMemcpyAsync(ctx->file_buf.p, file, file_bytes, cudaMemcpyHostToDevice, ctx->stream));
        const eds_status irc = eds_msa_index_host(file, file_bytes, &idx);
        if (irc != EDS_OK) {
            cudaStreamSynchronize(ctx->stream);
            throw edsb::BadMsa(g_last_error);
        }
        const double t1 = now();
        if (trace) EDSB_CUDA(cudaStreamSynchronize(ctx->stream));
        const double t2 = now();
        eds_msa_view v;
        memset(&v, 0, sizeof(v));
        v.text = ctx->file_buf.as<uint8_t>();
        v.text_bytes = file_bytes;
        v.row_start = idx.row_start;
        v.n_rows = idx.n_rows;
        v.line_width = idx.line_width;
        v.total_cols = idx.n_cols;
        v.col_begin = 0;
        v.col_count = idx.n_cols;
        v.own_begin = 0;
        v.own_end = idx.n_cols;
        eds_buffer de, ds;
        ctx->msa->transform(v, l, leds, &de, &ds, stats);
        const double t3 = now();
        eds_out->data = view ? to_host_view(ctx, 0, de) : to_host(ctx, de);
        eds_out->bytes = de.bytes;
        seds_out->data = view ? to_host_view(ctx, 1, ds) : to_host(ctx, ds);
        seds_out->bytes = ds.bytes;
        if (trace)
            fprintf(stderr, "[edsb trace] index (under the copy) %.3f ms, H2D %.3f ms (%.1f GB/s), transform %.3f ms, D2H %.3f ms\n", t1 - t0,
                    t2 - t0, file_bytes / (t2 - t0) / 1e6, t3 - t2, now() - t3);
    });
    eds_msa_index_free(&idx);
    if (rc != EDS_OK) {
        if (view) {
            if (eds_out) *eds_out = eds_buffer{nullptr, 0};
            if (seds_out) *seds_out = eds_buffer{nullptr, 0};
        } else {
            if (eds_out) eds_buffer_free_host(eds_out);
            if (seds_out) eds_buffer_free_host(seds_out);
        }
    }
    return rc;
}

eds_status eds_msa_transform_host(eds_ctx* ctx, const uint8_t* file, uint64_t file_bytes, uint32_t l, int leds,
                                  eds_buffer* eds_out, eds_buffer* seds_out, eds_msa_stats* stats) {
    return msa_transform_host_impl(ctx, file, file_bytes, l, leds, eds_out, seds_out, stats, false);
}

eds_status eds_msa_transform_host_view(eds_ctx* ctx, const uint8_t* file, uint64_t file_bytes, uint32_t l, int leds,
                                       eds_buffer* eds_out, eds_buffer* seds_out, eds_msa_stats* stats) {
    return msa_transform_host_impl(ctx, file, file_bytes, l, leds, eds_out, seds_out, stats, true);
}

eds_status eds_msa_conserved_bits(eds_ctx* ctx, const eds_msa_view* view, uint8_t* out_bits, uint64_t out_bytes) {
    return guarded([&] {
        use_device(ctx);
        if (!view || !out_bits) throw std::invalid_argument("eds_msa_conserved_bits: null argument");
        ctx->msa->conserved_bits(*view, out_bits, out_bytes);
    });
}

eds_status eds_msa_synth_device(eds_ctx* ctx, uint32_t n_rows, uint64_t total_cols, uint32_t line_width,
                                uint64_t col_begin, uint64_t col_count, uint64_t seed, uint32_t variable_ppm,
                                eds_msa_view* view) {
    return guarded([&] {
        use_device(ctx);
        if (!view) throw std::invalid_argument("eds_msa_synth_device: null view");
        edsb::msa_synth(ctx, n_rows, total_cols, line_width, col_begin, col_count, seed, variable_ppm, view);
    });
}

void eds_msa_synth_free(eds_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    ctx->synth_text.release();
    ctx->synth_rows.clear();
}

eds_status eds_buffer_to_host(eds_ctx* ctx, const eds_buffer* device_buf, eds_buffer* host_out) {
    return guarded([&] {
        use_device(ctx);
        if (!device_buf || !host_out) throw std::invalid_argument("eds_buffer_to_host: null argument");
        host_out->data = to_host(ctx, *device_buf);
        host_out->bytes = device_buf->bytes;
    });
}

eds_status eds_buffer_to_host_view(eds_ctx* ctx, int slot, const eds_buffer* device_buf, eds_buffer* host_out) {
    return guarded([&] {
        use_device(ctx);
        if (!device_buf || !host_out || slot < 0 || slot > 1) throw std::invalid_argument("eds_buffer_to_host_view: bad argument");
        host_out->data = to_host_view(ctx, slot, *device_buf);
        host_out->bytes = device_buf->bytes;
    });
}

void eds_buffer_free_host(eds_buffer* buf) {
    if (!buf) return;
    free(buf->data);
    buf->data = nullptr;
    buf->bytes = 0;
}

eds_status eds_leds_merge_host(eds_ctx* ctx, const uint8_t* eds_in, uint64_t eds_bytes, const uint8_t* seds_in,
                               uint64_t seds_bytes, uint32_t l, int compact, uint64_t max_output_bytes,
                               eds_buffer* leds_out, eds_buffer* seds_out, uint32_t* rounds_out) {
    if (leds_out) *leds_out = eds_buffer{nullptr, 0};
    if (seds_out) *seds_out = eds_buffer{nullptr, 0};
    eds_status rc = guarded([&] {
        use_device(ctx);
        if (!eds_in || !leds_out || !seds_out) throw std::invalid_argument("eds_leds_merge_host: null argument");
        ctx->leds->merge_host(eds_in, eds_bytes, seds_in, seds_bytes, l, compact != 0, max_output_bytes, leds_out,
                              seds_out, rounds_out);
    });
    if (rc != EDS_OK) {
        if (leds_out) eds_buffer_free_host(leds_out);
        if (seds_out) eds_buffer_free_host(seds_out);
    }
    return rc;
}

eds_status eds_leds_merge_host_view(eds_ctx* ctx, const uint8_t* eds_in, uint64_t eds_bytes, const uint8_t* seds_in,
                                    uint64_t seds_bytes, uint32_t l, int compact, uint64_t max_output_bytes,
                                    eds_buffer* leds_out, eds_buffer* seds_out, uint32_t* rounds_out) {
    if (leds_out) *leds_out = eds_buffer{nullptr, 0};
    if (seds_out) *seds_out = eds_buffer{nullptr, 0};
    eds_status rc = guarded([&] {
        use_device(ctx);
        if (!eds_in || !leds_out || !seds_out) throw std::invalid_argument("eds_leds_merge_host_view: null argument");
        ctx->leds->merge_host(eds_in, eds_bytes, seds_in, seds_bytes, l, compact != 0, max_output_bytes, leds_out, seds_out,
                              rounds_out, nullptr, false,
                              [ctx](int which, uint64_t bytes) -> uint8_t* { return view_slot(ctx, which, bytes); });
    });
    if (rc != EDS_OK) {
        if (leds_out) *leds_out = eds_buffer{nullptr, 0};
        if (seds_out) *seds_out = eds_buffer{nullptr, 0};
    }
    return rc;
}

eds_status eds_leds_merge_device_in(eds_ctx* ctx, const uint8_t* eds_dev, uint64_t eds_bytes, const uint8_t* seds_dev,
                                    uint64_t seds_bytes, uint32_t l, int compact, uint64_t max_output_bytes,
                                    eds_buffer* leds_out, eds_buffer* seds_out, uint32_t* rounds_out) {
    if (leds_out) *leds_out = eds_buffer{nullptr, 0};
    if (seds_out) *seds_out = eds_buffer{nullptr, 0};
    eds_status rc = guarded([&] {
        use_device(ctx);
        if (!eds_dev || !leds_out || !seds_out) throw std::invalid_argument("eds_leds_merge_device_in: null argument");
        if ((reinterpret_cast<uintptr_t>(eds_dev) | reinterpret_cast<uintptr_t>(seds_dev)) & 15u)
            throw std::invalid_argument("eds_leds_merge_device_in: inputs must be 16-byte aligned");
        ctx->leds->merge_host(eds_dev, eds_bytes, seds_dev, seds_bytes, l, compact != 0, max_output_bytes, leds_out, seds_out,
                              rounds_out, nullptr, true);
    });
    if (rc != EDS_OK) {
        if (leds_out) eds_buffer_free_host(leds_out);
        if (seds_out) eds_buffer_free_host(seds_out);
    }
    return rc;
}

eds_status eds_genrandomeds_device(eds_ctx* ctx, uint64_t ref_size, uint32_t variability_ppm, uint32_t paths, uint64_t seed,
                                   eds_buffer* eds_out, eds_buffer* seds_out) {
    if (eds_out) *eds_out = eds_buffer{nullptr, 0};
    if (seds_out) *seds_out = eds_buffer{nullptr, 0};
    return guarded([&] {
        use_device(ctx);
        if (!eds_out || !seds_out) throw std::invalid_argument("eds_genrandomeds_device: null argument");
        ctx->leds->genrandomeds(ref_size, variability_ppm, paths, seed, eds_out, seds_out);
    });
}

eds_status eds_parse_host(eds_ctx* ctx, const uint8_t* eds, uint64_t eds_bytes, const uint8_t* seds, uint64_t seds_bytes,
                          eds_parsed* out) {
    if (out) memset(out, 0, sizeof(*out));
    return guarded([&] {
        use_device(ctx);
        if ((!eds && eds_bytes) || !out) throw std::invalid_argument("eds_parse_host: null argument");
        static const uint8_t nothing = 0;
        ctx->leds->merge_host(eds ? eds : &nothing, eds_bytes, seds, seds_bytes, 0, false, 0, nullptr, nullptr, nullptr, nullptr, false,
                              {}, out);
    });
}

void eds_parsed_free(eds_parsed* p) {
    if (!p) return;
    free(p->text);
    free(p->str_start);
    free(p->str_end);
    free(p->sym_first);
    free(p->src_off);
    free(p->src_ids);
    memset(p, 0, sizeof(*p));
}

eds_status eds_merge_adjacent_host(eds_ctx* ctx, const uint8_t* eds, uint64_t eds_bytes, const uint8_t* seds, uint64_t seds_bytes,
                                   uint64_t pos1, eds_buffer* eds_out, eds_buffer* seds_out) {
    if (eds_out) *eds_out = eds_buffer{nullptr, 0};
    if (seds_out) *seds_out = eds_buffer{nullptr, 0};
    eds_status rc = guarded([&] {
        use_device(ctx);
        if (!eds || !eds_out || !seds_out) throw std::invalid_argument("eds_merge_adjacent_host: null argument");
        if (pos1 >= 0xfffffffeull) throw std::out_of_range("eds_merge_adjacent_host: position out of range");
        ctx->leds->merge_host(eds, eds_bytes, seds, seds_bytes, 0, false, 0, eds_out, seds_out, nullptr, nullptr, false, {}, nullptr,
                              &pos1);
    });
    if (rc != EDS_OK) {
        if (eds_out) eds_buffer_free_host(eds_out);
        if (seds_out) eds_buffer_free_host(seds_out);
    }
    return rc;
}

eds_status eds_is_leds_host(eds_ctx* ctx, const uint8_t* eds_in, uint64_t eds_bytes, uint32_t l, int* is_leds_out) {
    return guarded([&] {
        use_device(ctx);
        if (!eds_in || !is_leds_out) throw std::invalid_argument("eds_is_leds_host: null argument");
        *is_leds_out = 1;
        if (l == 0) return;
        ctx->leds->merge_host(eds_in, eds_bytes, nullptr, 0, l, true, 0, nullptr, nullptr, nullptr, is_leds_out);
    });
}

eds_status eds_vcf_transform_device(eds_ctx* ctx, const uint8_t* vcf, uint64_t vcf_bytes, const uint8_t* fasta,
                                    uint64_t fasta_bytes, eds_buffer* eds_out, eds_buffer* seds_out,
                                    eds_vcf_stats* stats) {
    return guarded([&] {
        use_device(ctx);
        if (!vcf || !fasta || !eds_out || !seds_out) throw std::invalid_argument("eds_vcf_transform_device: null argument");
        if (stats) memset(stats, 0, sizeof(*stats));
        ctx->vcf->transform_device(vcf, vcf_bytes, fasta, fasta_bytes, eds_out, seds_out, stats, nullptr);
    });
}

}  // extern "C"

namespace {
eds_status vcf_transform_host_impl(eds_ctx* ctx, const uint8_t* vcf, uint64_t vcf_bytes, const uint8_t* fasta,
                                  uint64_t fasta_bytes, uint32_t l, eds_buffer* eds_out, eds_buffer* seds_out,
                                  eds_vcf_stats* stats, uint64_t** sv_lines, uint64_t* n_sv_lines, bool view) {
    if (sv_lines) *sv_lines = nullptr;
    if (n_sv_lines) *n_sv_lines = 0;
    if (eds_out) *eds_out = eds_buffer{nullptr, 0};
    if (seds_out) *seds_out = eds_buffer{nullptr, 0};
    if (stats) memset(stats, 0, sizeof(*stats));
    eds_status rc = guarded([&] {
        use_device(ctx);
        if ((!vcf && vcf_bytes) || (!fasta && fasta_bytes) || !eds_out || !seds_out)
            throw std::invalid_argument("eds_vcf_transform_host: null argument");
        const bool trace = getenv("EDSB_TRACE_HOST") != nullptr;
        auto t0 = std::chrono::steady_clock::now();
        auto lap = [&](const char* what) {
            if (!trace) return;
            cudaStreamSynchronize(ctx->stream);
            const auto t1 = std::chrono::steady_clock::now();
            fprintf(stderr, "[edsb] %-22s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
            t0 = t1;
        };
        ctx->vcf_in[0].reserve(vcf_bytes + 16);
        ctx->vcf_in[1].reserve(fasta_bytes + 16);
        if (vcf_bytes) EDSB_CUDA(cudaMemcpyAsync(ctx->vcf_in[0].p, vcf, vcf_bytes, cudaMemcpyHostToDevice, ctx->stream));
        if (fasta_bytes) EDSB_CUDA(cudaMemcpyAsync(ctx->vcf_in[1].p, fasta, fasta_bytes, cudaMemcpyHostToDevice, ctx->stream));
        lap("H2D");
        std::vector<uint64_t> sv;
        eds_buffer d_eds{nullptr, 0}, d_seds{nullptr, 0};
        ctx->vcf->transform_device(ctx->vcf_in[0].as<uint8_t>(), vcf_bytes, ctx->vcf_in[1].as<uint8_t>(), fasta_bytes, &d_eds,
                                   &d_seds, stats, &sv);
        if (sv_lines && n_sv_lines && !sv.empty()) {
            *sv_lines = static_cast<uint64_t*>(malloc(sv.size() * sizeof(uint64_t)));
            if (!*sv_lines) throw std::bad_alloc();
            memcpy(*sv_lines, sv.data(), sv.size() * sizeof(uint64_t));
            *n_sv_lines = sv.size();
        }
        lap("front end");
        if (l == 0) {
            eds_out->data = view ? to_host_view(ctx, 0, d_eds) : to_host(ctx, d_eds);
            eds_out->bytes = d_eds.bytes;
            seds_out->data = view ? to_host_view(ctx, 1, d_seds) : to_host(ctx, d_seds);
            seds_out->bytes = d_seds.bytes;
            lap("D2H");
        } else {
            // parse_vcf_to_leds_streaming :750-752: LINEAR merge of the text just produced, still in HBM
            uint32_t rounds = 0;
            const uint32_t launches = ctx->clock.launches;
            std::function<uint8_t*(int, uint64_t)> sink;
            if (view)
                sink = [ctx](int which, uint64_t bytes) -> uint8_t* { return view_slot(ctx, which, bytes); };
            ctx->leds->merge_host(d_eds.data, d_eds.bytes, d_seds.data, d_seds.bytes, l, true, 0, eds_out, seds_out, &rounds,
                                  nullptr, true, sink);
            if (stats) {
                stats->leds_rounds = rounds;
                stats->gpu_launches = launches + ctx->clock.launches;
            }
        }
    });
    if (rc != EDS_OK) {
        if (view) {
            if (eds_out) *eds_out = eds_buffer{nullptr, 0};
            if (seds_out) *seds_out = eds_buffer{nullptr, 0};
        } else {
            if (eds_out) eds_buffer_free_host(eds_out);
            if (seds_out) eds_buffer_free_host(seds_out);
        }
    }
    return rc;
}
}  // namespace

extern "C" {

eds_status eds_vcf_transform_host(eds_ctx* ctx, const uint8_t* vcf, uint64_t vcf_bytes, const uint8_t* fasta,
                                  uint64_t fasta_bytes, uint32_t l, eds_buffer* eds_out, eds_buffer* seds_out,
                                  eds_vcf_stats* stats, uint64_t** sv_lines, uint64_t* n_sv_lines) {
    return vcf_transform_host_impl(ctx, vcf, vcf_bytes, fasta, fasta_bytes, l, eds_out, seds_out, stats, sv_lines, n_sv_lines, false);
}

eds_status eds_vcf_transform_host_view(eds_ctx* ctx, const uint8_t* vcf, uint64_t vcf_bytes, const uint8_t* fasta,
                                       uint64_t fasta_bytes, uint32_t l, eds_buffer* eds_out, eds_buffer* seds_out,
                                       eds_vcf_stats* stats, uint64_t** sv_lines, uint64_t* n_sv_lines) {
    return vcf_transform_host_impl(ctx, vcf, vcf_bytes, fasta, fasta_bytes, l, eds_out, seds_out, stats, sv_lines, n_sv_lines, true);
}

eds_status eds_device_upload(eds_ctx* ctx, const uint8_t* host, uint64_t bytes, uint8_t** device_out) {
    return guarded([&] {
        use_device(ctx);
        if (!device_out || (!host && bytes)) throw std::invalid_argument("eds_device_upload: null argument");
        void* p = nullptr;
        EDSB_CUDA(cudaMalloc(&p, ((bytes + 16 + 255) / 256) * 256));
        if (bytes) {
            cudaError_t e = cudaMemcpyAsync(p, host, bytes, cudaMemcpyHostToDevice, ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            if (e != cudaSuccess) {
                cudaFree(p);
                throw edsb::CudaError(std::string("host to device copy: ") + cudaGetErrorString(e));
            }
        }
        *device_out = static_cast<uint8_t*>(p);
    });
}

void eds_device_free(eds_ctx* ctx, uint8_t* device) {
    if (!ctx || !device) return;
    cudaSetDevice(ctx->device);
    cudaFree(device);
}

}  // extern "C"
