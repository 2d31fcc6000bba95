// Context, device-memory helpers, CUDA-event timing and the synthetic workload generator kernel.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/mc_synth.h"
#include "mc_common.cuh"

extern "C" int mc_ctx_create(int device, mc_ctx** out) {
    try {
        if (!out) mc_throw(MC_ERR_ARG, "null argument");
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n == 0) {
            cudaGetLastError();
            mc_throw(MC_ERR_CUDA, "no CUDA device available (%s): libmcgpu has no CPU fallback", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        }
        if (device < 0 || device >= n) mc_throw(MC_ERR_ARG, "device %d out of range (0..%d)", device, n - 1);
        // Host synchronisation policy: MC_SYNC=spin|block|yield|auto (default: the driver's choice).
        {
            unsigned flags = cudaDeviceScheduleAuto;   // (spinning was measured no better on the shared boxes: profiles/r2_NOTES.md)
            if (const char* e = getenv("MC_SYNC")) {
                if (!strcmp(e, "spin")) flags = cudaDeviceScheduleSpin;
                else if (!strcmp(e, "block")) flags = cudaDeviceScheduleBlockingSync;
                else if (!strcmp(e, "yield")) flags = cudaDeviceScheduleYield;
                else if (!strcmp(e, "auto")) flags = cudaDeviceScheduleAuto;
            }
            cudaSetDevice(device);
            cudaSetDeviceFlags(flags);
            cudaGetLastError();
        }
        MC_CUDA(cudaSetDevice(device));
        mc_ctx* c = new mc_ctx();
        c->device = device;
        cudaDeviceProp prop;
        MC_CUDA(cudaGetDeviceProperties(&prop, device));
        c->sm_count = prop.multiProcessorCount;
        if (prop.major < 10) {
            delete c;
            mc_throw(MC_ERR_CUDA, "device %d is sm_%d%d; libmcgpu is built for sm_100a only", device, prop.major, prop.minor);
        }
        // The count table is probed one 32-byte sector at a time at random addresses.  With the default L2 fetch
        // granularity every miss pulls a whole 128-byte line from HBM (measured: 4 DRAM sectors per probe, ncu
        // profiles/r1a); ask for sector-sized fills instead.  MC_L2_FETCH=32|64|128 overrides (for experiments).
        {
            size_t gran = 32;
            if (const char* e = getenv("MC_L2_FETCH")) gran = (size_t)atoi(e);
            if (gran == 32 || gran == 64 || gran == 128) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
            cudaGetLastError();
        }
        // tables, rehash targets, bins and temporaries come from the context's own block arena (mc_dev_alloc): cudaMalloc /
        // cudaFree of multi-GB blocks cost 20-800 ms each on this part (measured), and the stream-ordered driver pool of
        // round 1 still went back to the driver whenever a step's largest blocks missed its cache (profiles/r1d_NOTES.md)
        MC_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        MC_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        MC_CUDA(cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
        for (auto& ev : c->ev) MC_CUDA(cudaEventCreate(&ev));
        for (auto& ev : c->stage_ev) MC_CUDA(cudaEventCreate(&ev));
        MC_CUDA(cudaMallocHost(&c->h_pinned_small, 4096));
        *out = c;
        return MC_OK;
    } catch (const McError& e) {
        mc_set_error(e.what());
        return e.code;
    }
}

extern "C" void mc_ctx_destroy(mc_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(c->copy_stream);
    cudaStreamSynchronize(c->aux_stream);
    mc_resolve_timers(c);
    for (int i = 0; i < 2; i++) {
        cudaFree(c->d_chunk[i]);
        cudaFreeHost(c->h_stage[i]);
    }
    for (int i = 0; i < 4; i++) cudaFree(c->d_scratch[i]);
    mc_arena_trim(c);
    for (auto& kv : c->arena_live) cudaFree(kv.first);   // tables the caller never freed die with their context
    c->arena_live.clear();
    for (auto& ev : c->event_pool) cudaEventDestroy(ev);
    cudaFreeHost(c->h_pinned_small);
    for (auto& ev : c->ev) cudaEventDestroy(ev);
    for (auto& ev : c->stage_ev) cudaEventDestroy(ev);
    cudaStreamDestroy(c->stream);
    cudaStreamDestroy(c->copy_stream);
    cudaStreamDestroy(c->aux_stream);
    delete c;
}

extern "C" int mc_ctx_set_chunk_bytes(mc_ctx* c, size_t bytes) {
    if (!c || bytes < 4096 || bytes > ((size_t)3 << 30)) {
        mc_set_error("chunk size must be in [4 KiB, 3 GiB]");
        return MC_ERR_ARG;
    }
    c->chunk_bytes = bytes;
    return MC_OK;
}
extern "C" uint64_t mc_ctx_launches(const mc_ctx* c) { return c ? c->launches : 0; }
extern "C" int mc_ctx_trim(mc_ctx* c) {
    if (!c) {
        mc_set_error("null argument");
        return MC_ERR_ARG;
    }
    cudaSetDevice(c->device);
    mc_arena_trim(c);
    return MC_OK;
}
extern "C" int mc_ctx_arena_stats(mc_ctx* c, uint64_t* cached_bytes, uint64_t* driver_allocs) {
    if (!c) {
        mc_set_error("null argument");
        return MC_ERR_ARG;
    }
    if (cached_bytes) *cached_bytes = mc_arena_cached_bytes(c);
    if (driver_allocs) *driver_allocs = c->arena_driver_allocs;
    return MC_OK;
}

void* mc_scratch(mc_ctx* ctx, int slot, size_t bytes) {
    if (bytes > ctx->d_scratch_cap[slot]) {
        MC_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->d_scratch[slot]);
        ctx->d_scratch[slot] = nullptr;
        ctx->d_scratch_cap[slot] = 0;
        size_t cap = bytes + bytes / 8 + 4096;
        cudaError_t e = cudaMalloc(&ctx->d_scratch[slot], cap);
        if (e != cudaSuccess) {
            cudaGetLastError();
            mc_throw(MC_ERR_NOMEM, "cannot allocate %.1f MiB of device scratch", (double)cap / (1 << 20));
        }
        ctx->d_scratch_cap[slot] = cap;
        ctx->scratch_allocs++;
    }
    return ctx->d_scratch[slot];
}

#define MC_API_TRY try {
#define MC_API_END                  \
    return MC_OK;                   \
    }                               \
    catch (const McError& e) {      \
        mc_set_error(e.what());     \
        return e.code;              \
    }

extern "C" int mc_device_alloc(mc_ctx* c, size_t bytes, void** p) {
    MC_API_TRY
    if (!c || !p) mc_throw(MC_ERR_ARG, "null argument");
    MC_CUDA(cudaSetDevice(c->device));
    MC_CUDA(cudaMalloc(p, bytes ? bytes : 1));
    MC_API_END
}
extern "C" int mc_device_free(mc_ctx* c, void* p) {
    MC_API_TRY
    if (!c) mc_throw(MC_ERR_ARG, "null argument");
    MC_CUDA(cudaSetDevice(c->device));
    MC_CUDA(cudaStreamSynchronize(c->stream));
    MC_CUDA(cudaFree(p));
    MC_API_END
}
extern "C" int mc_memcpy_d2h(mc_ctx* c, void* dst, const void* src, size_t bytes) {
    MC_API_TRY
    MC_CUDA(cudaSetDevice(c->device));
    MC_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
    MC_CUDA(cudaStreamSynchronize(c->stream));
    MC_API_END
}
extern "C" int mc_memcpy_h2d(mc_ctx* c, void* dst, const void* src, size_t bytes) {
    MC_API_TRY
    MC_CUDA(cudaSetDevice(c->device));
    MC_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
    MC_CUDA(cudaStreamSynchronize(c->stream));
    MC_API_END
}
extern "C" int mc_ctx_sync(mc_ctx* c) {
    MC_API_TRY
    MC_CUDA(cudaSetDevice(c->device));
    MC_CUDA(cudaStreamSynchronize(c->stream));
    MC_CUDA(cudaStreamSynchronize(c->copy_stream));
    MC_API_END
}
extern "C" int mc_event_record(mc_ctx* c, int slot) {
    MC_API_TRY
    if (slot < 0 || slot >= 16) mc_throw(MC_ERR_ARG, "event slot out of range");
    MC_CUDA(cudaEventRecord(c->ev[slot], c->stream));
    MC_API_END
}
extern "C" int mc_event_elapsed_ms(mc_ctx* c, int a, int b, float* ms) {
    MC_API_TRY
    if (a < 0 || a >= 16 || b < 0 || b >= 16 || !ms) mc_throw(MC_ERR_ARG, "bad event slot");
    MC_CUDA(cudaEventSynchronize(c->ev[b]));
    MC_CUDA(cudaEventElapsedTime(ms, c->ev[a], c->ev[b]));
    MC_API_END
}
extern "C" int mc_ctx_stage_ms(mc_ctx* c, float* ms3) {
    MC_API_TRY
    if (!c || !ms3) mc_throw(MC_ERR_ARG, "null argument");
    mc_resolve_timers(c);
    for (int i = 0; i < 3; i++) ms3[i] = c->stage_ms[i];
    MC_API_END
}
extern "C" int mc_ctx_stage_reset(mc_ctx* c) {
    MC_API_TRY
    if (!c) mc_throw(MC_ERR_ARG, "null argument");
    mc_resolve_timers(c);
    for (auto& v : c->stage_ms) v = 0;
    for (auto& v : c->host_ms) v = 0;
    MC_API_END
}
extern "C" int mc_ctx_host_ms(mc_ctx* c, double* ms8) {
    MC_API_TRY
    if (!c || !ms8) mc_throw(MC_ERR_ARG, "null argument");
    for (int i = 0; i < MC_N_STAGES; i++) ms8[i] = c->host_ms[i];
    MC_API_END
}
extern "C" int mc_ctx_stage_ms_all(mc_ctx* c, float* ms8) {
    MC_API_TRY
    if (!c || !ms8) mc_throw(MC_ERR_ARG, "null argument");
    mc_resolve_timers(c);
    for (int i = 0; i < MC_N_STAGES; i++) ms8[i] = c->stage_ms[i];
    MC_API_END
}

// --------------------------------------------------------------------------------------------
// synthetic reads on the device (bench/test utility; twin of oracle/mc_synth_cpu.cpp)
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_synth(const mc_synth_params* __restrict__ p, uint64_t first, uint64_t count, uint8_t* out) {
    const uint32_t rb = mc_synth_record_bytes(p);
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x)
        mc_synth_write_record(p, first + i, out + i * rb);
}

extern "C" int mc_synth_generate_device(mc_ctx* c, const void* params, size_t params_size, uint64_t first, uint64_t count, void* d_out) {
    MC_API_TRY
    if (!c || !params || !d_out) mc_throw(MC_ERR_ARG, "null argument");
    if (params_size != sizeof(mc_synth_params)) mc_throw(MC_ERR_ARG, "mc_synth_params size mismatch (%zu vs %zu)", params_size, sizeof(mc_synth_params));
    MC_CUDA(cudaSetDevice(c->device));
    mc_synth_params* d_p;
    MC_CUDA(cudaMalloc(&d_p, sizeof(mc_synth_params)));
    MC_CUDA(cudaMemcpyAsync(d_p, params, sizeof(mc_synth_params), cudaMemcpyHostToDevice, c->stream));
    uint64_t blocks = (count + 127) / 128;
    if (blocks > (uint64_t)c->sm_count * 32) blocks = (uint64_t)c->sm_count * 32;
    if (blocks < 1) blocks = 1;
    k_synth<<<(int)blocks, 128, 0, c->stream>>>(d_p, first, count, (uint8_t*)d_out);
    MC_CUDA(cudaGetLastError());
    MC_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_p);
    MC_API_END
}
