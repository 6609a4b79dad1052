// ctx.cu -- context, device arena and memory plumbing of the C ABI (include/lele_b200.h).
// Maps lele's host-side static buffer arena (src/tensor.rs TensorView over Vec<f32>,
// kernels/utils.rs:10 ensure_capacity, <Model>Workspace in src/compiler/mod.rs:1057-1092)
// onto HBM: every workspace Vec / the weights blob gets a grow-only device mirror.
#include "common.cuh"
#include <stdlib.h>
#include <stdarg.h>

static thread_local char g_err[1024] = "";

void lb_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* lele_b200_last_error(void) { return g_err; }

extern "C" int lele_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int lele_b200_ctx_create(int device, void* stream, lele_b200_ctx** out) {
    LB_REQUIRE(out != nullptr, "ctx_create: out is NULL");
    int n = 0;
    LB_CHECK_CUDA(cudaGetDeviceCount(&n));
    LB_REQUIRE(device >= 0 && device < n, "ctx_create: device %d out of range (have %d)", device, n);
    LB_CHECK_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    LB_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        lb_set_error("ctx_create: device %d is sm_%d%d; this library is built for sm_100a only", device,
                     prop.major, prop.minor);
        return LELE_B200_ERR_UNSUPPORTED;
    }
    lele_b200_ctx* c = new lele_b200_ctx();
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    if (stream) { c->stream = (cudaStream_t)stream; c->own_stream = false; }
    else {
        cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { delete c; lb_set_error("cudaStreamCreate: %s", cudaGetErrorString(e)); return LELE_B200_ERR_CUDA; }
        c->own_stream = true;
    }
    *out = c;
    return LELE_B200_OK;
}

extern "C" int lele_b200_ctx_destroy(lele_b200_ctx* ctx) {
    if (!ctx) return LELE_B200_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto& kv : ctx->arena) cudaFree(kv.second.dptr);
    for (auto& kv : ctx->tables) cudaFree(kv.second);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return LELE_B200_OK;
}

extern "C" int lele_b200_sync(lele_b200_ctx* ctx) {
    LB_REQUIRE(ctx, "sync: NULL ctx");
    LB_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return LELE_B200_OK;
}

extern "C" unsigned long long lele_b200_launch_count(const lele_b200_ctx* ctx) { return ctx ? ctx->launches : 0ull; }

extern "C" int lele_b200_malloc(lele_b200_ctx* ctx, size_t nbytes, void** dptr) {
    LB_REQUIRE(ctx && dptr, "malloc: NULL argument");
    LB_CHECK_CUDA(cudaSetDevice(ctx->device));
    LB_CHECK_CUDA(cudaMalloc(dptr, nbytes ? nbytes : 16));
    return LELE_B200_OK;
}
extern "C" int lele_b200_free(lele_b200_ctx* ctx, void* dptr) {
    LB_REQUIRE(ctx, "free: NULL ctx");
    if (dptr) { LB_CHECK_CUDA(cudaStreamSynchronize(ctx->stream)); LB_CHECK_CUDA(cudaFree(dptr)); }
    return LELE_B200_OK;
}
extern "C" int lele_b200_memset(lele_b200_ctx* ctx, void* dptr, int value, size_t nbytes) {
    LB_REQUIRE(ctx, "memset: NULL ctx");
    LB_CHECK_CUDA(cudaMemsetAsync(dptr, value, nbytes, ctx->stream));
    return LELE_B200_OK;
}
extern "C" int lele_b200_h2d(lele_b200_ctx* ctx, void* dst, const void* src, size_t nbytes) {
    LB_REQUIRE(ctx, "h2d: NULL ctx");
    if (nbytes) LB_CHECK_CUDA(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyHostToDevice, ctx->stream));
    return LELE_B200_OK;
}
extern "C" int lele_b200_d2h(lele_b200_ctx* ctx, void* dst, const void* src, size_t nbytes) {
    LB_REQUIRE(ctx, "d2h: NULL ctx");
    if (nbytes) LB_CHECK_CUDA(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDeviceToHost, ctx->stream));
    return LELE_B200_OK;
}
extern "C" int lele_b200_d2d(lele_b200_ctx* ctx, void* dst, const void* src, size_t nbytes) {
    LB_REQUIRE(ctx, "d2d: NULL ctx");
    if (nbytes) LB_CHECK_CUDA(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return LELE_B200_OK;
}

extern "C" int lele_b200_arena_bind(lele_b200_ctx* ctx, const void* host_base, size_t nbytes, void** dptr) {
    LB_REQUIRE(ctx && host_base && dptr, "arena_bind: NULL argument");
    auto it = ctx->arena.find(host_base);
    if (it != ctx->arena.end() && it->second.bytes >= nbytes) { *dptr = it->second.dptr; return LELE_B200_OK; }
    void* p = nullptr;
    LB_CHECK_CUDA(cudaSetDevice(ctx->device));
    LB_CHECK_CUDA(cudaMalloc(&p, nbytes ? nbytes : 16));
    if (it != ctx->arena.end()) {  // grow: keep contents, like Vec::reserve
        LB_CHECK_CUDA(cudaMemcpyAsync(p, it->second.dptr, it->second.bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        LB_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFree(it->second.dptr);
        it->second = {p, nbytes};
    } else ctx->arena[host_base] = {p, nbytes};
    *dptr = p;
    return LELE_B200_OK;
}
extern "C" int lele_b200_arena_release(lele_b200_ctx* ctx, const void* host_base) {
    LB_REQUIRE(ctx, "arena_release: NULL ctx");
    auto it = ctx->arena.find(host_base);
    if (it != ctx->arena.end()) {
        LB_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFree(it->second.dptr);
        ctx->arena.erase(it);
    }
    return LELE_B200_OK;
}

bool lb_env_flag(const char* name, int dflt) {
    const char* e = getenv(name);
    if (!e || !e[0]) return dflt != 0;
    return e[0] != '0';
}

bool lb_pdl_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("LELE_B200_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
    return v != 0;
}

int lb_scratch(lele_b200_ctx* ctx, size_t bytes, void** out) {
    if (bytes > ctx->scratch_bytes) {
        LB_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->scratch) cudaFree(ctx->scratch);
        ctx->scratch = nullptr; ctx->scratch_bytes = 0;
        size_t want = bytes + (bytes >> 2) + 4096;
        LB_CHECK_CUDA(cudaMalloc(&ctx->scratch, want));
        ctx->scratch_bytes = want;
    }
    *out = ctx->scratch;
    return LELE_B200_OK;
}

int lb_table(lele_b200_ctx* ctx, const std::string& key, const void* host, size_t bytes, void** out) {
    auto it = ctx->tables.find(key);
    if (it != ctx->tables.end()) { *out = it->second; return LELE_B200_OK; }
    void* p = nullptr;
    LB_CHECK_CUDA(cudaMalloc(&p, bytes ? bytes : 16));
    // synchronous copy: `host` is usually a temporary
    LB_CHECK_CUDA(cudaMemcpyAsync(p, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    LB_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->tables[key] = p;
    *out = p;
    return LELE_B200_OK;
}
