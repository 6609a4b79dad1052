// ctx.cu -- context, device arena and memory plumbing of the C ABI (include/lele_b200.h).
// Maps lele's host-side static buffer arena (src/tensor.rs TensorView over Vec<f32>,
// kernels/utils.rs:10 ensure_capacity, <Model>Workspace in src/compiler/mod.rs:1057-1092)
// onto HBM: every workspace Vec / the weights blob gets a grow-only device mirror.
#include "common.cuh"
#include <stdlib.h>
#include <stdarg.h>
#include <algorithm>
#include <map>
#include <mutex>

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
    if (cudaMalloc((void**)&c->dev_err, sizeof(int)) != cudaSuccess || cudaMemset(c->dev_err, 0, sizeof(int)) != cudaSuccess) {
        cudaGetLastError();
        if (c->own_stream) cudaStreamDestroy(c->stream);
        delete c;
        lb_set_error("ctx_create: cannot allocate the device error word");
        return LELE_B200_ERR_CUDA;
    }
    *out = c;
    return LELE_B200_OK;
}

extern "C" int lele_b200_ctx_destroy(lele_b200_ctx* ctx) {
    if (!ctx) return LELE_B200_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    lb_flush_deferred_free(ctx);
    for (auto& kv : ctx->arena) cudaFree(kv.second.dptr);
    for (auto& kv : ctx->tables) cudaFree(kv.second);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->scratch2) cudaFree(ctx->scratch2);
    if (ctx->dev_err) cudaFree(ctx->dev_err);
    if (ctx->ev) cudaEventDestroy(ctx->ev);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return LELE_B200_OK;
}

extern "C" int lele_b200_sync(lele_b200_ctx* ctx) {
    LB_REQUIRE(ctx, "sync: NULL ctx");
    LB_ENTER(ctx);
    LB_REQUIRE(!lb_stream_capturing(ctx), "sync: the context's stream is being captured into a graph (nothing on a captured path may join the stream)");
    if (ctx->dev_err_armed) {
        int flag = 0;
        LB_CHECK_CUDA(cudaMemcpyAsync(&flag, ctx->dev_err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        LB_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
        ctx->dev_err_armed = false;
        lb_flush_deferred_free(ctx);
        if (flag) {
            cudaMemsetAsync(ctx->dev_err, 0, sizeof(int), ctx->stream);
            lb_set_error("%s: index out of range for the gathered axis (the reference panics, manipulation.rs:589 / conv2d.rs:1438)",
                         flag == 2 ? "gather_elements" : "gather");
            return LELE_B200_ERR_ARG;
        }
        return LELE_B200_OK;
    }
    LB_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    lb_flush_deferred_free(ctx);
    return LELE_B200_OK;
}

extern "C" unsigned long long lele_b200_launch_count(const lele_b200_ctx* ctx) { return ctx ? ctx->launches : 0ull; }

extern "C" int lele_b200_malloc(lele_b200_ctx* ctx, size_t nbytes, void** dptr) {
    LB_REQUIRE(ctx && dptr, "malloc: NULL argument");
    LB_ENTER(ctx);
    LB_CHECK_CUDA(cudaSetDevice(ctx->device));
    LB_CHECK_CUDA(cudaMalloc(dptr, nbytes ? nbytes : 16));
    return LELE_B200_OK;
}
extern "C" int lele_b200_free(lele_b200_ctx* ctx, void* dptr) {
    LB_REQUIRE(ctx, "free: NULL ctx");
    LB_ENTER(ctx);
    if (dptr) {
        lb_tmap_forget_range(ctx, dptr, 0);          // descriptors of a freed tensor must not outlive it
        if (lb_stream_capturing(ctx)) { ctx->deferred_free.push_back(dptr); return LELE_B200_OK; }   // (e.g. a host GC running during a capture)
        LB_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
        lb_flush_deferred_free(ctx);
        LB_CHECK_CUDA(cudaFree(dptr));
    }
    return LELE_B200_OK;
}
// page-locked host memory: what a host hands to h2d / d2h when the copies are to run asynchronously at PCIe rate (a pageable buffer is
// staged by the driver at a fraction of it)
extern "C" int lele_b200_malloc_host(lele_b200_ctx* ctx, size_t nbytes, void** hptr) {
    LB_REQUIRE(ctx && hptr, "malloc_host: NULL argument");
    LB_ENTER(ctx);
    LB_CHECK_CUDA(cudaMallocHost(hptr, nbytes ? nbytes : 16));
    return LELE_B200_OK;
}
extern "C" int lele_b200_free_host(lele_b200_ctx* ctx, void* hptr) {
    LB_REQUIRE(ctx, "free_host: NULL ctx");
    LB_ENTER(ctx);
    if (hptr) {
        if (lb_stream_capturing(ctx)) { lb_set_error("free_host: the context's stream is being captured"); return LELE_B200_ERR_ARG; }
        LB_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
        LB_CHECK_CUDA(cudaFreeHost(hptr));
    }
    return LELE_B200_OK;
}
extern "C" int lele_b200_memset(lele_b200_ctx* ctx, void* dptr, int value, size_t nbytes) {
    LB_REQUIRE(ctx, "memset: NULL ctx");
    LB_ENTER(ctx);
    LB_CHECK_CUDA(cudaMemsetAsync(dptr, value, nbytes, ctx->stream));
    return LELE_B200_OK;
}
extern "C" int lele_b200_h2d(lele_b200_ctx* ctx, void* dst, const void* src, size_t nbytes) {
    LB_REQUIRE(ctx, "h2d: NULL ctx");
    LB_ENTER(ctx);
    if (nbytes) LB_CHECK_CUDA(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyHostToDevice, ctx->stream));
    return LELE_B200_OK;
}
extern "C" int lele_b200_d2h(lele_b200_ctx* ctx, void* dst, const void* src, size_t nbytes) {
    LB_REQUIRE(ctx, "d2h: NULL ctx");
    LB_ENTER(ctx);
    if (nbytes) LB_CHECK_CUDA(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDeviceToHost, ctx->stream));
    return LELE_B200_OK;
}
extern "C" int lele_b200_d2d(lele_b200_ctx* ctx, void* dst, const void* src, size_t nbytes) {
    LB_REQUIRE(ctx, "d2d: NULL ctx");
    LB_ENTER(ctx);
    if (nbytes) LB_CHECK_CUDA(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return LELE_B200_OK;
}

extern "C" int lele_b200_arena_bind(lele_b200_ctx* ctx, const void* host_base, size_t nbytes, void** dptr) {
    LB_REQUIRE(ctx && host_base && dptr, "arena_bind: NULL argument");
    LB_ENTER(ctx);
    auto it = ctx->arena.find(host_base);
    if (it != ctx->arena.end() && it->second.bytes >= nbytes) { *dptr = it->second.dptr; return LELE_B200_OK; }
    void* p = nullptr;
    LB_CHECK_CUDA(cudaSetDevice(ctx->device));
    LB_CHECK_CUDA(cudaMalloc(&p, nbytes ? nbytes : 16));
    if (it != ctx->arena.end()) {  // grow: keep contents, like Vec::reserve
        LB_CHECK_CUDA(cudaMemcpyAsync(p, it->second.dptr, it->second.bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        LB_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
        lb_tmap_forget_range(ctx, it->second.dptr, it->second.bytes);
        cudaFree(it->second.dptr);
        it->second = {p, nbytes};
    } else ctx->arena[host_base] = {p, nbytes};
    *dptr = p;
    return LELE_B200_OK;
}
extern "C" int lele_b200_arena_release(lele_b200_ctx* ctx, const void* host_base) {
    LB_REQUIRE(ctx, "arena_release: NULL ctx");
    LB_ENTER(ctx);
    auto it = ctx->arena.find(host_base);
    if (it != ctx->arena.end()) {
        LB_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
        lb_tmap_forget_range(ctx, it->second.dptr, it->second.bytes);
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
        if (ctx->scratch) { lb_tmap_forget_range(ctx, ctx->scratch, ctx->scratch_bytes); cudaFree(ctx->scratch); }
        ctx->scratch = nullptr; ctx->scratch_bytes = 0;
        size_t want = bytes + (bytes >> 2) + 4096;
        LB_CHECK_CUDA(cudaMalloc(&ctx->scratch, want));
        ctx->scratch_bytes = want;
    }
    *out = ctx->scratch;
    return LELE_B200_OK;
}

int lb_scratch2(lele_b200_ctx* ctx, size_t bytes, void** out) {
    if (bytes > ctx->scratch2_bytes) {
        LB_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->scratch2) { lb_tmap_forget_range(ctx, ctx->scratch2, ctx->scratch2_bytes); cudaFree(ctx->scratch2); }
        ctx->scratch2 = nullptr; ctx->scratch2_bytes = 0;
        size_t want = bytes + (bytes >> 2) + 4096;
        LB_CHECK_CUDA(cudaMalloc(&ctx->scratch2, want));
        ctx->scratch2_bytes = want;
    }
    *out = ctx->scratch2;
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


bool lb_stream_capturing(lele_b200_ctx* ctx) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(ctx->stream, &st) != cudaSuccess) { cudaGetLastError(); return ctx->capturing; }
    return st != cudaStreamCaptureStatusNone;
}

void lb_flush_deferred_free(lele_b200_ctx* ctx) {
    for (void* p : ctx->deferred_free) cudaFree(p);
    ctx->deferred_free.clear();
}

int lb_enter(lele_b200_ctx* ctx) {
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess || cur != ctx->device) LB_CHECK_CUDA(cudaSetDevice(ctx->device));
    return LELE_B200_OK;
}

// The opt-in is a property of (device, kernel), shared by every context on that device, and must only ever be RAISED: a second
// context asking for less would otherwise lower it under a context that already launches with more.  One process-wide table,
// keyed by (device, function), guarded for hosts that drive several contexts from several threads.
int lb_func_smem(lele_b200_ctx* ctx, const void* func, size_t bytes) {
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, size_t> granted;
    std::lock_guard<std::mutex> lock(mu);
    size_t& cur = granted[std::make_pair(ctx->device, func)];
    if (cur >= bytes && cur != 0) return LELE_B200_OK;
    LB_CHECK_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes ? bytes : 16)));
    cur = bytes ? bytes : 16;
    return LELE_B200_OK;
}

static unsigned long long lb_tmap_hash(const unsigned long long (&key)[10]) {
    unsigned long long h = 0x746d6170ull;
    for (int i = 0; i < 10; ++i) h = lb_hash_mix(h, key[i]);
    return h;
}
bool lb_tmap_lookup(lele_b200_ctx* ctx, const unsigned long long (&key)[10], void* blob128) {
    auto it = ctx->tmaps.find(lb_tmap_hash(key));
    if (it == ctx->tmaps.end() || memcmp(it->second.key, key, sizeof(key)) != 0) return false;   // a hash collision is a miss
    it->second.stamp = ++ctx->tmap_clock;
    memcpy(blob128, it->second.blob, 128);
    return true;
}
void lb_tmap_store(lele_b200_ctx* ctx, const unsigned long long (&key)[10], const void* blob128) {
    if (ctx->tmaps.size() >= LB_TMAP_CACHE_MAX) {        // drop the least recently used half
        std::vector<unsigned long long> stamps;
        stamps.reserve(ctx->tmaps.size());
        for (auto& kv : ctx->tmaps) stamps.push_back(kv.second.stamp);
        std::nth_element(stamps.begin(), stamps.begin() + stamps.size() / 2, stamps.end());
        const unsigned long long cut = stamps[stamps.size() / 2];
        for (auto it = ctx->tmaps.begin(); it != ctx->tmaps.end();) it = it->second.stamp < cut ? ctx->tmaps.erase(it) : ++it;
    }
    lele_b200_ctx::TmapEntry e;
    memcpy(e.key, key, sizeof(key));
    e.stamp = ++ctx->tmap_clock;
    memcpy(e.blob, blob128, 128);
    ctx->tmaps[lb_tmap_hash(key)] = e;                   // a colliding older entry is replaced
}
void lb_tmap_forget_range(lele_b200_ctx* ctx, const void* base, size_t bytes) {
    const unsigned long long lo = (unsigned long long)(uintptr_t)base, hi = lo + (bytes ? bytes : 1);
    for (auto it = ctx->tmaps.begin(); it != ctx->tmaps.end();) {
        const unsigned long long p = it->second.key[1];
        it = (p >= lo && p < hi) ? ctx->tmaps.erase(it) : ++it;
    }
}

// ---- stream fork / join and CUDA-graph capture of a sequence of C-ABI calls ----------------------------------------------
// A replayed model.rs is a few hundred short launches per forward; after its first (arena-sizing) forward nothing on its path
// allocates or synchronises, so the whole call sequence -- over several contexts (streams) of one device when independent
// inputs are processed side by side -- is captured once and replayed as one graph (the analogue of what the SenseVoice runner
// does internally, csrc/sensevoice.cu).
struct lele_b200_graph { cudaGraphExec_t exec = nullptr; unsigned long long launches = 0; };

static int lb_ctx_event(lele_b200_ctx* ctx) {
    if (!ctx->ev) LB_CHECK_CUDA(cudaEventCreateWithFlags(&ctx->ev, cudaEventDisableTiming));
    return LELE_B200_OK;
}

extern "C" int lele_b200_stream_fork(lele_b200_ctx* ctx, lele_b200_ctx* lane) {
    LB_REQUIRE(ctx && lane, "stream_fork: NULL argument");
    LB_ENTER(ctx);
    LB_REQUIRE(ctx->device == lane->device, "stream_fork: contexts live on different devices (%d, %d)", ctx->device, lane->device);
    if (ctx == lane || ctx->stream == lane->stream) return LELE_B200_OK;
    int rc = lb_ctx_event(ctx);
    if (rc) return rc;
    LB_CHECK_CUDA(cudaEventRecord(ctx->ev, ctx->stream));
    LB_CHECK_CUDA(cudaStreamWaitEvent(lane->stream, ctx->ev, 0));
    return LELE_B200_OK;
}

extern "C" int lele_b200_stream_join(lele_b200_ctx* ctx, lele_b200_ctx* lane) {
    LB_REQUIRE(ctx && lane, "stream_join: NULL argument");
    LB_ENTER(ctx);
    LB_REQUIRE(ctx->device == lane->device, "stream_join: contexts live on different devices (%d, %d)", ctx->device, lane->device);
    if (ctx == lane || ctx->stream == lane->stream) return LELE_B200_OK;
    int rc = lb_ctx_event(lane);
    if (rc) return rc;
    LB_CHECK_CUDA(cudaEventRecord(lane->ev, lane->stream));
    LB_CHECK_CUDA(cudaStreamWaitEvent(ctx->stream, lane->ev, 0));
    return LELE_B200_OK;
}

extern "C" int lele_b200_capture_begin(lele_b200_ctx* ctx) {
    LB_REQUIRE(ctx, "capture_begin: NULL ctx");
    LB_ENTER(ctx);
    LB_REQUIRE(!ctx->capturing, "capture_begin: this context is already capturing");
    LB_CHECK_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
    ctx->capturing = true;
    ctx->capture_l0 = ctx->launches;
    return LELE_B200_OK;
}

extern "C" int lele_b200_capture_end(lele_b200_ctx* ctx, unsigned long long lane_launches, lele_b200_graph** out) {
    LB_REQUIRE(ctx && out, "capture_end: NULL argument");
    LB_ENTER(ctx);
    LB_REQUIRE(ctx->capturing, "capture_end: no capture in progress on this context");
    ctx->capturing = false;
    cudaGraph_t graph = nullptr;
    cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
    const unsigned long long captured = ctx->launches - ctx->capture_l0 + lane_launches;
    ctx->launches = ctx->capture_l0;            // captured, not executed
    if (ce != cudaSuccess) {
        cudaGetLastError();
        lb_set_error("capture_end: %s (a call on the captured path synchronised or used an un-forked stream)", cudaGetErrorString(ce));
        return LELE_B200_ERR_CUDA;
    }
    lele_b200_graph* g = new lele_b200_graph();
    ce = cudaGraphInstantiate(&g->exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) { delete g; lb_set_error("capture_end: cudaGraphInstantiate: %s", cudaGetErrorString(ce)); return LELE_B200_ERR_CUDA; }
    g->launches = captured;
    cudaGraphUpload(g->exec, ctx->stream);
    cudaGetLastError();
    *out = g;
    return LELE_B200_OK;
}

extern "C" int lele_b200_graph_launch(lele_b200_ctx* ctx, lele_b200_graph* g) {
    LB_REQUIRE(ctx && g && g->exec, "graph_launch: NULL argument");
    LB_ENTER(ctx);
    LB_CHECK_CUDA(cudaGraphLaunch(g->exec, ctx->stream));
    ctx->launches += g->launches;
    return LELE_B200_OK;
}

extern "C" int lele_b200_graph_destroy(lele_b200_ctx* ctx, lele_b200_graph* g) {
    if (!g) return LELE_B200_OK;
    if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
    if (g->exec) cudaGraphExecDestroy(g->exec);
    delete g;
    return LELE_B200_OK;
}
