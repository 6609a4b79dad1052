// comm.cu -- the two collectives of the clip-sharded path behind the C ABI: one broadcast of the weights blob from
// rank 0 at start-up and one gather of each rank's greedy ids per batch, over NCCL (NVLink 5 / NVSwitch on the B200 box), so that a
// host without torch.distributed (the Rust host of INTEGRATION.md, one process per GPU) can shard clips across the 8 GPUs of a box.
// There is no collective inside the forward pass -- clips are independent end to end -- hence nothing to fuse with a kernel.
//
// NCCL is bound at run time (dlopen, local scope): the library links against nothing but the CUDA runtime, a single-GPU host never
// needs NCCL, and inside a process that already carries an NCCL (torch's bundled one) that loaded copy is reused (RTLD_NOLOAD probe
// by soname first).  LELE_B200_NCCL_LIB names the file explicitly -- the Python package sets it to the pip-installed NCCL that torch
// itself loads, so that importing torch AFTER the first collective still finds the NCCL it was built against.  The 128-byte unique id is created on rank 0 (lele_b200_comm_unique_id) and handed to the other ranks
// by the host's own means (a file, a socket, MPI, torch.distributed's store ...) -- rendez-vous is host plumbing, not part of the path.
#include "common.cuh"
#include <dlfcn.h>
#include <stdlib.h>

namespace {
struct NcclId { char internal[128]; };
typedef void* NcclComm;
typedef int (*FnGetUniqueId)(NcclId*);
typedef int (*FnCommInitRank)(NcclComm*, int, NcclId, int);
typedef int (*FnCommDestroy)(NcclComm);
typedef int (*FnBroadcast)(const void*, void*, size_t, int /*dtype*/, int /*root*/, NcclComm, cudaStream_t);
typedef int (*FnSend)(const void*, size_t, int, int, NcclComm, cudaStream_t);
typedef int (*FnRecv)(void*, size_t, int, int, NcclComm, cudaStream_t);
typedef int (*FnGroup)(void);
typedef const char* (*FnErrStr)(int);
typedef int (*FnGetVersion)(int*);
constexpr int NCCL_UINT8 = 1;   // ncclUint8 (nccl.h ncclDataType_t)

struct NcclApi {
    void* handle = nullptr;
    FnGetUniqueId get_unique_id = nullptr;
    FnCommInitRank comm_init_rank = nullptr;
    FnCommDestroy comm_destroy = nullptr;
    FnBroadcast broadcast = nullptr;
    FnSend send = nullptr;
    FnRecv recv = nullptr;
    FnGroup group_start = nullptr, group_end = nullptr;
    FnErrStr err_str = nullptr;
    FnGetVersion get_version = nullptr;
};

NcclApi* nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.handle ? &api : nullptr;
    tried = true;
    api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL | RTLD_NOLOAD);      // an NCCL this process already runs
    const char* names[] = {getenv("LELE_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        if (api.handle) break;
        if (!n || !n[0]) continue;
        api.handle = dlopen(n, RTLD_NOW | RTLD_LOCAL);
    }
    if (!api.handle) { lb_set_error("comm: cannot load NCCL (%s); set LELE_B200_NCCL_LIB", dlerror()); return nullptr; }
    bool ok = true;
    auto sym = [&](const char* s) { void* p = dlsym(api.handle, s); if (!p) ok = false; return p; };
    api.get_unique_id = (FnGetUniqueId)sym("ncclGetUniqueId");
    api.comm_init_rank = (FnCommInitRank)sym("ncclCommInitRank");
    api.comm_destroy = (FnCommDestroy)sym("ncclCommDestroy");
    api.broadcast = (FnBroadcast)sym("ncclBroadcast");
    api.send = (FnSend)sym("ncclSend");
    api.recv = (FnRecv)sym("ncclRecv");
    api.group_start = (FnGroup)sym("ncclGroupStart");
    api.group_end = (FnGroup)sym("ncclGroupEnd");
    api.err_str = (FnErrStr)sym("ncclGetErrorString");
    api.get_version = (FnGetVersion)sym("ncclGetVersion");
    if (!ok) { lb_set_error("comm: the loaded NCCL lacks a required symbol"); dlclose(api.handle); api.handle = nullptr; return nullptr; }
    return &api;
}
}  // namespace

struct lele_b200_comm {
    NcclComm comm = nullptr;
    int world = 1, rank = 0, device = 0;
};

#define LB_NCCL(api, expr)                                                                                   \
    do {                                                                                                     \
        int _r = (expr);                                                                                     \
        if (_r != 0) { lb_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, (api)->err_str(_r)); return LELE_B200_ERR_CUDA; } \
    } while (0)

extern "C" int lele_b200_comm_nccl_version(void) {
    NcclApi* api = nccl_api();
    int v = 0;
    if (!api || api->get_version(&v) != 0) return 0;
    return v;
}

extern "C" int lele_b200_comm_unique_id(void* id128_host) {
    LB_REQUIRE(id128_host, "comm_unique_id: NULL argument");
    NcclApi* api = nccl_api();
    if (!api) return LELE_B200_ERR_UNSUPPORTED;
    NcclId id;
    LB_NCCL(api, api->get_unique_id(&id));
    memcpy(id128_host, &id, sizeof(id));
    return LELE_B200_OK;
}

extern "C" int lele_b200_comm_create(lele_b200_ctx* ctx, const void* id128_host, int world, int rank, lele_b200_comm** out) {
    LB_REQUIRE(ctx && id128_host && out, "comm_create: NULL argument");
    LB_ENTER(ctx);
    LB_REQUIRE(world >= 1 && rank >= 0 && rank < world, "comm_create: rank %d outside a world of %d", rank, world);
    NcclApi* api = nccl_api();
    if (!api) return LELE_B200_ERR_UNSUPPORTED;
    NcclId id;
    memcpy(&id, id128_host, sizeof(id));
    lele_b200_comm* c = new lele_b200_comm();
    c->world = world; c->rank = rank; c->device = ctx->device;
    int r = api->comm_init_rank(&c->comm, world, id, rank);
    if (r != 0) { lb_set_error("comm_create: ncclCommInitRank -> %s", api->err_str(r)); delete c; return LELE_B200_ERR_CUDA; }
    *out = c;
    return LELE_B200_OK;
}

extern "C" int lele_b200_comm_destroy(lele_b200_comm* c) {
    if (!c) return LELE_B200_OK;
    NcclApi* api = nccl_api();
    if (api && c->comm) api->comm_destroy(c->comm);
    delete c;
    return LELE_B200_OK;
}

extern "C" int lele_b200_comm_rank(const lele_b200_comm* c) { return c ? c->rank : 0; }
extern "C" int lele_b200_comm_world(const lele_b200_comm* c) { return c ? c->world : 1; }

// in-place broadcast of `nbytes` at dptr from `root` (the weights blob: ~236 MB, once per process), on the context stream
extern "C" int lele_b200_comm_broadcast(lele_b200_ctx* ctx, lele_b200_comm* c, void* dptr, size_t nbytes, int root) {
    LB_REQUIRE(ctx && c && dptr, "comm_broadcast: NULL argument");
    LB_ENTER(ctx);
    LB_REQUIRE(root >= 0 && root < c->world, "comm_broadcast: root %d outside the world of %d", root, c->world);
    if (c->world == 1 || nbytes == 0) return LELE_B200_OK;
    NcclApi* api = nccl_api();
    if (!api) return LELE_B200_ERR_UNSUPPORTED;
    LB_NCCL(api, api->broadcast(dptr, dptr, nbytes, NCCL_UINT8, root, c->comm, ctx->stream));
    return LELE_B200_OK;
}

// gather: every rank contributes `nbytes` from send_dev; on `root`, recv_dev [world][nbytes] receives them in rank order
// (recv_dev is ignored elsewhere).  One grouped send/recv exchange on the context stream -- the output gather of SURVEY 8e:
// ids [clips_per_rank, T'] i32 = 69 KB per rank for the headline workload.
extern "C" int lele_b200_comm_gather(lele_b200_ctx* ctx, lele_b200_comm* c, const void* send_dev, void* recv_dev, size_t nbytes, int root) {
    LB_REQUIRE(ctx && c && send_dev, "comm_gather: NULL argument");
    LB_ENTER(ctx);
    LB_REQUIRE(root >= 0 && root < c->world, "comm_gather: root %d outside the world of %d", root, c->world);
    LB_REQUIRE(c->rank != root || recv_dev, "comm_gather: the root needs a receive buffer");
    if (nbytes == 0) return LELE_B200_OK;
    if (c->world == 1) {
        if (recv_dev != send_dev) LB_CHECK_CUDA(cudaMemcpyAsync(recv_dev, send_dev, nbytes, cudaMemcpyDeviceToDevice, ctx->stream));
        return LELE_B200_OK;
    }
    NcclApi* api = nccl_api();
    if (!api) return LELE_B200_ERR_UNSUPPORTED;
    if (c->rank == root) {
        LB_CHECK_CUDA(cudaMemcpyAsync((char*)recv_dev + (size_t)root * nbytes, send_dev, nbytes, cudaMemcpyDeviceToDevice, ctx->stream));
        LB_NCCL(api, api->group_start());
        for (int r = 0; r < c->world; ++r)
            if (r != root) LB_NCCL(api, api->recv((char*)recv_dev + (size_t)r * nbytes, nbytes, NCCL_UINT8, r, c->comm, ctx->stream));
        LB_NCCL(api, api->group_end());
    } else {
        LB_NCCL(api, api->send(send_dev, nbytes, NCCL_UINT8, root, c->comm, ctx->stream));
    }
    return LELE_B200_OK;
}
