// redsec_b200/csrc/comm.cpp -- the exchange step between layers: NCCL all-gather of the ranks' output
// ciphertexts over NVLink, issued on the engine's stream so that a sharded network runs without host synchronisation.
//
// The reference has no exchange step at all: with NUM_GPUS > 1 every OpenMP thread fills only its own loop iterations of its
// own replica enc_segs[idx] and nothing merges the replicas (lib/GPU/Layer.cuh:15,22-26; lib/GPU/BinFunc_gpu.cu:599-621).
// Here every rank computes a block of output channels and the blocks are all-gathered, so each GPU holds the full layer
// output again -- the replicated layout the reference's types assume, made correct.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy torch has already loaded in a torchrun process, the system
// one in a plain C++ process such as the drop-in drivers), so the engine library itself carries no NCCL link dependency and a
// single-GPU user never touches it.  Two ways to form a group:
//   rs_comm_init_rank  one process per GPU (torchrun): rank 0 makes an id with rs_comm_unique_id, the caller distributes it
//   rs_comm_init_all   one process, several GPUs (the reference's model: one host thread per device)
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/redsec_b200.h"

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};

NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) { api.error = std::string("cannot load libnccl.so.2: ") + dlerror(); return; }
        auto sym = [&](const char* name) -> void* {
            void* p = dlsym(api.handle, name);
            if (!p && api.error.empty()) api.error = std::string("libnccl lacks ") + name;
            return p;
        };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(sym("ncclCommInitAll"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    });
    return api;
}

std::string g_comm_error;
int comm_fail(int code, const std::string& msg) { g_comm_error = msg; return code; }

}  // namespace

struct rs_comm {
    rs_ctx* ctx = nullptr;
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
};

static_assert(sizeof(ncclUniqueId) == RS_COMM_ID_BYTES, "RS_COMM_ID_BYTES must match ncclUniqueId");

extern "C" {

const char* rs_comm_last_error(void) { return g_comm_error.c_str(); }

int rs_comm_unique_id(uint8_t* id) {
    if (!id) return comm_fail(RS_ERR_ARG, "rs_comm_unique_id: NULL");
    NcclApi& api = nccl();
    if (!api.error.empty()) return comm_fail(RS_ERR_STATE, api.error);
    ncclUniqueId u;
    ncclResult_t r = api.GetUniqueId(&u);
    if (r != ncclSuccess) return comm_fail(RS_ERR_CUDA, std::string("ncclGetUniqueId: ") + api.GetErrorString(r));
    memcpy(id, &u, sizeof(u));
    return RS_OK;
}

int rs_comm_init_rank(rs_ctx* ctx, rs_comm** out, const uint8_t* id, int rank, int world) {
    if (!ctx || !out || !id || world < 1 || rank < 0 || rank >= world) return comm_fail(RS_ERR_ARG, "rs_comm_init_rank: bad argument");
    *out = nullptr;
    NcclApi& api = nccl();
    if (!api.error.empty()) return comm_fail(RS_ERR_STATE, api.error);
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(rs_ctx_device(ctx));
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    ncclComm_t c = nullptr;
    ncclResult_t r = api.CommInitRank(&c, world, u, rank);
    cudaSetDevice(prev);
    if (r != ncclSuccess) return comm_fail(RS_ERR_CUDA, std::string("ncclCommInitRank: ") + api.GetErrorString(r));
    rs_comm* cm = new rs_comm();
    cm->ctx = ctx; cm->comm = c; cm->rank = rank; cm->world = world;
    rs_ctx_retain(ctx);
    *out = cm;
    return RS_OK;
}

int rs_comm_init_all(rs_ctx** ctxs, int n, rs_comm** comms_out) {
    if (!ctxs || !comms_out || n < 1) return comm_fail(RS_ERR_ARG, "rs_comm_init_all: bad argument");
    NcclApi& api = nccl();
    if (!api.error.empty()) return comm_fail(RS_ERR_STATE, api.error);
    std::vector<int> devs(n);
    for (int i = 0; i < n; i++) {
        if (!ctxs[i]) return comm_fail(RS_ERR_ARG, "rs_comm_init_all: NULL context");
        devs[i] = rs_ctx_device(ctxs[i]);
    }
    std::vector<ncclComm_t> comms(n, nullptr);
    ncclResult_t r = api.CommInitAll(comms.data(), n, devs.data());
    if (r != ncclSuccess) return comm_fail(RS_ERR_CUDA, std::string("ncclCommInitAll: ") + api.GetErrorString(r));
    for (int i = 0; i < n; i++) {
        rs_comm* cm = new rs_comm();
        cm->ctx = ctxs[i]; cm->comm = comms[i]; cm->rank = i; cm->world = n;
        rs_ctx_retain(ctxs[i]);
        comms_out[i] = cm;
    }
    return RS_OK;
}

int rs_comm_destroy(rs_comm* c) {
    if (!c) return RS_OK;
    if (c->comm) {
        rs_sync(c->ctx);
        nccl().CommDestroy(c->comm);
    }
    rs_ctx_release(c->ctx);
    delete c;
    return RS_OK;
}

int rs_comm_rank(const rs_comm* c) { return c ? c->rank : 0; }
int rs_comm_world(const rs_comm* c) { return c ? c->world : 1; }
rs_ctx* rs_comm_ctx(const rs_comm* c) { return c ? c->ctx : nullptr; }

// out[world][rows_per_rank] <- every rank's in[rows_per_rank] (rows of RS_LWE_STRIDE words), ordered on the context's
// current lane; no host synchronisation
int rs_allgather(rs_comm* c, uint32_t* out_dev, const uint32_t* in_dev, size_t rows_per_rank) {
    if (!c || !out_dev || !in_dev) return comm_fail(RS_ERR_ARG, "rs_allgather: NULL argument");
    if (rows_per_rank == 0) return RS_OK;
    NcclApi& api = nccl();
    void* stream = nullptr;
    if (rs_get_stream(c->ctx, &stream) != RS_OK) return comm_fail(RS_ERR_STATE, "rs_allgather: no stream");
    int prev = 0;
    cudaGetDevice(&prev);
    const int dev = rs_ctx_device(c->ctx);
    if (prev != dev) cudaSetDevice(dev);
    ncclResult_t r = api.AllGather(in_dev, out_dev, rows_per_rank * RS_LWE_STRIDE, ncclUint32, c->comm, (cudaStream_t)stream);
    if (prev != dev) cudaSetDevice(prev);
    if (r != ncclSuccess) return comm_fail(RS_ERR_CUDA, std::string("ncclAllGather: ") + api.GetErrorString(r));
    return RS_OK;
}

}  // extern "C"
