// dropin/src/facade.cpp -- libredcufhe.so: the redcufhe:: names of dropin/include/REDcuFHE/redcufhe_gpu.cuh on the C-ABI.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <vector>
#include <cstring>

#include "REDcuFHE/redcufhe_gpu.cuh"

namespace redcufhe {
namespace {
std::mutex g_mu;
std::map<int, rs_ctx*> g_ctx;   // device -> engine context
std::vector<rs_comm*> g_comm;   // single-process NCCL group over devices 0..n-1 (NUM_GPUS > 1)

[[noreturn]] void die(const char* what, rs_ctx* ctx) {
    fprintf(stderr, "redcufhe facade: %s: %s\n", what, rs_last_error(ctx));
    exit(1);
}
}  // namespace

PubKey::~PubKey() { free(bsk); free(ksk); }

void ReadPubKeyFromFile(PubKey& key, const char* path) {
    key.bsk = static_cast<uint32_t*>(malloc(RS_BSK_WORDS * sizeof(uint32_t)));
    key.ksk = static_cast<uint32_t*>(malloc(RS_KSK_WORDS * sizeof(uint32_t)));
    if (!key.bsk || !key.ksk || rs_read_eval_key(path, key.bsk, key.ksk) != RS_OK) {
        fprintf(stderr, "redcufhe facade: cannot read evaluation key %s\n", path);
        exit(1);
    }
}

void Initialize(PubKey& key) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(g_mu);
    if (g_ctx.count(dev)) return;
    rs_ctx* ctx = nullptr;
    if (rs_ctx_create(&ctx, dev) != RS_OK) die("rs_ctx_create", nullptr);
    if (rs_load_eval_key(ctx, key.bsk, key.ksk) != RS_OK) die("rs_load_eval_key", ctx);
    g_ctx[dev] = ctx;
}

rs_ctx* CurrentContext() {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_ctx.find(dev);
    if (it == g_ctx.end()) {
        fprintf(stderr, "redcufhe facade: Initialize(PubKey&) has not been called on device %d\n", dev);
        exit(1);
    }
    return it->second;
}

rs_ctx* ContextOf(int device) {
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_ctx.find(device);
    return it == g_ctx.end() ? nullptr : it->second;
}

rs_comm* CommunicatorOf(int device, int n) {
    if (n <= 1) return nullptr;
    std::lock_guard<std::mutex> lock(g_mu);
    if (g_comm.empty()) {
        std::vector<rs_ctx*> ctxs(n, nullptr);
        for (int d = 0; d < n; d++) {
            auto it = g_ctx.find(d);
            if (it == g_ctx.end()) {
                fprintf(stderr, "redcufhe facade: NUM_GPUS = %d but Initialize(PubKey&) was not called on device %d\n", n, d);
                exit(1);
            }
            ctxs[d] = it->second;
        }
        g_comm.assign(n, nullptr);
        if (rs_comm_init_all(ctxs.data(), n, g_comm.data()) != RS_OK) {
            fprintf(stderr, "redcufhe facade: rs_comm_init_all: %s\n", rs_comm_last_error());
            exit(1);
        }
    }
    return (device >= 0 && device < (int)g_comm.size()) ? g_comm[device] : nullptr;
}

void Copy(Ctxt& out, const Ctxt& in, Stream) { memcpy(out.lwe, in.lwe, sizeof(out.lwe)); out.variance = in.variance; }

void ReadCtxtFromFileRed(Ctxt& ct, std::ifstream& in) {
    in.read(reinterpret_cast<char*>(ct.lwe), sizeof(ct.lwe));
    in.read(reinterpret_cast<char*>(&ct.variance), sizeof(ct.variance));
    if (!in) {
        fprintf(stderr, "redcufhe facade: short read on the ciphertext file\n");
        exit(1);
    }
}

void WriteCtxtToFileRed(Ctxt& ct, const char* path) {
    if (rs_write_ctxt(path, ct.lwe, 1, ct.variance, /*append=*/1) != RS_OK) {
        fprintf(stderr, "redcufhe facade: cannot append to %s\n", path);
        exit(1);
    }
}

void Synchronize() {
    std::lock_guard<std::mutex> lock(g_mu);
    for (auto& kv : g_ctx)
        if (rs_sync(kv.second) != RS_OK) die("rs_sync", kv.second);
}

void CuCheckError() {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        fprintf(stderr, "CUDA error: %s\n", cudaGetErrorString(e));
        exit(1);
    }
}

void CleanUp() {
    std::lock_guard<std::mutex> lock(g_mu);
    for (rs_comm* c : g_comm) rs_comm_destroy(c);
    g_comm.clear();
    // the generated drivers never delete their layers (main.cu exits right after CleanUp), so a context may still be
    // referenced: rs_ctx_destroy then refuses and the process exit reclaims it
    for (auto& kv : g_ctx) { rs_sync(kv.second); rs_ctx_destroy(kv.second); }
    g_ctx.clear();
}

}  // namespace redcufhe
