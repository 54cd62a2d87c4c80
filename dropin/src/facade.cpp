// dropin/src/facade.cpp -- libredcufhe.so: the redcufhe:: names of dropin/include/REDcuFHE/redcufhe_gpu.cuh on the C-ABI.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>

#include "REDcuFHE/redcufhe_gpu.cuh"

namespace redcufhe {
namespace {
std::mutex g_mu;
std::map<int, rs_ctx*> g_ctx;   // device -> engine context

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
    for (auto& kv : g_ctx) rs_ctx_destroy(kv.second);
    g_ctx.clear();
}

}  // namespace redcufhe
