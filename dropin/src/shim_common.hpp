// dropin/src/shim_common.hpp -- helpers shared by the drop-in shims (not part of the reference's surface).
#pragma once
#include <omp.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "Layer.cuh"

namespace shim {

inline void check(int rc, const char* what, rs_ctx* ctx) {
    if (rc != RS_OK) {
        fprintf(stderr, "redsec drop-in: %s failed: %s\n", what, ctx ? rs_last_error(ctx) : rs_comm_last_error());
        exit(1);
    }
}
inline size_t dim_count(const tDimensions& d) { return (size_t)d.hw.h * d.hw.w * d.in_dep; }
inline rs_ctx* ctx_of(int g) {
    rs_ctx* c = redcufhe::ContextOf(g);
    if (!c) { fprintf(stderr, "redsec drop-in: Initialize(PubKey&) has not been called on device %d\n", g); exit(1); }
    return c;
}
inline rs_comm* comm_of(int g) { return redcufhe::CommunicatorOf(g, NUM_GPUS); }
// forms the NCCL group now (all contexts exist once net.cu's Initialize loop has run): layer / stage constructors call this so
// that the group is never created in the middle of an inference with kernels in flight
inline void ensure_group() { if (NUM_GPUS > 1) (void)redcufhe::CommunicatorOf(0, NUM_GPUS); }

// one host thread per GPU, as the reference's `omp_set_num_threads(NUM_GPUS); #pragma omp parallel for` (lib/GPU/BinFunc_gpu.cu:116-138)
template <class F>
inline void for_each_gpu(F&& body) {
    if (NUM_GPUS == 1) { body(0); return; }
#pragma omp parallel for num_threads(NUM_GPUS) schedule(static, 1)
    for (int g = 0; g < NUM_GPUS; g++) {
        cudaSetDevice(g);
        body(g);
    }
}

// host Ctxt arrays (filled by main.cu) -> one device batch on GPU g
inline redsec::Batch upload_host(const tMultiBit* arr, size_t count, int g) {
    rs_ctx* ctx = ctx_of(g);
    std::vector<uint32_t> wire(count * RS_LWE_WORDS);
    for (size_t i = 0; i < count; i++) memcpy(&wire[i * RS_LWE_WORDS], arr[i].ctxt[0].lwe, sizeof(uint32_t) * RS_LWE_WORDS);
    redsec::Batch b;
    b.count = count;
    check(rs_lwe_alloc(ctx, count, &b.dev), "rs_lwe_alloc", ctx);
    check(rs_lwe_upload(ctx, b.dev, wire.data(), count), "rs_lwe_upload", ctx);
    return b;
}
inline redsec::Batch upload_host(const tBit* arr, size_t count, int g) {
    rs_ctx* ctx = ctx_of(g);
    std::vector<uint32_t> wire(count * RS_LWE_WORDS);
    for (size_t i = 0; i < count; i++) memcpy(&wire[i * RS_LWE_WORDS], arr[i].lwe, sizeof(uint32_t) * RS_LWE_WORDS);
    redsec::Batch b;
    b.count = count;
    check(rs_lwe_alloc(ctx, count, &b.dev), "rs_lwe_alloc", ctx);
    check(rs_lwe_upload(ctx, b.dev, wire.data(), count), "rs_lwe_upload", ctx);
    return b;
}

template <class Packed>
inline Packed* new_packed(uint32_t len) {
    Packed* p = new Packed();
    for (int g = 0; g < NUM_GPUS; g++) { p->enc_segs[g] = nullptr; p->dev[g] = redsec::Batch(); }
    p->size = (uint8_t)(len > 255 ? 255 : len);
    p->len = len;
    p->pending_sign = false;
    p->shard_c0 = 0; p->shard_cl = 0;
    return p;
}
inline void free_host(tMultiBitPacked* p) {
    for (int g = 0; g < NUM_GPUS; g++) {
        if (!p->enc_segs[g]) continue;
        for (uint32_t i = 0; i < p->len; i++) delete[] p->enc_segs[g][i].ctxt;
        delete[] p->enc_segs[g];
        p->enc_segs[g] = nullptr;
    }
}
inline void free_host(tBitPacked* p) {
    for (int g = 0; g < NUM_GPUS; g++) { delete[] p->enc_segs[g]; p->enc_segs[g] = nullptr; }
}

// device batch of GPU g for a packed array: what a previous stage left there, or the host ciphertexts uploaded now
template <class Packed>
inline redsec::Batch device_input(Packed* p, size_t count, int g) {
    if (p->dev[g].dev) return p->dev[g];
    if (!p->enc_segs[g]) { fprintf(stderr, "redsec drop-in: input array has neither device nor host ciphertexts for GPU %d\n", g); exit(1); }
    return upload_host(p->enc_segs[g], count, g);
}

// the network output (a layer without activation, nets/*/net.cu:118): also materialise the host view main.cu:82 reads
inline void download_to_host(tMultiBitPacked* r) {
    for (int g = 0; g < NUM_GPUS; g++) {
        if (!r->dev[g].dev) continue;
        rs_ctx* ctx = ctx_of(g);
        std::vector<uint32_t> wire(r->dev[g].count * RS_LWE_WORDS);
        check(rs_lwe_download(ctx, wire.data(), r->dev[g].dev, r->dev[g].count), "rs_lwe_download", ctx);
        if (!r->enc_segs[g]) {
            r->enc_segs[g] = new tMultiBit[r->len];
            for (uint32_t i = 0; i < r->len; i++) { r->enc_segs[g][i].ctxt = new tBit[1](); r->enc_segs[g][i].size = 1; r->enc_segs[g][i].gpu_id = (uint8_t)g; }
        }
        for (size_t i = 0; i < r->dev[g].count && i < r->len; i++)
            memcpy(r->enc_segs[g][i].ctxt[0].lwe, &wire[i * RS_LWE_WORDS], sizeof(uint32_t) * RS_LWE_WORDS);
    }
}

}  // namespace shim
