// dropin/src/layers_shim.cpp -- the object the reference Makefile picks up as $(LIB_DIR)/GPU/*.o: IntLayer / BinLayer /
// mbit_calloc_global with the reference's GPU signatures, forwarding to redsec::Layer (one batched bootstrap per layer).
#include <cstdlib>
#include <cstring>
#include <vector>

#include "BinLayer.cuh"
#include "IntLayer.cuh"

namespace {

size_t dim_count(const tDimensions& d) { return (size_t)d.hw.h * d.hw.w * d.in_dep; }

void check(int rc, const char* what) {
    if (rc != RS_OK) {
        fprintf(stderr, "redsec drop-in: %s failed: %s\n", what, rs_last_error(redcufhe::CurrentContext()));
        exit(1);
    }
}

// network input: host Ctxt objects (filled by main.cu) -> one device batch
redsec::Batch upload_multibit(const tMultiBit* arr, size_t count) {
    rs_ctx* ctx = redcufhe::CurrentContext();
    std::vector<uint32_t> wire(count * RS_LWE_WORDS);
    for (size_t i = 0; i < count; i++) memcpy(&wire[i * RS_LWE_WORDS], arr[i].ctxt[0].lwe, sizeof(uint32_t) * RS_LWE_WORDS);
    redsec::Batch b;
    b.count = count;
    check(rs_lwe_alloc(ctx, count, &b.dev), "rs_lwe_alloc");
    check(rs_lwe_upload(ctx, b.dev, wire.data(), count), "rs_lwe_upload");
    return b;
}
redsec::Batch upload_bits(const tBit* arr, size_t count) {
    rs_ctx* ctx = redcufhe::CurrentContext();
    std::vector<uint32_t> wire(count * RS_LWE_WORDS);
    for (size_t i = 0; i < count; i++) memcpy(&wire[i * RS_LWE_WORDS], arr[i].lwe, sizeof(uint32_t) * RS_LWE_WORDS);
    redsec::Batch b;
    b.count = count;
    check(rs_lwe_alloc(ctx, count, &b.dev), "rs_lwe_alloc");
    check(rs_lwe_upload(ctx, b.dev, wire.data(), count), "rs_lwe_upload");
    return b;
}

void free_multibit_host(tMultiBitPacked* p, size_t count) {
    if (p->enc_segs[0]) {
        for (size_t i = 0; i < count; i++) delete[] p->enc_segs[0][i].ctxt;
        delete[] p->enc_segs[0];
    }
}

// result of a layer: device batch; for a layer without activation (the network output, net.cu:118) also the host view
// main.cu:82 reads, laid out as tMultiBitPacked
tBitPacked* wrap_output(redsec::Batch out, bool is_network_output) {
    if (!is_network_output) {
        tBitPacked* r = new tBitPacked();
        r->enc_segs[0] = nullptr;
        r->size = (uint8_t)out.count;      // the reference stores uint8_t(len) too (lib/GPU/Layer.cu:47)
        r->dev = out;
        return r;
    }
    rs_ctx* ctx = redcufhe::CurrentContext();
    std::vector<uint32_t> wire(out.count * RS_LWE_WORDS);
    check(rs_lwe_download(ctx, wire.data(), out.dev, out.count), "rs_lwe_download");
    tMultiBitPacked* r = nullptr;
    mbit_calloc_global(&r, (uint32_t)out.count, 1);
    for (size_t i = 0; i < out.count; i++) memcpy(r->enc_segs[0][i].ctxt[0].lwe, &wire[i * RS_LWE_WORDS], sizeof(uint32_t) * RS_LWE_WORDS);
    r->dev = out;
    return reinterpret_cast<tBitPacked*>(r);
}

}  // namespace

void mbit_calloc_global(tMultiBitPacked** ret, uint32_t len, uint8_t bits) {
    tMultiBitPacked* p = new tMultiBitPacked();
    p->enc_segs[0] = new tMultiBit[len];
    for (uint32_t i = 0; i < len; i++) {
        p->enc_segs[0][i].ctxt = new tBit[bits]();
        p->enc_segs[0][i].size = bits;
        p->enc_segs[0][i].gpu_id = 0;
    }
    p->size = (uint8_t)len;
    *ret = p;
}
void bit_calloc_global(tBitPacked** ret, uint32_t len) {
    tBitPacked* p = new tBitPacked();
    p->enc_segs[0] = new tBit[len]();
    p->size = (uint8_t)len;
    *ret = p;
}
void print_status(const char* msg) { fputs(msg, stdout); fflush(stdout); }

// ---------------------------------------------------------------------------------------------- IntLayer
IntLayer::IntLayer(eConvType ec, uint16_t dep, ePoolType ep, eQuantType eq, tNetParams* np)
    : impl_(new redsec::Layer(redcufhe::CurrentContext(), /*int_inputs=*/true, ec, dep, ep, eq, np)), quant_(eq) {
    memset(&in_dim, 0, sizeof(in_dim));
    memset(&out_dim, 0, sizeof(out_dim));
}
IntLayer::~IntLayer() { delete impl_; }
tDimensions* IntLayer::prep(FILE* fd, tDimensions* dim) {
    tDimensions work = *dim;            // the engine's prep updates its argument in place; callers chain our own out_dim
    tDimensions* r = impl_->prep(fd, &work);
    in_dim = impl_->in_dim;
    out_dim = impl_->out_dim;
    return r ? &out_dim : nullptr;
}
tBitPacked* IntLayer::execute(tMultiBitPacked* p_in) {
    const size_t count = dim_count(in_dim);
    redsec::Batch in = p_in->dev.dev ? p_in->dev : upload_multibit(p_in->enc_segs[0], count);
    free_multibit_host(p_in, count);
    delete p_in;
    return wrap_output(impl_->execute(in), quant_ == E_ACTIVATION_NONE);
}
void IntLayer::export_weights(FILE*) { printf("Weight Convert not defined\r\n"); }
void IntLayer::set_print_layer(uint8_t) {}

// ---------------------------------------------------------------------------------------------- BinLayer
BinLayer::BinLayer(eConvType ec, uint16_t dep, ePoolType ep, eQuantType eq, tNetParams* np)
    : impl_(new redsec::Layer(redcufhe::CurrentContext(), /*int_inputs=*/false, ec, dep, ep, eq, np)), quant_(eq) {
    memset(&in_dim, 0, sizeof(in_dim));
    memset(&out_dim, 0, sizeof(out_dim));
}
BinLayer::~BinLayer() { delete impl_; }
tDimensions* BinLayer::prep(FILE* fd, tDimensions* dim) {
    tDimensions work = *dim;            // the engine's prep updates its argument in place; callers chain our own out_dim
    tDimensions* r = impl_->prep(fd, &work);
    in_dim = impl_->in_dim;
    out_dim = impl_->out_dim;
    return r ? &out_dim : nullptr;
}
tBitPacked* BinLayer::execute(tBitPacked* p_in) {
    const size_t count = dim_count(in_dim);
    redsec::Batch in = p_in->dev.dev ? p_in->dev : upload_bits(p_in->enc_segs[0], count);
    delete[] p_in->enc_segs[0];
    delete p_in;
    return wrap_output(impl_->execute(in), quant_ == E_ACTIVATION_NONE);
}
void BinLayer::export_weights(FILE*) { printf("Weight Convert not defined\r\n"); }
void BinLayer::set_print_layer(uint8_t) {}
