// dropin/src/layers_shim.cpp -- IntLayer / BinLayer with the reference's GPU signatures (lib/GPU/IntLayer.cuh:16-34,
// BinLayer.cuh:16-34), forwarding to redsec::Layer (one batched bootstrap per layer).  With NUM_GPUS > 1 there is one engine
// layer per device, one host thread per device (as lib/GPU/BinFunc_gpu.cu:116-138), and each layer is neuron-sharded:
// every GPU computes its block of output channels and the blocks are all-gathered over NCCL on the engine streams
// (redsec::Layer::execute_sharded), so afterwards every GPU holds the full layer output -- the reference's replicated
// enc_segs[NUM_GPUS] layout.
#include "BinLayer.cuh"
#include "IntLayer.cuh"
#include "shim_common.hpp"

namespace {

tDimensions* prep_all(redsec::Layer** impl, FILE* fd, tDimensions* dim, tDimensions* in_dim, tDimensions* out_dim) {
    const long pos = ftell(fd);
    tDimensions* r = nullptr;
    for (int g = 0; g < NUM_GPUS; g++) {          // every device's layer reads the same blocks of the weights file
        fseek(fd, pos, SEEK_SET);
        tDimensions work = *dim;                  // the engine's prep updates its argument in place; callers chain our own out_dim
        r = impl[g]->prep(fd, &work);
        if (!r) return nullptr;
    }
    *in_dim = impl[0]->in_dim;
    *out_dim = impl[0]->out_dim;
    return out_dim;
}

template <class In>
tBitPacked* execute_all(redsec::Layer** impl, In* p_in, size_t in_count, bool is_network_output) {
    redsec::Batch out[NUM_GPUS];
    shim::for_each_gpu([&](int g) {
        redsec::Batch in = shim::device_input(p_in, in_count, g);
        out[g] = impl[g]->execute_sharded(in, shim::comm_of(g));
        if (!out[g].dev) shim::check(RS_ERR_STATE, "layer forward", shim::ctx_of(g));
    });
    shim::free_host(p_in);
    delete p_in;
    tMultiBitPacked* r = shim::new_packed<tMultiBitPacked>((uint32_t)out[0].count);
    for (int g = 0; g < NUM_GPUS; g++) r->dev[g] = out[g];
    if (is_network_output) shim::download_to_host(r);
    return reinterpret_cast<tBitPacked*>(r);      // same layout (net.cu:118 casts the last layer's result back)
}

}  // namespace

// ---------------------------------------------------------------------------------------------- IntLayer
IntLayer::IntLayer(eConvType ec, uint16_t dep, ePoolType ep, eQuantType eq, tNetParams* np) : quant_(eq) {
    shim::ensure_group();
    for (int g = 0; g < NUM_GPUS; g++) impl_[g] = new redsec::Layer(shim::ctx_of(g), /*int_inputs=*/true, ec, dep, ep, eq, np);
    memset(&in_dim, 0, sizeof(in_dim));
    memset(&out_dim, 0, sizeof(out_dim));
}
IntLayer::~IntLayer() { for (int g = 0; g < NUM_GPUS; g++) delete impl_[g]; }
tDimensions* IntLayer::prep(FILE* fd, tDimensions* dim) { return prep_all(impl_, fd, dim, &in_dim, &out_dim); }
tBitPacked* IntLayer::execute(tMultiBitPacked* p_in) { return execute_all(impl_, p_in, shim::dim_count(in_dim), quant_ == E_ACTIVATION_NONE); }
void IntLayer::export_weights(FILE*) { printf("Weight Convert not defined\r\n"); }
void IntLayer::set_print_layer(uint8_t) {}

// ---------------------------------------------------------------------------------------------- BinLayer
BinLayer::BinLayer(eConvType ec, uint16_t dep, ePoolType ep, eQuantType eq, tNetParams* np) : quant_(eq) {
    shim::ensure_group();
    for (int g = 0; g < NUM_GPUS; g++) impl_[g] = new redsec::Layer(shim::ctx_of(g), /*int_inputs=*/false, ec, dep, ep, eq, np);
    memset(&in_dim, 0, sizeof(in_dim));
    memset(&out_dim, 0, sizeof(out_dim));
}
BinLayer::~BinLayer() { for (int g = 0; g < NUM_GPUS; g++) delete impl_[g]; }
tDimensions* BinLayer::prep(FILE* fd, tDimensions* dim) { return prep_all(impl_, fd, dim, &in_dim, &out_dim); }
tBitPacked* BinLayer::execute(tBitPacked* p_in) { return execute_all(impl_, p_in, shim::dim_count(in_dim), quant_ == E_ACTIVATION_NONE); }
void BinLayer::export_weights(FILE*) { printf("Weight Convert not defined\r\n"); }
void BinLayer::set_print_layer(uint8_t) {}
