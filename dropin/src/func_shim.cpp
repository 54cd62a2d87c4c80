// dropin/src/func_shim.cpp -- {Bin,Int}Func::{Convolution,SumPooling,MaxPooling,Quantize} of lib/GPU/BinFunc_gpu.cuh:16-147 and
// lib/GPU/IntFunc_gpu.cuh:8-124 over the engine's batched stages.  Ownership follows the reference: every execute() deletes
// its input array and returns a new one (lib/GPU/BinFunc_gpu.cu:219-222,626-628).
//
// Two things differ from a per-ciphertext implementation and are invisible to a caller that composes the stages in the
// reference's order (conv -> [sum-pool] -> quantize -> [max-pool], lib/GPU/BinLayer.cu:114-203):
//  * Quantize::execute adds the bias and returns the PRE-activations with `pending_sign` set; the ONE batched sign bootstrap is
//    issued by whoever consumes the bits, with the encoding that consumer needs: MaxPooling::execute bootstraps at 1/8, runs
//    the OR tree and emits 1/4096 (SURVEY H2 / defect R3); anything else (the next Convolution, the final read-out) at 1/4096.
//  * With NUM_GPUS > 1 a Convolution computes on GPU g only the g-th block of output channels; pooling / quantize / max-pool
//    run on that block (they are per channel), and the next consumer that needs all channels all-gathers the blocks over NCCL
//    (redsec::gather_channels).  Layers without a convolution run whole on every GPU.
#include "BinFunc_gpu.cuh"
#include "IntFunc_gpu.cuh"
#include "shim_common.hpp"

namespace {

constexpr uint32_t kUnit = 1u << 20, kEighth = 1u << 29;

template <class Stage, class... A>
void make_stages(Stage** st, A... args) {
    shim::ensure_group();
    for (int g = 0; g < NUM_GPUS; g++) st[g] = new Stage(shim::ctx_of(g), args...);
}
template <class Stage>
void drop_stages(Stage** st) { for (int g = 0; g < NUM_GPUS; g++) delete st[g]; }

// run `prep` on every device's stage with the same file position and the same input dimensions
template <class F>
tDimensions* prep_all(FILE* fd, tDimensions* dim, F&& prep_one) {
    const long pos = fd ? ftell(fd) : 0;
    const tDimensions in = *dim;
    tDimensions out = in;
    for (int g = 0; g < NUM_GPUS; g++) {
        if (fd) fseek(fd, pos, SEEK_SET);
        out = in;
        if (!prep_one(g, &out)) return nullptr;
    }
    *dim = out;
    return dim;
}

// resolve what a consumer needs from an array produced by an earlier stage: the pending sign bootstrap (at 1/4096) and, when
// all channels are needed, the gather of the per-GPU channel blocks
template <class Packed>
void materialize(Packed* p, bool need_all_channels) {
    shim::for_each_gpu([&](int g) {
        if (!p->dev[g].dev) return;
        rs_ctx* ctx = shim::ctx_of(g);
        if (p->pending_sign) shim::check(redsec::QuantizeStage::sign_bootstrap(ctx, p->dev[g], kUnit), "sign bootstrap", ctx);
        if (need_all_channels && p->shard_cl > 0) {
            if (getenv("RS_SHIM_DEBUG")) {      // checksum of this GPU's channel block before the all-gather
                std::vector<uint32_t> wire(p->dev[g].count * RS_LWE_WORDS);
                rs_lwe_download(ctx, wire.data(), p->dev[g].dev, p->dev[g].count);
                uint64_t sum = 0;
                for (size_t i = 0; i < wire.size(); i++) sum = sum * 1000003u + wire[i];
                fprintf(stderr, "shim: block of GPU %d before the all-gather: %zu rows, checksum %016llx\n", g, p->dev[g].count, (unsigned long long)sum);
            }
            p->dev[g] = redsec::gather_channels(ctx, shim::comm_of(g), p->dev[g], p->shard_cl);
            if (!p->dev[g].dev) shim::check(RS_ERR_STATE, "gather_channels", nullptr);
        }
    });
    p->pending_sign = false;
    if (need_all_channels) p->shard_cl = 0;
    if (need_all_channels && getenv("RS_SHIM_DEBUG"))       // checksum of the full activation array on every GPU
        for (int g = 0; g < NUM_GPUS; g++) {
            if (!p->dev[g].dev) continue;
            std::vector<uint32_t> wire(p->dev[g].count * RS_LWE_WORDS);
            rs_lwe_download(shim::ctx_of(g), wire.data(), p->dev[g].dev, p->dev[g].count);
            uint64_t sum = 0;
            for (size_t i = 0; i < wire.size(); i++) sum = sum * 1000003u + wire[i];
            fprintf(stderr, "shim: full array on GPU %d: %zu rows, checksum %016llx\n", g, p->dev[g].count, (unsigned long long)sum);
        }
    if (p->dev[0].dev) { p->len = (uint32_t)p->dev[0].count; p->size = (uint8_t)(p->len > 255 ? 255 : p->len); }
}

template <class Out, class In, class F>
Out* run_stage(In* in, size_t host_count, F&& body) {
    redsec::Batch out[NUM_GPUS];
    shim::for_each_gpu([&](int g) {
        redsec::Batch b = shim::device_input(in, host_count, g);
        out[g] = body(g, b);
        if (!out[g].dev) shim::check(RS_ERR_STATE, "Func stage", shim::ctx_of(g));
    });
    Out* r = shim::new_packed<Out>((uint32_t)out[0].count);
    r->shard_c0 = in->shard_c0; r->shard_cl = in->shard_cl; r->pending_sign = in->pending_sign;
    for (int g = 0; g < NUM_GPUS; g++) r->dev[g] = out[g];
    shim::free_host(in);
    delete in;
    return r;
}

}  // namespace

// ---------------------------------------------------------------------------------------------- BinFunc
BinFunc::Convolution::Convolution(uint32_t out_depth, tConvParams* p) { make_stages(st_, false, out_depth, *p); memset(&out_dim_, 0, sizeof(out_dim_)); }
BinFunc::Convolution::~Convolution() { drop_stages(st_); }
tDimensions* BinFunc::Convolution::prep(FILE* fd, tDimensions* dim) {
    tDimensions* r = prep_all(fd, dim, [&](int g, tDimensions* d) { return st_[g]->prep(fd, d) != nullptr; });
    if (r) out_dim_ = *r;
    return r;
}
void BinFunc::Convolution::get_outhw(tRectangle* ret_dim) { *ret_dim = out_dim_.hw; }
void BinFunc::Convolution::get_outdep(uint32_t* ret_dep) { *ret_dep = out_dim_.in_dep; }
template <class Out, class In>
static Out* conv_execute(redsec::ConvStage** st, In* in) {
    materialize(in, /*need_all_channels=*/true);
    const int depth = st[0]->out_depth();
    const bool shard = NUM_GPUS > 1 && depth % NUM_GPUS == 0;
    const int cl = shard ? depth / NUM_GPUS : depth;
    Out* r = run_stage<Out>(in, in->len, [&](int g, redsec::Batch b) { return shard ? st[g]->execute(b, g * cl, (g + 1) * cl) : st[g]->execute(b); });
    r->shard_c0 = 0; r->shard_cl = shard ? cl : 0;
    return r;
}
tMultiBitPacked* BinFunc::Convolution::execute(tBitPacked* in) { return conv_execute<tMultiBitPacked>(st_, in); }

BinFunc::SumPooling::SumPooling(tPoolParams* p) { make_stages(st_, *p); }
BinFunc::SumPooling::~SumPooling() { drop_stages(st_); }
tDimensions* BinFunc::SumPooling::prep(tDimensions* dim) { return prep_all(nullptr, dim, [&](int g, tDimensions* d) { return st_[g]->prep(d) != nullptr; }); }
tMultiBitPacked* BinFunc::SumPooling::execute(tMultiBitPacked* in) {
    materialize(in, false);
    return run_stage<tMultiBitPacked>(in, in->len, [&](int g, redsec::Batch b) { return st_[g]->execute(b); });
}

BinFunc::MaxPooling::MaxPooling(tPoolParams* p) { make_stages(st_, *p); }
BinFunc::MaxPooling::~MaxPooling() { drop_stages(st_); }
tDimensions* BinFunc::MaxPooling::prep(tDimensions* dim) { return prep_all(nullptr, dim, [&](int g, tDimensions* d) { return st_[g]->prep(d) != nullptr; }); }
tBitPacked* BinFunc::MaxPooling::execute(tBitPacked* in) {
    if (!in->pending_sign) {
        // a bit at +-1/4096 cannot be re-bootstrapped on its own (the rounding to 2N adds sigma ~ 7.7/4096), and an OR
        // gate needs +-1/8 inputs: the pool must receive the sign stage's deferred output
        fprintf(stderr, "redsec drop-in: MaxPooling::execute needs the array returned by Quantize::execute (lib/GPU/BinLayer.cu:176-190)\n");
        exit(1);
    }
    in->pending_sign = false;      // consumed here: sign bootstrap at 1/8 + OR tree + final level at 1/4096
    return run_stage<tBitPacked>(in, in->len, [&](int g, redsec::Batch b) { return st_[g]->execute_from_preact(b); });
}

BinFunc::Quantize::Quantize(tQParams* q) { make_stages(st_, false, *q); }
BinFunc::Quantize::~Quantize() { drop_stages(st_); }
tDimensions* BinFunc::Quantize::prep(FILE* fd, tDimensions* dim, tMultiBitPacked*, uint16_t* p_slope) {
    return prep_all(fd, dim, [&](int g, tDimensions* d) { return st_[g]->prep(fd, d, p_slope != nullptr) != nullptr; });
}
template <class Out, class In>
static Out* quant_bias(redsec::QuantizeStage** st, In* in, bool defer_sign) {
    materialize(in, false);
    Out* r = run_stage<Out>(in, in->len, [&](int g, redsec::Batch b) { return st[g]->add_bias(b, in->shard_cl > 0 ? g * in->shard_cl : 0); });
    r->pending_sign = defer_sign;
    return r;
}
tBitPacked* BinFunc::Quantize::execute(tMultiBitPacked* in, tMultiBitPacked*) { return quant_bias<tBitPacked>(st_, in, true); }
tMultiBitPacked* BinFunc::Quantize::add_bias(tMultiBitPacked* in, tMultiBitPacked*) {
    tMultiBitPacked* r = quant_bias<tMultiBitPacked>(st_, in, false);
    materialize(r, true);              // an activation-free layer is a network output: all channels, on the host as well
    shim::download_to_host(r);
    return r;
}
tFixedPointPacked* BinFunc::Quantize::relu_shift(tMultiBitPacked*, tMultiBitPacked*, uint16_t*) {
    fprintf(stderr, "redsec drop-in: BinFunc::Quantize::relu_shift is not provided (no shipped net uses it; DESIGN.md 9)\n");
    exit(1);
}

// ---------------------------------------------------------------------------------------------- IntFunc
IntFunc::Convolution::Convolution(uint16_t out_depth, tConvParams* p) { make_stages(st_, true, (uint32_t)out_depth, *p); }
IntFunc::Convolution::~Convolution() { drop_stages(st_); }
tDimensions* IntFunc::Convolution::prep(FILE* fd, tDimensions* dim) { return prep_all(fd, dim, [&](int g, tDimensions* d) { return st_[g]->prep(fd, d) != nullptr; }); }
tFixedPointPacked* IntFunc::Convolution::execute(tFixedPointPacked* in) { return conv_execute<tFixedPointPacked>(st_, in); }

IntFunc::SumPooling::SumPooling(tPoolParams* p) { make_stages(st_, *p); }
IntFunc::SumPooling::~SumPooling() { drop_stages(st_); }
tDimensions* IntFunc::SumPooling::prep(tDimensions* dim) { return prep_all(nullptr, dim, [&](int g, tDimensions* d) { return st_[g]->prep(d) != nullptr; }); }
tFixedPointPacked* IntFunc::SumPooling::execute(tFixedPointPacked* in) {
    materialize(in, false);
    return run_stage<tFixedPointPacked>(in, in->len, [&](int g, redsec::Batch b) { return st_[g]->execute(b); });
}

IntFunc::Quantize::Quantize(tQParams* q) : relu_(q->shift_bits >= 2 && q->shift_bits <= 8) { make_stages(st_, true, *q); }
IntFunc::Quantize::~Quantize() { drop_stages(st_); }
tDimensions* IntFunc::Quantize::prep(FILE* fd, tDimensions* dim, tMultiBitPacked**, uint16_t* p_slope) {
    const bool slope = relu_ && p_slope != nullptr;     // lib/IntFunc.cpp:800-803: only a ReLU reads the slope block
    return prep_all(fd, dim, [&](int g, tDimensions* d) { return st_[g]->prep(fd, d, slope) != nullptr; });
}
tBitPacked* IntFunc::Quantize::execute(tFixedPointPacked* in, tFixedPointPacked*) { return quant_bias<tBitPacked>(st_, in, true); }
tFixedPointPacked* IntFunc::Quantize::add_bias(tFixedPointPacked* in, tMultiBitPacked*) {
    tFixedPointPacked* r = quant_bias<tFixedPointPacked>(st_, in, false);
    materialize(r, true);
    shim::download_to_host(r);
    return r;
}
tFixedPointPacked* IntFunc::Quantize::relu_shift(tFixedPointPacked* in, tMultiBitPacked*, uint16_t*) {
    materialize(in, false);
    return run_stage<tFixedPointPacked>(in, in->len, [&](int g, redsec::Batch b) { return st_[g]->relu_shift(b, in->shard_cl > 0 ? g * in->shard_cl : 0); });
}
