// redsec_b200/host/redsec_layers.hpp -- host-side mirror of REDsec's layer API for the encrypted GPU path.
//
// Same names, constructor arguments, prep/execute sequencing and ownership rules as the reference's
// lib/GPU/{Layer,IntLayer,BinLayer}.cuh (IntLayer/BinLayer(eConvType, uint16_t, ePoolType, eQuantType, tNetParams*),
// tDimensions* prep(FILE*, tDimensions*), execute(...), public in_dim/out_dim), but activations are
// device-resident LWE batches (redsec::Batch) instead of arrays of individually allocated ciphertexts, and
// every layer issues ONE batched bootstrap launch (plus the OR-tree levels of a max-pool) through the C-ABI
// in include/redsec_b200.h.  Implementation: redsec_b200/csrc/layers.cpp (host C++ only; all device work goes
// through rs_* calls).
#pragma once
#include <cstdint>
#include <cstdio>
#include <memory>
#include <vector>

#include "../../include/redsec_b200.h"

// ---- enums and parameter structs: the API surface of lib/Layer.h:38-177 / lib/GPU/Layer.cuh:42-151
enum eConvType { E_NO_CONV, E_CONV, E_FC, E_FC_FINAL, NUM_CONVS };
enum ePoolType { E_NO_POOL, E_MAXPOOL, E_SUMPOOL, NUM_POOLS };
enum eBiasType { E_NO_BIAS, E_BIAS, E_BNORM, NUM_BIASES };
enum eQuantType { E_ACTIVATION_NONE, E_ACTIVATION_SIGN, E_ACTIVATION_RELU, NUM_ACTIVATIONS };

#define SIZE_EMPTY 1
#define SINGLE_BIT 1
#define MSG_SPACE 4096

struct tRectangle { int16_t h, w; };
struct tDimensions {
    tRectangle hw;
    uint32_t in_dep;
    uint8_t in_bits, out_bits, filter_bits, bias_bits;
    uint32_t up_bound;
    float scale;
};
struct tConvParams { tRectangle window; uint16_t w_max; bool same_pad; float tern_thresh; tRectangle stride; };   // lib/GPU/Layer.cuh:107-114
struct tBNormParams { bool use_scale; float eps; };
struct tPoolParams { tRectangle window; bool same_pad; tRectangle stride; };
struct tQParams { uint8_t shift_bits; };
struct tNetParams {
    tConvParams conv;
    tPoolParams pool;
    tBNormParams bnorm;
    tQParams quant;
    eBiasType e_bias;
    uint16_t version;
};

namespace redsec {

// A batch of LWE ciphertexts on the device (rows of RS_LWE_STRIDE words).  Ownership follows the reference:
// every execute() consumes (frees) its input and returns a freshly allocated batch (lib/BinFunc.cpp:327-329,1073).
struct Batch {
    uint32_t* dev = nullptr;
    size_t count = 0;
};

struct ShardSpec {   // neuron sharding across ranks: contiguous blocks of output channels
    int rank = 0, world = 1;
};

class LayerImpl;

// Common implementation; IntLayer / BinLayer below differ only in the conv semantics for zero weights and padding
// (lib/IntFunc.cpp:268,277 vs lib/BinFunc.cpp:265,280).
class Layer {
public:
    Layer(rs_ctx* ctx, bool int_inputs, eConvType ec, uint16_t depth, ePoolType ep, eQuantType eq, tNetParams* np);
    ~Layer();
    tDimensions* prep(FILE* fd, tDimensions* dim);       // lib/BinLayer.cpp:89-113, lib/IntLayer.cpp:93-117
    Batch execute(Batch in);                              // lib/BinLayer.cpp:122-127 (run(E_EXEC))
    // Sharded execution: computes this rank's output-channel slice [begin,end) of the layer for all pixels,
    // bootstraps and max-pools it locally; the caller all-gathers the slices (rows [pixel][c_local]) and
    // calls interleave_shards() to restore the canonical (h,w,c) order.  A layer without a conv stage (input layers)
    // is sharded by output pixel instead when its pixel count divides the world size: [begin,end) is then the full
    // channel range, the result holds out_count()/world rows, and the all-gather of those blocks is already canonical.
    Batch execute_shard(const Batch& in, ShardSpec shard, int* ch_begin, int* ch_end);
    // The whole exchange step in one call, all on the context's stream (no host synchronisation): this rank's slice
    // (execute_shard), NCCL all-gather of the slices (rs_allgather), interleave back to (h,w,c).  Consumes `in` and returns
    // the FULL layer output on every rank -- the replicated layout of the reference's tBitPacked::enc_segs[NUM_GPUS]
    // (lib/GPU/Layer.cuh:22-26).  comm == nullptr or world 1: same as execute().
    Batch execute_sharded(Batch in, rs_comm* comm);
    Batch forward_sharded(const Batch& in, rs_comm* comm);   // same without consuming `in`
    // How the layer is split over `world` ranks: 0 = computed whole on every rank, 1 = by output-channel block, 2 = by output pixel
    int shard_mode(int world) const;
    // uploads every device table execute()/execute_shard() needs for this rank's slice now instead of on first use
    int build_tables(ShardSpec shard = ShardSpec());
    size_t out_count() const;                             // ciphertexts in the full output
    int out_channels() const;
    size_t bootstraps() const;                            // PBS issued per execute (sign + OR tree, or one per ReLU neuron)
    // IntFunc conv convention (call before prep): false (default) = the reference's encrypted branch, true = its plaintext
    // twin (weight -1 contributes -x-1, zero weight 0); Net::prep switches every layer of a ReLU net to the twin.
    void set_int_conv_twin(bool on);
    bool is_relu() const;
    tDimensions in_dim{}, out_dim{};
private:
    std::unique_ptr<LayerImpl> impl_;
};

class IntLayer : public Layer {   // lib/GPU/IntLayer.cuh:16-34
public:
    IntLayer(rs_ctx* ctx, eConvType ec, uint16_t depth, ePoolType ep, eQuantType eq, tNetParams* np)
        : Layer(ctx, true, ec, depth, ep, eq, np) {}
};
class BinLayer : public Layer {   // lib/GPU/BinLayer.cuh:16-34
public:
    BinLayer(rs_ctx* ctx, eConvType ec, uint16_t depth, ePoolType ep, eQuantType eq, tNetParams* np)
        : Layer(ctx, false, ec, depth, ep, eq, np) {}
};

// A sequential network = the generated HeBNN classes of nets/*/net.cu (init + run).
class Net {
public:
    explicit Net(rs_ctx* ctx) : ctx_(ctx) {}
    Layer* add(bool int_layer, eConvType ec, uint16_t depth, ePoolType ep, eQuantType eq, tNetParams* np);
    int prep(FILE* weights, tDimensions* input_dim);      // HeBNN::init (nets/mnist/sign1024x1/net.cpp:46-113)
    Batch run(Batch in);                                  // HeBNN::run (net.cpp:117-131); consumes in
    Batch run_sharded(Batch in, rs_comm* comm);           // same, every layer neuron-sharded over the communicator's ranks
    size_t num_layers() const { return layers_.size(); }
    Layer* layer(size_t i) { return layers_[i].get(); }
    size_t bootstraps() const;
private:
    rs_ctx* ctx_;
    std::vector<std::unique_ptr<Layer>> layers_;
};

// ---- Func-level stages: what {Bin,Int}Func::{Convolution,SumPooling,Quantize,MaxPooling} (lib/GPU/BinFunc_gpu.cuh:16-147,
// lib/GPU/IntFunc_gpu.cuh) are in the reference, as batched device stages.  A caller that composes its own network from Func
// objects (the way lib/GPU/BinLayer.cu:114-203 and IntLayer.cu:90-170 do) gets the same batched launches as redsec::Layer.
// Every execute() consumes its input batch and returns a new one (reference ownership rule).
// Channel slices: with several GPUs a caller computes output channels [c0,c1) of a conv on each device, runs the pooling /
// quantize / max-pool stages on the slice (they are per channel) and restores the full layer with gather_channels() before
// the next conv.  c1 < 0 means "all channels".
class ConvStage {          // Convolution::prep / execute: ternary conv or FC on LWE rows, NO bias (Quantize adds it)
public:
    ConvStage(rs_ctx* ctx, bool int_inputs, uint32_t out_depth, const tConvParams& conv);
    ~ConvStage();
    tDimensions* prep(FILE* fd, tDimensions* dim);        // lib/BinFunc.cpp:76-133: dimension pass + ternary block
    Batch execute(Batch in, int c0 = 0, int c1 = -1);     // lib/BinFunc.cpp:142-330 / lib/IntFunc.cpp:152-319; rows (oh,ow,c0..c1)
    int out_depth() const;
private:
    std::unique_ptr<LayerImpl> impl_;
};
class SumPoolStage {       // SumPooling::prep / execute (lib/IntFunc.cpp:598-700)
public:
    SumPoolStage(rs_ctx* ctx, const tPoolParams& pool);
    ~SumPoolStage();
    tDimensions* prep(tDimensions* dim);
    Batch execute(Batch in);                              // channel count inferred from in.count (full layer or a slice)
private:
    std::unique_ptr<LayerImpl> impl_;
};
class QuantizeStage {      // Quantize::prep / execute / add_bias / relu_shift (lib/BinFunc.cpp:985-1162, lib/IntFunc.cpp:800-973)
public:
    QuantizeStage(rs_ctx* ctx, bool int_inputs, const tQParams& q);
    ~QuantizeStage();
    tDimensions* prep(FILE* fd, tDimensions* dim, bool read_slope);   // bias block (+ slope block when read_slope && shift_bits > 1)
    Batch add_bias(Batch in, int c0 = 0);                 // (0,bias[c0+c]) + in, no bootstrap; in place
    // sign activation, split in two so the consumer decides the output encoding: pre_sign() adds the bias and returns the
    // pre-activations; sign_bootstrap() is the ONE batched bootstrap, with mu = 1/4096 when a conv / the client reads the bits
    // and mu = 1/8 when a MaxPoolStage follows (an OR gate needs +-1/8 inputs).
    Batch pre_sign(Batch in, int c0 = 0);
    static int sign_bootstrap(rs_ctx* ctx, Batch& pre, uint32_t mu);
    Batch relu_shift(Batch in, int c0 = 0);               // IntFunc only: one test-vector bootstrap per neuron
    const std::vector<int32_t>& bias() const;
    int channels() const;
private:
    std::unique_ptr<LayerImpl> impl_;
};
class MaxPoolStage {       // MaxPooling::prep / execute (lib/BinFunc.cpp:836-925) as an OR tree
public:
    MaxPoolStage(rs_ctx* ctx, const tPoolParams& pool);
    ~MaxPoolStage();
    tDimensions* prep(tDimensions* dim);
    // `pre` holds the PRE-activations of the sign stage (QuantizeStage::pre_sign); this stage issues the sign bootstrap at 1/8,
    // the OR tree, and emits bits at +-1/4096 -- the same block-pipelined launches as redsec::Layer
    Batch execute_from_preact(Batch pre);
private:
    std::unique_ptr<LayerImpl> impl_;
};
// restores the full (h,w,C) layer from every rank's channel slice [pixel][c_local]: all-gather + interleave; consumes `part`
Batch gather_channels(rs_ctx* ctx, rs_comm* comm, Batch part, int c_local);

}  // namespace redsec
