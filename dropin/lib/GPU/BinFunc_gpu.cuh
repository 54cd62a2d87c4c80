// dropin/lib/GPU/BinFunc_gpu.cuh -- BinFunc::{Convolution,SumPooling,MaxPooling,Quantize} with the reference's GPU signatures
// (lib/GPU/BinFunc_gpu.cuh:16-147) as batched device stages (redsec::ConvStage / SumPoolStage / QuantizeStage / MaxPoolStage).
// A net composed from these objects -- e.g. by the reference's own lib/GPU/BinLayer.cu:114-203, which dropin/build.sh compiles
// UNMODIFIED against this header -- runs the same batched bootstrap launches as redsec::Layer and yields the same ciphertexts.
// Weight-convert members (extract_*, export_weights, BatchNorm) are offline tooling and are not provided.
#pragma once
#include <cstdio>
#include "Layer.cuh"
#include "REDcuFHE/redcufhe_gpu.cuh"
namespace BinFunc
{
    class Convolution {
    public:
        Convolution(uint32_t out_depth, tConvParams* in_params);
        ~Convolution();
        tDimensions* prep(FILE* fd_filt, tDimensions* ret_dim);
        tMultiBitPacked* execute(tBitPacked* p_inputs);          // consumes p_inputs
        void extract_bias(FILE*, tMultiBitPacked*, eBiasType) {}   // weight-convert only
        void export_weights(FILE*) { printf("Weight Convert not defined\r\n"); }
        void get_outhw(tRectangle* ret_dim);
        void get_outdep(uint32_t* ret_dep);
    private:
        redsec::ConvStage* st_[NUM_GPUS];
        tDimensions out_dim_;
    };
    class SumPooling {
    public:
        SumPooling(tPoolParams* in_params);
        ~SumPooling();
        tDimensions* prep(tDimensions* ret_dim);
        tMultiBitPacked* execute(tMultiBitPacked* p_inputs);
        void extract_bias(tMultiBitPacked*) {}
    private:
        redsec::SumPoolStage* st_[NUM_GPUS];
    };
    class MaxPooling {
    public:
        MaxPooling(tPoolParams* in_params);
        ~MaxPooling();
        tDimensions* prep(tDimensions* ret_dim);
        tBitPacked* execute(tBitPacked* p_inputs);               // takes Quantize::execute's result (sign bootstrap still pending)
    private:
        redsec::MaxPoolStage* st_[NUM_GPUS];
    };
    class Quantize {
    public:
        Quantize(tQParams* qparam);
        ~Quantize();
        tDimensions* prep(FILE* fd_bias, tDimensions* ret_dim, tMultiBitPacked* p_bias, uint16_t* p_slope);
        tBitPacked* execute(tMultiBitPacked* p_inputs, tMultiBitPacked* p_bias);
        tMultiBitPacked* add_bias(tMultiBitPacked* p_inputs, tMultiBitPacked* p_bias);
        tFixedPointPacked* relu_shift(tMultiBitPacked* p_inputs, tMultiBitPacked* p_bias, uint16_t* p_slope);
        void extract_bias(tMultiBitPacked*, uint16_t*) {}
        void export_weights(FILE*, tMultiBitPacked*, uint16_t*) { printf("Weight Convert not defined\r\n"); }
    private:
        redsec::QuantizeStage* st_[NUM_GPUS];
    };
}
