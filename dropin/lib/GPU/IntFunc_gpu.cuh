// dropin/lib/GPU/IntFunc_gpu.cuh -- IntFunc::{Convolution,SumPooling,Quantize} with the reference's GPU signatures
// (lib/GPU/IntFunc_gpu.cuh:8-124) as batched device stages; see BinFunc_gpu.cuh.  IntFunc convolutions follow the reference's
// CPU convention for zero weights / padding (-1/4096, lib/IntFunc.cpp:268,277).
#pragma once
#include <cstdio>
#include "Layer.cuh"
#include "REDcuFHE/redcufhe_gpu.cuh"
namespace IntFunc
{
    class Convolution {
    public:
        Convolution(uint16_t out_depth, tConvParams* in_params);
        ~Convolution();
        tDimensions* prep(FILE* fd_filt, tDimensions* ret_dim);
        tFixedPointPacked* execute(tFixedPointPacked* p_inputs);
        void export_weights(FILE*) { printf("Weight Convert not defined\r\n"); }
    private:
        redsec::ConvStage* st_[NUM_GPUS];
    };
    class SumPooling {
    public:
        SumPooling(tPoolParams* in_params);
        ~SumPooling();
        tDimensions* prep(tDimensions* ret_dim);
        tFixedPointPacked* execute(tFixedPointPacked* p_inputs);
        void extract_bias(tFixedPointPacked*) {}
    private:
        redsec::SumPoolStage* st_[NUM_GPUS];
    };
    class Quantize {
    public:
        Quantize(tQParams* qparam);
        ~Quantize();
        tDimensions* prep(FILE* fd_bias, tDimensions* ret_dim, tMultiBitPacked** p_bias, uint16_t* p_slope);
        tBitPacked* execute(tFixedPointPacked* p_inputs, tFixedPointPacked* p_bias);
        tFixedPointPacked* add_bias(tFixedPointPacked* p_inputs, tMultiBitPacked* p_bias);
        tFixedPointPacked* relu_shift(tFixedPointPacked* p_inputs, tMultiBitPacked* p_bias, uint16_t* p_slope);
        void export_weights(FILE*, tMultiBitPacked*, uint16_t*) { printf("Weight Convert not defined\r\n"); }
    private:
        redsec::QuantizeStage* st_[NUM_GPUS];
        bool relu_;
    };
}
