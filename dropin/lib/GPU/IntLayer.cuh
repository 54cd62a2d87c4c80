// dropin/lib/GPU/IntLayer.cuh -- IntLayer with the reference's GPU signature (lib/GPU/IntLayer.cuh:16-34) over redsec::Layer.
#pragma once
#include "Layer.cuh"

class IntLayer {
public:
    IntLayer(eConvType ec, uint16_t dep, ePoolType ep, eQuantType eq, tNetParams* np);
    ~IntLayer();
    tDimensions* prep(FILE* fd, tDimensions* dim);
    tBitPacked* execute(tMultiBitPacked* p_in);      // consumes p_in (callee frees input, lib/GPU/IntFunc_gpu.cu:455-458)
    void export_weights(FILE* fd);
    void set_print_layer(uint8_t i);
    tDimensions in_dim, out_dim;
private:
    redsec::Layer* impl_[NUM_GPUS];
    eQuantType quant_;
};
