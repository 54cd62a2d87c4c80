// dropin/lib/GPU/BinLayer.cuh -- BinLayer with the reference's GPU signature (lib/GPU/BinLayer.cuh:16-34) over redsec::Layer.
#pragma once
#include "Layer.cuh"

class BinLayer {
public:
    BinLayer(eConvType ec, uint16_t dep, ePoolType ep, eQuantType eq, tNetParams* np);
    ~BinLayer();
    tDimensions* prep(FILE* fd, tDimensions* dim);
    tBitPacked* execute(tBitPacked* p_in);           // consumes p_in; the last layer's result is a tMultiBitPacked (net.cu:118)
    void export_weights(FILE* fd);
    void set_print_layer(uint8_t i);
    tDimensions in_dim, out_dim;
private:
    redsec::Layer* impl_[NUM_GPUS];
    eQuantType quant_;
};
