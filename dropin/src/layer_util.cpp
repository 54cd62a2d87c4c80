// dropin/src/layer_util.cpp -- the allocation helpers and print_status of lib/GPU/Layer.cu:7-194 for the facade's packed arrays.
#include "shim_common.hpp"

void print_status(const char* msg) { fputs(msg, stdout); fflush(stdout); }
uint64_t get_size(tRectangle* ws, uint16_t in_dep, uint16_t out_dep) { return (uint64_t)ws->h * ws->w * in_dep * out_dep; }
void netParamsCpy(tNetParams* dest, tNetParams* src) { *dest = *src; }     // the reference omits `quant`
void* arr_calloc(uint32_t len, uint8_t type_size) { return calloc(len, type_size); }

void bit_calloc(tBit** ret, uint32_t len) { *ret = new tBit[len](); }
void mbit_calloc(tMultiBit** ret, uint32_t len, uint8_t bits) {
    *ret = new tMultiBit[len];
    for (uint32_t i = 0; i < len; i++) { (*ret)[i].size = bits; (*ret)[i].ctxt = new tBit[bits](); (*ret)[i].gpu_id = 0; }
}
void fixpt_calloc(tFixedPoint** ret, uint32_t len, uint8_t) { mbit_calloc(ret, len, 1); }

void bit_calloc_global(tBitPacked** ret, uint32_t len) {
    tBitPacked* p = shim::new_packed<tBitPacked>(len);
    for (int g = 0; g < NUM_GPUS; g++) p->enc_segs[g] = new tBit[len]();
    *ret = p;
}
void mbit_calloc_global(tMultiBitPacked** ret, uint32_t len, uint8_t bits) {
    tMultiBitPacked* p = shim::new_packed<tMultiBitPacked>(len);
    for (int g = 0; g < NUM_GPUS; g++) {
        p->enc_segs[g] = new tMultiBit[len];
        for (uint32_t i = 0; i < len; i++) {
            p->enc_segs[g][i].ctxt = new tBit[bits]();
            p->enc_segs[g][i].size = bits;
            p->enc_segs[g][i].gpu_id = (uint8_t)g;
        }
    }
    *ret = p;
}
void fixpt_calloc_global(tFixedPointPacked** ret, uint32_t len, uint8_t bits) { mbit_calloc_global(ret, len, bits ? bits : 1); }

void bit_free(uint32_t, tBit* to_free) { delete[] to_free; }
void mbit_free(uint32_t len, tMultiBit* to_free) {
    for (uint32_t i = 0; i < len; i++) delete[] to_free[i].ctxt;
    delete[] to_free;
}
void fixpt_free(uint32_t len, tFixedPoint* to_free) { mbit_free(len, to_free); }
static void free_dev(redsec::Batch* dev) {
    for (int g = 0; g < NUM_GPUS; g++)
        if (dev[g].dev) { rs_lwe_free(shim::ctx_of(g), dev[g].dev); dev[g] = redsec::Batch(); }
}
void bit_free_global(tBitPacked* to_free) { shim::free_host(to_free); free_dev(to_free->dev); delete to_free; }
void mbit_free_global(uint32_t, tMultiBitPacked* to_free) { shim::free_host(to_free); free_dev(to_free->dev); delete to_free; }
void fixpt_free_global(uint32_t len, tMultiBitPacked* to_free) { mbit_free_global(len, to_free); }
