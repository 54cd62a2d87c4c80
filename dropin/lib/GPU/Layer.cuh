// dropin/lib/GPU/Layer.cuh -- the names nets/*/{main,net}.cu and lib/GPU/{Bin,Int}Layer.cu expect from the reference's
// lib/GPU/Layer.cuh:13-177 (macros, enums, parameter structs, packed ciphertext arrays, tActParams, the *_calloc_global
// helpers, print_status), over the B200 engine.  Enums and parameter structs come from the engine's own host header; only
// the packed-array types live here.
#pragma once
#include <omp.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "REDcuFHE/redcufhe_gpu.cuh"
#include "redsec_layers.hpp"

typedef redcufhe::PubKey TFheGateBootstrappingCloudKeySet;   // lib/GPU/Layer.cuh:7

// The reference hard-codes `#define NUM_GPUS 1` and is edited by hand for more (lib/GPU/Layer.cuh:15).  Here it is a
// compile-time option (-DNUM_GPUS=8, dropin/build.sh REDSEC_NUM_GPUS=8): one engine context per device, one host thread
// per device, layers neuron-sharded with an NCCL all-gather between them, every GPU holding the full activations again
// after each layer -- the replicated enc_segs[NUM_GPUS] layout of the reference, made correct.
#ifndef NUM_GPUS
#define NUM_GPUS 1
#endif
#define MULTIBIT_BITS 12    // CPU-path message space 2^12 (lib/Layer.h:33), which the weights files assume
#define FIXEDPOINT_BITS 12

typedef float tFloat;
typedef redcufhe::Ctxt tBit;
// Array-of-ciphertext views with the reference's field names.  `dev[g]` is GPU g's device-resident batch: between layers only
// `dev` is populated; the host arrays exist for the network input (filled by main.cu) and the final output.
// `size` keeps the reference's uint8_t field (it stores uint8_t(len), lib/GPU/Layer.cu:47) for source
// compatibility only; the authoritative count is `len`.
struct tMultiBit {
    tBit* ctxt;
    uint32_t size;
    uint8_t gpu_id;
};
struct tBitPacked {
    tBit* enc_segs[NUM_GPUS];
    uint8_t size;
    uint32_t len;
    redsec::Batch dev[NUM_GPUS];
    bool pending_sign;     // dev[] holds PRE-activations: the sign bootstrap is issued by the consumer, which knows the encoding it
                           // needs (1/8 in front of a max-pool, 1/4096 otherwise) -- Func-level composition only
    int shard_c0, shard_cl;   // dev[g] holds channels [g*shard_cl, (g+1)*shard_cl) of the layer only (0 = full layer everywhere)
};
struct tMultiBitPacked {       // same layout as tBitPacked: net.cu:118 casts the last layer's result between the two
    tMultiBit* enc_segs[NUM_GPUS];
    uint8_t size;
    uint32_t len;
    redsec::Batch dev[NUM_GPUS];
    bool pending_sign;
    int shard_c0, shard_cl;
};
typedef tMultiBit tFixedPoint;
typedef tMultiBitPacked tFixedPointPacked;

typedef enum _ACTION { E_INIT, E_PREP, E_EXEC, E_PREP_BIAS, E_EXPORT, NUM_ACTIONS } eAction;   // lib/GPU/Layer.cuh:75-83
typedef union _ACT_PARAMS {                                                                      // lib/GPU/Layer.cuh:151-156
    tDimensions* d;
    tBitPacked* b;
    tFixedPointPacked* fp;
} tActParams;

uint64_t get_size(tRectangle* ws, uint16_t in_dep, uint16_t out_dep);
void netParamsCpy(tNetParams* dest, tNetParams* src);
void* arr_calloc(uint32_t len, uint8_t type_size);
void bit_calloc(tBit** ret, uint32_t len);
void bit_calloc_global(tBitPacked** ret, uint32_t len);
void mbit_calloc(tMultiBit** ret, uint32_t len, uint8_t bits);
void mbit_calloc_global(tMultiBitPacked** ret, uint32_t len, uint8_t bits);   // main.cu:58-60
void fixpt_calloc(tFixedPoint** ret, uint32_t len, uint8_t bits);
void fixpt_calloc_global(tFixedPointPacked** ret, uint32_t len, uint8_t bits);
void bit_free(uint32_t len, tBit* to_free);
void bit_free_global(tBitPacked* to_free);
void mbit_free(uint32_t len, tMultiBit* to_free);
void mbit_free_global(uint32_t len, tMultiBitPacked* to_free);
void fixpt_free(uint32_t len, tFixedPoint* to_free);
void fixpt_free_global(uint32_t len, tMultiBitPacked* to_free);
void print_status(const char* msg);
