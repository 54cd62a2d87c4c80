// dropin/lib/GPU/Layer.cuh -- the names nets/*/{main,net}.cu expect from lib/GPU/Layer.cuh:42-151 (enums, parameter structs,
// packed ciphertext arrays, mbit_calloc_global, print_status), over the B200 engine.  Enums and parameter structs come from
// the engine's own host header; only the packed-array types live here.
#pragma once
#include <omp.h>

#include <cstdint>
#include <cstdio>

#include "REDcuFHE/redcufhe_gpu.cuh"
#include "redsec_layers.hpp"

#define NUM_GPUS 1          // one process per GPU here; multi-GPU runs shard layers across processes
#define MULTIBIT_BITS 12    // CPU-path message space 2^12 (lib/Layer.h:33), which the weights files assume

typedef redcufhe::Ctxt tBit;
// Array-of-ciphertext views with the reference's field names.  `dev` is the engine's device-resident batch: between layers
// only `dev` is populated; the host arrays exist for the network input (filled by main.cu) and the final output.
struct tMultiBit {
    tBit* ctxt;
    uint32_t size;
    uint8_t gpu_id;
};
struct tBitPacked {
    tBit* enc_segs[NUM_GPUS];
    uint8_t size;
    redsec::Batch dev;
};
struct tMultiBitPacked {       // same layout as tBitPacked: net.cu:118 casts the last layer's result between the two
    tMultiBit* enc_segs[NUM_GPUS];
    uint8_t size;
    redsec::Batch dev;
};
typedef tMultiBit tFixedPoint;
typedef tMultiBitPacked tFixedPointPacked;

void mbit_calloc_global(tMultiBitPacked** ret, uint32_t len, uint8_t bits);   // main.cu:58-60
void bit_calloc_global(tBitPacked** ret, uint32_t len);
void print_status(const char* msg);
