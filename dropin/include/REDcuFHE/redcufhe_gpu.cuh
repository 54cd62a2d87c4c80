// dropin/include/REDcuFHE/redcufhe_gpu.cuh -- source-level facade of the (RED)cuFHE entry points that REDsec's generated
// GPU drivers call directly (SURVEY.md 8b "B-outer"): nets/*/net.cu:43-49 (PubKey, ReadPubKeyFromFile, Initialize) and
// nets/*/main.cu:58-84 (ReadCtxtFromFileRed, WriteCtxtToFileRed, Synchronize, CuCheckError, CleanUp).  Behind it sits the
// B200 engine's C-ABI (include/redsec_b200.h); the library built from dropin/src/facade.cpp is named libredcufhe.so so the
// reference's own link line (`-lredcufhe`, nets/*/Makefile:27-29) resolves to it unchanged.
//
// Differences a maintainer should know: ciphertext and key FILES are this repo's formats (upstream layouts
// are unverifiable here), and the layer semantics follow the reference's CPU path (inputs 2p-255, mu = 1/4096).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <fstream>
#include <iostream>   // the generated drivers use std::cout without including it themselves (main.cu:78)

#include "redsec_b200.h"

namespace redcufhe {

// One LWE sample in wire order (a[0..349], b) plus the variance field of the file record.  Host memory: the engine keeps
// activations on the device and only the network input / output ever exist as Ctxt objects.
struct Ctxt {
    uint32_t lwe[RS_LWE_WORDS];
    double variance = 0.0;
};

// Evaluation key as read from eval.key (torus32 BSK + KSK, host).  Initialize() converts it to the device layouts.
struct PubKey {
    uint32_t* bsk = nullptr;
    uint32_t* ksk = nullptr;
    ~PubKey();
};

void ReadPubKeyFromFile(PubKey& key, const char* path);   // nets/*/net.cu:43
void Initialize(PubKey& key);                             // once per device, after cudaSetDevice (net.cu:45-48)
void ReadCtxtFromFileRed(Ctxt& ct, std::ifstream& in);    // main.cu:67: next record of image.ctxt
void WriteCtxtToFileRed(Ctxt& ct, const char* path);      // main.cu:82: appends one record
void Synchronize();                                       // main.cu:74
void CuCheckError();                                      // main.cu:75: aborts with the engine's last error, as the macro did
void CleanUp();                                           // main.cu:84
rs_ctx* CurrentContext();                                 // the engine context of the calling thread's current device (shim-internal)

}  // namespace redcufhe
