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

typedef uint32_t Torus;                                   // torus32, as (RED)cuFHE's
inline Torus ModSwitchToTorus(int32_t mu, int32_t space) { return rs_modswitch_to_torus32(mu, space); }   // lib/GPU/gates.cu:125

// (RED)cuFHE's per-ciphertext launches take a Stream (lib/GPU/gates.cuh).  The engine batches and orders its own work, so the
// facade's Stream only has to exist and be creatable / destroyable the way lib/GPU/BinFunc_gpu.cu:116-138 does it.
class Stream {
public:
    Stream() = default;
    Stream(int) {}                                        // the reference passes a plain int in places
    void Create() {}
    void Destroy() {}
    cudaStream_t st() const { return nullptr; }
};

void Copy(Ctxt& out, const Ctxt& in, Stream st = Stream());        // host copy of one sample
inline void Copy(Ctxt& out, const Ctxt* in, Stream st = Stream()) { Copy(out, *in, st); }   // lib/GPU/BinOps_gpu.cu:124 passes a pointer
void Not(Ctxt& out, const Ctxt& in, Stream st = Stream());         // LWE negation (NotOp, lib/GPU/gates.cu:110-122)
inline void Not(Ctxt& out, const Ctxt* in, Stream st = Stream()) { Not(out, *in, st); }

void ReadPubKeyFromFile(PubKey& key, const char* path);   // nets/*/net.cu:43
void Initialize(PubKey& key);                             // once per device, after cudaSetDevice (net.cu:45-48)
void ReadCtxtFromFileRed(Ctxt& ct, std::ifstream& in);    // main.cu:67: next record of image.ctxt
void WriteCtxtToFileRed(Ctxt& ct, const char* path);      // main.cu:82: appends one record
void Synchronize();                                       // main.cu:74
void CuCheckError();                                      // main.cu:75: aborts with the engine's last error, as the macro did
void CleanUp();                                           // main.cu:84
rs_ctx* CurrentContext();                                 // the engine context of the calling thread's current device (shim-internal)
rs_ctx* ContextOf(int device);                            // ... of a given device (NULL if Initialize was not called on it)
// communicator of `device` in the single-process group over devices 0..n-1 (created on first use; NULL when n == 1)
rs_comm* CommunicatorOf(int device, int n);

}  // namespace redcufhe
