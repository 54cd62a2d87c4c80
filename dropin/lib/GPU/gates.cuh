// dropin/lib/GPU/gates.cuh -- the per-ciphertext gate / arithmetic entry points of the reference's lib/GPU/gates.cuh:10-29,
// same names and argument meaning, over the engine's batch API with count = 1.  Every call uploads its operands, runs the
// batched kernel on one ciphertext and downloads the result: correct, and as slow per call as the reference's own
// H2D -> Bootstrap -> D2H (lib/GPU/gates.cu:124-130); code that wants throughput uses the Func / Layer classes, which batch.
// Encodings follow the reference's CPU path: binarize emits +-1/4096 (lib/BinOps_enc.cpp:182-186), unbinarize
// +-1/2048 (:188-192), gates take and emit +-1/8 (lib/GPU/gates.cu:246-286).
#pragma once
#include "REDcuFHE/redcufhe_gpu.cuh"
#include <vector>

void CtxtCopyD2H(const redcufhe::Ctxt& c, redcufhe::Stream st);   // no-ops: a facade Ctxt is host memory
void CtxtCopyH2D(const redcufhe::Ctxt& c, redcufhe::Stream st);
void redsec_binarize_bootstrap(redcufhe::Ctxt& out, redcufhe::Stream st);
void redsec_unbinarize_bootstrap(redcufhe::Ctxt& out, redcufhe::Stream st);
void redsec_unbinarize_bootstrap_inv(redcufhe::Ctxt& out, redcufhe::Stream st);
void bootsNAND(redcufhe::Ctxt& out, const redcufhe::Ctxt& in0, const redcufhe::Ctxt& in1, redcufhe::Stream st);
void bootsOR(redcufhe::Ctxt& out, const redcufhe::Ctxt& in0, const redcufhe::Ctxt& in1, redcufhe::Stream st);
void bootsAND(redcufhe::Ctxt& out, const redcufhe::Ctxt& in0, const redcufhe::Ctxt& in1, redcufhe::Stream st);
void bootsNOR(redcufhe::Ctxt& out, const redcufhe::Ctxt& in0, const redcufhe::Ctxt& in1, redcufhe::Stream st);
void bootsXOR(redcufhe::Ctxt& out, const redcufhe::Ctxt& in0, const redcufhe::Ctxt& in1, redcufhe::Stream st);
void bootsXNOR(redcufhe::Ctxt& out, const redcufhe::Ctxt& in0, const redcufhe::Ctxt& in1, redcufhe::Stream st);
void levelNOT(redcufhe::Ctxt& out, const redcufhe::Ctxt& in0, redcufhe::Stream st);
void NoiselessTrivial(redcufhe::Ctxt& result, redcufhe::Torus mu);
void levelCONSTANT(redcufhe::Ctxt& result, int32_t value);
void add_int(redcufhe::Ctxt& sum, const redcufhe::Ctxt& a, const redcufhe::Ctxt& b, redcufhe::Stream st);
void mul_int(redcufhe::Ctxt& prod, const redcufhe::Ctxt& a, uint16_t b);
void sub_int(redcufhe::Ctxt& res, const redcufhe::Ctxt& a, const redcufhe::Ctxt& b, redcufhe::Stream st);
void bootstrapped_full_adder(redcufhe::Ctxt& sum, redcufhe::Ctxt& carry_out, redcufhe::Ctxt& temp_a, redcufhe::Ctxt& temp_b,
                             const redcufhe::Ctxt& a, const redcufhe::Ctxt& b, const redcufhe::Ctxt& carry_in, redcufhe::Stream st);
