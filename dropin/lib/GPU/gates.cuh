// dropin/lib/GPU/gates.cuh -- the per-ciphertext gate / arithmetic entry points of the reference's lib/GPU/gates.cuh:10-29,
// same names and argument meaning, over the engine's batch API with count = 1.  Every call uploads its operands, runs the
// batched kernel on one ciphertext and downloads the result: correct, and as slow per call as the reference's own
// H2D -> Bootstrap -> D2H (lib/GPU/gates.cu:124-130); code that wants throughput uses the Func / Layer classes, which batch.
// Encodings follow the reference's CPU path: binarize emits +-1/4096 (lib/BinOps_enc.cpp:182-186), unbinarize
// +-1/2048 (:188-192), gates take and emit +-1/8 (lib/GPU/gates.cu:246-286).
#pragma once
#include "REDcuFHE/redcufhe_gpu.cuh"
#include <vector>

namespace redsec_facade {
using Ctxt = redcufhe::Ctxt;
using Stream = redcufhe::Stream;
}

// ---- two-input gates, inputs and output at +-1/8: one rs_gate_batch of count 1 each
#define RS_FACADE_GATE2(name) void name(redsec_facade::Ctxt& out, const redsec_facade::Ctxt& lhs, const redsec_facade::Ctxt& rhs, redsec_facade::Stream st)
RS_FACADE_GATE2(bootsAND);
RS_FACADE_GATE2(bootsNAND);
RS_FACADE_GATE2(bootsOR);
RS_FACADE_GATE2(bootsNOR);
RS_FACADE_GATE2(bootsXOR);
RS_FACADE_GATE2(bootsXNOR);
#undef RS_FACADE_GATE2
void bootstrapped_full_adder(redsec_facade::Ctxt& sum, redsec_facade::Ctxt& carry_out, redsec_facade::Ctxt& scratch_a, redsec_facade::Ctxt& scratch_b,
                             const redsec_facade::Ctxt& lhs, const redsec_facade::Ctxt& rhs, const redsec_facade::Ctxt& carry_in,
                             redsec_facade::Stream st);

// ---- sign bootstraps, in place: one rs_pbs_batch of count 1 each
void redsec_binarize_bootstrap(redsec_facade::Ctxt& io, redsec_facade::Stream st);           // -> +-1/4096
void redsec_unbinarize_bootstrap(redsec_facade::Ctxt& io, redsec_facade::Stream st);         // -> +-1/2048
void redsec_unbinarize_bootstrap_inv(redsec_facade::Ctxt& io, redsec_facade::Stream st);     // -> -+1/2048

// ---- bootstrap-free: one rs_lwe_axpby of count 1 each (NoiselessTrivial / levelCONSTANT only fill the host words)
void levelNOT(redsec_facade::Ctxt& out, const redsec_facade::Ctxt& value, redsec_facade::Stream st);
void add_int(redsec_facade::Ctxt& out, const redsec_facade::Ctxt& lhs, const redsec_facade::Ctxt& rhs, redsec_facade::Stream st);
void sub_int(redsec_facade::Ctxt& out, const redsec_facade::Ctxt& lhs, const redsec_facade::Ctxt& rhs, redsec_facade::Stream st);
void mul_int(redsec_facade::Ctxt& out, const redsec_facade::Ctxt& value, uint16_t plain_factor);
void NoiselessTrivial(redsec_facade::Ctxt& out, redcufhe::Torus mu);
void levelCONSTANT(redsec_facade::Ctxt& out, int32_t value);

// ---- no-ops here: a facade Ctxt lives in host memory, the engine moves data itself
void CtxtCopyH2D(const redsec_facade::Ctxt& c, redsec_facade::Stream st);
void CtxtCopyD2H(const redsec_facade::Ctxt& c, redsec_facade::Stream st);
