// dropin/lib/GPU/IntOps_gpu.cuh -- the IntOps:: surface the reference's GPU layer code calls (prototypes as at
// lib/GPU/IntOps_gpu.cuh:6-29, so that callers compile unchanged), implemented in dropin/src/ops_shim.cpp as count-1 calls on the
// engine's batch API.  Grouped by what a call costs here:
#pragma once
#include "Layer.cuh"
#include "gates.cuh"

namespace IntOps {
using Stream = redcufhe::Stream;
using Ctxt = redcufhe::Ctxt;

// ---- bootstraps
void binarize_int(tBit* io, Stream st);                                          // one sign bootstrap, in place, -> +-1/4096
void relu(tFixedPoint* out, tFixedPoint* value, uint8_t value_bits, Stream st);  // value_bits - 1 AND gates with the top slice

// ---- bootstrap-free: one rs_lwe_axpby of count 1, or copies / negations of bit slices
void add(tFixedPoint* out, const tFixedPoint* lhs, const tFixedPoint* rhs, Stream st);
void subtract(tFixedPoint* out, const tFixedPoint* lhs, const tFixedPoint* rhs, Stream st);
void add_pc_ints(Ctxt& out, Ctxt& value, const uint16_t* plain_addend, uint8_t value_bits, Stream st);     // addend in units of 1/4096
void multiply_pc_ints(Ctxt& out, Ctxt& value, const uint16_t* plain_factor, uint8_t value_bits, uint8_t factor_bits, Stream st);
void binarize(tBit* sign_out, const tFixedPoint* value, Stream st);              // copy of the top slice of a bit-sliced value
void invert(tFixedPoint* out, const tFixedPoint* value, const uint8_t* keep, Stream st);          // *keep == 1: copy, else NOT, per slice
void shift(tFixedPoint* out, tFixedPoint* value, uint8_t value_bits, uint8_t by_bits, Stream st); // arithmetic right shift of the slices
}  // namespace IntOps
