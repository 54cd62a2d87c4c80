// dropin/lib/GPU/BinOps_gpu.cuh -- the BinOps:: surface of the reference's GPU build (prototypes as at
// lib/GPU/BinOps_gpu.cuh:7-47 so that lib/GPU/BinLayer.cu and the net drivers compile unchanged); bodies in
// dropin/src/ops_shim.cpp, each a count-1 call on the engine's batch API (gates.cuh says what that costs).
#pragma once
#include <cstdio>
#include "Layer.cuh"
#include "gates.cuh"

namespace BinOps {
using Stream = redcufhe::Stream;
using Ctxt = redcufhe::Ctxt;

// ---- weight-file readers (host only; the var_prep.dat block formats of SURVEY 5.4)
void get_filters(FILE* weights, tBitPacked** packed_out, uint32_t count);            // float weights -> trivial +-1/8 bits
void get_bitfilters(FILE* weights, tBitPacked** packed_out, uint32_t count);         // packed bits, MSB first
void get_ternfilters(FILE* weights, uint8_t* sign_out, uint8_t* nonzero_out, uint32_t count, float threshold);
void get_intfilters(FILE* weights, tMultiBitPacked** packed_out, uint32_t count);    // int32 block -> trivial samples, units of 1/4096
void get_intfilters_ptxt(FILE* weights, uint16_t* values_out, uint32_t count);

// ---- one bootstrap per call
void binarize_int(Ctxt& io, Stream st);                   // sign, -> +-1/4096 (lib/BinOps_enc.cpp:182-186)
void unbinarize_int(Ctxt& io, Stream st);                 // sign, -> +-1/2048 (:188-192)
void unbinarize_int_inv(Ctxt& io, Stream st);             // the same with the sign flipped
void max(tBit* out, const tBit* lhs, const tBit* rhs, Stream st);            // OR gate at +-1/8

// ---- gate-level bit-sliced arithmetic (one bootstrap per gate)
void add_bit(tMultiBit* sum_out, const tBit* lhs, const tBit* rhs, Stream st);                          // XOR + AND
void add(tMultiBit* sum_out, const tMultiBit* lhs, const tMultiBit* rhs, uint8_t bits, Stream st);     // ripple carry of full adders
void inc(tMultiBit* sum_out, const tMultiBit* value, const tBit* carry_in, Stream st);                 // XOR / AND chain
void relu(tFixedPoint* out, tMultiBit* value, uint8_t value_bits, Stream st);                          // value_bits - 1 AND gates with the top slice

// ---- bootstrap-free: one rs_lwe_axpby of count 1, or copies / negations
void int_add(Ctxt& out, const Ctxt& lhs, const Ctxt& rhs, Stream st);
void add_pc_ints(Ctxt& out, Ctxt& value, const uint16_t* plain_addend, uint8_t value_bits, Stream st);   // addend in units of 1/4096
void multiply_pc_ints(Ctxt& out, Ctxt& value, const uint16_t* plain_factor, uint8_t value_bits, uint8_t factor_bits, Stream st);
void multiply(tBit* out, const tBit* value, const uint8_t plain_bit, Stream st);      // XNOR with a plaintext bit: NOT or copy
void binarize(tBit* sign_out, const tMultiBit* value);                               // copy of the top slice
void shift(tMultiBit* out, tMultiBit* value, uint8_t value_bits, uint8_t by_bits, Stream st);            // arithmetic right shift of the slices
}  // namespace BinOps
