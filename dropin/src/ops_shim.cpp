// dropin/src/ops_shim.cpp -- the per-ciphertext entry points of lib/GPU/gates.cu and the BinOps:: / IntOps:: namespaces
// (lib/GPU/BinOps_gpu.cuh:7-47, lib/GPU/IntOps_gpu.cuh:6-29; CPU twins lib/BinOps_enc.h:8-49, lib/IntOps_enc.h:9-32) as
// count-1 calls on the engine's batch API: upload the operands, one batched kernel launch over ONE ciphertext, download.
// Correct and deliberately simple -- the reference itself copies every operand to the device and back around each call
// (lib/GPU/gates.cu:124-130,182-193).  Throughput comes from the Func / Layer classes, which batch a whole layer.
#include "BinOps_gpu.cuh"
#include "IntOps_gpu.cuh"
#include "shim_common.hpp"

using redcufhe::Ctxt;
using redcufhe::Stream;
using redcufhe::Torus;

namespace {

constexpr uint32_t kEighth = 1u << 29;

struct Dev1 {      // one device row, freed on scope exit
    rs_ctx* ctx; uint32_t* p = nullptr;
    explicit Dev1(rs_ctx* c) : ctx(c) { shim::check(rs_lwe_alloc(ctx, 1, &p), "rs_lwe_alloc", ctx); }
    ~Dev1() { rs_lwe_free(ctx, p); }
    void up(const Ctxt& c) { shim::check(rs_lwe_upload(ctx, p, c.lwe, 1), "rs_lwe_upload", ctx); }
    void down(Ctxt& c) { shim::check(rs_lwe_download(ctx, c.lwe, p, 1), "rs_lwe_download", ctx); }
};

void bootstrap1(Ctxt& io, uint32_t mu) {
    rs_ctx* ctx = redcufhe::CurrentContext();
    Dev1 d(ctx);
    d.up(io);
    shim::check(rs_pbs_batch(ctx, d.p, d.p, 1, mu), "rs_pbs_batch", ctx);
    d.down(io);
}
void gate1(int gate, Ctxt& out, const Ctxt& a, const Ctxt& b) {
    rs_ctx* ctx = redcufhe::CurrentContext();
    Dev1 da(ctx), db(ctx);
    da.up(a); db.up(b);
    shim::check(rs_gate_batch(ctx, gate, da.p, da.p, db.p, 1, kEighth), "rs_gate_batch", ctx);
    da.down(out);
}
void axpby1(Ctxt& out, const Ctxt& a, const Ctxt* b, uint32_t m0, uint32_t m1) {
    rs_ctx* ctx = redcufhe::CurrentContext();
    Dev1 da(ctx), db(ctx);
    da.up(a);
    if (b) db.up(*b);
    shim::check(rs_lwe_axpby(ctx, da.p, da.p, b ? db.p : nullptr, 1, m0, m1, 0u), "rs_lwe_axpby", ctx);
    da.down(out);
}

}  // namespace

namespace redcufhe {
void Not(Ctxt& out, const Ctxt& in, Stream) { axpby1(out, in, nullptr, 0xFFFFFFFFu, 0u); }
}

// ---------------------------------------------------------------------------------------------- gates.cuh
void CtxtCopyD2H(const Ctxt&, Stream) {}
void CtxtCopyH2D(const Ctxt&, Stream) {}
// encodings of the reference's CPU path (lib/BinOps_enc.cpp:182-192): binarize -> +-1/4096, unbinarize -> +-1/2048
void redsec_binarize_bootstrap(Ctxt& out, Stream) { bootstrap1(out, rs_modswitch_to_torus32(1, 4096)); }
void redsec_unbinarize_bootstrap(Ctxt& out, Stream) { bootstrap1(out, rs_modswitch_to_torus32(1, 2048)); }
void redsec_unbinarize_bootstrap_inv(Ctxt& out, Stream) { bootstrap1(out, rs_modswitch_to_torus32(-1, 2048)); }
void bootsNAND(Ctxt& out, const Ctxt& a, const Ctxt& b, Stream) { gate1(RS_GATE_NAND, out, a, b); }
void bootsOR(Ctxt& out, const Ctxt& a, const Ctxt& b, Stream) { gate1(RS_GATE_OR, out, a, b); }
void bootsAND(Ctxt& out, const Ctxt& a, const Ctxt& b, Stream) { gate1(RS_GATE_AND, out, a, b); }
void bootsNOR(Ctxt& out, const Ctxt& a, const Ctxt& b, Stream) { gate1(RS_GATE_NOR, out, a, b); }
void bootsXOR(Ctxt& out, const Ctxt& a, const Ctxt& b, Stream) { gate1(RS_GATE_XOR, out, a, b); }
void bootsXNOR(Ctxt& out, const Ctxt& a, const Ctxt& b, Stream) { gate1(RS_GATE_XNOR, out, a, b); }
void levelNOT(Ctxt& out, const Ctxt& in0, Stream st) { redcufhe::Not(out, in0, st); }
void NoiselessTrivial(Ctxt& result, Torus mu) {            // lib/GPU/gates.cu:146-151
    memset(result.lwe, 0, sizeof(uint32_t) * RS_LWE_N);
    result.lwe[RS_LWE_N] = mu;
}
void levelCONSTANT(Ctxt& result, int32_t value) {          // lib/GPU/gates.cu:153-156: +-1/8
    NoiselessTrivial(result, value ? kEighth : 0u - kEighth);
}
void add_int(Ctxt& sum, const Ctxt& a, const Ctxt& b, Stream) { axpby1(sum, a, &b, 1u, 1u); }
void sub_int(Ctxt& res, const Ctxt& a, const Ctxt& b, Stream) { axpby1(res, a, &b, 1u, 0xFFFFFFFFu); }
void mul_int(Ctxt& prod, const Ctxt& a, uint16_t b) { axpby1(prod, a, nullptr, (uint32_t)b, 0u); }
void bootstrapped_full_adder(Ctxt& sum, Ctxt& carry_out, Ctxt& temp_a, Ctxt& temp_b, const Ctxt& a, const Ctxt& b,
                             const Ctxt& carry_in, Stream st) {       // lib/GPU/gates.cu:204-244
    bootsXOR(temp_a, a, b, st);
    bootsXOR(sum, temp_a, carry_in, st);
    bootsAND(temp_a, carry_in, temp_a, st);
    bootsAND(temp_b, a, b, st);
    bootsOR(carry_out, temp_a, temp_b, st);
}

// ---------------------------------------------------------------------------------------------- BinOps
namespace BinOps {
void multiply(tBit* result, const tBit* a, const uint8_t b, Stream st) {      // XNOR with a plaintext bit: negate or copy
    if (b == 0) redcufhe::Not(*result, *a, st); else redcufhe::Copy(*result, *a, st);
}
void multiply_pc_ints(Ctxt& result, Ctxt& in1, const uint16_t* multicand, uint8_t, uint8_t, Stream) { mul_int(result, in1, *multicand); }
void add_bit(tMultiBit* result, const tBit* a, const tBit* b, Stream st) {
    result->ctxt = new Ctxt[2];
    result->size = 2;
    bootsXOR(result->ctxt[0], *a, *b, st);
    bootsAND(result->ctxt[1], *a, *b, st);
}
void add(tMultiBit* result, const tMultiBit* a, const tMultiBit* b, uint8_t bits, Stream st) {   // ripple-carry, 3*bits-1 bootstraps
    result->size = bits;
    result->ctxt = new Ctxt[bits];
    std::vector<Ctxt> carry(bits), aa(bits), bb(bits);
    Ctxt t0, t1;
    for (int i = 0; i < bits; i++) {
        if (i >= (int)a->size) levelCONSTANT(aa[i], 0); else redcufhe::Copy(aa[i], a->ctxt[i], st);
        if (i >= (int)b->size) levelCONSTANT(bb[i], 0); else redcufhe::Copy(bb[i], b->ctxt[i], st);
    }
    levelCONSTANT(carry[0], 0);
    for (int i = 0; i < bits - 1; i++) bootstrapped_full_adder(result->ctxt[i], carry[i + 1], t0, t1, aa[i], bb[i], carry[i], st);
    bootsXOR(t0, aa[bits - 1], bb[bits - 1], st);
    bootsXOR(result->ctxt[bits - 1], carry[bits - 1], t0, st);
}
void add_pc_ints(Ctxt& result, Ctxt& in1, const uint16_t* addend, uint8_t, Stream st) {
    Ctxt enc;
    NoiselessTrivial(enc, rs_modswitch_to_torus32((int32_t)(*addend & 0xFFFF), MSG_SPACE));   // plaintext addend in units of 1/4096
    add_int(result, in1, enc, st);
}
void int_add(Ctxt& result, const Ctxt& a, const Ctxt& b, Stream st) { add_int(result, a, b, st); }
void inc(tMultiBit* result, const tMultiBit* a, const tBit* b, Stream st) {
    std::vector<Ctxt> carry(a->size);
    result->size = a->size;
    result->ctxt = new Ctxt[result->size];
    redcufhe::Copy(carry[0], *b, st);
    for (uint32_t i = 0; i + 1 < a->size; i++) {
        bootsXOR(result->ctxt[i], carry[i], a->ctxt[i], st);
        bootsAND(carry[i + 1], carry[i], a->ctxt[i], st);
    }
    bootsXOR(result->ctxt[result->size - 1], carry[a->size - 1], a->ctxt[a->size - 1], st);
}
void max(tBit* result, const tBit* a, const tBit* b, Stream st) { bootsOR(*result, *a, *b, st); }
void shift(tMultiBit* result, tMultiBit* in1, uint8_t in_bits, uint8_t shift_bits, Stream st) {
    if (result->size != in_bits) { result->size = in_bits; result->ctxt = new Ctxt[in_bits]; }
    for (int i = 0; i < in_bits; i++)
        redcufhe::Copy(result->ctxt[i], (i + shift_bits) > (in_bits - 1) ? in1->ctxt[in_bits - 1] : in1->ctxt[i + shift_bits], st);
}
void relu(tFixedPoint* result, tMultiBit* in1, uint8_t in_bits, Stream st) {
    if (result->size != in_bits) { result->size = in_bits; result->ctxt = new Ctxt[in_bits]; }
    for (uint8_t i = 0; i + 1 < in_bits; i++) bootsAND(result->ctxt[i], in1->ctxt[i], in1->ctxt[in_bits - 1], st);
}
void binarize_int(Ctxt& result, Stream st) { redsec_binarize_bootstrap(result, st); }
void binarize(tBit* result, const tMultiBit* a) { redcufhe::Copy(*result, a->ctxt[a->size - 1]); }
void unbinarize_int(Ctxt& result, Stream st) { redsec_unbinarize_bootstrap(result, st); }
void unbinarize_int_inv(Ctxt& result, Stream st) { redsec_unbinarize_bootstrap_inv(result, st); }

// ---- weight-file readers (lib/GPU/BinOps_gpu.cu:191-331; format SURVEY.md 5.4)
void get_filters(FILE* fd_in, tBitPacked** p_filt_b, uint32_t len) {           // float weights -> trivial +-1/8 bits
    std::vector<float> f(len);
    if (fread(f.data(), sizeof(float), len, fd_in) != len) { printf("Bad Weights File. Exiting...\r\n"); return; }
    bit_calloc_global(p_filt_b, len);
    for (int g = 0; g < NUM_GPUS; g++)
        for (uint32_t j = 0; j < len; j++) levelCONSTANT((*p_filt_b)->enc_segs[g][j], f[j] < 0 ? 0 : 1);
}
void get_bitfilters(FILE* fd_in, tBitPacked** p_filt_b, uint32_t len) {        // packed bits, MSB first
    std::vector<uint8_t> pack((len + 7) / 8);
    if (fread(pack.data(), 1, pack.size(), fd_in) != pack.size()) { printf("Bad Weights File. Exiting...\r\n"); return; }
    bit_calloc_global(p_filt_b, len);
    for (int g = 0; g < NUM_GPUS; g++)
        for (uint32_t j = 0; j < len; j++) levelCONSTANT((*p_filt_b)->enc_segs[g][j], (pack[j >> 3] >> (7 - (j & 7))) & 1);
}
void get_intfilters(FILE* fd_in, tMultiBitPacked** p_filt_mb, uint32_t len) {  // int32 block -> trivial samples in units of 1/4096
    uint8_t tag = 0;
    std::vector<int32_t> v(len);
    if (fread(&tag, 1, 1, fd_in) != 1 || (tag != 3 && tag != 4) || fread(v.data(), 4, len, fd_in) != len) { printf("Bad Weights File. Exiting...\r\n"); return; }
    mbit_calloc_global(p_filt_mb, len, 1);
    for (int g = 0; g < NUM_GPUS; g++)
        for (uint32_t i = 0; i < len; i++) NoiselessTrivial((*p_filt_mb)->enc_segs[g][i].ctxt[0], rs_modswitch_to_torus32(v[i], MSG_SPACE));
}
void get_intfilters_ptxt(FILE* fd_in, uint16_t* p_filt_mb, uint32_t len) {     // (the reference reallocates its by-value argument; here the caller's array is filled)
    uint8_t tag = 0;
    std::vector<int32_t> v(len);
    if (fread(&tag, 1, 1, fd_in) != 1 || (tag != 3 && tag != 4) || fread(v.data(), 4, len, fd_in) != len) { printf("Bad Weights File. Exiting...\r\n"); return; }
    if (p_filt_mb) for (uint32_t i = 0; i < len; i++) p_filt_mb[i] = (uint16_t)(v[i] & 0xFFFF);
}
void get_ternfilters(FILE* fd_in, uint8_t* p_filt_b, uint8_t* p_tern, uint32_t len, float) {
    uint8_t tag = 0;
    if (fread(&tag, 1, 1, fd_in) != 1 || (tag != 1 && tag != 2)) { printf("Bad Weights File. Exiting...\r\n"); return; }
    const int nbits = tag == 1 ? 1 : 2;
    std::vector<uint8_t> pack(((size_t)len * nbits + 7) / 8);
    if (fread(pack.data(), 1, pack.size(), fd_in) != pack.size()) { printf("Bad Weights File. Exiting...\r\n"); return; }
    for (uint32_t i = 0; i < len; i++) {
        const size_t bit = (size_t)i * nbits;
        p_filt_b[i] = (pack[bit >> 3] >> (7 - (bit & 7))) & 1;
        if (p_tern) p_tern[i] = nbits == 2 ? (pack[(bit + 1) >> 3] >> (7 - ((bit + 1) & 7))) & 1 : 0;
    }
}
}  // namespace BinOps

// ---------------------------------------------------------------------------------------------- IntOps
namespace IntOps {
void add(tFixedPoint* result, const tFixedPoint* a, const tFixedPoint* b, Stream st) { add_int(result->ctxt[0], a->ctxt[0], b->ctxt[0], st); }
void subtract(tFixedPoint* result, const tFixedPoint* a, const tFixedPoint* b, Stream st) { sub_int(result->ctxt[0], a->ctxt[0], b->ctxt[0], st); }
void binarize(tBit* result, const tFixedPoint* a, Stream st) { redcufhe::Copy(*result, a->ctxt[a->size - 1], st); }
void binarize_int(tBit* result, Stream st) { redsec_binarize_bootstrap(*result, st); }
void invert(tFixedPoint* result, const tFixedPoint* a, const uint8_t* b, Stream st) {
    result->size = a->size;
    result->ctxt = new Ctxt[result->size];
    for (uint32_t i = 0; i < result->size; i++) {
        if (*b == 1) redcufhe::Copy(result->ctxt[i], a->ctxt[i], st); else redcufhe::Not(result->ctxt[i], a->ctxt[i], st);
    }
}
void multiply_pc_ints(Ctxt& result, Ctxt& in1, const uint16_t* multicand, uint8_t, uint8_t, Stream) { mul_int(result, in1, *multicand); }
void add_pc_ints(Ctxt& result, Ctxt& in1, const uint16_t* addend, uint8_t, Stream st) {
    Ctxt enc;
    NoiselessTrivial(enc, rs_modswitch_to_torus32((int32_t)(*addend & 0xFF), MSG_SPACE));
    add_int(result, in1, enc, st);
}
void relu(tFixedPoint* result, tFixedPoint* in1, uint8_t input_bits, Stream st) { BinOps::relu(result, in1, input_bits, st); }
void shift(tFixedPoint* result, tFixedPoint* in1, uint8_t input_bits, uint8_t shift_bits, Stream st) {
    for (int i = 0; i < input_bits; i++)
        redcufhe::Copy(result->ctxt[i], (i + shift_bits) > (input_bits - 1) ? in1->ctxt[input_bits - 1] : in1->ctxt[i + shift_bits], st);
}
}  // namespace IntOps
