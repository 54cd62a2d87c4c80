// dropin/src/ops_check.cpp -- exercises the Ops-level surface (gates.cuh, BinOps::, IntOps::) the way a maintainer's own code
// would: reads ../client/eval.key and ../client/ops_in.ctxt (5 ciphertexts: gate bits a, b, c at +-1/8 and integers x, y in
// units of 1/4096), applies one call of each kind, appends the results to ../client/ops_out.ctxt.  tests/test_gpu_dropin.py
// compares every output ciphertext with the CPU oracle.
#include <fstream>

#include "BinOps_gpu.cuh"
#include "IntOps_gpu.cuh"

using namespace redcufhe;

int main() {
    cudaSetDevice(0);
    PubKey bk;
    ReadPubKeyFromFile(bk, "../client/eval.key");
    Initialize(bk);
    std::ifstream in("../client/ops_in.ctxt");
    Ctxt a, b, c, x, y;
    ReadCtxtFromFileRed(a, in); ReadCtxtFromFileRed(b, in); ReadCtxtFromFileRed(c, in);
    ReadCtxtFromFileRed(x, in); ReadCtxtFromFileRed(y, in);
    Stream st;
    st.Create();
    std::vector<Ctxt> out;
    auto emit = [&](const Ctxt& v) { out.push_back(v); };
    Ctxt r, r2, t0, t1;
    bootsNAND(r, a, b, st); emit(r);
    bootsOR(r, a, b, st); emit(r);
    bootsAND(r, a, b, st); emit(r);
    bootsNOR(r, a, b, st); emit(r);
    bootsXOR(r, a, b, st); emit(r);
    bootsXNOR(r, a, b, st); emit(r);
    add_int(r, x, y, st); emit(r);
    sub_int(r, x, y, st); emit(r);
    mul_int(r, x, 3); emit(r);
    levelNOT(r, x, st); emit(r);
    BinOps::int_add(r, x, y, st); BinOps::binarize_int(r, st); emit(r);
    Copy(r, x, st); BinOps::unbinarize_int(r, st); emit(r);
    bootstrapped_full_adder(r, r2, t0, t1, a, b, c, st); emit(r); emit(r2);
    BinOps::max(&r, &a, &b, st); emit(r);
    BinOps::multiply(&r, &a, 0, st); emit(r);
    tFixedPoint fx{&x, 1, 0}, fy{&y, 1, 0}, fr{&r, 1, 0};
    IntOps::subtract(&fr, &fx, &fy, st); emit(r);
    uint16_t addend = 17;
    BinOps::add_pc_ints(r, x, &addend, 8, st); emit(r);
    st.Destroy();
    Synchronize();
    CuCheckError();
    for (auto& v : out) WriteCtxtToFileRed(v, "../client/ops_out.ctxt");
    printf("ops_check: %zu results\n", out.size());
    CleanUp();
    return 0;
}
