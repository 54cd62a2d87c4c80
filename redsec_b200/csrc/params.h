// redsec_b200/csrc/params.h -- compile-time parameters of the reference keyset.
// Reference: client/gen_secure_keyset.cpp:70-92 (redsec_params_small_v2):
//   n=350, N=1024, k=1, bk_l=10, bk_Bgbit=3, ks_t=9, ks_basebit=3.
#pragma once
#include <cstdint>

namespace rs {
constexpr int LWE_N = 350;          // LWE dimension n
constexpr int LWE_WORDS = 351;      // wire format: a[0..n-1], b
constexpr int LWE_STRIDE = 352;     // device row stride in words (16-byte aligned rows; word 351 is zero padding)
constexpr int N = 1024;             // TLWE polynomial degree
constexpr int NH = 512;             // N/2 complex points per polynomial in Fourier form
constexpr int BK_L = 10;            // gadget levels
constexpr int BK_BGBIT = 3;         // log2 gadget base
constexpr int BK_ROWS = 2 * BK_L;   // (k+1)*l rows per TGSW sample
constexpr int KS_T = 9;
constexpr int KS_BASEBIT = 3;
constexpr int KS_BASE = 8;
constexpr int EXT_STRIDE = 1028;    // extracted LWE (dimension N): a[0..N-1], b, padding
constexpr uint32_t DECOMP_OFFSET =  // sum_{p=1..l} (Bg/2) << (32 - p*Bgbit)
    (4u << 29) + (4u << 26) + (4u << 23) + (4u << 20) + (4u << 17) + (4u << 14) + (4u << 11) + (4u << 8) + (4u << 5) + (4u << 2);
constexpr uint32_t KS_PREC_OFFSET = 1u << (32 - (1 + KS_BASEBIT * KS_T));

// Fourier-domain BSK: [n][BK_ROWS][2 output polys][NH] complex double; one (i,row) slab = 16 KiB
constexpr size_t BSK_ROW_BYTES = 2ull * NH * 16;
constexpr size_t BSK_F_BYTES = (size_t)LWE_N * BK_ROWS * BSK_ROW_BYTES;   // 114,688,000 B
constexpr size_t KSK_DEV_WORDS = (size_t)N * KS_T * KS_BASE * LWE_STRIDE; // padded rows
}  // namespace rs
