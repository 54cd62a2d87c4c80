/*
 * oracle/tfhe_oracle.h -- CPU restatement of the TFHE gate-bootstrap path that REDsec reduces to.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (redsec_b200/csrc) never links or calls anything in this directory.
 *
 * PARITY UNPINNED: the arithmetic lives in TFHE v1.1 (github.com/tfhe/tfhe, linked by the
 * reference as -ltfhe-spqlios-fma, lib/Makefile:3) which is NOT vendored under /root/reference and
 * is not installed; the reference holds no golden ciphertexts or key fixtures.
 * This file restates TFHE's published algorithm and anchors on the
 * reference's call sites:
 *   - tfhe_bootstrap_FFT(result, bk->bkFFT, mu, x)      lib/BinOps_enc.cpp:182-192
 *   - bootsOR / gate constants                          lib/BinOps_enc.cpp:164-167, lib/GPU/gates.cu:246-286
 *   - lweAddTo / lweSubTo / lweNoiselessTrivial         lib/BinOps_enc.cpp:121-141, lib/BinFunc.cpp:207-208
 *   - parameter set redsec_params_small_v2              client/gen_secure_keyset.cpp:70-92
 *   - encrypt / decrypt encodings                       client/encrypt_image.cpp:76-77, client/decrypt_image.cpp:52-58
 * The spec is "exact-integer TFHE semantics": the external product is the exact negacyclic
 * integer convolution mod 2^32.  Two implementations live here and must
 * agree bit-for-bit: an integer schoolbook one (orc_*_exact) and a double-precision FFT one
 * with round-to-nearest (orc_*_fft) that is also the timed CPU baseline.
 */
#ifndef TFHE_ORACLE_H
#define TFHE_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* reference keyset: client/gen_secure_keyset.cpp:70-92 (redsec_params_small_v2) */
#define ORC_n        350
#define ORC_N        1024
#define ORC_K        1
#define ORC_L        10
#define ORC_BGBIT    3
#define ORC_KS_T     9
#define ORC_KS_BASEBIT 3
#define ORC_KS_BASE  8
#define ORC_LWE_WORDS (ORC_n + 1)            /* a[0..n-1], b */
#define ORC_BSK_WORDS ((size_t)ORC_n * 2 * ORC_L * 2 * ORC_N)
#define ORC_KSK_WORDS ((size_t)ORC_N * ORC_KS_T * ORC_KS_BASE * ORC_LWE_WORDS)
#define ORC_BSK_FFT_DOUBLES ((size_t)ORC_n * 2 * ORC_L * 2 * ORC_N) /* N/2 complex per poly */

/* gate ids shared with the product C-ABI (include/redsec_b200.h) */
enum { ORC_GATE_NAND = 0, ORC_GATE_OR = 1, ORC_GATE_AND = 2, ORC_GATE_NOR = 3, ORC_GATE_XOR = 4, ORC_GATE_XNOR = 5 };

/* torus helpers (TFHE modSwitchToTorus32 / modSwitchFromTorus32) */
uint32_t orc_modswitch_to_torus32(int32_t mu, int32_t msize);
int32_t  orc_modswitch_from_torus32(uint32_t phase, int32_t msize);

/* keygen: deterministic from seed (spec in oracle/tfhe_oracle.c header comment) */
void orc_keygen(uint64_t seed, int32_t *lwe_key /*[n]*/, int32_t *tlwe_key /*[N]*/,
                uint32_t *bsk /*[n][2l][2][N]*/, uint32_t *ksk /*[N][t][base][n+1]*/);

/* lweSymEncrypt / phase / decrypt (client/encrypt_image.cpp:77, decrypt_image.cpp:52) */
void orc_lwe_encrypt(uint32_t *ct /*[count][n+1]*/, const uint32_t *mu /*[count]*/, int count, double alpha,
                     const int32_t *lwe_key, uint64_t seed);
void orc_lwe_phase(uint32_t *phase /*[count]*/, const uint32_t *ct, int count, const int32_t *lwe_key);
void orc_lwe_trivial(uint32_t *ct, uint32_t mu);

/* BSK -> double Fourier form used by the *_fft variant */
void orc_bsk_to_fft(const uint32_t *bsk, double *bsk_fft);

/* blind rotate: acc_out = 2 polys of N torus32 (a-poly then b-poly) */
void orc_blind_rotate_exact(uint32_t *acc_out, const uint32_t *lwe_in, uint32_t mu, const uint32_t *bsk);
/* err_stats (optional, may be NULL): [0] = max |x - rint(x)| over all inverse-FFT outputs, [1] = sum, [2] = count */
void orc_blind_rotate_fft(uint32_t *acc_out, const uint32_t *lwe_in, uint32_t mu, const double *bsk_fft, double *err_stats);

void orc_sample_extract(uint32_t *ext /*[N+1]*/, const uint32_t *acc /*[2][N]*/);
void orc_keyswitch(uint32_t *lwe_out /*[n+1]*/, const uint32_t *ext /*[N+1]*/, const uint32_t *ksk);

/* full programmable bootstrap = tfhe_bootstrap_FFT; exact!=0 selects the integer schoolbook external product */
void orc_pbs_batch(uint32_t *out, const uint32_t *in, int count, uint32_t mu,
                   const uint32_t *bsk, const double *bsk_fft, const uint32_t *ksk, int exact, int threads,
                   double *err_stats);

/* programmable bootstrap with per-ciphertext test vectors luts[lut_mod][N] (ciphertext c uses row c % lut_mod):
 * the correct encrypted form of the DoReFa ReLU of lib/IntFunc.cpp:934-973 (row f4) */
void orc_pbs_lut_batch(uint32_t *out, const uint32_t *in, int count, const uint32_t *luts, int lut_mod,
                       const uint32_t *bsk, const double *bsk_fft, const uint32_t *ksk, int exact, int threads);

/* gate linear part (lib/GPU/gates.cu:44-108): out = (0,fix) +/- in0 +/- in1 (x2 for XOR/XNOR) */
void orc_gate_linear(int op, uint32_t *out, const uint32_t *in0, const uint32_t *in1, int count);
void orc_gate_batch(int op, uint32_t *out, const uint32_t *in0, const uint32_t *in1, int count, uint32_t mu,
                    const uint32_t *bsk, const double *bsk_fft, const uint32_t *ksk, int exact, int threads);

/* ternary linear layer on LWE arrays: out[o] = (0,bias[o]) + sum_k sign[k]*in[col[k]], k in [rowptr[o],rowptr[o+1]) */
void orc_lwe_lincomb(uint32_t *out, int out_count, const uint32_t *in, const int32_t *rowptr, const int32_t *col,
                     const int8_t *sign, const uint32_t *bias);

int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
