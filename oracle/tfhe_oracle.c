/*
 * oracle/tfhe_oracle.c -- CPU restatement of TFHE's gate bootstrap with REDsec's parameters.
 *
 * TEST INFRASTRUCTURE ONLY (see tfhe_oracle.h).  PARITY UNPINNED: TFHE v1.1 is an un-vendored
 * dependency of the reference; no upstream fixture exists in the tree.
 *
 * Algorithm sources (upstream TFHE v1.1, restated from its published algorithm):
 *   modSwitch{To,From}Torus32            -> orc_modswitch_*            (A.1)
 *   tfhe_bootstrap_FFT                   -> orc_pbs_batch              (A.2; called at lib/BinOps_enc.cpp:185,191)
 *   tGswTorus32PolynomialDecompH         -> decomp_digit()             (A.2 step 3)
 *   tLweExtractLweSample                 -> orc_sample_extract         (A.2 step 4)
 *   lweKeySwitch                         -> orc_keyswitch              (A.2 step 5)
 *   bootsNAND/OR/AND/NOR/XOR/XNOR        -> orc_gate_linear            (A.3; constants lib/GPU/gates.cu:246-286)
 *   lweSymEncrypt / lwePhase             -> orc_lwe_encrypt / _phase   (A.4; client/encrypt_image.cpp:77)
 *
 * Deterministic RNG spec (shared, independently re-implemented, by the product keygen in
 * redsec_b200/csrc/client.cpp so both sides can produce the identical keyset from a seed):
 *   splitmix64 seeds xoshiro256**; stream(seed, domain, index) starts splitmix64 at
 *   seed + 0x632BE59BD9B4E019*(domain+1) + 0xD1342543DE82EF95*index.
 *   u32 = next()>>32; bit = next()>>63; uniform01 = ((next()>>11)+1)*2^-53;
 *   gauss(sigma) = sigma*sqrt(-2 ln u1)*cos(2 pi u2); torus32 = dtot32(x) = (int32)(int64)((x-(int64)x)*2^32).
 *   domains: 1 lwe_key (n bits), 2 tlwe_key (N bits), 3 bsk row (index=i*2l+r: N u32 mask, then N gaussians
 *   sigma=2^-30), 4 ksk (index=i: for j<t, h<base: n u32 mask then 1 gaussian sigma=2^-25), 5 encryption
 *   (index=sample: n u32 mask then 1 gaussian).
 */
#include "tfhe_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define n_   ORC_n
#define N_   ORC_N
#define L_   ORC_L
#define NH   (ORC_N / 2)
#define ROWS (2 * ORC_L)

/* ------------------------------------------------------------------ RNG */
typedef struct { uint64_t s[4]; } rng_t;
static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t splitmix64(uint64_t *x) {
    uint64_t z = (*x += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static void rng_stream(rng_t *r, uint64_t seed, uint64_t domain, uint64_t index) {
    uint64_t x = seed + 0x632BE59BD9B4E019ULL * (domain + 1) + 0xD1342543DE82EF95ULL * index;
    for (int i = 0; i < 4; i++) r->s[i] = splitmix64(&x);
}
static inline uint64_t rng_next(rng_t *r) {
    uint64_t *s = r->s;
    uint64_t result = rotl64(s[1] * 5, 7) * 9;
    uint64_t t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl64(s[3], 45);
    return result;
}
static inline uint32_t rng_u32(rng_t *r) { return (uint32_t)(rng_next(r) >> 32); }
static inline int32_t rng_bit(rng_t *r) { return (int32_t)(rng_next(r) >> 63); }
static inline double rng_uniform01(rng_t *r) { return (double)((rng_next(r) >> 11) + 1) * 0x1.0p-53; }
static inline double rng_gauss(rng_t *r, double sigma) {
    double u1 = rng_uniform01(r), u2 = rng_uniform01(r);
    return sigma * sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925286766559 * u2);
}
static inline uint32_t dtot32(double d) { return (uint32_t)(int32_t)(int64_t)((d - (double)(int64_t)d) * 4294967296.0); }

/* ------------------------------------------------------------------ torus helpers (A.1) */
uint32_t orc_modswitch_to_torus32(int32_t mu, int32_t msize) {
    uint64_t interv = ((UINT64_C(1) << 63) / (uint64_t)msize) * 2;
    uint64_t phase64 = (uint64_t)(int64_t)mu * interv;
    return (uint32_t)(phase64 >> 32);
}
int32_t orc_modswitch_from_torus32(uint32_t phase, int32_t msize) {
    uint64_t interv = ((UINT64_C(1) << 63) / (uint64_t)msize) * 2;
    uint64_t half = interv / 2;
    uint64_t phase64 = ((uint64_t)phase << 32) + half;
    return (int32_t)(phase64 / interv);
}

/* ------------------------------------------------------------------ keygen (A.4) */
/* b[j] += sum_m s[m] * a[(j-m) mod^- N]   (negacyclic a*s with binary s) */
static void negacyclic_mul_binary_add(uint32_t *b, const uint32_t *a, const int32_t *s) {
    uint32_t *ext = (uint32_t *)malloc(2 * N_ * sizeof(uint32_t));
    for (int x = 0; x < N_; x++) { ext[x] = (uint32_t)(0u - a[x]); ext[x + N_] = a[x]; }
    for (int m = 0; m < N_; m++) {
        if (!s[m]) continue;
        const uint32_t *e = ext + (N_ - m);
        for (int j = 0; j < N_; j++) b[j] += e[j];
    }
    free(ext);
}

void orc_keygen(uint64_t seed, int32_t *lwe_key, int32_t *tlwe_key, uint32_t *bsk, uint32_t *ksk) {
    rng_t r;
    rng_stream(&r, seed, 1, 0);
    for (int i = 0; i < n_; i++) lwe_key[i] = rng_bit(&r);
    rng_stream(&r, seed, 2, 0);
    for (int i = 0; i < N_; i++) tlwe_key[i] = rng_bit(&r);
    const double bk_sigma = 0x1.0p-30, ks_sigma = 0x1.0p-25;
    #pragma omp parallel for schedule(dynamic, 8)
    for (int row = 0; row < n_ * ROWS; row++) {
        int i = row / ROWS, rr = row % ROWS, c = rr / L_, p = rr % L_;
        rng_t q; rng_stream(&q, seed, 3, (uint64_t)row);
        uint32_t *a = bsk + (size_t)row * 2 * N_, *b = a + N_;
        for (int j = 0; j < N_; j++) a[j] = rng_u32(&q);
        for (int j = 0; j < N_; j++) b[j] = dtot32(rng_gauss(&q, bk_sigma));
        negacyclic_mul_binary_add(b, a, tlwe_key);
        uint32_t h = (uint32_t)lwe_key[i] << (32 - (p + 1) * ORC_BGBIT);
        (c == 0 ? a : b)[0] += h;
    }
    #pragma omp parallel for schedule(dynamic, 8)
    for (int i = 0; i < N_; i++) {
        rng_t q; rng_stream(&q, seed, 4, (uint64_t)i);
        for (int j = 0; j < ORC_KS_T; j++)
            for (int h = 0; h < ORC_KS_BASE; h++) {
                uint32_t *ct = ksk + (((size_t)i * ORC_KS_T + j) * ORC_KS_BASE + h) * ORC_LWE_WORDS;
                uint32_t acc = 0;
                for (int x = 0; x < n_; x++) { ct[x] = rng_u32(&q); if (lwe_key[x]) acc += ct[x]; }
                uint32_t msg = ((uint32_t)(h * tlwe_key[i])) << (32 - (j + 1) * ORC_KS_BASEBIT);
                ct[n_] = acc + dtot32(rng_gauss(&q, ks_sigma)) + msg;
            }
    }
}

void orc_lwe_encrypt(uint32_t *ct, const uint32_t *mu, int count, double alpha, const int32_t *lwe_key, uint64_t seed) {
    for (int c = 0; c < count; c++) {
        rng_t q; rng_stream(&q, seed, 5, (uint64_t)c);
        uint32_t *s = ct + (size_t)c * ORC_LWE_WORDS, acc = 0;
        for (int x = 0; x < n_; x++) { s[x] = rng_u32(&q); if (lwe_key[x]) acc += s[x]; }
        s[n_] = acc + dtot32(rng_gauss(&q, alpha)) + mu[c];
    }
}
void orc_lwe_phase(uint32_t *phase, const uint32_t *ct, int count, const int32_t *lwe_key) {
    for (int c = 0; c < count; c++) {
        const uint32_t *s = ct + (size_t)c * ORC_LWE_WORDS; uint32_t acc = 0;
        for (int x = 0; x < n_; x++) if (lwe_key[x]) acc += s[x];
        phase[c] = s[n_] - acc;
    }
}
void orc_lwe_trivial(uint32_t *ct, uint32_t mu) { memset(ct, 0, n_ * sizeof(uint32_t)); ct[n_] = mu; }

/* ------------------------------------------------------------------ negacyclic FFT (N/2 complex, folded + twisted) */
/* Four polynomials at a time: every value is a vector of 4 doubles (GCC vector extension; AVX2 + FMA with -march=x86-64-v3),
 * lane = polynomial, so each butterfly is plain SIMD with a broadcast scalar twiddle and no shuffles.  The 20 digit rows of a
 * blind-rotate step are 5 such batches; the Fourier key is stored interleaved the same way ([i][batch][out poly][slot][re4,im4]).
 * 512 complex points = radix-4 DIF passes of span 512, 128, 32, 8 and one radix-2 pass; the output is in digit-reversed
 * order, which only has to match between the digits, the key (both through fft4_fwd) and the mirrored inverse fft4_inv.
 * This is what makes the port a fair CPU baseline: same operation count as TFHE's SPQLIOS-FMA transforms (vectorised,
 * precomputed twiddles, no allocation in the loop), 2.5-3x the speed of the scalar radix-2 code it replaces. */
typedef double v4d __attribute__((vector_size(32), aligned(32)));
typedef double v4du __attribute__((vector_size(32), aligned(8)));     /* the Fourier key lives in a caller-allocated buffer (numpy: 16-byte aligned) */
#define NGRP (ROWS / 4)                   /* batches of 4 rows per blind-rotate step */
static double tw_re[NH], tw_im[NH];       /* W_512^j = exp(-2 pi i j/512) */
static double twist_re[NH], twist_im[NH]; /* omega^j = exp(i pi j / 1024) */
static int fft_ready = 0;
static void fft_init(void) {
    if (fft_ready) return;
    #pragma omp critical(orc_fft_init)
    if (!fft_ready) {
        const long double PI = 3.14159265358979323846264338327950288L;
        for (int j = 0; j < NH; j++) {
            tw_re[j] = (double)cosl(-2.0L * PI * j / NH); tw_im[j] = (double)sinl(-2.0L * PI * j / NH);
            twist_re[j] = (double)cosl(PI * j / N_);      twist_im[j] = (double)sinl(PI * j / N_);
        }
        fft_ready = 1;
    }
}
static inline v4d bc(double x) { return (v4d){x, x, x, x}; }
/* forward: natural order in -> digit-reversed order out */
static void fft4_fwd(v4d *re, v4d *im) {
    for (int L = NH; L >= 8; L >>= 2) {
        const int q = L >> 2, stride = NH / L;
        for (int s = 0; s < NH; s += L) {
            v4d *r0 = re + s, *r1 = r0 + q, *r2 = r1 + q, *r3 = r2 + q;
            v4d *i0 = im + s, *i1 = i0 + q, *i2 = i1 + q, *i3 = i2 + q;
            for (int j = 0; j < q; j++) {
                const v4d t0r = r0[j] + r2[j], t0i = i0[j] + i2[j], t1r = r0[j] - r2[j], t1i = i0[j] - i2[j];
                const v4d t2r = r1[j] + r3[j], t2i = i1[j] + i3[j], t3r = r1[j] - r3[j], t3i = i1[j] - i3[j];
                r0[j] = t0r + t2r; i0[j] = t0i + t2i;
                const v4d y2r = t0r - t2r, y2i = t0i - t2i;
                const v4d y1r = t1r + t3i, y1i = t1i - t3r;           /* t1 - i*t3 */
                const v4d y3r = t1r - t3i, y3i = t1i + t3r;           /* t1 + i*t3 */
                const v4d w1r = bc(tw_re[j * stride]), w1i = bc(tw_im[j * stride]);
                const v4d w2r = bc(tw_re[2 * j * stride]), w2i = bc(tw_im[2 * j * stride]);
                const v4d w3r = bc(tw_re[3 * j * stride]), w3i = bc(tw_im[3 * j * stride]);
                r1[j] = y1r * w1r - y1i * w1i; i1[j] = y1r * w1i + y1i * w1r;
                r2[j] = y2r * w2r - y2i * w2i; i2[j] = y2r * w2i + y2i * w2r;
                r3[j] = y3r * w3r - y3i * w3i; i3[j] = y3r * w3i + y3i * w3r;
            }
        }
    }
    for (int s = 0; s < NH; s += 2) {
        const v4d ar = re[s], ai = im[s], br = re[s + 1], bi = im[s + 1];
        re[s] = ar + br; im[s] = ai + bi; re[s + 1] = ar - br; im[s + 1] = ai - bi;
    }
}
/* inverse (unscaled, x512): digit-reversed in -> natural out; the exact mirror of fft4_fwd */
static void fft4_inv(v4d *re, v4d *im) {
    for (int s = 0; s < NH; s += 2) {
        const v4d ar = re[s], ai = im[s], br = re[s + 1], bi = im[s + 1];
        re[s] = ar + br; im[s] = ai + bi; re[s + 1] = ar - br; im[s + 1] = ai - bi;
    }
    for (int L = 8; L <= NH; L <<= 2) {
        const int q = L >> 2, stride = NH / L;
        for (int s = 0; s < NH; s += L) {
            v4d *r0 = re + s, *r1 = r0 + q, *r2 = r1 + q, *r3 = r2 + q;
            v4d *i0 = im + s, *i1 = i0 + q, *i2 = i1 + q, *i3 = i2 + q;
            for (int j = 0; j < q; j++) {
                const v4d w1r = bc(tw_re[j * stride]), w1i = bc(-tw_im[j * stride]);            /* conjugate twiddles */
                const v4d w2r = bc(tw_re[2 * j * stride]), w2i = bc(-tw_im[2 * j * stride]);
                const v4d w3r = bc(tw_re[3 * j * stride]), w3i = bc(-tw_im[3 * j * stride]);
                const v4d u0r = r0[j], u0i = i0[j];
                const v4d u1r = r1[j] * w1r - i1[j] * w1i, u1i = r1[j] * w1i + i1[j] * w1r;
                const v4d u2r = r2[j] * w2r - i2[j] * w2i, u2i = r2[j] * w2i + i2[j] * w2r;
                const v4d u3r = r3[j] * w3r - i3[j] * w3i, u3i = r3[j] * w3i + i3[j] * w3r;
                const v4d e0r = u0r + u2r, e0i = u0i + u2i, e1r = u0r - u2r, e1i = u0i - u2i;
                const v4d o0r = u1r + u3r, o0i = u1i + u3i, o1r = u1r - u3r, o1i = u1i - u3i;
                r0[j] = e0r + o0r; i0[j] = e0i + o0i;          /* a = u0 + u1 + u2 + u3 */
                r2[j] = e0r - o0r; i2[j] = e0i - o0i;          /* c = u0 - u1 + u2 - u3 */
                r1[j] = e1r - o1i; i1[j] = e1i + o1r;          /* b = u0 + i*u1 - u2 - i*u3 */
                r3[j] = e1r + o1i; i3[j] = e1i - o1r;          /* d = u0 - i*u1 - u2 + i*u3 */
            }
        }
    }
}

/* bsk_fft layout (doubles): [i][batch g][out poly co][slot k][re4, im4], lane l of batch g = TGSW row 4g+l */
static inline size_t bskf_index(int i, int g, int co) { return ((((size_t)i * NGRP + g) * 2 + co) * NH) * 8; }
void orc_bsk_to_fft(const uint32_t *bsk, double *bsk_fft) {
    fft_init();
    #pragma omp parallel for schedule(static)
    for (int ig = 0; ig < n_ * NGRP; ig++) {
        const int i = ig / NGRP, g = ig % NGRP;
        v4d re[NH], im[NH];
        for (int co = 0; co < 2; co++) {
            for (int j = 0; j < NH; j++) {
                double x[4], y[4];
                for (int l = 0; l < 4; l++) {
                    const int32_t *p = (const int32_t *)bsk + (((size_t)i * ROWS + 4 * g + l) * 2 + co) * N_;
                    x[l] = (double)p[j]; y[l] = (double)p[j + NH];
                }
                const v4d xv = {x[0], x[1], x[2], x[3]}, yv = {y[0], y[1], y[2], y[3]};
                re[j] = xv * bc(twist_re[j]) - yv * bc(twist_im[j]);
                im[j] = xv * bc(twist_im[j]) + yv * bc(twist_re[j]);
            }
            fft4_fwd(re, im);
            v4du *dst = (v4du *)(bsk_fft + bskf_index(i, g, co));
            for (int k = 0; k < NH; k++) { dst[2 * k] = re[k]; dst[2 * k + 1] = im[k]; }
        }
    }
}

/* ------------------------------------------------------------------ blind rotate (A.2 steps 1-3) */
/* result = X^a * src over Z[X]/(X^N+1), a in [0,2N) */
static void poly_mul_by_xai(uint32_t *res, int a, const uint32_t *src) {
    if (a < N_) {
        for (int j = 0; j < a; j++) res[j] = 0u - src[j - a + N_];
        for (int j = a; j < N_; j++) res[j] = src[j - a];
    } else {
        int aa = a - N_;
        for (int j = 0; j < aa; j++) res[j] = src[j - aa + N_];
        for (int j = aa; j < N_; j++) res[j] = 0u - src[j - aa];
    }
}
static inline uint32_t decomp_offset(void) {
    uint32_t off = 0;
    for (int p = 1; p <= L_; p++) off += (uint32_t)(ORC_KS_BASE / 2) << (32 - p * ORC_BGBIT); /* Bg/2 = 4 */
    return off;
}
static inline int32_t decomp_digit(uint32_t v_plus_offset, int p /*0-based*/) {
    int decal = 32 - (p + 1) * ORC_BGBIT;
    return (int32_t)((v_plus_offset >> decal) & 7u) - 4;
}
/* tv == NULL: the constant test vector mu*(1+X+...+X^(N-1)) of tfhe_bootstrap_FFT; otherwise a caller-supplied
 * test vector (programmable bootstrap: coefficient j is the output for a phase that rounds to j/2N) */
static void init_acc(uint32_t *acc, int *bara, const uint32_t *lwe_in, uint32_t mu, const uint32_t *tv_in) {
    int barb = orc_modswitch_from_torus32(lwe_in[n_], 2 * N_);
    for (int i = 0; i < n_; i++) bara[i] = orc_modswitch_from_torus32(lwe_in[i], 2 * N_) % (2 * N_);
    barb %= 2 * N_;
    uint32_t tv[N_];
    for (int j = 0; j < N_; j++) tv[j] = tv_in ? tv_in[j] : mu;
    memset(acc, 0, N_ * sizeof(uint32_t));
    poly_mul_by_xai(acc + N_, (2 * N_ - barb) % (2 * N_), tv);
}

static void blind_rotate_exact_tv(uint32_t *acc, const uint32_t *lwe_in, uint32_t mu, const uint32_t *tv, const uint32_t *bsk) {
    int bara[n_];
    init_acc(acc, bara, lwe_in, mu, tv);
    const uint32_t off = decomp_offset();
    uint32_t tmp[2 * N_], ext[2 * N_], res[2 * N_];
    for (int i = 0; i < n_; i++) {
        int a = bara[i];
        if (a == 0) continue;
        for (int c = 0; c < 2; c++) {
            poly_mul_by_xai(tmp + c * N_, a, acc + c * N_);
            for (int j = 0; j < N_; j++) tmp[c * N_ + j] -= acc[c * N_ + j];
        }
        memset(res, 0, 2 * N_ * sizeof(uint32_t));
        for (int c = 0; c < 2; c++)
            for (int p = 0; p < L_; p++) {
                const uint32_t *row = bsk + ((size_t)i * ROWS + c * L_ + p) * 2 * N_;
                for (int co = 0; co < 2; co++) {
                    const uint32_t *B = row + co * N_;
                    for (int x = 0; x < N_; x++) { ext[x] = 0u - B[x]; ext[x + N_] = B[x]; }
                    uint32_t *out = res + co * N_;
                    for (int m = 0; m < N_; m++) {
                        uint32_t d = (uint32_t)decomp_digit(tmp[c * N_ + m] + off, p);
                        if (!d) continue;
                        const uint32_t *e = ext + (N_ - m);
                        for (int j = 0; j < N_; j++) out[j] += d * e[j];
                    }
                }
            }
        for (int j = 0; j < 2 * N_; j++) acc[j] += res[j];
    }
}

void orc_blind_rotate_exact(uint32_t *acc, const uint32_t *lwe_in, uint32_t mu, const uint32_t *bsk) {
    blind_rotate_exact_tv(acc, lwe_in, mu, NULL, bsk);
}
static void blind_rotate_fft_tv(uint32_t *acc, const uint32_t *lwe_in, uint32_t mu, const uint32_t *tv, const double *bsk_fft,
                                double *err_stats);
void orc_blind_rotate_fft(uint32_t *acc, const uint32_t *lwe_in, uint32_t mu, const double *bsk_fft, double *err_stats) {
    blind_rotate_fft_tv(acc, lwe_in, mu, NULL, bsk_fft, err_stats);
}
static void blind_rotate_fft_tv(uint32_t *acc, const uint32_t *lwe_in, uint32_t mu, const uint32_t *tv, const double *bsk_fft,
                                double *err_stats) {
    fft_init();
    int bara[n_];
    init_acc(acc, bara, lwe_in, mu, tv);
    const uint32_t off = decomp_offset();
    uint32_t tmp[2 * N_];
    v4d fr[NH], fi[NH], accr[2][NH], acci[2][NH];      /* all on the stack: no allocation inside the loop */
    double emax = 0, esum = 0, ecnt = 0;
    for (int i = 0; i < n_; i++) {
        int a = bara[i];
        if (a == 0) continue;
        for (int c = 0; c < 2; c++) {
            poly_mul_by_xai(tmp + c * N_, a, acc + c * N_);
            for (int j = 0; j < N_; j++) tmp[c * N_ + j] = tmp[c * N_ + j] - acc[c * N_ + j] + off;
        }
        memset(accr, 0, sizeof(accr)); memset(acci, 0, sizeof(acci));
        for (int g = 0; g < NGRP; g++) {
            /* rows 4g..4g+3 (row r = c*l + p): gadget digits of the four rows in the four lanes, folded and twisted */
            int cc[4], sh[4];
            for (int l = 0; l < 4; l++) { const int r = 4 * g + l; cc[l] = r / L_; sh[l] = 32 - ((r % L_) + 1) * ORC_BGBIT; }
            for (int j = 0; j < NH; j++) {
                double x[4], y[4];
                for (int l = 0; l < 4; l++) {
                    const uint32_t *t = tmp + cc[l] * N_;
                    x[l] = (double)((int32_t)((t[j] >> sh[l]) & 7u) - 4);
                    y[l] = (double)((int32_t)((t[j + NH] >> sh[l]) & 7u) - 4);
                }
                const v4d xv = {x[0], x[1], x[2], x[3]}, yv = {y[0], y[1], y[2], y[3]};
                fr[j] = xv * bc(twist_re[j]) - yv * bc(twist_im[j]);
                fi[j] = xv * bc(twist_im[j]) + yv * bc(twist_re[j]);
            }
            fft4_fwd(fr, fi);
            for (int co = 0; co < 2; co++) {
                const v4du *B = (const v4du *)(bsk_fft + bskf_index(i, g, co));
                v4d *ar = accr[co], *ai = acci[co];
                for (int k = 0; k < NH; k++) {
                    const v4d br = B[2 * k], bi = B[2 * k + 1];
                    ar[k] += fr[k] * br - fi[k] * bi;
                    ai[k] += fr[k] * bi + fi[k] * br;
                }
            }
        }
        /* lane sums (the four rows of every batch) -> lanes 0,1 of one inverse transform = the two output polynomials */
        for (int k = 0; k < NH; k++) {
            const v4d r0 = accr[0][k], i0 = acci[0][k], r1 = accr[1][k], i1 = acci[1][k];
            fr[k] = (v4d){r0[0] + r0[1] + r0[2] + r0[3], r1[0] + r1[1] + r1[2] + r1[3], 0.0, 0.0};
            fi[k] = (v4d){i0[0] + i0[1] + i0[2] + i0[3], i1[0] + i1[1] + i1[2] + i1[3], 0.0, 0.0};
        }
        fft4_inv(fr, fi);
        for (int co = 0; co < 2; co++)
            for (int j = 0; j < NH; j++) {
                double zr = fr[j][co] * (1.0 / NH), zi = fi[j][co] * (1.0 / NH);
                double x = zr * twist_re[j] + zi * twist_im[j];   /* z * conj(omega^j) */
                double y = zi * twist_re[j] - zr * twist_im[j];
                double rx = nearbyint(x), ry = nearbyint(y);
                if (err_stats) {
                    double ex = fabs(x - rx), ey = fabs(y - ry);
                    if (ex > emax) emax = ex; if (ey > emax) emax = ey;
                    esum += ex + ey; ecnt += 2;
                }
                acc[co * N_ + j] += (uint32_t)(int64_t)rx;
                acc[co * N_ + j + NH] += (uint32_t)(int64_t)ry;
            }
    }
    if (err_stats) { if (emax > err_stats[0]) err_stats[0] = emax; err_stats[1] += esum; err_stats[2] += ecnt; }
}

/* ------------------------------------------------------------------ extract + keyswitch (A.2 steps 4-5) */
void orc_sample_extract(uint32_t *ext, const uint32_t *acc) {
    ext[0] = acc[0];
    for (int j = 1; j < N_; j++) ext[j] = 0u - acc[N_ - j];
    ext[N_] = acc[N_];
}
void orc_keyswitch(uint32_t *out, const uint32_t *ext, const uint32_t *ksk) {
    const uint32_t prec_offset = 1u << (32 - (1 + ORC_KS_BASEBIT * ORC_KS_T));
    memset(out, 0, n_ * sizeof(uint32_t));
    out[n_] = ext[N_];
    for (int i = 0; i < N_; i++) {
        uint32_t aibar = ext[i] + prec_offset;
        for (int j = 0; j < ORC_KS_T; j++) {
            uint32_t aij = (aibar >> (32 - (j + 1) * ORC_KS_BASEBIT)) & (ORC_KS_BASE - 1);
            if (!aij) continue;
            const uint32_t *row = ksk + (((size_t)i * ORC_KS_T + j) * ORC_KS_BASE + aij) * ORC_LWE_WORDS;
            for (int x = 0; x < ORC_LWE_WORDS; x++) out[x] -= row[x];
        }
    }
}

void orc_pbs_batch(uint32_t *out, const uint32_t *in, int count, uint32_t mu, const uint32_t *bsk, const double *bsk_fft,
                   const uint32_t *ksk, int exact, int threads, double *err_stats) {
    fft_init();
    double emax = 0, esum = 0, ecnt = 0;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#endif
    /* loop shape follows lib/BinFunc.cpp:1056-1071: omp parallel for over neurons, one bootstrap per iteration */
    #pragma omp parallel for schedule(dynamic, 1) num_threads(threads) reduction(max : emax) reduction(+ : esum, ecnt)
    for (int c = 0; c < count; c++) {
        uint32_t acc[2 * N_], ext[N_ + 1], res[ORC_LWE_WORDS];
        double st[3] = {0, 0, 0};
        if (exact) orc_blind_rotate_exact(acc, in + (size_t)c * ORC_LWE_WORDS, mu, bsk);
        else orc_blind_rotate_fft(acc, in + (size_t)c * ORC_LWE_WORDS, mu, bsk_fft, err_stats ? st : NULL);
        orc_sample_extract(ext, acc);
        orc_keyswitch(res, ext, ksk);
        memcpy(out + (size_t)c * ORC_LWE_WORDS, res, sizeof(res));
        if (st[0] > emax) emax = st[0]; esum += st[1]; ecnt += st[2];
    }
    if (err_stats) { if (emax > err_stats[0]) err_stats[0] = emax; err_stats[1] += esum; err_stats[2] += ecnt; }
}

/* programmable bootstrap with caller-supplied test vectors: ciphertext c uses luts[(c % lut_mod)][N]; the output is
 * LWE(tv[j]) for a phase that rounds to j/2N in [0,1/2) and LWE(-tv[j-N]) in [1/2,1).  This is the correct encrypted
 * form of the DoReFa ReLU staircase of lib/IntFunc.cpp:934-973 (defect R6). */
void orc_pbs_lut_batch(uint32_t *out, const uint32_t *in, int count, const uint32_t *luts, int lut_mod, const uint32_t *bsk,
                       const double *bsk_fft, const uint32_t *ksk, int exact, int threads) {
    fft_init();
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#endif
    #pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
    for (int c = 0; c < count; c++) {
        uint32_t acc[2 * N_], ext[N_ + 1], res[ORC_LWE_WORDS];
        const uint32_t *tv = luts + (size_t)(c % lut_mod) * N_;
        if (exact) blind_rotate_exact_tv(acc, in + (size_t)c * ORC_LWE_WORDS, 0, tv, bsk);
        else blind_rotate_fft_tv(acc, in + (size_t)c * ORC_LWE_WORDS, 0, tv, bsk_fft, NULL);
        orc_sample_extract(ext, acc);
        orc_keyswitch(res, ext, ksk);
        memcpy(out + (size_t)c * ORC_LWE_WORDS, res, sizeof(res));
    }
}

/* ------------------------------------------------------------------ gates (A.3) */
void orc_gate_linear(int op, uint32_t *out, const uint32_t *in0, const uint32_t *in1, int count) {
    /* fix on b; sign s and factor f:  out = (0,fix) + s*f*(in0+in1) */
    static const int32_t fix8[6] = {1, 1, -1, -1, 2, -2}; /* in eighths: XOR +1/4, XNOR -1/4 */
    static const int sgn[6] = {-1, 1, 1, -1, 1, -1};
    static const int fac[6] = {1, 1, 1, 1, 2, 2};
    uint32_t fix = orc_modswitch_to_torus32(fix8[op], 8);
    uint32_t m = (uint32_t)(sgn[op] * fac[op]);
    for (int c = 0; c < count; c++) {
        const uint32_t *a = in0 + (size_t)c * ORC_LWE_WORDS, *b = in1 + (size_t)c * ORC_LWE_WORDS;
        uint32_t *o = out + (size_t)c * ORC_LWE_WORDS;
        for (int x = 0; x < ORC_LWE_WORDS; x++) o[x] = m * (a[x] + b[x]);
        o[n_] += fix;
    }
}
void orc_gate_batch(int op, uint32_t *out, const uint32_t *in0, const uint32_t *in1, int count, uint32_t mu,
                    const uint32_t *bsk, const double *bsk_fft, const uint32_t *ksk, int exact, int threads) {
    uint32_t *lin = (uint32_t *)malloc((size_t)count * ORC_LWE_WORDS * sizeof(uint32_t));
    orc_gate_linear(op, lin, in0, in1, count);
    orc_pbs_batch(out, lin, count, mu, bsk, bsk_fft, ksk, exact, threads, NULL);
    free(lin);
}

void orc_lwe_lincomb(uint32_t *out, int out_count, const uint32_t *in, const int32_t *rowptr, const int32_t *col,
                     const int8_t *sign, const uint32_t *bias) {
    #pragma omp parallel for schedule(static)
    for (int o = 0; o < out_count; o++) {
        uint32_t *dst = out + (size_t)o * ORC_LWE_WORDS;
        memset(dst, 0, ORC_LWE_WORDS * sizeof(uint32_t));
        if (bias) dst[n_] = bias[o];
        for (int k = rowptr[o]; k < rowptr[o + 1]; k++) {
            const uint32_t *src = in + (size_t)col[k] * ORC_LWE_WORDS;
            uint32_t s = (uint32_t)(int32_t)sign[k];
            for (int x = 0; x < ORC_LWE_WORDS; x++) dst[x] += s * src[x];
        }
    }
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
