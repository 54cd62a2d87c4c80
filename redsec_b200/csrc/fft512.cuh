// redsec_b200/csrc/fft512.cuh -- negacyclic size-1024 real transform as a folded, twisted 512-point
// complex FP64 FFT, held in registers by a 64-thread group (8 complex points per thread, three
// radix-8 passes, two shared-memory exchanges).
//
// Replaces (from scratch) the role of TFHE's spqlios FFT inside tfhe_bootstrap_FFT, which the
// reference calls at lib/BinOps_enc.cpp:185,191; and of redcufhe's NTT (lib/GPU/gates.cuh:7).
//
// Index algebra (forward, W = exp(-2*pi*i/512), omega = exp(i*pi/1024)):
//   z[j] = (p[j] + i*p[j+512]) * omega^j,  Z[k] = sum_j z[j] W^{jk} = p(omega * W^k)   (a root of X^1024+1)
//   j = t + 64q,  k = r + 8*r2 + 64*r3
//   pass 1 (thread t):            a_r[t]     = omega^t W^{t r} * DFT8_q( z[t+64q] * omega^{64q} )[r]
//   pass 2 (thread u=t2+8r):      b_rr2[t2]  = W64^{t2 r2} * DFT8_q2( a_r[t2+8q2] )[r2]
//   pass 3 (thread v=r2+8r):      Z[r+8r2+64r3] = DFT8_t2( b_rr2[t2] )[r3]
// The inverse is the exact mirror (conjugate constants, reversed exchanges, scale 1/512).
// Output "slot" layout of the forward transform: thread v holds slots r3*64+v, r3=0..7; the
// Fourier-domain bootstrapping key is produced by the same routine so layouts agree by construction.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace rs {

// COMPACT_H: keep only h[1], h[2], h[4] in registers and rebuild h[3,5,6,7] with 4 complex products per
// transform (+16 DFMA, -16 registers); used by the 6-group variant whose budget is 168 registers/thread.
template <bool COMPACT_H>
struct TwiddlesT {
    double2 g[8];   // g[r]  = exp(i*pi*t*(1-4r)/1024)    (twist merged with pass-1 twiddle)
    double2 h[8];   // h[r2] = exp(-2*pi*i*t2*r2/64), t2 = t & 7   (h[0] = 1 unused)
};
template <>
struct TwiddlesT<true> {
    double2 g[8];
    double2 h1, h2, h4;
};
using Twiddles = TwiddlesT<false>;

__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cmul_conj(double2 a, double2 b) {  // a * conj(b)
    return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
// multiply by i*S (S = +1 or -1)
template <int S>
__device__ __forceinline__ double2 mul_i(double2 a) {
    return S > 0 ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x);
}

__device__ __forceinline__ void group_sync(int group) {  // named barrier over the 64 threads of one group
    asm volatile("bar.sync %0, 64;" ::"r"(group + 1) : "memory");
}

__device__ __forceinline__ double2 unit_root(double turns_times_two) {   // exp(i*pi*x)
    double s, c;
    sincospi(turns_times_two, &s, &c);
    return make_double2(c, s);
}
__device__ __forceinline__ void make_twiddles(TwiddlesT<false>& tw, int t) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
        tw.g[r] = unit_root((double)(t * (1 - 4 * r)) / 1024.0);
        tw.h[r] = unit_root(-(double)((t & 7) * r) / 32.0);
    }
}
__device__ __forceinline__ void make_twiddles(TwiddlesT<true>& tw, int t) {
#pragma unroll
    for (int r = 0; r < 8; r++) tw.g[r] = unit_root((double)(t * (1 - 4 * r)) / 1024.0);
    tw.h1 = unit_root(-(double)(t & 7) / 32.0);
    tw.h2 = unit_root(-(double)((t & 7) * 2) / 32.0);
    tw.h4 = unit_root(-(double)((t & 7) * 4) / 32.0);
}
__device__ __forceinline__ void expand_h(double2 (&h)[8], const TwiddlesT<false>& tw) {
#pragma unroll
    for (int r = 1; r < 8; r++) h[r] = tw.h[r];
}
__device__ __forceinline__ void expand_h(double2 (&h)[8], const TwiddlesT<true>& tw) {
    h[1] = tw.h1; h[2] = tw.h2; h[4] = tw.h4;
    h[3] = cmul(tw.h1, tw.h2); h[5] = cmul(tw.h1, tw.h4); h[6] = cmul(tw.h2, tw.h4); h[7] = cmul(h[3], tw.h4);
}

// 8-point DFT in registers, W8 = exp(S*2*pi*i/8); natural order in, natural order out.
template <int S>
__device__ __forceinline__ void dft8(double2 (&v)[8]) {
    const double hs = 0.70710678118654752440;
    double2 a0 = cadd(v[0], v[4]), a1 = cadd(v[1], v[5]), a2 = cadd(v[2], v[6]), a3 = cadd(v[3], v[7]);
    double2 b0 = csub(v[0], v[4]), b1 = csub(v[1], v[5]), b2 = csub(v[2], v[6]), b3 = csub(v[3], v[7]);
    // b1 *= (hs, S*hs); b2 *= i*S; b3 *= (-hs, S*hs)
    double2 u1 = make_double2(b1.x - S * b1.y, S * b1.x + b1.y);     // (1 + iS) * b1
    double2 u3 = make_double2(-b3.x - S * b3.y, S * b3.x - b3.y);    // (-1 + iS) * b3
    b2 = mul_i<S>(b2);
    // even outputs: DFT4(a)
    double2 e0 = cadd(a0, a2), e1 = csub(a0, a2), o0 = cadd(a1, a3), o1 = mul_i<S>(csub(a1, a3));
    v[0] = cadd(e0, o0); v[4] = csub(e0, o0); v[2] = cadd(e1, o1); v[6] = csub(e1, o1);
    // odd outputs: DFT4(b0, hs*u1, b2, hs*u3)
    double2 f0 = cadd(b0, b2), f1 = csub(b0, b2);
    double2 s = cadd(u1, u3), d = mul_i<S>(csub(u1, u3));
    v[1] = make_double2(f0.x + hs * s.x, f0.y + hs * s.y);
    v[5] = make_double2(f0.x - hs * s.x, f0.y - hs * s.y);
    v[3] = make_double2(f1.x + hs * d.x, f1.y + hs * d.y);
    v[7] = make_double2(f1.x - hs * d.x, f1.y - hs * d.y);
}

// omega^{64q} = exp(i*pi*q/16)
__device__ __forceinline__ double2 twist_const(int q) {
    const double c[8] = {1.0, 0.98078528040323044913, 0.92387953251128675613, 0.83146961230254523708,
                         0.70710678118654752440, 0.55557023301960222474, 0.38268343236508977173, 0.19509032201612826785};
    const double s[8] = {0.0, 0.19509032201612826785, 0.38268343236508977173, 0.55557023301960222474,
                         0.70710678118654752440, 0.83146961230254523708, 0.92387953251128675613, 0.98078528040323044913};
    return make_double2(c[q], s[q]);
}

constexpr int FFT_BUF1 = 512;          // complex entries, exchange 1: [r][t]
constexpr int FFT_BUF2 = 576;          // complex entries, exchange 2: [r][r2*9 + t2] (padded, conflict-free)

// Forward transform.  In: v[q] = p[t+64q] + i*p[t+64q+512].  Out: v[r3] = Z[slot r3*64 + tid].
template <class TW>
__device__ __forceinline__ void fft512_fwd(double2 (&v)[8], const TW& tw, double2* buf1, double2* buf2,
                                           int t, int group) {
#pragma unroll
    for (int q = 1; q < 8; q++) v[q] = cmul(v[q], twist_const(q));
    dft8<-1>(v);
#pragma unroll
    for (int r = 0; r < 8; r++) buf1[r * 64 + t] = cmul(v[r], tw.g[r]);
    group_sync(group);
    const int t2 = t & 7, rr = t >> 3;
#pragma unroll
    for (int q2 = 0; q2 < 8; q2++) v[q2] = buf1[rr * 64 + t2 + 8 * q2];
    dft8<-1>(v);
    buf2[rr * 72 + t2] = v[0];
    {
        double2 h[8];
        expand_h(h, tw);
#pragma unroll
        for (int r2 = 1; r2 < 8; r2++) buf2[rr * 72 + r2 * 9 + t2] = cmul(v[r2], h[r2]);
    }
    group_sync(group);
    // thread v = r2' + 8*rr with r2' = t2 (same lane bits, different meaning)
#pragma unroll
    for (int x = 0; x < 8; x++) v[x] = buf2[rr * 72 + t2 * 9 + x];
    dft8<-1>(v);
}

// Inverse transform (scaled by 1/512).  In: v[r3] = A[slot r3*64 + tid].
// Out: v[q] = z[t+64q] with Re -> coefficient t+64q, Im -> coefficient t+64q+512.
template <class TW>
__device__ __forceinline__ void fft512_inv(double2 (&v)[8], const TW& tw, double2* buf1, double2* buf2,
                                           int t, int group) {
    const int t2 = t & 7, rr = t >> 3;
    dft8<+1>(v);
#pragma unroll
    for (int x = 0; x < 8; x++) buf2[rr * 72 + t2 * 9 + x] = v[x];
    group_sync(group);
    v[0] = buf2[rr * 72 + t2];
    {
        double2 h[8];
        expand_h(h, tw);
#pragma unroll
        for (int r2 = 1; r2 < 8; r2++) v[r2] = cmul_conj(buf2[rr * 72 + r2 * 9 + t2], h[r2]);
    }
    dft8<+1>(v);
#pragma unroll
    for (int q2 = 0; q2 < 8; q2++) buf1[rr * 64 + t2 + 8 * q2] = v[q2];
    group_sync(group);
#pragma unroll
    for (int r = 0; r < 8; r++) v[r] = cmul_conj(buf1[r * 64 + t], tw.g[r]);
    dft8<+1>(v);
    v[0] = make_double2(v[0].x * (1.0 / 512.0), v[0].y * (1.0 / 512.0));
#pragma unroll
    for (int q = 1; q < 8; q++) {
        double2 c = twist_const(q);
        c.x *= (1.0 / 512.0); c.y *= (1.0 / 512.0);
        v[q] = cmul_conj(v[q], c);
    }
}

}  // namespace rs
