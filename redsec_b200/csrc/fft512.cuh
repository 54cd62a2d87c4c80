// redsec_b200/csrc/fft512.cuh -- negacyclic size-1024 real transform as a folded, twisted 512-point
// complex FP64 FFT, held in registers by a 64-thread group (8 complex points per thread, three
// radix-8 passes; exchange 1 through shared memory, exchange 2 through warp shuffles).
//
// Replaces (from scratch) the role of TFHE's spqlios FFT inside tfhe_bootstrap_FFT, which the
// reference calls at lib/BinOps_enc.cpp:185,191; and of redcufhe's NTT (lib/GPU/gates.cuh:7).
//
// Index algebra (forward, W = exp(-2*pi*i/512), W8 = W^64, W64 = W^8, omega = exp(i*pi/1024)):
//   z[j] = (p[j] + i*p[j+512]) * omega^j,  Z[k] = sum_j z[j] W^{jk} = p(omega * W^k)   (a root of X^1024+1)
//   j = t + 64q,  t = t2 + 8*q2,  k = r + 8*r2 + 64*r3
//   pass 1 (thread t):          a_r[t]    = omega^t W^{t r} * DFT8_q( z[t+64q] * omega^{64q} )[r]
//   pass 2 (thread u=t2+8r):    b_rr2[t2] = W64^{t2 r2} * DFT8_q2( a_r[t2+8q2] )[r2]
//   pass 3 (thread v=r2+8r):    Z[r+8r2+64r3] = DFT8_t2( b_rr2[t2] )[r3]
//
// Exchange 2 (b_rr2[t2]: thread t2 -> thread r2, same r) is an 8x8 transpose inside each group of 8 consecutive
// lanes.  It is done with 7 rounds of 16-byte shuffles and NO dynamic register indexing by rotating both sides:
//   * the pass-2 inputs a_r[t2+8q2] are also multiplied by W8^{q2*t2} (merged into the g twiddle), which rotates the
//     pass-2 output so that register k of thread t2 holds frequency r2 = (t2+k) mod 8;
//   * round k sends register k to lane (t2+k) mod 8, so lane r2 receives y[k] = b_rr2[(r2-k) mod 8];
//   * DFT8 of the reversed+rotated sequence y is W8^{r2 r3} * DFT8^+(y)[r3]  (DFT8^+ = conjugate kernel).
// The unit-modulus factor phi = W8^{r2 r3} is NOT applied: the transform returns V' = conj(phi) * Z.  The blind
// rotation multiplies V' by the TRUE Fourier bootstrapping key, so its accumulators hold F' = conj(phi) * F, which is
// exactly the pre-rotated input the mirrored inverse needs.  Only bsk_to_fourier_kernel applies phi (once, at key load).
// The inverse is the exact mirror (conjugate constants, reversed exchanges, scale 1/512).
// Twiddles are applied on the consuming side of each exchange (pass-1 twiddles by the reader of exchange 1, pass-2
// twiddles by the receiver of exchange 2) so that they fuse into the first butterfly stage of the next DFT8 as FMAs.
// Output "slot" layout of the forward transform: thread v holds slots r3*64+v, r3=0..7.
//
// Exchange 1 crosses the two warps of a group and goes through shared memory.  The caller alternates between two
// exchange buffers from one transform to the next, so ONE named barrier per transform is enough (a warp can only
// reach the writes of transform n+2 after the other warp has arrived at the barrier of transform n+1, i.e. after its
// reads of transform n).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace rs {

struct Twiddles {
    double2 g[8];   // g[q2] = exp(i*pi*t*(1-4r)/1024) * exp(-2*pi*i*q2*t2/8), t = t2+8*q2   (twist, pass-1 twiddle, rotation;
                    //         applied by the READER of exchange 1: thread u = t2+8r)
    double2 h[8];   // h[k]  = exp(-2*pi*i*((r2-k)&7)*r2/64)                                 (pass-2 twiddle, applied by the
                    //         RECEIVER of exchange 2: thread v = r2+8r, register k came from lane t2 = (r2-k)&7)
};

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
__device__ __forceinline__ void make_twiddles(Twiddles& tw, int tid) {
    const int lo = tid & 7, r = tid >> 3;    // lo plays t2 (as exchange-1 reader) and r2 (as exchange-2 receiver)
#pragma unroll
    for (int k = 0; k < 8; k++) {
        // exponents are reduced with integers first, so sincospi sees small exact arguments
        const int t = lo + 8 * k;                                               // writer of element q2 = k
        const int e1024 = (t * (1 - 4 * r) - 256 * ((k * lo) & 7)) % 2048;      // units of pi/1024
        tw.g[k] = unit_root((double)e1024 / 1024.0);
        const int e32 = (((lo - k) & 7) * lo) & 63;                             // units of -pi/32
        tw.h[k] = unit_root(-(double)e32 / 32.0);
    }
}

// 8-point DFT, W8 = exp(S*2*pi*i/8), natural order in and out, split in two so that twiddles can be fused into the
// first butterfly stage.  dft8_tail takes a[q] = x[q] + x[q+4], b[q] = x[q] - x[q+4] (q < 4): 36 FP64 instructions.
template <int S>
__device__ __forceinline__ void dft8_tail(const double2 (&a)[4], const double2 (&b)[4], double2 (&v)[8]) {
    const double hs = 0.70710678118654752440;
    // b1 *= (hs, S*hs); b2 *= i*S; b3 *= (-hs, S*hs)
    const double2 u1 = make_double2(b[1].x - S * b[1].y, S * b[1].x + b[1].y);     // (1 + iS) * b1
    const double2 u3 = make_double2(-b[3].x - S * b[3].y, S * b[3].x - b[3].y);    // (-1 + iS) * b3
    const double2 b2 = mul_i<S>(b[2]);
    // even outputs: DFT4(a)
    const double2 e0 = cadd(a[0], a[2]), e1 = csub(a[0], a[2]), o0 = cadd(a[1], a[3]), o1 = mul_i<S>(csub(a[1], a[3]));
    v[0] = cadd(e0, o0); v[4] = csub(e0, o0); v[2] = cadd(e1, o1); v[6] = csub(e1, o1);
    // odd outputs: DFT4(b0, hs*u1, b2, hs*u3)
    const double2 f0 = cadd(b[0], b2), f1 = csub(b[0], b2);
    const double2 s = cadd(u1, u3), d = mul_i<S>(csub(u1, u3));
    v[1] = make_double2(fma(hs, s.x, f0.x), fma(hs, s.y, f0.y));
    v[5] = make_double2(fma(-hs, s.x, f0.x), fma(-hs, s.y, f0.y));
    v[3] = make_double2(fma(hs, d.x, f1.x), fma(hs, d.y, f1.y));
    v[7] = make_double2(fma(-hs, d.x, f1.x), fma(-hs, d.y, f1.y));
}
template <int S>
__device__ __forceinline__ void dft8(double2 (&v)[8]) {
    double2 a[4], b[4];
#pragma unroll
    for (int q = 0; q < 4; q++) { a[q] = cadd(v[q], v[q + 4]); b[q] = csub(v[q], v[q + 4]); }
    dft8_tail<S>(a, b, v);
}
// DFT8 of x[q]*w[q] with the twiddle products fused into the first butterfly stage:
//   p = x_q w_q (2 DMUL + 2 DFMA), a = p + x_{q+4} w_{q+4} (4 DFMA), b = 2p - a (2 DFMA): 40 instead of 48 instructions.
// W0_ONE: w[0] == 1 and is not read.
template <int S, bool W0_ONE, class WFn>
__device__ __forceinline__ void dft8_twiddled(double2 (&v)[8], WFn w) {
    double2 a[4], b[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const double2 p = (q == 0 && W0_ONE) ? v[0] : cmul(v[q], w(q));
        const double2 y = v[q + 4], c = w(q + 4);
        a[q].x = fma(y.x, c.x, fma(-y.y, c.y, p.x));
        a[q].y = fma(y.x, c.y, fma(y.y, c.x, p.y));
        b[q].x = fma(2.0, p.x, -a[q].x);
        b[q].y = fma(2.0, p.y, -a[q].y);
    }
    dft8_tail<S>(a, b, v);
}

// omega^{64q} = exp(i*pi*q/16)
__device__ __forceinline__ double2 twist_const(int q) {
    const double c[8] = {1.0, 0.98078528040323044913, 0.92387953251128675613, 0.83146961230254523708,
                         0.70710678118654752440, 0.55557023301960222474, 0.38268343236508977173, 0.19509032201612826785};
    const double s[8] = {0.0, 0.19509032201612826785, 0.38268343236508977173, 0.55557023301960222474,
                         0.70710678118654752440, 0.83146961230254523708, 0.92387953251128675613, 0.98078528040323044913};
    return make_double2(c[q], s[q]);
}

constexpr int FFT_BUF = 512;           // complex entries of one exchange buffer: [r][t]

// 8x8 transpose inside each group of 8 consecutive lanes: v[k] <- register k of lane (lane-k) mod 8.
__device__ __forceinline__ void rotate_exchange(double2 (&v)[8], int lo) {
#pragma unroll
    for (int k = 1; k < 8; k++) {
        const int src = (lo - k) & 7;
        v[k].x = __shfl_sync(0xffffffffu, v[k].x, src, 8);
        v[k].y = __shfl_sync(0xffffffffu, v[k].y, src, 8);
    }
}

// Forward transform.  In: v[q] = p[t+64q] + i*p[t+64q+512].  Out: v[r3] = conj(phi) * Z[slot r3*64 + tid].
__device__ __forceinline__ void fft512_fwd(double2 (&v)[8], const Twiddles& tw, double2* buf, int t, int group) {
    dft8_twiddled<-1, true>(v, [](int q) { return twist_const(q); });
#pragma unroll
    for (int r = 0; r < 8; r++) buf[r * 64 + t] = v[r];
    group_sync(group);
    const int lo = t & 7, rr = t >> 3;
#pragma unroll
    for (int q2 = 0; q2 < 8; q2++) v[q2] = buf[rr * 64 + lo + 8 * q2];
    dft8_twiddled<-1, false>(v, [&](int q) { return tw.g[q]; });
    rotate_exchange(v, lo);
    dft8_twiddled<+1, false>(v, [&](int k) { return tw.h[k]; });
}

// Inverse transform (scaled by 1/512).  In: v[r3] = conj(phi) * A[slot r3*64 + tid].
// Out: v[q] = z[t+64q] with Re -> coefficient t+64q, Im -> coefficient t+64q+512.
__device__ __forceinline__ void fft512_inv(double2 (&v)[8], const Twiddles& tw, double2* buf, int t, int group) {
    const int lo = t & 7, rr = t >> 3;
    dft8<+1>(v);
    v[0] = cmul_conj(v[0], tw.h[0]);
#pragma unroll
    for (int k = 1; k < 8; k++) v[k] = cmul_conj(v[k], tw.h[8 - k]);
    rotate_exchange(v, lo);
    dft8<-1>(v);
#pragma unroll
    for (int q2 = 0; q2 < 8; q2++) buf[rr * 64 + lo + 8 * q2] = cmul_conj(v[q2], tw.g[q2]);
    group_sync(group);
#pragma unroll
    for (int r = 0; r < 8; r++) v[r] = buf[r * 64 + t];
    dft8<+1>(v);
    v[0] = make_double2(v[0].x * (1.0 / 512.0), v[0].y * (1.0 / 512.0));
#pragma unroll
    for (int q = 1; q < 8; q++) {
        double2 c = twist_const(q);
        c.x *= (1.0 / 512.0); c.y *= (1.0 / 512.0);
        v[q] = cmul_conj(v[q], c);
    }
}

// phi(thread, r3) = exp(-2*pi*i*(t&7)*r3/8): the factor fft512_fwd leaves out (see header)
__device__ __forceinline__ double2 fwd_phase(int t, int r3) {
    return unit_root(-(double)(((t & 7) * r3) & 7) / 4.0);
}

}  // namespace rs
