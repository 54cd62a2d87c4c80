// redsec_b200/csrc/blind_rotate.cuh -- batched blind rotation + sample extract for sm_100a.
//
// Replaces, for a whole batch of ciphertexts in one launch, what the reference does one ciphertext at a
// time through redcufhe::Bootstrap (lib/GPU/gates.cu:124-130) / tfhe_bootstrap_FFT (lib/BinOps_enc.cpp:185):
//   modswitch to 2N  ->  acc = X^{-b} * testvector(mu)  ->  for i<n: acc += BK_i (x) ((X^{a_i}-1) acc)
//   -> sample-extract coefficient 0.
// Layout per CTA: GROUPS independent 64-thread groups (one ciphertext each).  Thread 0 of group 0 is
// also the producer: it streams the Fourier-domain bootstrapping key, one 16 KiB (i,row) slab at a time,
// with 1-D TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx) into a STAGES-deep shared-memory
// ring; all groups of the CTA consume the same slab (full/empty mbarrier pipeline), so each BSK byte is
// fetched once per CTA per step and reused GROUPS times.  (A dedicated producer warp would be the 4k+1-th
// warp of the CTA and push one SM sub-partition to an extra resident warp, which cuts the register
// budget of every thread from 255/168 to 168/128; hence the in-line producer.)
#pragma once
#include "fft512.cuh"
#include "params.h"

namespace rs {

// ---------------------------------------------------------------- mbarrier / TMA helpers (inline PTX)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------- shared-memory plan
template <int GROUPS, int STAGES>
struct BrSmem {
    static constexpr int kStageBytes = (int)BSK_ROW_BYTES;                   // 16 KiB
    static constexpr int kAccBytes = 2 * N * 4;                              // 8 KiB: a-poly, b-poly (torus32)
    static constexpr int kBuf1Bytes = FFT_BUF1 * 16;
    static constexpr int kBuf2Bytes = FFT_BUF2 * 16;
    static constexpr int kBaraBytes = 352 * 2;
    static constexpr int kGroupBytes = kAccBytes + kBuf1Bytes + kBuf2Bytes + kBaraBytes;   // 26,304 B
    static constexpr int kStagesOff = 0;
    static constexpr int kGroupsOff = STAGES * kStageBytes;
    static constexpr int kBarOff = kGroupsOff + GROUPS * kGroupBytes;
    static constexpr int kTotal = kBarOff + 2 * STAGES * 8;
};

__device__ __forceinline__ uint32_t modswitch_2N(uint32_t x) {   // modSwitchFromTorus32(x, 2N) mod 2N
    return ((x + (1u << 20)) >> 21) & (2 * N - 1);
}

// coefficient j of (X^a - 1) * poly, a in [0, 2N)
__device__ __forceinline__ uint32_t rot_diff(const uint32_t* poly, int j, int a) {
    int idx = (j - a) & (2 * N - 1);
    uint32_t v = poly[idx & (N - 1)];
    return (idx < N ? v : 0u - v) - poly[j];
}

template <int GROUPS, int STAGES>
__global__ void __launch_bounds__(GROUPS * 64, 1)
blind_rotate_kernel(const uint32_t* __restrict__ lwe_in,    // [count][LWE_STRIDE]
                    int count, uint32_t mu,
                    const double2* __restrict__ bsk_f,       // [n][BK_ROWS][2][NH]
                    uint32_t* __restrict__ ext_out)          // [count][EXT_STRIDE]
{
    using S = BrSmem<GROUPS, STAGES>;
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kBarOff);
    uint64_t* empty_bar = full_bar + STAGES;

    const int first_ct = blockIdx.x * GROUPS;
    const int active = min(GROUPS, count - first_ct);
    constexpr int kTotalRows = LWE_N * BK_ROWS;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], active * 64);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5;
    const bool producer = threadIdx.x == 0;
    const uint8_t* bsk_bytes = reinterpret_cast<const uint8_t*>(bsk_f);
    if (producer) {   // prologue: fill STAGES-1 slots
        for (int r = 0; r < STAGES - 1; r++) {
            mbar_arrive_expect_tx(&full_bar[r], S::kStageBytes);
            tma_load_1d(smem + S::kStagesOff + r * S::kStageBytes, bsk_bytes + (size_t)r * S::kStageBytes, S::kStageBytes, &full_bar[r]);
        }
    }
    const int g = warp >> 1;
    if (g >= active) return;
    const int t = threadIdx.x & 63;
    const int ct = first_ct + g;

    uint8_t* gbase = smem + S::kGroupsOff + g * S::kGroupBytes;
    uint32_t* acc = reinterpret_cast<uint32_t*>(gbase);
    double2* buf1 = reinterpret_cast<double2*>(gbase + S::kAccBytes);
    double2* buf2 = reinterpret_cast<double2*>(gbase + S::kAccBytes + S::kBuf1Bytes);
    uint16_t* bara = reinterpret_cast<uint16_t*>(gbase + S::kAccBytes + S::kBuf1Bytes + S::kBuf2Bytes);

    // ---- modswitch and accumulator init (step 2)
    const uint32_t* lwe = lwe_in + (size_t)ct * LWE_STRIDE;
    for (int i = t; i < LWE_N; i += 64) bara[i] = (uint16_t)modswitch_2N(lwe[i]);
    const int barb = (int)modswitch_2N(lwe[LWE_N]);
    for (int j = t; j < N; j += 64) {
        acc[j] = 0;
        acc[N + j] = (((j + barb) & (2 * N - 1)) < N) ? mu : 0u - mu;   // X^{2N-barb} * (mu + mu X + ...)
    }
    TwiddlesT<(GROUPS > 4)> tw;
    make_twiddles(tw, t);
    group_sync(g);

    int rc = 0;   // BSK slab counter (same sequence in every group and in the producer)
#pragma unroll 1
    for (int i = 0; i < LWE_N; i++) {
        const int a = bara[i];
        double2 f0[8], f1[8];   // Fourier accumulators for the two output polynomials
#pragma unroll
        for (int x = 0; x < 8; x++) { f0[x] = make_double2(0.0, 0.0); f1[x] = make_double2(0.0, 0.0); }

#pragma unroll 1
        for (int c = 0; c < 2; c++) {
            uint32_t src[16];   // (X^a - 1)*acc_c at this thread's 16 coefficients, plus the decomposition offset
#pragma unroll
            for (int q = 0; q < 8; q++) {
                src[2 * q] = rot_diff(acc + c * N, t + 64 * q, a) + DECOMP_OFFSET;
                src[2 * q + 1] = rot_diff(acc + c * N, t + 64 * q + NH, a) + DECOMP_OFFSET;
            }
#pragma unroll 1
            for (int p = 0; p < BK_L; p++) {
                if (producer) {   // refill the slot released by slab rc-1 with slab rc+STAGES-1
                    const int nr = rc + STAGES - 1;
                    if (nr < kTotalRows) {
                        const int ns = nr % STAGES;
                        if (rc > 0) mbar_wait(&empty_bar[ns], ((rc - 1) / STAGES) & 1);
                        mbar_arrive_expect_tx(&full_bar[ns], S::kStageBytes);
                        tma_load_1d(smem + S::kStagesOff + ns * S::kStageBytes, bsk_bytes + (size_t)nr * S::kStageBytes,
                                    S::kStageBytes, &full_bar[ns]);
                    }
                }
                const int sh = 32 - (p + 1) * BK_BGBIT;
                double2 v[8];
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    v[q].x = (double)((int)((src[2 * q] >> sh) & 7u) - 4);
                    v[q].y = (double)((int)((src[2 * q + 1] >> sh) & 7u) - 4);
                }
                fft512_fwd(v, tw, buf1, buf2, t, g);

                const int s = rc % STAGES;
                mbar_wait(&full_bar[s], (rc / STAGES) & 1);
                const double2* B = reinterpret_cast<const double2*>(smem + S::kStagesOff + s * S::kStageBytes);
#pragma unroll
                for (int x = 0; x < 8; x++) {
                    const double2 b0 = B[x * 64 + t], b1 = B[NH + x * 64 + t];
                    f0[x].x += v[x].x * b0.x - v[x].y * b0.y;
                    f0[x].y += v[x].x * b0.y + v[x].y * b0.x;
                    f1[x].x += v[x].x * b1.x - v[x].y * b1.y;
                    f1[x].y += v[x].x * b1.y + v[x].y * b1.x;
                }
                mbar_arrive(&empty_bar[s]);
                rc++;
            }
        }
        group_sync(g);   // last forward exchange reads finished before the inverse reuses buf2
        // ---- inverse transforms, round to nearest, accumulate into acc (exact integers mod 2^32)
        fft512_inv(f0, tw, buf1, buf2, t, g);
#pragma unroll
        for (int q = 0; q < 8; q++) {
            acc[t + 64 * q] += (uint32_t)__double2ll_rn(f0[q].x);
            acc[t + 64 * q + NH] += (uint32_t)__double2ll_rn(f0[q].y);
        }
        fft512_inv(f1, tw, buf1, buf2, t, g);
#pragma unroll
        for (int q = 0; q < 8; q++) {
            acc[N + t + 64 * q] += (uint32_t)__double2ll_rn(f1[q].x);
            acc[N + t + 64 * q + NH] += (uint32_t)__double2ll_rn(f1[q].y);
        }
        group_sync(g);
    }

    // ---- sample extract: a'[0]=acc_a[0], a'[j]=-acc_a[N-j], b'=acc_b[0]
    uint32_t* ext = ext_out + (size_t)ct * EXT_STRIDE;
    for (int j = t; j < N; j += 64) ext[j] = (j == 0) ? acc[0] : 0u - acc[N - j];
    if (t == 0) { ext[N] = acc[N]; ext[N + 1] = 0; ext[N + 2] = 0; ext[N + 3] = 0; }
}

// ---------------------------------------------------------------- BSK -> Fourier-domain device layout (north-star item (c))
// One 64-thread group per polynomial; same forward routine as the blind rotation, so the slot layout matches.
__global__ void __launch_bounds__(64)
bsk_to_fourier_kernel(const int32_t* __restrict__ bsk, double2* __restrict__ bsk_f, int npolys) {
    __shared__ double2 buf1[FFT_BUF1];
    __shared__ double2 buf2[FFT_BUF2];
    const int t = threadIdx.x;
    Twiddles tw;
    make_twiddles(tw, t);
    for (int poly = blockIdx.x; poly < npolys; poly += gridDim.x) {
        const int32_t* p = bsk + (size_t)poly * N;
        double2 v[8];
#pragma unroll
        for (int q = 0; q < 8; q++) v[q] = make_double2((double)p[t + 64 * q], (double)p[t + 64 * q + NH]);
        fft512_fwd(v, tw, buf1, buf2, t, 0);
        double2* o = bsk_f + (size_t)poly * NH;
#pragma unroll
        for (int x = 0; x < 8; x++) o[x * 64 + t] = v[x];
        __syncthreads();
    }
}

}  // namespace rs
