// redsec_b200/csrc/keyswitch_mma.cuh -- the LWE keyswitch as an exact integer GEMM on the 5th-generation tensor cores
// (tcgen05.mma kind::i8, accumulators in tensor memory).  Variant 2 of the keyswitch (rs_set_ks_variant) and the default for
// batches of 2 048 ciphertexts and more: 5.3 ms per 2^16 against 24.9 ms for the shared-memory gather kernel of
// lwe_kernels.cuh, which keeps the small batches.
//
// lweKeySwitch (TFHE's lweKeySwitch behind tfhe_bootstrap_FFT, lib/BinOps_enc.cpp:185):
//     out = (0, b') - sum_{i<1024, j<9} KSK[i][j][ digit_j(a'_i) ]            (digit 0 contributes nothing)
// is a contraction of a ONE-HOT matrix with the key:
//     A[ct][(i,j,d)] = [digit_j(a'_i of ct) == d]            (u8, 0/1;   M = ciphertexts, K = 1024*9*8 = 73 728)
//     B[(i,j,d)][(word, limb)] = byte `limb` of KSK[i][j][d][word]   (u8;  N = 352*4 = 1 408; rows d = 0 are zero)
//     S[ct][(word, limb)] = sum_K A*B  <= 9 216 * 255 < 2^22         (s32 accumulators, exact)
//     out[ct][word] = [word == 350] * b' - (S0 + 2^8 S1 + 2^16 S2 + 2^24 S3)   mod 2^32
// so it is bit-exact by construction.  One CTA = 256 ciphertexts (two M = 128 tiles, both accumulators in TMEM: 2 x 256
// columns = all 512) x 64 output words (N = 256 columns); both M tiles use every key tile, which halves the key traffic per
// ciphertext.  One pipeline stage = two K-steps of 32 bytes (a K-step = 4 (i,j) pairs x 8 digit slots):
//   warp 0 lane 0      streams the stage's 16 KiB key tile with one 1-D TMA bulk copy through a 10-stage ring (the key is
//                      pre-tiled in the canonical no-swizzle K-major layout by kskb_build_kernel at first use);
//   warps 1..8         build the one-hot A tile in shared memory, one ciphertext per thread (digit -> 1 << 8*digit), from the
//                      transposed a' array (coalesced; eight coefficients in registers, the next eight requested a group
//                      ahead), fence.proxy.async, arrive;
//   warp 9 lane 0      issues the four tcgen05.mma of the stage (two K-steps x two M tiles) and a tcgen05.commit onto each
//                      ring's empty barrier;
// after 1 152 stages warps 1..8 read their accumulator rows with tcgen05.ld, recombine the limbs and store the output words
// (every output word has exactly one writer: no atomics, no initialising pass).
#pragma once
#include "blind_rotate_tm.cuh"   // tcgen05 / TMEM helpers
#include "lwe_kernels.cuh"

namespace rs {

constexpr int KM_CTS = 256;                         // ciphertexts per CTA
constexpr int KM_WORDS = 64;                        // output words per CTA
constexpr int KM_NCOL = KM_WORDS * 4;               // accumulator columns
constexpr int KM_NT = 6;                            // N tiles (352 words padded to 384)
constexpr int KM_PAIRS = 4;                         // (i,j) pairs per MMA (K = 32 bytes = 4 pairs x 8 digit slots)
constexpr int KM_STEPS = N * KS_T / KM_PAIRS;       // 2 304 MMA K-steps; the tiled key is laid out per K-step
constexpr int KM_BSTEP = KM_NCOL * 32;              // 8 KiB of key per K-step
constexpr int KM_KK = 2;                            // K-steps per pipeline stage (one barrier round trip per 64 bytes of K)
constexpr int KM_STAGE_STEPS = KM_STEPS / KM_KK;    // 1 152 stages
constexpr int KM_BSTAGE = KM_KK * KM_BSTEP;         // 16 KiB key tile per stage
constexpr int KM_ASTAGE = KM_KK * KM_CTS * 32;      // 16 KiB one-hot tile per stage
constexpr int KM_SA = 3;                            // one-hot ring: produced on the SM, three stages are plenty
constexpr int KM_SB = 10;                           // key ring: 160 KiB in flight hides the L2 latency under load
constexpr size_t KSKB_BYTES = (size_t)KM_NT * KM_STEPS * KM_BSTEP;      // 113 MB
struct KmSmem {
    static constexpr int kAOff = 0;
    static constexpr int kBOff = KM_SA * KM_ASTAGE;
    static constexpr int kBarOff = kBOff + KM_SB * KM_BSTAGE;
    // barriers: b_full[SB], b_empty[SB], a_full[SA], a_empty[SA], acc_full
    static constexpr int kBFull = 0, kBEmpty = KM_SB, kAFull = 2 * KM_SB, kAEmpty = 2 * KM_SB + KM_SA, kAccFull = 2 * KM_SB + 2 * KM_SA;
    static constexpr int kTmemPtrOff = kBarOff + (kAccFull + 1) * 8;
    static constexpr int kTotal = kTmemPtrOff + 16;
};
static_assert(KmSmem::kTotal > 116 * 1024, "the tile rings must also keep a second CTA (and its 512-column TMEM allocation) off the SM");

// tiled key: [nt][step][kc 2][n/8 32][n%8 8][16 B]; byte b of a chunk = K index kc*16 + b = pair (kc*2 + b/8), digit b%8
__global__ void kskb_build_kernel(const uint32_t* __restrict__ ksk /*[N][t][8][LWE_STRIDE], the padded device table*/, uint8_t* __restrict__ kskb) {
    const size_t chunks = KSKB_BYTES / 16;
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < chunks; c += (size_t)gridDim.x * blockDim.x) {
        const int nr = (int)(c % 8), ng = (int)((c / 8) % 32), kc = (int)((c / 256) % 2);
        const size_t s_nt = c / 512;
        const int step = (int)(s_nt % KM_STEPS), nt = (int)(s_nt / KM_STEPS);
        const int n = ng * 8 + nr, wl = n >> 2, limb = n & 3, word = nt * KM_WORDS + wl;
        uint32_t out[4] = {0, 0, 0, 0};
        if (word < LWE_WORDS) {
#pragma unroll
            for (int b = 0; b < 16; b++) {
                const int pair = step * KM_PAIRS + kc * 2 + (b >> 3), d = b & 7;
                const int i = pair / KS_T, j = pair % KS_T;
                if (d) {
                    const uint32_t v = __ldcg(ksk + (((size_t)i * KS_T + j) * KS_BASE + d) * LWE_STRIDE + word);   // written by ksk_pad_kernel, possibly twice (key reload): L2 only
                    out[b >> 2] |= ((v >> (8 * limb)) & 255u) << (8 * (b & 3));
                }
            }
        }
        reinterpret_cast<uint4*>(kskb)[c] = make_uint4(out[0], out[1], out[2], out[3]);
    }
}

// a'_i + rounding offset, transposed to [i][ciphertext] (stride = count padded to 256), and b'
__global__ void ks_mma_prep_kernel(const uint32_t* __restrict__ ext, int count, int stride, uint32_t* __restrict__ abar_t,
                                   uint32_t* __restrict__ bprime) {
    __shared__ uint32_t tile[32][33];
    const int ct0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int ct = ct0 + r;
        tile[r][threadIdx.x] = ct < count ? __ldcg(ext + (size_t)ct * EXT_STRIDE + i0 + threadIdx.x) + KS_PREC_OFFSET : 0u;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int ct = ct0 + threadIdx.x;
        if (ct < stride) abar_t[(size_t)(i0 + r) * stride + ct] = tile[threadIdx.x][r];
    }
    if (blockIdx.y == 0 && threadIdx.y == 0) {
        const int ct = ct0 + threadIdx.x;
        if (ct < count) bprime[ct] = __ldcg(ext + (size_t)ct * EXT_STRIDE + N);
    }
    __threadfence();
}

__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}

__global__ void __launch_bounds__(320, 1)
keyswitch_mma_kernel(const uint32_t* __restrict__ abar_t, const uint32_t* __restrict__ bprime, int count, int stride,
                     const uint8_t* __restrict__ kskb, uint32_t* __restrict__ lwe_out) {
    using S = KmSmem;
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar = smem_base + S::kBarOff;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int first = blockIdx.x * KM_CTS, nt = blockIdx.y;

    if (threadIdx.x == 0) {
        for (int s = 0; s < KM_SB; s++) { mbar_init(bar + (S::kBFull + s) * 8, 1); mbar_init(bar + (S::kBEmpty + s) * 8, 1); }
        for (int s = 0; s < KM_SA; s++) {
            mbar_init(bar + (S::kAFull + s) * 8, 8);       // one arrival per generator warp
            mbar_init(bar + (S::kAEmpty + s) * 8, 1);      // tcgen05.commit
        }
        mbar_init(bar + S::kAccFull * 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc_512(smem_base + S::kTmemPtrOff);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + S::kTmemPtrOff);

    if (warp == 0) {
        // ===================================================================== key-tile producer (TMA)
        if (lane == 0) {
            const uint8_t* src = kskb + (size_t)nt * KM_STEPS * KM_BSTEP;
#pragma unroll 1
            for (int s = 0; s < KM_STAGE_STEPS; s++) {
                const int st = s % KM_SB;
                if (s >= KM_SB) mbar_wait_thread(bar + (S::kBEmpty + st) * 8, ((s - KM_SB) / KM_SB) & 1);
                mbar_arrive_expect_tx(bar + (S::kBFull + st) * 8, KM_BSTAGE);
                tma_load_1d(smem_base + S::kBOff + st * KM_BSTAGE, src + (size_t)s * KM_BSTAGE, KM_BSTAGE, bar + (S::kBFull + st) * 8);
            }
        }
    } else if (warp == 9) {
        // ===================================================================== MMA issuer (one thread)
        if (lane == 0) {
            // instruction descriptor: D = s32 (2 << 4), A and B unsigned 8 bit (0), both K-major, N = 256 (>> 3 at bit 17), M = 128 (>> 4 at bit 24)
            const uint32_t idesc = (2u << 4) | ((uint32_t)(KM_NCOL >> 3) << 17) | ((128u >> 4) << 24);
            // canonical no-swizzle K-major tiles: core matrix = 8 rows x 16 bytes = 128 contiguous bytes; the tiles are stored
            // [k chunk][row/8][row%8][16 B]: the descriptor's leading byte offset is the distance between core matrices adjacent in K
            // ((rows/8)*128 bytes), its stride byte offset the distance between 8-row groups (128 bytes) -- measured: the other
            // assignment gives wrong sums
            constexpr uint32_t a_k = 16 * 128, b_k = 32 * 128, mn = 128;
            constexpr uint32_t a_mt = KM_KK * 2 * a_k;          // bytes of one M tile's one-hot rows in a stage
#pragma unroll 1
            for (int s = 0; s < KM_STAGE_STEPS; s++) {
                const int sa = s % KM_SA, sb = s % KM_SB;
                mbar_wait_thread(bar + (S::kBFull + sb) * 8, (s / KM_SB) & 1);
                mbar_wait_thread(bar + (S::kAFull + sa) * 8, (s / KM_SA) & 1);
                tc_fence_after();
                const uint32_t a_addr = smem_base + S::kAOff + sa * KM_ASTAGE, b_addr = smem_base + S::kBOff + sb * KM_BSTAGE;
#pragma unroll
                for (int kk = 0; kk < KM_KK; kk++) {
                    const uint64_t db = tc_smem_desc(b_addr + kk * 2 * b_k, b_k, mn);
                    tc_mma_i8(tmem, tc_smem_desc(a_addr + kk * 2 * a_k, a_k, mn), db, idesc, (s | kk) > 0);
                    tc_mma_i8(tmem + KM_NCOL, tc_smem_desc(a_addr + a_mt + kk * 2 * a_k, a_k, mn), db, idesc, (s | kk) > 0);
                }
                tc_commit(bar + (S::kAEmpty + sa) * 8);         // both rings' stages are free once these MMAs have read them
                tc_commit(bar + (S::kBEmpty + sb) * 8);
            }
            tc_commit(bar + S::kAccFull * 8);
        }
    } else {
        // ===================================================================== one-hot generators, then epilogue (warps 1..8)
        const int g = (warp - 1) * 32 + lane;          // ciphertext of this CTA handled by this thread
        const int mt = g >> 7, m = g & 127;
        const uint32_t* col = abar_t + first + g;      // a'_i of this ciphertext: abar_t[i * stride + first + g] (rows beyond count are zero)
        // 8 coefficients = 72 (i,j) pairs = 9 pipeline stages per group: the group's 8 words sit in registers and the next group's
        // are requested a whole group (~9 stages) ahead, so the L2 latency of these per-thread loads never stalls a stage
        static_assert(KM_KK * KM_PAIRS == 8 && KS_T == 9, "the group structure below assumes 8 pairs per stage and 9 digits per coefficient");
        uint32_t cur[8], nxt[8];
#pragma unroll
        for (int k = 0; k < 8; k++) cur[k] = __ldcg(col + (size_t)k * stride);
#pragma unroll 1
        for (int grp = 0; grp < N / 8; grp++) {
#pragma unroll
            for (int k = 0; k < 8; k++) nxt[k] = grp + 1 < N / 8 ? __ldcg(col + (size_t)((grp + 1) * 8 + k) * stride) : 0u;
#pragma unroll
            for (int ss = 0; ss < 9; ss++) {
                const int s = grp * 9 + ss, sa = s % KM_SA;
                uint64_t oh[8];
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const int p = ss * 8 + q, il = p / KS_T, j = p % KS_T;           // compile-time after unrolling
                    const uint32_t d = (cur[il] >> (32 - (j + 1) * KS_BASEBIT)) & (KS_BASE - 1);
                    oh[q] = 1ull << (8 * d);             // digit 0 selects the all-zero key row
                }
                if (s >= KM_SA) mbar_wait_warp(bar + (S::kAEmpty + sa) * 8, ((s - KM_SA) / KM_SA) & 1);
                uint8_t* a = smem + S::kAOff + sa * KM_ASTAGE + mt * (KM_ASTAGE / 2) + (m >> 3) * 128 + (m & 7) * 16;
#pragma unroll
                for (int kc = 0; kc < 4; kc++)           // K chunk kc = pairs 2kc, 2kc+1
                    *reinterpret_cast<uint4*>(a + kc * 16 * 128) = make_uint4((uint32_t)oh[2 * kc], (uint32_t)(oh[2 * kc] >> 32),
                                                                                (uint32_t)oh[2 * kc + 1], (uint32_t)(oh[2 * kc + 1] >> 32));
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy stores -> visible to the tensor core
                __syncwarp();
                if (lane == 0) mbar_arrive(bar + (S::kAFull + sa) * 8);
            }
#pragma unroll
            for (int k = 0; k < 8; k++) cur[k] = nxt[k];
        }
        // ---- epilogue: accumulator row -> limbs -> words.  A warp may only touch the TMEM lanes of its quarter (warp % 4)
        mbar_wait_warp(bar + S::kAccFull * 8, 0);
        tc_fence_after();
        const int emt = (warp - 1) >> 2, quarter = warp & 3;
        const int row = quarter * 32 + lane, ct = first + emt * 128 + row;
        const uint32_t bp = ct < count ? __ldcg(bprime + ct) : 0u;
#pragma unroll 1
        for (int c16 = 0; c16 < KM_NCOL / 16; c16++) {
            uint32_t r[16];
            tmem_ld16(tmem + ((uint32_t)(quarter * 32) << 16) + emt * KM_NCOL + c16 * 16, r);
            tmem_ld_wait16(r);
            uint32_t o[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int word = nt * KM_WORDS + c16 * 4 + k;
                const uint32_t sum = r[4 * k] + (r[4 * k + 1] << 8) + (r[4 * k + 2] << 16) + (r[4 * k + 3] << 24);
                o[k] = word < LWE_WORDS ? (word == LWE_N ? bp : 0u) - sum : 0u;
            }
            const int word0 = nt * KM_WORDS + c16 * 4;
            if (ct < count && word0 < LWE_STRIDE)
                *reinterpret_cast<uint4*>(lwe_out + (size_t)ct * LWE_STRIDE + word0) = make_uint4(o[0], o[1], o[2], o[3]);
        }
        __threadfence();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc_512(tmem);
}

}  // namespace rs
