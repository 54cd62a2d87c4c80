// redsec_b200/csrc/blind_rotate_tm.cuh -- warp-specialised blind rotation with the Fourier bootstrapping key served
// to the multiply-accumulate from TENSOR MEMORY instead of shared memory.
//
// Same computation and the same arithmetic per row as blind_rotate_ws.cuh (the role of tfhe_bootstrap_FFT at
// lib/BinOps_enc.cpp:185 / redcufhe::Bootstrap at lib/GPU/gates.cu:124-130 for a whole batch), so results stay
// bit-identical (tests/test_gpu_pbs.py::test_variants_agree).  What changes is the path of the BSK operand:
//
//   HBM/L2 --TMA (cp.async.bulk)--> smem slab --tcgen05.cp 64x128b.warpx2::02_13--> TMEM --tcgen05.ld--> registers
//
// Why: ncu on the ws kernel (profiles/r1_blind_rotate_ws_ncu_summary.txt) shows the LSU/shared data pipe at 67 % of
// its wavefront peak next to an FP64 pipe at 61 %: the two are co-limiters, and 44 % of those wavefronts are the four
// ciphertexts of a CTA each re-reading the same 16 KiB BSK slab with ld.shared.  tcgen05.ld runs on its own pipe
// (measured, scripts/probes/tmem_probe.cu: 257 B/clk/SM for 8 warps, and 121 + 121 B/clk when interleaved with LDS.128,
// i.e. the two paths do not share bandwidth), so the slab is read from shared memory ONCE per CTA (by the copy engine)
// and the per-ciphertext reads move off the LSU pipe.  No tensor-core MMA is involved; TMEM is used as a broadcast
// operand buffer.
//
// tcgen05.cp layout (measured with the same probe): with a no-swizzle descriptor, shape 64x128b reads row r's 16 bytes
// from  start + (r/8)*SBO + (r%8)*16, so SBO = 128 makes it a plain contiguous [64][16 B] block -- exactly one
// (polynomial, x) plane  B[o*512 + x*64 + u]  of the existing slab layout, u = thread-column.  The .warpx2::02_13
// multicast writes rows 0..31 to TMEM lanes 0..31 AND 64..95 and rows 32..63 to lanes 32..63 AND 96..127, which is the
// lane quadrant (warp id % 4) of the back warp that owns thread-column u for ciphertexts {0,2} and {1,3}.
// One slab = 16 such copies into 64 TMEM columns, ordered x-major so one tcgen05.ld.32x32b.x16 fetches
// (B0[x], B1[x], B0[x+1], B1[x+1]).  512 columns = an 8-deep TMEM ring.
//
// One CTA = 16 warps, 4 ciphertexts:
//   warps 0..7   BACK  (two per ciphertext): passes 2+3 of the forward transform, MAC against TMEM, inverse transforms
//   warps 8..11  FRONT (one per ciphertext): rotate/decompose, pass 1, exchange ring            (as in the ws kernel)
//   warp  12     TMA producer (one lane): slab requests
//   warps 13..15 copy producers (one lane each): smem->TMEM copies, tcgen05.commit onto the ring barriers
//   (a CTA's register allocation is rounded to 4 warps anyway: 13 warps get the budget of 16)
// Registers (setmaxnreg): launch at 128 x 512 = 65536; back 184, front 120, the rest 24 (8*32*184 + 4*32*120 + 4*32*24 = 65536).
#pragma once
#include <type_traits>
#include "blind_rotate.cuh"
#include "blind_rotate_ws.cuh"

namespace rs {

// ---------------------------------------------------------------- tcgen05 helpers (inline PTX)
__device__ __forceinline__ void tmem_alloc_512(uint32_t smem_result_addr) {       // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_result_addr) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_512(uint32_t taddr) {               // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// shared-memory matrix descriptor, no swizzle: start address, leading/stride byte offsets (>>4), sm_100 version bit
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
           ((uint64_t)1 << 46);
}
__device__ __forceinline__ void tc_cp_64x128b_02_13(uint32_t taddr, uint64_t desc) {
    asm volatile("tcgen05.cp.cta_group::1.64x128b.warpx2::02_13 [%0], %1;" ::"r"(taddr), "l"(desc) : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {    // arrives (count 1) when all prior tcgen05 async ops of this thread are done
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
// tcgen05.wait::ld that the compiler cannot move the consumers of r[] above: r is an in/out operand of the wait
__device__ __forceinline__ void tmem_ld_wait16(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :: "memory");
}

// single-thread wait for hand-offs that are expected to last: polls with a suspend-time hint so the spinning lane does
// not take issue slots from the back warps of its SM sub-partition
__device__ __forceinline__ void mbar_wait_thread_long(uint32_t bar, uint32_t parity) {
    if (mbar_try(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_hint(bar, parity, 4000u)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}

template <int STAGES, int XSLOTS>
struct TmSmem {
    static constexpr int kCts = 4;
    static constexpr int kTStages = 8;                                      // 512 TMEM columns / 64 per slab
    static constexpr int kStageBytes = (int)BSK_ROW_BYTES;                  // 16 KiB BSK slab
    static constexpr int kAccBytes = 2 * N * 4;
    static constexpr int kBaraBytes = 352 * 2;
    static constexpr int kSlotBytes = FFT_BUF * 16;
    static constexpr int kCtBytes = kAccBytes + kBaraBytes + XSLOTS * kSlotBytes;
    static constexpr int kStagesOff = 0;
    static constexpr int kCtOff = STAGES * kStageBytes;
    static constexpr int kBarOff = kCtOff + kCts * kCtBytes;
    // barriers: bsk_full[STAGES], bsk_empty[STAGES], tm_full[8], tm_empty[8], x_full[4][XSLOTS], x_empty[4][XSLOTS], acc_ready[4]
    static constexpr int kBskFull = 0, kBskEmpty = STAGES, kTmFull = 2 * STAGES, kTmEmpty = kTmFull + kTStages,
                         kXFull = kTmEmpty + kTStages, kXEmpty = kXFull + kCts * XSLOTS, kAccReady = kXEmpty + kCts * XSLOTS,
                         kNumBars = kAccReady + kCts;
    static constexpr int kTmemPtrOff = kBarOff + kNumBars * 8;
    static constexpr int kTotal = kTmemPtrOff + 8;
    static_assert(kCtBytes % 16 == 0, "ciphertext block must stay 16-byte aligned");
    static_assert(XSLOTS >= 2, "inverse hand-off uses slots 0 and 1");
};

template <int STAGES, int XSLOTS, int FRONT_REGS, int BACK_REGS>
__global__ void __launch_bounds__(512, 1)
blind_rotate_tm_kernel(const uint32_t* __restrict__ lwe_in,    // [count][LWE_STRIDE]
                       int count, uint32_t mu,
                       const double2* __restrict__ bsk_f,       // [n][BK_ROWS][2][NH]
                       uint32_t* __restrict__ ext_out)          // [count][EXT_STRIDE]
{
    static_assert(2 * BACK_REGS + FRONT_REGS + 24 <= 512, "per-SMSP register file: 2 back + 1 front + 1 producer warp");
    using S = TmSmem<STAGES, XSLOTS>;
    constexpr int TS = S::kTStages;
    constexpr int kTotalRows = LWE_N * BK_ROWS;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_base = smem_base + S::kBarOff;

    // balanced partition: CTA b owns ciphertexts [b*count/grid, (b+1)*count/grid) -- 3 or 4 each when the host sizes the grid
    // as whole waves of SMs (launch_blind_rotate), fewer for small batches so that every SM gets work
    const int first_ct = (int)((long long)blockIdx.x * count / gridDim.x);
    const int active = (int)((long long)(blockIdx.x + 1) * count / gridDim.x) - first_ct;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(bar_base + (S::kBskFull + s) * 8, 1);                // TMA transaction
            mbar_init(bar_base + (S::kBskEmpty + s) * 8, 3);               // tcgen05.commit of each copy warp
        }
        for (int s = 0; s < TS; s++) {
            mbar_init(bar_base + (S::kTmFull + s) * 8, 3);                 // tcgen05.commit of each copy warp
            mbar_init(bar_base + (S::kTmEmpty + s) * 8, active * 2);       // one arrival per back warp
        }
        for (int j = 0; j < S::kCts; j++) {
            for (int x = 0; x < XSLOTS; x++) {
                mbar_init(bar_base + (S::kXFull + j * XSLOTS + x) * 8, 1);
                mbar_init(bar_base + (S::kXEmpty + j * XSLOTS + x) * 8, 2);
            }
            mbar_init(bar_base + (S::kAccReady + j) * 8, 2);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 12) tmem_alloc_512(smem_base + S::kTmemPtrOff);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + S::kTmemPtrOff);

    if (warp >= 12) {
        // =========================================================================================== PRODUCER warps
        reg_dealloc<24>();
        if (warp == 12 && lane == 0) {
            // ---- TMA: request slab s into smem stage s % STAGES once the copy of slab s-STAGES has drained it
            const uint8_t* bsk_bytes = reinterpret_cast<const uint8_t*>(bsk_f);
#pragma unroll 1
            for (int s = 0; s < kTotalRows; s++) {
                const int ns = s % STAGES;
                if (s >= STAGES) mbar_wait_thread_long(bar_base + (S::kBskEmpty + ns) * 8, ((s - STAGES) / STAGES) & 1);
                mbar_arrive_expect_tx(bar_base + (S::kBskFull + ns) * 8, S::kStageBytes);
                tma_load_1d(smem_base + S::kStagesOff + ns * S::kStageBytes, bsk_bytes + (size_t)s * S::kStageBytes, S::kStageBytes,
                            bar_base + (S::kBskFull + ns) * 8);
            }
        } else if (warp >= 13 && lane == 0) {
            // ---- smem -> TMEM: slab c into TMEM stage c % 8 once every back warp has read slab c-8 out of it.  One
            // tcgen05.cp costs its issuing thread ~70 cycles of dependent uniform-datapath work (measured,
            // scripts/probes/tmem_probe.cu) while the copy unit itself sustains one per ~20, so the 16 copies of a slab are
            // split over the three copy warps (6 + 5 + 5); each commits onto both ring barriers (count 3).
            const int k0 = (warp == 13) ? 0 : (warp == 14 ? 6 : 11), k1 = (warp == 13) ? 6 : (warp == 14 ? 11 : 16);
#pragma unroll 1
            for (int c = 0; c < kTotalRows; c++) {
                const int cs = c % STAGES, ts = c % TS;
                mbar_wait_thread_long(bar_base + (S::kBskFull + cs) * 8, (c / STAGES) & 1);
                if (c >= TS) mbar_wait_thread_long(bar_base + (S::kTmEmpty + ts) * 8, ((c - TS) / TS) & 1);
                tc_fence_after();
                // descriptor of plane (o, x) = descriptor of the stage + its byte offset >> 4 (start-address field, no carry:
                // shared-memory addresses stay below 2^18)
                const uint64_t desc0 = tc_smem_desc(smem_base + S::kStagesOff + cs * S::kStageBytes, 128, 128);
                const uint32_t dst = tmem_base + ts * 64;
#pragma unroll 1
                for (int k = k0; k < k1; k++) {   // k = x*2 + o  <-  plane (o, x) of the slab: B[o*512 + x*64 + u]
                    const int x = k >> 1, o = k & 1;
                    tc_cp_64x128b_02_13(dst + k * 4, desc0 + (uint64_t)(o * NH + x * 64));
                }
                tc_commit(bar_base + (S::kTmFull + ts) * 8);
                tc_commit(bar_base + (S::kBskEmpty + cs) * 8);
            }
        }
    } else if (warp >= 8) {
        // =========================================================================================== FRONT warp
        reg_dealloc<FRONT_REGS>();
        const int j = warp - 8;
        if (j < active) {
            const int ct = first_ct + j;
            uint8_t* cbase = smem + S::kCtOff + j * S::kCtBytes;
            uint32_t* acc = reinterpret_cast<uint32_t*>(cbase);
            uint16_t* bara = reinterpret_cast<uint16_t*>(cbase + S::kAccBytes);
            double2* ring = reinterpret_cast<double2*>(cbase + S::kAccBytes + S::kBaraBytes);
            const uint32_t xfull = bar_base + (S::kXFull + j * XSLOTS) * 8, xempty = bar_base + (S::kXEmpty + j * XSLOTS) * 8;
            const uint32_t accready = bar_base + (S::kAccReady + j) * 8;

            // ---- modswitch and accumulator init (step 2)
            const uint32_t* lwe = lwe_in + (size_t)ct * LWE_STRIDE;
            for (int i = lane; i < LWE_N; i += 32) bara[i] = (uint16_t)modswitch_2N(__ldcg(lwe + i));
            const int barb = (int)modswitch_2N(__ldcg(lwe + LWE_N));
            for (int k = lane; k < N; k += 32) {
                acc[k] = 0;
                acc[N + k] = (((k + barb) & (2 * N - 1)) < N) ? mu : 0u - mu;   // X^{2N-barb} * (mu + mu X + ...)
            }
            __syncwarp();

            int rowc = 0;
#pragma unroll 1
            for (int i = 0; i < LWE_N; i++) {
                if (i > 0) mbar_wait_warp_long(accready, (i - 1) & 1);   // back warps have added step i-1 into the accumulator
                const int a = bara[i];
#pragma unroll 1
                for (int c = 0; c < 2; c++) {
                    uint32_t src[2][16];   // (X^a - 1)*acc_c + decomposition offset at this lane's 2 x 16 coefficients
#pragma unroll
                    for (int hf = 0; hf < 2; hf++) {
#pragma unroll
                        for (int q = 0; q < 8; q++) {
                            src[hf][2 * q] = rot_diff(acc + c * N, lane + 32 * hf + 64 * q, a) + DECOMP_OFFSET;
                            src[hf][2 * q + 1] = rot_diff(acc + c * N, lane + 32 * hf + 64 * q + NH, a) + DECOMP_OFFSET;
                        }
                    }
#pragma unroll 1
                    for (int p = 0; p < BK_L; p++) {
                        const int slot = rowc % XSLOTS;
                        if (rowc >= XSLOTS) mbar_wait_warp(xempty + slot * 8, ((rowc - XSLOTS) / XSLOTS) & 1);
                        double2* buf = ring + slot * FFT_BUF;
                        const DigitLevel dl = digit_level(p);
#pragma unroll
                        for (int hf = 0; hf < 2; hf++) {
                            double2 v[8];
#pragma unroll
                            for (int q = 0; q < 8; q++) {
                                v[q].x = digit_scaled(src[hf][2 * q], dl);
                                v[q].y = digit_scaled(src[hf][2 * q + 1], dl);
                            }
                            dft8_twiddled<-1, true>(v, [](int q) { return twist_const(q); });
#pragma unroll
                            for (int r = 0; r < 8; r++) buf[r * 64 + lane + 32 * hf] = v[r];
                        }
                        __syncwarp();
                        if (lane == 0) mbar_arrive(xfull + slot * 8);
                        rowc++;
                    }
                }
            }
            mbar_wait_warp_long(accready, (LWE_N - 1) & 1);     // the last step's accumulator update
            // ---- sample extract: a'[0]=acc_a[0], a'[k]=-acc_a[N-k], b'=acc_b[0]
            uint32_t* ext = ext_out + (size_t)ct * EXT_STRIDE;
            for (int k = lane; k < N; k += 32) ext[k] = (k == 0) ? acc[0] : 0u - acc[N - k];
            if (lane == 0) { ext[N] = acc[N]; ext[N + 1] = 0; ext[N + 2] = 0; ext[N + 3] = 0; }
        }
    } else {
        // =========================================================================================== BACK warp
        reg_alloc<BACK_REGS>();
        const int j = warp >> 1;
        if (j < active) {
            const int u = lane + 32 * (warp & 1);      // thread-column of the 64-wide transform layout
            const int lo = u & 7, rr = u >> 3;
            uint8_t* cbase = smem + S::kCtOff + j * S::kCtBytes;
            double2* ring = reinterpret_cast<double2*>(cbase + S::kAccBytes + S::kBaraBytes);
            const uint32_t xfull = bar_base + (S::kXFull + j * XSLOTS) * 8, xempty = bar_base + (S::kXEmpty + j * XSLOTS) * 8;
            const uint32_t accready = bar_base + (S::kAccReady + j) * 8;
            uint32_t* acc = reinterpret_cast<uint32_t*>(cbase);
            const uint32_t tm_lane = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16);    // this warp's TMEM lane quadrant
            Twiddles tw;
            make_twiddles(tw, u);

            int rowc = 0;
#pragma unroll 1
            for (int i = 0; i < LWE_N; i++) {
                double2 f0[8], f1[8];   // Fourier accumulators for the two output polynomials
#pragma unroll
                for (int x = 0; x < 8; x++) { f0[x] = make_double2(0.0, 0.0); f1[x] = make_double2(0.0, 0.0); }
#pragma unroll 1
                for (int row = 0; row < 2 * BK_L; row++) {
                    const int slot = rowc % XSLOTS;
                    const int ts = rowc % TS;
                    const uint32_t tmfull = bar_base + (S::kTmFull + ts) * 8;
                    const bool slab_ready = __all_sync(0xffffffffu, mbar_test(tmfull, (rowc / TS) & 1));
                    mbar_wait_warp_long(xfull + slot * 8, (rowc / XSLOTS) & 1);
                    const double2* buf = ring + slot * FFT_BUF;
                    double2 v[8];
#pragma unroll
                    for (int q2 = 0; q2 < 8; q2++) v[q2] = buf[rr * 64 + lo + 8 * q2];
                    dft8_twiddled<-1, false>(v, [&](int q) { return tw.g[q]; });
                    __syncwarp();
                    if (lane == 0) mbar_arrive(xempty + slot * 8);      // the row has been consumed into registers
                    rotate_exchange(v, lo);
                    dft8_twiddled<+1, false>(v, [&](int k) { return tw.h[k]; });

                    if (!slab_ready) mbar_wait_warp(tmfull, (rowc / TS) & 1);
                    tc_fence_after();
                    const uint32_t tsrc = tm_lane + ts * 64;
                    uint32_t ra[16], rb[16];
                    auto mac2 = [&](const uint32_t (&r)[16], auto xc) {   // slots x0, x0+1: r = (B0[x0], B1[x0], B0[x0+1], B1[x0+1])
                        constexpr int x0 = decltype(xc)::value;
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            const double2 b0 = make_double2(__hiloint2double((int)r[8 * h + 1], (int)r[8 * h + 0]),
                                                            __hiloint2double((int)r[8 * h + 3], (int)r[8 * h + 2]));
                            const double2 b1 = make_double2(__hiloint2double((int)r[8 * h + 5], (int)r[8 * h + 4]),
                                                            __hiloint2double((int)r[8 * h + 7], (int)r[8 * h + 6]));
                            const double2 w = v[x0 + h];
                            f0[x0 + h].x = fma(-w.y, b0.y, fma(w.x, b0.x, f0[x0 + h].x));
                            f0[x0 + h].y = fma(w.y, b0.x, fma(w.x, b0.y, f0[x0 + h].y));
                            f1[x0 + h].x = fma(-w.y, b1.y, fma(w.x, b1.x, f1[x0 + h].x));
                            f1[x0 + h].y = fma(w.y, b1.x, fma(w.x, b1.y, f1[x0 + h].y));
                        }
                    };
                    tmem_ld16(tsrc + 0, ra);
                    tmem_ld_wait16(ra);
                    tmem_ld16(tsrc + 16, rb);
                    mac2(ra, std::integral_constant<int, 0>{});
                    tmem_ld_wait16(rb);
                    tmem_ld16(tsrc + 32, ra);
                    mac2(rb, std::integral_constant<int, 2>{});
                    tmem_ld_wait16(ra);
                    tmem_ld16(tsrc + 48, rb);
                    mac2(ra, std::integral_constant<int, 4>{});
                    tmem_ld_wait16(rb);
                    // the slab is in registers: release the TMEM stage
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_base + (S::kTmEmpty + ts) * 8);
                    mac2(rb, std::integral_constant<int, 6>{});
                    rowc++;
                }
                // ---- inverse transforms, round to nearest, accumulate into acc (exact integers mod 2^32); see the ws kernel
                group_sync(j);      // both back warps are done reading the step's forward rows
                auto inverse_head = [&](double2 (&f)[8], int poly) {
                    dft8<+1>(f);
                    f[0] = cmul_conj(f[0], tw.h[0]);
#pragma unroll
                    for (int k = 1; k < 8; k++) f[k] = cmul_conj(f[k], tw.h[8 - k]);
                    rotate_exchange(f, lo);
                    dft8<-1>(f);
                    double2* buf = ring + poly * FFT_BUF;
#pragma unroll
                    for (int q2 = 0; q2 < 8; q2++) buf[rr * 64 + lo + 8 * q2] = cmul_conj(f[q2], tw.g[q2]);
                };
                inverse_head(f0, 0);
                inverse_head(f1, 1);
                group_sync(j);
#pragma unroll 1
                for (int poly = 0; poly < 2; poly++) {
                    const double2* buf = ring + poly * FFT_BUF;
                    double2 v[8];
#pragma unroll
                    for (int r = 0; r < 8; r++) v[r] = buf[r * 64 + u];
                    dft8<+1>(v);
                    v[0] = make_double2(v[0].x * (1.0 / 512.0), v[0].y * (1.0 / 512.0));
#pragma unroll
                    for (int q = 1; q < 8; q++) {
                        double2 w = twist_const(q);
                        w.x *= (1.0 / 512.0); w.y *= (1.0 / 512.0);
                        v[q] = cmul_conj(v[q], w);
                    }
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        acc[poly * N + u + 64 * q] += (uint32_t)__double2ll_rn(v[q].x);
                        acc[poly * N + u + 64 * q + NH] += (uint32_t)__double2ll_rn(v[q].y);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(accready);
            }
        }
    }
    // ---- common tail: TMEM is released by the warp that allocated it once every role is done with it
    tc_fence_before();
    __syncthreads();
    if (warp == 12) {
        tc_fence_after();
        tmem_dealloc_512(tmem_base);
    }
}

}  // namespace rs
