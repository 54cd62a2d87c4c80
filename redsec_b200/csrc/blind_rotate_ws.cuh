// redsec_b200/csrc/blind_rotate_ws.cuh -- warp-specialised batched blind rotation + sample extract for sm_100a.
//
// Same computation as blind_rotate.cuh (the role of tfhe_bootstrap_FFT at lib/BinOps_enc.cpp:185 /
// redcufhe::Bootstrap at lib/GPU/gates.cu:124-130 for a whole batch), re-cut as a three-stage pipeline so that every SM
// sub-partition has THREE resident warps with different phases instead of two in near lock-step:
//
//   PRODUCER warp: TMA (cp.async.bulk)  --BSK ring-->  BACK warps  <--exchange ring--  FRONT warps
//
// One CTA = 4 warpgroups, 4 ciphertexts (the default, PRODUCER = true; 512 threads):
//   * warpgroup 3 = the BSK producer: one lane of warp 12 streams the 16 KiB Fourier-key slabs in order through the STAGES-deep
//     ring (wait bsk_empty of the stage's previous occupant, expect_tx, cp.async.bulk).  setmaxnreg is a warpgroup instruction,
//     so a whole warpgroup is launched for it; warps 13-15 shrink to 24 registers and leave.
//   * warpgroup 2 = 4 FRONT warps, one per ciphertext.  A front warp owns the torus32 accumulator (shared memory): per
//     blind-rotate step it forms (X^a - 1)*acc, gadget-decomposes it, and for each of the 20 (polynomial, level) rows
//     runs pass 1 of the forward transform (twist + radix-8) for all 64 thread-columns (two halves of 32) and writes the
//     result into a 3-slot exchange ring.  It is the lighter role on purpose: the ring stays full and the back warps, which
//     carry the FP64 bulk, rarely wait (measured: only at the first row of a step, profiles/r2_ws_phase_timers.txt).
//   * warpgroups 0,1 = 8 BACK warps, two per ciphertext.  A back warp reads a row from the exchange ring, runs passes 2
//     and 3 (twiddles fused as FMAs, exchange 2 through shuffles), multiplies by the BSK slab and accumulates in
//     registers (Fourier accumulators, 64 registers).  After 20 rows the pair runs both inverse transforms (their
//     exchange 1 goes through the ring slot of the step's last row, released afterwards), rounds, adds into the
//     accumulator and signals acc_ready per polynomial, polynomial 0 first, so the front warp restarts early.
//   The two back warps of a ciphertext never synchronise with each other during the 20 rows (exchange 1 is now
//   front -> back, exchange 2 is intra-warp); all hand-offs are mbarriers, so the roles drift freely within the rings.
//   * Registers are redistributed with setmaxnreg: launch at 128/thread (16 warps = the whole file), producer warpgroup
//     drops to 24, front warps to 120, back warps grow to 184: 8*32*184 + 4*32*120 + 4*32*24 = 65536.
//
// PRODUCER = false is the round-1 shape (12 warps, launch at 168, back 192 / front 120), kept behind RS_WS_PRODUCER=0 for A/B
// runs: there is no producer warp and the front warps request the slabs, in order, through a shared claim counter at the end of
// each of their rows.  That duty cost a front warp ~520 of its ~2200 cycles per row and made it the critical path of a lone
// ciphertext; the producer warpgroup costs the back warps 8 registers and nothing else.  Measured, same inputs, bit-identical
// outputs (profiles/r2_ws_producer_ab.log): 2^16 bootstraps 842 -> 791 ms, one wave of 592 7.75 -> 7.18 ms, a single
// ciphertext 3.79 -> 3.33 ms.
//
// Row-split mode (SPLIT = 2; 4 exists but is slower, see api.cu) for batches below two ciphertexts per SM, which are
// latency-bound (a ciphertext alone on an SM needs 6.1 ms un-split: 7 000 rows one after the other): the 4 slots of a CTA
// then hold up to 4/SPLIT ciphertexts, and the SPLIT slots of a ciphertext take its rows r = k (mod SPLIT) of every step,
// each with its own front warp, back-warp pair, exchange ring and partial Fourier accumulators.  After the rows of a step
// the partial sums meet in shared memory; part 0 adds up polynomial 0 and part 1 polynomial 1 and each runs one inverse
// transform.  The summation order differs from the un-split kernel, the result does not: the inverse transform is rounded
// to the exact integer convolution either way (pre-rounding error ~1e-3 against the 0.5 bound), so ciphertexts stay
// bit-identical (tests/test_gpu_pbs.py compares counts 1..300 with the oracle: split 2 up to 296, un-split above).
//
// See fft512.cuh for the transform algebra; the arithmetic per row is identical to blind_rotate_kernel, so results are
// bit-identical (tests/test_gpu_pbs.py::test_variants_agree).
#pragma once
#include "blind_rotate.cuh"

namespace rs {

#ifdef RS_WS_PROF
// per-warp phase timers of CTA 0 (debug builds only: RS_NVCC_EXTRA=-DRS_WS_PROF, read with scripts/ws_prof.py):
// [warp][phase] accumulated clock64 cycles.  Results: profiles/r1_ws_phase_timers.txt
__device__ long long g_ws_prof[12][8];
// where in a step the hand-off waits fall (warp 0 = back, warp 8 = front, CTA 0): [0][row] back warp waiting for row `row` of a step,
// [1][row] front warp waiting for a free ring slot before row `row`, [2][c] front warp waiting for accumulator polynomial c
__device__ long long g_ws_rowwait[3][20];
#define WSP_ROWWAIT_BEGIN long long wsp_rw = clock64()
#define WSP_ROWWAIT_END(kind, idx) do { if (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 8)) g_ws_rowwait[kind][idx] += clock64() - wsp_rw; } while (0)
#define WSP_DECL long long wsp_t = clock64(), wsp_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define WSP(phase) do { const long long n_ = clock64(); wsp_acc[phase] += n_ - wsp_t; wsp_t = n_; } while (0)
#define WSP_FLUSH() do { if (blockIdx.x == 0 && lane == 0) for (int q_ = 0; q_ < 8; q_++) g_ws_prof[warp][q_] = wsp_acc[q_]; } while (0)
#else
#define WSP_DECL
#define WSP(phase) do { } while (0)
#define WSP_FLUSH() do { } while (0)
#define WSP_ROWWAIT_BEGIN do { } while (0)
#define WSP_ROWWAIT_END(kind, idx) do { } while (0)
#endif

#ifdef RS_ERR_STATS
// pre-rounding error of the FFT external product (debug builds only: RS_NVCC_EXTRA=-DRS_ERR_STATS, scripts/phase_error.py):
// histogram of |x - rint(x)| over every inverse-transform output, bins [0,1e-6) [1e-6,1e-5) ... [1e-2,1e-1) [1e-1,0.5], and the maximum
__device__ unsigned long long g_err_hist[8];
__device__ unsigned long long g_err_max_bits;
__device__ __forceinline__ void err_stat(double x) {
    const double e = fabs(x - rint(x));
    const int bin = e < 1e-6 ? 0 : e < 1e-5 ? 1 : e < 1e-4 ? 2 : e < 1e-3 ? 3 : e < 1e-2 ? 4 : e < 1e-1 ? 5 : 6;
    atomicAdd(&g_err_hist[bin], 1ull);
    atomicMax(&g_err_max_bits, (unsigned long long)__double_as_longlong(e));
}
#else
__device__ __forceinline__ void err_stat(double) {}
#endif

template <int STAGES, int XSLOTS>
struct WsSmem {
    static constexpr int kCts = 4;
    static constexpr int kStageBytes = (int)BSK_ROW_BYTES;                 // 16 KiB BSK slab
    static constexpr int kAccBytes = 2 * N * 4;                            // 8 KiB
    static constexpr int kBaraBytes = 352 * 2;
    static constexpr int kSlotBytes = FFT_BUF * 16;                        // 8 KiB exchange slot
    static constexpr int kCtBytes = kAccBytes + kBaraBytes + XSLOTS * kSlotBytes;
    static constexpr int kStagesOff = 0;
    static constexpr int kCtOff = STAGES * kStageBytes;
    static constexpr int kBarOff = kCtOff + kCts * kCtBytes;
    // barriers (8 B each): bsk_full[STAGES], bsk_empty[STAGES], x_full[4][XSLOTS], x_empty[4][XSLOTS], acc_ready[4][2]
    static constexpr int kBskFull = 0, kBskEmpty = STAGES, kXFull = 2 * STAGES, kXEmpty = kXFull + kCts * XSLOTS,
                         kAccReady = kXEmpty + kCts * XSLOTS, kNumBars = kAccReady + 2 * kCts;   // acc_ready[4][2]: one per accumulator polynomial
    static constexpr int kIssuedOff = kBarOff + kNumBars * 8;
    static constexpr int kStageSeqOff = kIssuedOff + 8;                    // int[STAGES]: last slab issued into each ring stage (row-split modes)
    static constexpr int kTotal = kStageSeqOff + ((STAGES * 4 + 7) & ~7);
    static_assert(kCtBytes % 16 == 0, "ciphertext block must stay 16-byte aligned");
    static_assert(XSLOTS >= 3, "the inverse holds one slot; the front warp needs two more to run ahead");
};

template <int N_REGS>
__device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N_REGS)); }
template <int N_REGS>
__device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N_REGS)); }

template <int STAGES, int XSLOTS, int SPLIT, bool STRESS = false, bool PRODUCER = false>
__global__ void __launch_bounds__(PRODUCER ? 512 : 384, 1)
blind_rotate_ws_kernel(const uint32_t* __restrict__ lwe_in,    // [count][LWE_STRIDE]
                       int count, uint32_t mu,
                       const double2* __restrict__ bsk_f,       // [n][BK_ROWS][2][NH]
                       uint32_t* __restrict__ ext_out,          // [count][EXT_STRIDE]
                       float l2_keep,                           // fraction of the BSK stream marked L2 evict_last (0 = no hint)
                       const uint32_t* __restrict__ lut,        // [lut_mod][N] test vectors (ciphertext c uses row c % lut_mod), or
                       int lut_mod,                             // nullptr: the constant test vector mu of the sign bootstrap
                       int lookahead,                           // producer warp: slabs requested ahead of the slowest consumer (1..STAGES)
                       unsigned* __restrict__ wave_done,        // wave gate (see below): ciphertexts finished so far in this launch, or nullptr
                       int wave_ctas, int gate_every)           // CTAs per wave (= SMs), gate every gate_every-th wave
{
    using S = WsSmem<STAGES, XSLOTS>;
    static_assert(SPLIT == 1 || SPLIT == 2 || SPLIT == 4, "a ciphertext is spread over 1, 2 or 4 slots");
    constexpr int AHEAD = 1;                       // a front warp that has produced the row of slab g makes sure slabs <= g+AHEAD+1 are requested
    constexpr int kTotalRows = LWE_N * BK_ROWS;
    constexpr int kRowsPerPart = BK_ROWS / SPLIT;  // rows a slot handles per blind-rotate step
    static_assert(STAGES > AHEAD + XSLOTS, "BSK ring must cover the front-to-back distance");
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_base = smem_base + S::kBarOff;
    int* issued = reinterpret_cast<int*>(smem + S::kIssuedOff);
    volatile int* stage_seq = reinterpret_cast<volatile int*>(smem + S::kStageSeqOff);

    // balanced partition: CTA b owns ciphertexts [b*count/grid, (b+1)*count/grid): 4 each (fewer in the last CTAs) for the
    // usual grid of ceil(count/4); at most 4/SPLIT each in the row-split modes, where the host spreads a small batch over all SMs
    const int first_ct = (int)((long long)blockIdx.x * count / gridDim.x);
    const int active = (int)((long long)(blockIdx.x + 1) * count / gridDim.x) - first_ct;   // ciphertexts of this CTA
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // Wave gate (un-split launches of many waves; OFF by default, RS_WS_GATE=n).  All CTAs walk the 115 MB Fourier key from slab 0; as
    // long as they do it together every slab leaves HBM once per wave and the other 147 CTAs find it in L2.  Nothing keeps them
    // together: the producer warp hides a CTA's own L2 misses and SMs finish their CTAs ~1 % apart, which is enough to spread the
    // CTAs of the next wave over more key than the L2 holds -- measured on 2^16-ciphertext launches: 73-220 GB of DRAM reads instead
    // of 7 (profiles/r2_traffic_ab.txt).  With the gate the CTAs of every n-th wave start only when all ciphertexts of the earlier waves
    // are finished (a counter the front warps bump), which lines the SMs up again: n = 1 restores 6.8 GB and a 99 % L2 hit rate and
    // costs 1.2 % of time (every SM waits for the slowest); n = 4 already loses most of the effect.  The launch is FP64-bound and HBM is
    // 4 % busy at worst, so time wins and the gate stays off.  The wait is bounded (200 us): a gate that cannot be met -- SMs shared
    // with another kernel -- is skipped, never a hang.
    if (SPLIT == 1 && wave_done != nullptr && threadIdx.x == 0) {
        const int wave = blockIdx.x / wave_ctas;
        if (wave > 0 && wave % gate_every == 0) {
            const unsigned need = (unsigned)((long long)wave * wave_ctas * count / gridDim.x);
            const long long t0 = clock64();
            while (*reinterpret_cast<volatile unsigned*>(wave_done) < need && clock64() - t0 < 400000LL) __nanosleep(100);
        }
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(bar_base + (S::kBskFull + s) * 8, 1);
            mbar_init(bar_base + (S::kBskEmpty + s) * 8, active * 2);      // one arrival per back warp that consumes the slab
        }
        for (int j = 0; j < S::kCts; j++) {
            for (int x = 0; x < XSLOTS; x++) {
                mbar_init(bar_base + (S::kXFull + j * XSLOTS + x) * 8, 1);   // the front warp
                mbar_init(bar_base + (S::kXEmpty + j * XSLOTS + x) * 8, 2);  // the two back warps
            }
            mbar_init(bar_base + (S::kAccReady + 2 * j) * 8, 2);           // the two back warps, accumulator polynomial 0
            mbar_init(bar_base + (S::kAccReady + 2 * j + 1) * 8, 2);       // ... polynomial 1
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        *issued = 0;
        for (int s = 0; s < STAGES; s++) stage_seq[s] = -1;
    }
    __syncthreads();

    if (PRODUCER && warp >= 12) {
        // =========================================================================================== PRODUCER warpgroup (16-warp build)
        // One lane of warp 12 streams the BSK slabs in order through the ring: no claim, no shared counter, and the front warps
        // lose their producer duty.  setmaxnreg is a warpgroup instruction, so the build launches a whole fourth warpgroup (warps
        // 13-15 leave at once).  Register pool: 512 x 128 = 65 536 = 8 x 32 x 184 (back) + 4 x 32 x 120 (front) + 4 x 32 x 24.
        reg_dealloc<24>();
        if (warp == 12 && lane == 0) {
            const uint8_t* bsk_bytes = reinterpret_cast<const uint8_t*>(bsk_f);
            const uint64_t l2pol = l2_policy_evict_last(l2_keep);
#pragma unroll 1
            for (int cur = 0; cur < kTotalRows; cur++) {
                const int ns = cur % STAGES;
                // slab cur goes out once every consumer has released slab cur - lookahead.  lookahead = STAGES uses the whole ring;
                // a shorter look-ahead leaves a CTA that runs ahead of the others exposed to its own L2 misses, which keeps the
                // CTAs of a long launch walking the key together (BSK streaming).  The barrier of slab cur - lookahead
                // cannot be a phase ahead of the one waited for: its stage's next occupant, cur - lookahead + STAGES, is not out yet.
                const int rel = cur - (SPLIT == 1 ? lookahead : STAGES);      // the row-split modes work on SPLIT slabs at once: whole ring
                if (rel >= 0) mbar_wait_thread(bar_base + (S::kBskEmpty + rel % STAGES) * 8, (rel / STAGES) & 1);
                mbar_arrive_expect_tx(bar_base + (S::kBskFull + ns) * 8, S::kStageBytes);
                if (l2_keep > 0.f)
                    tma_load_1d_hint(smem_base + S::kStagesOff + ns * S::kStageBytes, bsk_bytes + (size_t)cur * S::kStageBytes,
                                     S::kStageBytes, bar_base + (S::kBskFull + ns) * 8, l2pol);
                else
                    tma_load_1d(smem_base + S::kStagesOff + ns * S::kStageBytes, bsk_bytes + (size_t)cur * S::kStageBytes,
                                S::kStageBytes, bar_base + (S::kBskFull + ns) * 8);
            }
        }
        return;
    }
    if (warp >= 8) {
        // =========================================================================================== FRONT warp
        reg_dealloc<120>();
        const int j = warp - 8;            // slot
        const int jc = j / SPLIT;          // ciphertext of this CTA the slot works for
        const int part = j % SPLIT;        // rows r = part (mod SPLIT) of every step
        if (jc >= active) return;
        const int ct = first_ct + jc;
        uint8_t* cbase = smem + S::kCtOff + j * S::kCtBytes;
        uint8_t* pbase = smem + S::kCtOff + (jc * SPLIT) * S::kCtBytes;      // part 0's slot holds the accumulator
        uint32_t* acc = reinterpret_cast<uint32_t*>(pbase);
        uint16_t* bara = reinterpret_cast<uint16_t*>(cbase + S::kAccBytes);
        double2* ring = reinterpret_cast<double2*>(cbase + S::kAccBytes + S::kBaraBytes);
        const uint32_t xfull = bar_base + (S::kXFull + j * XSLOTS) * 8, xempty = bar_base + (S::kXEmpty + j * XSLOTS) * 8;
        const uint32_t accready = bar_base + (S::kAccReady + 2 * (jc * SPLIT)) * 8;   // [2]: per accumulator polynomial
        const uint8_t* bsk_bytes = reinterpret_cast<const uint8_t*>(bsk_f);
        const uint64_t l2pol = l2_policy_evict_last(l2_keep);

        // ---- modswitch and accumulator init (step 2)
        const uint32_t* lwe = lwe_in + (size_t)ct * LWE_STRIDE;
        for (int i = lane; i < LWE_N; i += 32) bara[i] = (uint16_t)modswitch_2N(__ldcg(lwe + i));
        if (part == 0) {
            const int barb = (int)modswitch_2N(__ldcg(lwe + LWE_N));
            const uint32_t* tv = lut ? lut + (size_t)(ct % lut_mod) * N : nullptr;
            for (int k = lane; k < N; k += 32) {
                acc[k] = 0;
                const int idx = (k + barb) & (2 * N - 1);                        // X^{2N-barb} * testvector
                const uint32_t v = tv ? __ldcg(tv + (idx & (N - 1))) : mu;                  // (mu + mu X + ... for the sign bootstrap)
                acc[N + k] = idx < N ? v : 0u - v;
            }
        }
        __syncwarp();
        if (SPLIT > 1) asm volatile("bar.sync %0, %1;" ::"r"(8 + jc), "n"(32 * SPLIT) : "memory");   // the ciphertext's front warps: acc is initialised

        // ---- BSK producer duty: slabs are claimed in order by whichever front warp gets here first.  It runs at the END of a row,
        // after the row has been handed to the back warps: the claim is a chain of shared-memory round trips (counter, CAS, stage
        // barrier, expect_tx, bulk copy) of ~700 cycles that would otherwise delay every row by as much
        auto request_slabs = [&](int upto) {
            if (PRODUCER) return;      // the 13-warp build has a warp for this
            if (lane == 0) {
                const int want = min(upto, kTotalRows - 1);
                int cur = *reinterpret_cast<volatile int*>(issued);
                while (cur <= want) {
                    const int prev = atomicCAS(issued, cur, cur + 1);
                    if (prev == cur) {
                        const int ns = cur % STAGES;
                        if (cur >= STAGES) {
                            // Row-split modes: a slot's rows are SPLIT slabs apart, so a claim can run STAGES or more slabs ahead
                            // of a lagging back-warp pair while the claim of this stage's PREVIOUS occupant (cur - STAGES) is
                            // still parked on another front warp.  A parity wait on the empty barrier would then alias (the
                            // barrier is two phases behind and the wanted parity equals that of an already completed phase) and
                            // the TMA would overwrite a slab that is still being read.  Issue per stage strictly in order: wait
                            // until the previous occupant has been ISSUED; from then on the barrier is at most one phase behind.
                            // Un-split, a front warp never claims beyond XSLOTS + AHEAD + 1 <= STAGES slabs past its own back
                            // warps (static_assert above), which rules the situation out, so the default path pays nothing.
                            if (SPLIT > 1) {
                                const long long t0 = clock64();
                                while (stage_seq[ns] != cur - STAGES)
                                    if (clock64() - t0 > 4000000000LL) __trap();
                                __threadfence_block();
                            }
                            mbar_wait_thread(bar_base + (S::kBskEmpty + ns) * 8, ((cur - STAGES) / STAGES) & 1);
                        }
                        mbar_arrive_expect_tx(bar_base + (S::kBskFull + ns) * 8, S::kStageBytes);
                        if (l2_keep > 0.f)
                            tma_load_1d_hint(smem_base + S::kStagesOff + ns * S::kStageBytes, bsk_bytes + (size_t)cur * S::kStageBytes,
                                             S::kStageBytes, bar_base + (S::kBskFull + ns) * 8, l2pol);
                        else
                            tma_load_1d(smem_base + S::kStagesOff + ns * S::kStageBytes, bsk_bytes + (size_t)cur * S::kStageBytes,
                                        S::kStageBytes, bar_base + (S::kBskFull + ns) * 8);
                        if (SPLIT > 1) { __threadfence_block(); stage_seq[ns] = cur; }
                        cur++;
                    } else {
                        cur = prev;
                    }
                }
            }
            __syncwarp();
        };
        request_slabs(part + AHEAD);     // the first row's slab (and its successor) before the pipeline starts

        int rowc = 0;    // rows produced by this front warp (SPLIT == 1: also the BSK slab index of the row)
        WSP_DECL;
#pragma unroll 1
        for (int i = 0; i < LWE_N; i++) {
            const int a = bara[i];
#pragma unroll 1
            for (int c = 0; c < 2; c++) {
                WSP(7);
                // the back warps have added step i-1 into accumulator polynomial c.  Polynomial 0 is released first, so the rows of
                // c = 0 are produced while the back warps still run the inverse transform of polynomial 1 (no bubble at the step boundary)
                { WSP_ROWWAIT_BEGIN;
                if (i > 0) mbar_wait_warp_long(accready + c * 8, (i - 1) & 1);
                WSP_ROWWAIT_END(2, c); }
                WSP(0);
                uint32_t src[2][16];   // (X^a - 1)*acc_c + decomposition offset at this lane's 2 x 16 coefficients
#pragma unroll
                for (int hf = 0; hf < 2; hf++) {
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        src[hf][2 * q] = rot_diff(acc + c * N, lane + 32 * hf + 64 * q, a) + DECOMP_OFFSET;
                        src[hf][2 * q + 1] = rot_diff(acc + c * N, lane + 32 * hf + 64 * q + NH, a) + DECOMP_OFFSET;
                    }
                }
                WSP(1);
#pragma unroll 1
                for (int p = 0; p < BK_L; p++) {
                    if (SPLIT > 1 && ((c * BK_L + p) % SPLIT) != part) continue;      // another slot's row
                    const int slab = SPLIT > 1 ? i * BK_ROWS + c * BK_L + p : rowc;     // BSK slab of this row
                    WSP(2);
                    // ---- ring slot: wait until both back warps have read its previous occupant
                    const int slot = rowc % XSLOTS;
                    { WSP_ROWWAIT_BEGIN;
                    if (rowc >= XSLOTS) mbar_wait_warp(xempty + slot * 8, ((rowc - XSLOTS) / XSLOTS) & 1);
                    WSP_ROWWAIT_END(1, c * BK_L + p); }
                    WSP(3);
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
                    WSP(4);
                    request_slabs(slab + AHEAD + 1);      // the next row's slab and the one after it
                    WSP(2);
                }
            }
        }
        WSP_FLUSH();
        if (part != 0) return;
        mbar_wait_warp_long(accready, (LWE_N - 1) & 1);     // the last step's accumulator update, both polynomials
        mbar_wait_warp_long(accready + 8, (LWE_N - 1) & 1);
        // ---- sample extract: a'[0]=acc_a[0], a'[k]=-acc_a[N-k], b'=acc_b[0]
        uint32_t* ext = ext_out + (size_t)ct * EXT_STRIDE;
        for (int k = lane; k < N; k += 32) ext[k] = (k == 0) ? acc[0] : 0u - acc[N - k];
        if (lane == 0) { ext[N] = acc[N]; ext[N + 1] = 0; ext[N + 2] = 0; ext[N + 3] = 0; }
        __threadfence();      // the extracted sample is read by the keyswitch launch that follows (RS_END_FENCE, lwe_kernels.cuh)
        if (SPLIT == 1 && wave_done != nullptr && lane == 0) atomicAdd(wave_done, 1u);
        return;
    }

    // =============================================================================================== BACK warp
    reg_alloc<PRODUCER ? 184 : 192>();
    const int j = warp >> 1;               // slot
    const int jc = j / SPLIT;
    const int part = j % SPLIT;
    if (jc >= active) return;
    const int u = lane + 32 * (warp & 1);      // thread-column of the 64-wide transform layout
    const int lo = u & 7, rr = u >> 3;
    uint8_t* cbase = smem + S::kCtOff + j * S::kCtBytes;
    uint8_t* pbase = smem + S::kCtOff + (jc * SPLIT) * S::kCtBytes;
    double2* ring = reinterpret_cast<double2*>(cbase + S::kAccBytes + S::kBaraBytes);
    const uint32_t xfull = bar_base + (S::kXFull + j * XSLOTS) * 8, xempty = bar_base + (S::kXEmpty + j * XSLOTS) * 8;
    const uint32_t accready = bar_base + (S::kAccReady + 2 * (jc * SPLIT)) * 8;   // [2]: per accumulator polynomial
    uint32_t* acc = reinterpret_cast<uint32_t*>(pbase);
    Twiddles tw;
    make_twiddles(tw, u);

    int rowc = 0;               // rows consumed by this warp
    uint32_t row_ready = 0;     // non-blocking test of the NEXT row's exchange slot, issued one row early (see below)
    WSP_DECL;
#pragma unroll 1
    for (int i = 0; i < LWE_N; i++) {
        double2 f0[8], f1[8];   // Fourier accumulators for the two output polynomials
#pragma unroll
        for (int x = 0; x < 8; x++) { f0[x] = make_double2(0.0, 0.0); f1[x] = make_double2(0.0, 0.0); }
#pragma unroll 1
        for (int row = 0; row < kRowsPerPart; row++) {
            const int slot = rowc % XSLOTS;
            const int slab = SPLIT > 1 ? i * BK_ROWS + row * SPLIT + part : rowc;    // this slot's row r = row*SPLIT + part of step i
            const int s = slab % STAGES;
            // stress instantiation (tests only, RS_WS_STRESS=1): one back-warp pair lags by ~a row per row, the situation the
            // in-order slab issue above exists for
            if (STRESS && j == 0 && (i & 3) != 3) __nanosleep(1500);
            // test the slab's barrier now and consume the answer after the transform (mbarrier round trip off the critical path)
            const bool slab_ready = __all_sync(0xffffffffu, mbar_test(bar_base + (S::kBskFull + s) * 8, (slab / STAGES) & 1));
            // the row's exchange slot was tested before the previous row's MAC; only a miss pays the mbarrier round trip here
            WSP(7);
            { WSP_ROWWAIT_BEGIN;
            if (!__all_sync(0xffffffffu, row_ready)) mbar_wait_warp_long(xfull + slot * 8, (rowc / XSLOTS) & 1);
            WSP_ROWWAIT_END(0, row); }
            WSP(0);
            const double2* buf = ring + slot * FFT_BUF;
            double2 v[8];
#pragma unroll
            for (int q2 = 0; q2 < 8; q2++) v[q2] = buf[rr * 64 + lo + 8 * q2];
            dft8_twiddled<-1, false>(v, [&](int q) { return tw.g[q]; });
            __syncwarp();
            // the row has been consumed into registers.  The LAST row of a step keeps its slot: the inverse transforms below use it
            // as their exchange buffer and release it afterwards, so the front warp can refill the other slots meanwhile
            if (lane == 0 && row != kRowsPerPart - 1) mbar_arrive(xempty + slot * 8);
            rotate_exchange(v, lo);
            // the first BSK operands are requested before pass 3 when the slab is already resident (the common case), so that
            // their shared-memory latency overlaps the butterflies instead of stalling the first FMA of the MAC
            const double2* B = reinterpret_cast<const double2*>(smem + S::kStagesOff + s * S::kStageBytes) + u;
            double2 b0, b1;
            if (slab_ready) { b0 = B[0]; b1 = B[NH]; }
            dft8_twiddled<+1, false>(v, [&](int k) { return tw.h[k]; });

            WSP(1);
            if (!slab_ready) {
                mbar_wait_warp(bar_base + (S::kBskFull + s) * 8, (slab / STAGES) & 1);
                b0 = B[0]; b1 = B[NH];
            }
            WSP(2);
            row_ready = mbar_test(xfull + ((rowc + 1) % XSLOTS) * 8, ((rowc + 1) / XSLOTS) & 1);   // consumed at the next row's start
#pragma unroll
            for (int x = 0; x < 8; x++) {
                double2 n0, n1;
                if (x < 7) { n0 = B[(x + 1) * 64]; n1 = B[NH + (x + 1) * 64]; }
                f0[x].x = fma(-v[x].y, b0.y, fma(v[x].x, b0.x, f0[x].x));
                f0[x].y = fma(v[x].y, b0.x, fma(v[x].x, b0.y, f0[x].y));
                f1[x].x = fma(-v[x].y, b1.y, fma(v[x].x, b1.x, f1[x].x));
                f1[x].y = fma(v[x].y, b1.x, fma(v[x].x, b1.y, f1[x].y));
                if (x < 7) { b0 = n0; b1 = n1; }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_base + (S::kBskEmpty + s) * 8);
            rowc++;
            WSP(3);
        }
        // ---- inverse transforms, round to nearest, accumulate into acc (exact integers mod 2^32).  Exchange 1 of the
        // inverse crosses the two back warps: it goes through the ring slot of the step's last forward row, which this pair
        // has not released yet.  Polynomial 0 is finished and published first (acc_ready[0]) so that the front warp produces
        // the next step's first rows into the two free slots while polynomial 1 is still being transformed.
        double2* ibuf = ring + ((rowc - 1) % XSLOTS) * FFT_BUF;
        if (SPLIT > 1) {
            // Row-split: the slots of a ciphertext hold partial sums.  Part 0 owns polynomial 0 and part 1 polynomial 1 (both inverse
            // transforms run at the same time).  Every part parks the partial sums it does not own in its own ring: polynomial 1
            // (polynomial 0 for part 1) in the slot BEFORE the held one -- that row was released, but the front warp cannot refill it
            // before it has seen acc_ready, which the owners signal only after both have read every partial sum -- and, for parts
            // 2 and 3, polynomial 0 in the held slot.  Then the parts meet the two owners at a named barrier.
            double2* pbuf = ring + ((rowc - 2 + XSLOTS) % XSLOTS) * FFT_BUF;
            group_sync(j);      // both back warps of the slot are done reading the step's forward rows
#pragma unroll
            for (int x = 0; x < 8; x++) {
                if (part != 1) pbuf[x * 64 + u] = f1[x]; else pbuf[x * 64 + u] = f0[x];
                if (part >= 2) ibuf[x * 64 + u] = f0[x];
            }
            __threadfence_block();
            if (part >= 2) {
                asm volatile("bar.arrive %0, %1;" ::"r"(12 + jc), "n"(64 * SPLIT) : "memory");
                if (lane == 0) mbar_arrive(xempty + ((rowc - 1) % XSLOTS) * 8);   // refills are ordered by acc_ready (see above)
                continue;
            }
            asm volatile("bar.sync %0, %1;" ::"r"(12 + jc), "n"(64 * SPLIT) : "memory");
            const int hslot = (rowc - 1) % XSLOTS, pslot = (rowc - 2 + XSLOTS) % XSLOTS;   // every part has consumed the same number of rows
#pragma unroll 1
            for (int k = 0; k < SPLIT; k++) {
                if (k == part) continue;
                const double2* oring = reinterpret_cast<const double2*>(pbase + k * S::kCtBytes + S::kAccBytes + S::kBaraBytes);
                // part 0 collects polynomial 0: parked in the pslot of part 1, in the held slot of parts 2, 3;
                // part 1 collects polynomial 1: parked in the pslot of every other part
                const double2* o = oring + ((part == 0 && k >= 2) ? hslot : pslot) * FFT_BUF;
                if (part == 0) {
#pragma unroll
                    for (int x = 0; x < 8; x++) { const double2 t = o[x * 64 + u]; f0[x].x += t.x; f0[x].y += t.y; }
                } else {
#pragma unroll
                    for (int x = 0; x < 8; x++) { const double2 t = o[x * 64 + u]; f1[x].x += t.x; f1[x].y += t.y; }
                }
            }
        }
        auto inverse_poly = [&](double2 (&f)[8], int poly) {
            group_sync(j);      // both back warps are done reading ibuf (the last forward row / the previous polynomial)
            dft8<+1>(f);
            f[0] = cmul_conj(f[0], tw.h[0]);
#pragma unroll
            for (int k = 1; k < 8; k++) f[k] = cmul_conj(f[k], tw.h[8 - k]);
            rotate_exchange(f, lo);
            dft8<-1>(f);
#pragma unroll
            for (int q2 = 0; q2 < 8; q2++) ibuf[rr * 64 + lo + 8 * q2] = cmul_conj(f[q2], tw.g[q2]);
            group_sync(j);
            double2 v[8];
#pragma unroll
            for (int r = 0; r < 8; r++) v[r] = ibuf[r * 64 + u];
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
                err_stat(v[q].x); err_stat(v[q].y);
                acc[poly * N + u + 64 * q] += (uint32_t)__double2ll_rn(v[q].x);
                acc[poly * N + u + 64 * q + NH] += (uint32_t)__double2ll_rn(v[q].y);
            }
            __syncwarp();
            if (SPLIT == 1 && lane == 0) mbar_arrive(accready + poly * 8);
        };
        if (SPLIT == 1) {
            inverse_poly(f0, 0);
            inverse_poly(f1, 1);
        } else {
            if (part == 0) inverse_poly(f0, 0); else inverse_poly(f1, 1);
            // both owners have read every parked partial sum and updated their accumulator polynomial: only now may the front warps
            // of the ciphertext refill the rings
            asm volatile("bar.sync %0, %1;" ::"r"(14 + jc), "n"(128) : "memory");
            if (lane == 0) mbar_arrive(accready + part * 8);
        }
        if (lane == 0) mbar_arrive(xempty + ((rowc - 1) % XSLOTS) * 8);    // the last row's slot, held for the inverse exchange
        WSP(4);
    }
    WSP_FLUSH();
}

}  // namespace rs
