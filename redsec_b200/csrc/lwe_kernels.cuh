// redsec_b200/csrc/lwe_kernels.cuh -- integer LWE kernels: keyswitch, gate pre-combination, ternary linear layers.
// All arithmetic is uint32 wrap-around (torus32), so results are bit-exact by construction.
// Loads of data that an EARLIER KERNEL wrote (ciphertext rows, extracted samples) use __ldcg (ld.global.cg: cached in L2 only).
// The compiler turns plain loads through const __restrict__ pointers into LDG.E.CONSTANT (the non-coherent L1 path), which is
// only coherent at kernel boundaries of an otherwise idle SM: with several lanes in flight (rs_lanes) an SM runs CTAs of other
// kernels back to back without its L1 being invalidated, and a buffer that is rewritten between two launches (the gate
// pre-combination scratch, the extracted-sample scratch) was then read stale by a later launch landing on the same SM.
// The same holds for the read-modify-write kernels (bias / constant add): a line they pull into an L1 can still be there when a
// LIBRARY kernel (NCCL's all-gather of a ciphertext buffer) later reads that address after other kernels rewrote it -- observed
// as intermittently wrong all-gathered rows with 2 GPUs.  No kernel here leaves ciphertext data in an L1.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "params.h"
#include "blind_rotate.cuh"   // mbarrier / TMA helpers

namespace rs {

// ---------------------------------------------------------------- keyswitch (TFHE lweKeySwitch)
// res = (0, b') - sum_{i<N, j<t} KS[i][j][digit_ij]  where digit_ij = ((a'_i + prec_offset) >> (32-(j+1)*basebit)) & 7,
// rows with digit 0 skipped (TFHE's `if (aij != 0)`).  Integer, bit-exact: uint32 adds commute, so any order and any
// split of the sum gives the same words.
//
// Device KSK layout: [N][t][7][LWE_STRIDE] -- the digit-0 rows are never read and are dropped, rows padded to 352 words,
// so the 3 x 7 rows of (i, j-triple) are one contiguous 29,568-byte block = one TMA bulk copy.
//
// keyswitch_tiled_kernel<TILE>: one CTA = TILE ciphertexts x `ir` consecutive coefficients i (grid.y = N/ir splits the
// sum; partial sums are combined with red.global.add into rows pre-set to (0, b') by keyswitch_init_kernel).
//   warp 11 (one lane) streams the KSK blocks of its i-range through a KS_STAGES-deep shared-memory ring (TMA +
//   full/empty mbarriers); warps 0..10 = 352 threads = 4 groups of 88 uint4-lanes; group g owns TILE/4 ciphertexts and
//   for each (i, j) subtracts the row its digit selects, read from SHARED memory (LDS.128, conflict-free).
// Traffic per launch: L2 -> smem = (count/TILE) * 90.8 MB (each KSK byte once per tile, instead of count * 11.3 MB of
// L1/L2 gathers in the un-tiled kernel below), smem -> registers = count * 11.3 MB.
constexpr int KS_DIGITS = KS_BASE - 1;                                       // rows kept per (i, j)
constexpr int KS_JGROUP = 3;                                                 // j's per ring stage
constexpr int KS_STAGE_BYTES = KS_JGROUP * KS_DIGITS * LWE_STRIDE * 4;       // 29,568
constexpr int KS_STAGES = 4;
constexpr int KS_IR_MAX = 128;                                               // coefficients per CTA (abar tile in smem)
constexpr size_t KSK_TILED_WORDS = (size_t)N * KS_T * KS_DIGITS * LWE_STRIDE;

template <int TILE>
struct KsSmem {
    static constexpr int kStagesOff = 0;
    static constexpr int kAbarOff = KS_STAGES * KS_STAGE_BYTES;
    static constexpr int kBarOff = kAbarOff + KS_IR_MAX * TILE * 4;
    static constexpr int kTotal = kBarOff + 2 * KS_STAGES * 8;
};

__global__ void keyswitch_init_kernel(const uint32_t* __restrict__ ext, int count, uint32_t* __restrict__ lwe_out) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)count * LWE_STRIDE;
    for (; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t c = idx / LWE_STRIDE;
        const int x = (int)(idx % LWE_STRIDE);
        lwe_out[idx] = (x == LWE_N) ? __ldcg(ext + c * EXT_STRIDE + N) : 0u;
    }
    __threadfence();      // RS_END_FENCE: see the header note
}

template <int TILE>
__global__ void __launch_bounds__(384, 1)
keyswitch_tiled_kernel(const uint32_t* __restrict__ ext,      // [count][EXT_STRIDE]
                       int count, int ir,                      // ir = coefficients per CTA, divides N, <= KS_IR_MAX
                       const uint32_t* __restrict__ ksk7,      // [N][t][7][LWE_STRIDE]
                       uint32_t* __restrict__ lwe_out)         // [count][LWE_STRIDE], pre-set to (0, b')
{
    using S = KsSmem<TILE>;
    constexpr int CPG = TILE / 4;                              // ciphertexts per 88-lane group
    static_assert(CPG % 4 == 0, "abar rows are read as uint4");
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_base = smem_base + S::kBarOff;          // full[KS_STAGES], empty[KS_STAGES]
    uint32_t* abar = reinterpret_cast<uint32_t*>(smem + S::kAbarOff);      // [ir][TILE]
    const int first = blockIdx.x * TILE;
    const int tile = min(TILE, count - first);
    const int i0 = blockIdx.y * ir;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nstages = ir * (KS_T / KS_JGROUP);

    if (threadIdx.x == 0) {
        for (int s = 0; s < KS_STAGES; s++) {
            mbar_init(bar_base + s * 8, 1);
            mbar_init(bar_base + (KS_STAGES + s) * 8, 11);     // one arrival per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // a'_i + prec_offset for the tile, transposed to [i][ciphertext] (coalesced reads of ir consecutive words per row)
    for (int idx = threadIdx.x; idx < ir * TILE; idx += blockDim.x) {
        const int c = idx / ir, il = idx % ir;
        abar[il * TILE + c] = (c < tile) ? __ldcg(ext + (size_t)(first + c) * EXT_STRIDE + i0 + il) + KS_PREC_OFFSET : 0u;   // 0 => all digits 0
    }
    __syncthreads();

    if (warp == 11) {
        if (lane == 0) {
            const uint8_t* src = reinterpret_cast<const uint8_t*>(ksk7) + (size_t)i0 * (KS_T / KS_JGROUP) * KS_STAGE_BYTES;
#pragma unroll 1
            for (int n = 0; n < nstages; n++) {
                const int s = n % KS_STAGES;
                if (n >= KS_STAGES) mbar_wait_thread(bar_base + (KS_STAGES + s) * 8, ((n - KS_STAGES) / KS_STAGES) & 1);
                mbar_arrive_expect_tx(bar_base + s * 8, KS_STAGE_BYTES);
                tma_load_1d(smem_base + S::kStagesOff + s * KS_STAGE_BYTES, src + (size_t)n * KS_STAGE_BYTES, KS_STAGE_BYTES, bar_base + s * 8);
            }
        }
        return;
    }

    const int g = threadIdx.x / (LWE_STRIDE / 4);              // 88-lane group 0..3
    const int l4 = threadIdx.x % (LWE_STRIDE / 4);             // uint4 lane within the row
    uint4 acc[CPG];
#pragma unroll
    for (int c = 0; c < CPG; c++) acc[c] = make_uint4(0, 0, 0, 0);

#pragma unroll 1
    for (int n = 0; n < nstages; n++) {
        const int s = n % KS_STAGES;
        const int il = n / (KS_T / KS_JGROUP), jg = n % (KS_T / KS_JGROUP);
        mbar_wait_warp(bar_base + s * 8, (n / KS_STAGES) & 1);
        const uint4* rows = reinterpret_cast<const uint4*>(smem + S::kStagesOff + s * KS_STAGE_BYTES) + l4;
        const uint4* ab4 = reinterpret_cast<const uint4*>(abar + il * TILE + g * CPG);
#pragma unroll
        for (int c4 = 0; c4 < CPG / 4; c4++) {
            const uint4 a4 = ab4[c4];
            const uint32_t aw[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
#pragma unroll
                for (int jj = 0; jj < KS_JGROUP; jj++) {
                    const int sh = 32 - (jg * KS_JGROUP + jj + 1) * KS_BASEBIT;
                    const uint32_t d = (aw[k] >> sh) & (KS_BASE - 1);
                    if (d) {
                        const uint4 v = rows[(jj * KS_DIGITS + (int)d - 1) * (LWE_STRIDE / 4)];
                        uint4& r = acc[c4 * 4 + k];
                        r.x -= v.x; r.y -= v.y; r.z -= v.z; r.w -= v.w;
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_base + (KS_STAGES + s) * 8);
    }
    uint32_t sink = 0;
#pragma unroll
    for (int c = 0; c < CPG; c++) {
        const int ct = g * CPG + c;
        if (ct < tile) {
            uint32_t* o = lwe_out + (size_t)(first + ct) * LWE_STRIDE + 4 * l4;
            // ATOM (value returned and consumed), not RED: the thread cannot retire before the L2 has performed its last partial sums
            sink ^= atomicAdd(o + 0, acc[c].x) ^ atomicAdd(o + 1, acc[c].y) ^ atomicAdd(o + 2, acc[c].z) ^ atomicAdd(o + 3, acc[c].w);
        }
    }
    if (sink == 0x9E3779B9u && count < 0) lwe_out[0] = sink;      // never true: keeps the returned values live
    __threadfence();      // RS_END_FENCE
}

// KSK host layout [N][t][8][351] -> tiled device layout [N][t][7][352] (digit-0 rows dropped, rows zero-padded)
__global__ void ksk_tile_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, size_t rows7) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = rows7 * LWE_STRIDE;
    for (; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t r7 = idx / LWE_STRIDE;
        const int x = (int)(idx % LWE_STRIDE);
        const size_t ij = r7 / KS_DIGITS;
        const int d = (int)(r7 % KS_DIGITS) + 1;
        dst[idx] = x < LWE_WORDS ? src[(ij * KS_BASE + d) * LWE_WORDS + x] : 0u;
    }
}

// Un-tiled reference kernel (first version; kept as variant for A/B checks): KS_TILE ciphertexts per CTA share every KSK
// row they pick through L1/L2; thread x owns output word x.  Device KSK layout: [N][t][base][LWE_STRIDE].
template <int KS_TILE>
__global__ void __launch_bounds__(LWE_STRIDE)
keyswitch_kernel(const uint32_t* __restrict__ ext,      // [count][EXT_STRIDE]
                 int count,
                 const uint32_t* __restrict__ ksk,      // [N][t][base][LWE_STRIDE]
                 uint32_t* __restrict__ lwe_out)        // [count][LWE_STRIDE]
{
    __shared__ uint32_t abar[KS_TILE][N];
    const int x = threadIdx.x;
    const int first = blockIdx.x * KS_TILE;
    const int tile = min(KS_TILE, count - first);
    for (int c = 0; c < tile; c++)
        for (int i = x; i < N; i += LWE_STRIDE) abar[c][i] = __ldcg(ext + (size_t)(first + c) * EXT_STRIDE + i) + KS_PREC_OFFSET;
    __syncthreads();
    uint32_t acc[KS_TILE];
#pragma unroll
    for (int c = 0; c < KS_TILE; c++) acc[c] = 0;
#pragma unroll 1
    for (int i = 0; i < N; i++) {
        const uint32_t* rows = ksk + (size_t)i * KS_T * KS_BASE * LWE_STRIDE + x;
#pragma unroll
        for (int c = 0; c < KS_TILE; c++) {
            if (c < tile) {
                const uint32_t ai = abar[c][i];
#pragma unroll
                for (int j = 0; j < KS_T; j++) {
                    const uint32_t d = (ai >> (32 - (j + 1) * KS_BASEBIT)) & (KS_BASE - 1);
                    if (d) acc[c] -= __ldg(rows + (j * KS_BASE + d) * LWE_STRIDE);
                }
            }
        }
    }
#pragma unroll
    for (int c = 0; c < KS_TILE; c++) {
        if (c < tile) {
            uint32_t v = acc[c];
            if (x == LWE_N) v += __ldcg(ext + (size_t)(first + c) * EXT_STRIDE + N);
            if (x > LWE_N) v = 0;
            lwe_out[(size_t)(first + c) * LWE_STRIDE + x] = v;
        }
    }
}

// ---------------------------------------------------------------- KSK host layout -> padded device layout
__global__ void ksk_pad_kernel(const uint32_t* __restrict__ src /*[rows][351]*/, uint32_t* __restrict__ dst /*[rows][352]*/,
                               size_t rows) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = rows * LWE_STRIDE;
    for (; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        size_t r = idx / LWE_STRIDE;
        int x = (int)(idx % LWE_STRIDE);
        dst[idx] = x < LWE_WORDS ? __ldcg(src + r * LWE_WORDS + x) : 0u;
    }
}

// wire rows (351 words) <-> device rows (352 words)
__global__ void lwe_pad_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, int count) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)count * LWE_STRIDE;
    for (; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        size_t r = idx / LWE_STRIDE;
        int x = (int)(idx % LWE_STRIDE);
        dst[idx] = x < LWE_WORDS ? __ldcg(src + r * LWE_WORDS + x) : 0u;
    }
    __threadfence();      // RS_END_FENCE
}
__global__ void lwe_unpad_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, int count) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)count * LWE_WORDS;
    for (; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        size_t r = idx / LWE_WORDS;
        int x = (int)(idx % LWE_WORDS);
        dst[idx] = __ldcg(src + r * LWE_STRIDE + x);
    }
}

// ---------------------------------------------------------------- gate linear part (lib/GPU/gates.cu:44-108, constants :246-286)
// out = (0, fix) + m * (in0 + in1) with m in {+1,-1,+2,-2}
__global__ void gate_linear_kernel(uint32_t* __restrict__ out, const uint32_t* __restrict__ in0,
                                   const uint32_t* __restrict__ in1, int count, uint32_t m, uint32_t fix) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)count * LWE_STRIDE;
    for (; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        int x = (int)(idx % LWE_STRIDE);
        uint32_t v = m * (__ldcg(in0 + idx) + __ldcg(in1 + idx));
        if (x == LWE_N) v += fix;
        if (x > LWE_N) v = 0;
        out[idx] = v;
    }
    __threadfence();      // RS_END_FENCE
}

// out = (0, fix) + m0 * in0 + m1 * in1: add_int / sub_int / mul_int / levelNOT of lib/GPU/gates.cu:110-122,158-202 for a batch
__global__ void lwe_axpby_kernel(uint32_t* __restrict__ out, const uint32_t* __restrict__ in0, const uint32_t* __restrict__ in1,
                                 size_t count, uint32_t m0, uint32_t m1, uint32_t fix) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = count * LWE_STRIDE;
    for (; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(idx % LWE_STRIDE);
        uint32_t v = m0 * __ldcg(in0 + idx) + (in1 ? m1 * __ldcg(in1 + idx) : 0u);
        if (x == LWE_N) v += fix;
        if (x > LWE_N) v = 0;
        out[idx] = v;
    }
    __threadfence();      // RS_END_FENCE
}

// ---------------------------------------------------------------- ternary linear layer on LWE rows (a10 / f1)
// out[o] = (0, bias[o]) + sum_{k in [rowptr[o], rowptr[o+1])} sign[k] * in[col[k]]
// Covers BinFunc/IntFunc Convolution::execute, SumPooling::execute, Quantize bias add, add_bias
// (lib/BinFunc.cpp:142-330,677-732,1044-1107; lib/IntFunc.cpp:152-319,643-700,860-889): +-1/0 weighted sums mod 2^32.
// One CTA (88 threads x uint4 = 352 words) per OUT_TILE outputs is not needed here: entries are CSR so each
// output is its own row; thread x owns 4 words.
__global__ void __launch_bounds__(LWE_STRIDE / 4)
lwe_lincomb_kernel(uint32_t* __restrict__ out, int out_count, const uint32_t* __restrict__ in,
                   const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const int8_t* __restrict__ sign,
                   const uint32_t* __restrict__ bias) {
    const int x = threadIdx.x;   // uint4 lane within the row
    for (int o = blockIdx.x; o < out_count; o += gridDim.x) {
        uint4 acc = make_uint4(0, 0, 0, 0);
        const int k0 = rowptr[o], k1 = rowptr[o + 1];
        for (int k = k0; k < k1; k++) {
            const uint4 v = __ldcg(reinterpret_cast<const uint4*>(in + (size_t)col[k] * LWE_STRIDE) + x);
            const uint32_t s = (uint32_t)(int32_t)sign[k];
            acc.x += s * v.x; acc.y += s * v.y; acc.z += s * v.z; acc.w += s * v.w;
        }
        if (bias && x == LWE_N / 4) acc.z += bias[o];   // word 350 = b
        reinterpret_cast<uint4*>(out + (size_t)o * LWE_STRIDE)[x] = acc;
    }
    __threadfence();      // RS_END_FENCE
}

// b word of every row += value: adds the trivial sample (0,value) (lweNoiselessTrivial + lweAddTo, lib/BinOps_enc.cpp:137-141)
__global__ void lwe_add_const_kernel(uint32_t* __restrict__ rows, size_t count, uint32_t value) {
    for (size_t c = blockIdx.x * (size_t)blockDim.x + threadIdx.x; c < count; c += (size_t)gridDim.x * blockDim.x)
        rows[c * LWE_STRIDE + LWE_N] = __ldcg(rows + c * LWE_STRIDE + LWE_N) + value;      // L2-only load: see the header note
    __threadfence();      // RS_END_FENCE
}

// b word of row r += bias[r % mod]: the per-channel bias add of Quantize::execute / add_bias (lib/BinFunc.cpp:1063-1065,1085-1107;
// rows are channel-fastest) as a stage of its own, for callers that compose Func objects instead of whole layers
__global__ void lwe_add_bias_kernel(uint32_t* __restrict__ rows, size_t count, const uint32_t* __restrict__ bias, int mod) {
    for (size_t c = blockIdx.x * (size_t)blockDim.x + threadIdx.x; c < count; c += (size_t)gridDim.x * blockDim.x)
        rows[c * LWE_STRIDE + LWE_N] = __ldcg(rows + c * LWE_STRIDE + LWE_N) + bias[c % (size_t)mod];
    __threadfence();      // RS_END_FENCE
}

// ---------------------------------------------------------------- [world][pixels][c_local] -> [pixels][world*c_local]
__global__ void lwe_interleave_kernel(uint4* __restrict__ out, const uint4* __restrict__ in, size_t pixels, int c_local, int world) {
    const size_t rows = pixels * (size_t)c_local * world;
    const int x = threadIdx.x;   // uint4 lane (88 per row)
    for (size_t row = blockIdx.x; row < rows; row += gridDim.x) {
        const size_t pix = row / ((size_t)c_local * world);
        const int ch = (int)(row % ((size_t)c_local * world));
        const int r = ch / c_local, c = ch % c_local;
        out[row * (LWE_STRIDE / 4) + x] = __ldcg(in + (((size_t)r * pixels + pix) * c_local + c) * (LWE_STRIDE / 4) + x);
    }
    __threadfence();      // RS_END_FENCE
}

// ---------------------------------------------------------------- ternary convolution / fully-connected layer on LWE rows
// out[(ph,pw,od)] = (0,bias[od]) + sum_{fh,fw,di} w[fh,fw,di,od] * in[(ph*st+fh-ofs, pw*st+fw-ofs, di)]
// Index conventions follow lib/BinFunc.cpp:388-404 (input/output (h,w,c) c-fastest; filter ((fh*W+fw)*in_dep+di)*OutDepth+od).
// INT_MODE reproduces lib/IntFunc.cpp:268,277: a zero (ternary) weight or a padded position contributes the
// trivial sample -1/4096 instead of 0 (lib/BinFunc.cpp:265,280 contributes 0).
// Weights are pre-packed on the host into tiles of CONV_OD_TILE output channels: wp[tile][k][CONV_OD_TILE] int8,
// k = (fh*win_w+fw)*in_dep+di, so one 16-byte broadcast load feeds 16 accumulators.
constexpr int CONV_OD_TILE = 16;
constexpr int CONV_THREADS = 96;   // 88 active lanes x uint4 = 352 words

struct ConvDesc {
    int in_h, in_w, in_dep;
    int out_h, out_w, out_dep;
    int win_h, win_w, stride_h, stride_w, ofs_h, ofs_w;
    int od_begin, od_end;     // channel slice computed by this launch (neuron sharding); od_begin % CONV_OD_TILE == 0
    uint32_t unit;            // torus value of one message unit (1/4096)
};

template <bool INT_MODE>
__global__ void __launch_bounds__(CONV_THREADS)
lwe_conv_kernel(uint32_t* __restrict__ out,             // [out_h*out_w][od_end-od_begin][LWE_STRIDE]
                const uint32_t* __restrict__ in,         // [in_h*in_w*in_dep][LWE_STRIDE]
                const int8_t* __restrict__ wp,           // [ceil(out_dep/16)][K][16]
                const uint32_t* __restrict__ bias,       // [out_dep] torus32 or nullptr
                ConvDesc d) {
    const int lane = threadIdx.x;
    if (lane >= LWE_STRIDE / 4) return;
    const int pix = blockIdx.x, ph = pix / d.out_w, pw = pix % d.out_w;
    const int tile = d.od_begin / CONV_OD_TILE + blockIdx.y;
    const int K = d.win_h * d.win_w * d.in_dep;
    const int8_t* wt = wp + (size_t)tile * K * CONV_OD_TILE;
    uint4 acc[CONV_OD_TILE];
    int skipped[CONV_OD_TILE];
#pragma unroll
    for (int o = 0; o < CONV_OD_TILE; o++) { acc[o] = make_uint4(0, 0, 0, 0); skipped[o] = 0; }
    int oob_terms = 0;
    for (int fh = 0; fh < d.win_h; fh++) {
        const int iy = ph * d.stride_h + fh - d.ofs_h;
        for (int fw = 0; fw < d.win_w; fw++) {
            const int ix = pw * d.stride_w + fw - d.ofs_w;
            if ((unsigned)iy >= (unsigned)d.in_h || (unsigned)ix >= (unsigned)d.in_w) { oob_terms += d.in_dep; continue; }
            const uint4* src = reinterpret_cast<const uint4*>(in + (size_t)((iy * d.in_w + ix) * d.in_dep) * LWE_STRIDE) + lane;
            const int8_t* wk = wt + (size_t)((fh * d.win_w + fw) * d.in_dep) * CONV_OD_TILE;
#pragma unroll 2
            for (int di = 0; di < d.in_dep; di++) {
                const uint4 v = __ldcg(src + (size_t)di * (LWE_STRIDE / 4));
                const uint4 wq = __ldg(reinterpret_cast<const uint4*>(wk + (size_t)di * CONV_OD_TILE));
                const uint32_t wwords[4] = {wq.x, wq.y, wq.z, wq.w};
#pragma unroll
                for (int o = 0; o < CONV_OD_TILE; o++) {
                    const int w8 = (int)(int8_t)((wwords[o >> 2] >> ((o & 3) * 8)) & 0xFF);
                    const uint32_t w = (uint32_t)w8;
                    acc[o].x += w * v.x; acc[o].y += w * v.y; acc[o].z += w * v.z; acc[o].w += w * v.w;
                    if (INT_MODE) skipped[o] += (w8 == 0);
                }
            }
        }
    }
    const int od_count = d.od_end - d.od_begin;
#pragma unroll
    for (int o = 0; o < CONV_OD_TILE; o++) {
        const int od = tile * CONV_OD_TILE + o;
        if (od >= d.od_end || od >= d.out_dep) continue;
        uint4 r = acc[o];
        if (lane == LWE_N / 4) {   // words 348..351: word 350 is b
            if (bias) r.z += bias[od];
            if (INT_MODE) r.z -= (uint32_t)(skipped[o] + oob_terms) * d.unit;
            r.w = 0;
        }
        reinterpret_cast<uint4*>(out + ((size_t)pix * od_count + (od - d.od_begin)) * LWE_STRIDE)[lane] = r;
    }
    __threadfence();      // RS_END_FENCE
}

}  // namespace rs
