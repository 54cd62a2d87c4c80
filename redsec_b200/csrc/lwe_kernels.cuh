// redsec_b200/csrc/lwe_kernels.cuh -- integer LWE kernels: keyswitch, gate pre-combination, ternary linear layers.
// All arithmetic is uint32 wrap-around (torus32), so results are bit-exact by construction.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "params.h"

namespace rs {

// ---------------------------------------------------------------- keyswitch (TFHE lweKeySwitch)
// res = (0, b') - sum_{i<N, j<t} KS[i][j][digit_ij]  where digit_ij = ((a'_i + prec_offset) >> (32-(j+1)*basebit)) & 7.
// KS_TILE ciphertexts per CTA share every KSK row they pick through L1/L2; thread x owns output word x
// of each ciphertext in the tile.  Device KSK layout: [N][t][base][LWE_STRIDE] (rows padded to 352 words).
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
        for (int i = x; i < N; i += LWE_STRIDE) abar[c][i] = ext[(size_t)(first + c) * EXT_STRIDE + i] + KS_PREC_OFFSET;
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
            if (x == LWE_N) v += ext[(size_t)(first + c) * EXT_STRIDE + N];
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
        dst[idx] = x < LWE_WORDS ? src[r * LWE_WORDS + x] : 0u;
    }
}

// wire rows (351 words) <-> device rows (352 words)
__global__ void lwe_pad_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, int count) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)count * LWE_STRIDE;
    for (; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        size_t r = idx / LWE_STRIDE;
        int x = (int)(idx % LWE_STRIDE);
        dst[idx] = x < LWE_WORDS ? src[r * LWE_WORDS + x] : 0u;
    }
}
__global__ void lwe_unpad_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, int count) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)count * LWE_WORDS;
    for (; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        size_t r = idx / LWE_WORDS;
        int x = (int)(idx % LWE_WORDS);
        dst[idx] = src[r * LWE_STRIDE + x];
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
        uint32_t v = m * (in0[idx] + in1[idx]);
        if (x == LWE_N) v += fix;
        if (x > LWE_N) v = 0;
        out[idx] = v;
    }
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
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(in + (size_t)col[k] * LWE_STRIDE) + x);
            const uint32_t s = (uint32_t)(int32_t)sign[k];
            acc.x += s * v.x; acc.y += s * v.y; acc.z += s * v.z; acc.w += s * v.w;
        }
        if (bias && x == LWE_N / 4) acc.z += bias[o];   // word 350 = b
        reinterpret_cast<uint4*>(out + (size_t)o * LWE_STRIDE)[x] = acc;
    }
}

}  // namespace rs
