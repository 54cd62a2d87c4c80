// scripts/probes/dmma_probe.cu -- is the FP64 tensor-core path (mma.sync m8n8k4 / m16n8k16 f64) a pipe of its own on B200, and
// at what rate?  Times (a) DFMA only, (b) DMMA only, (c) both in the same warps, (d) DMMA in half the warps and DFMA in the other
// half, 12 warps per SM as in the blind rotation.  If (c)/(d) take max(a,b) the two overlap and the radix-8 passes could be
// split between the pipes; if they take a+b the tensor path shares the FP64 units.
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
        : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
        : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// MODE bit0: DFMA work, bit1: DMMA m8n8k4 work, bit2: DMMA m16n8k16 work; SPLIT: even warps tensor, odd warps vector
template <int MODE, bool SPLIT>
__global__ void __launch_bounds__(384, 1) k(double* out, const double* in, int iters) {
    double f[8], g[8];
    double c8[4][2], c16[4][4], a16[8], b16[4];
    for (int i = 0; i < 8; i++) { f[i] = in[threadIdx.x + i]; g[i] = in[threadIdx.x + 8 + i]; a16[i] = in[threadIdx.x + 16 + i]; }
    for (int i = 0; i < 4; i++) { b16[i] = in[threadIdx.x + 24 + i]; c8[i][0] = c8[i][1] = 0; for (int j = 0; j < 4; j++) c16[i][j] = 0; }
    const int warp = threadIdx.x >> 5;
    const bool do_vec = (MODE & 1) && (!SPLIT || (warp & 1));
    const bool do_t8 = (MODE & 2) && (!SPLIT || !(warp & 1));
    const bool do_t16 = (MODE & 4) && (!SPLIT || !(warp & 1));
    for (int it = 0; it < iters; it++) {
        if (do_vec) {
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int i = 0; i < 8; i++) f[i] = fma(g[i], 1.0000001, f[i]);     // 32 DFMA (2 registers + immediate)
        }
        if (do_t8) {
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int i = 0; i < 4; i++) dmma884(c8[i], a16[i], b16[i]);         // 16 x (8*8*4 = 256 FMA)
        }
        if (do_t16) {
#pragma unroll
            for (int i = 0; i < 4; i++) dmma16816(c16[i], a16, b16);                // 4 x (16*8*16 = 2048 FMA)
        }
    }
    double s = 0;
    for (int i = 0; i < 8; i++) s += f[i];
    for (int i = 0; i < 4; i++) { s += c8[i][0] + c8[i][1]; for (int j = 0; j < 4; j++) s += c16[i][j]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, bool SPLIT>
void run(const char* name) {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int grid = p.multiProcessorCount, block = 384, iters = 4000;
    double *out, *in; cudaMalloc(&out, (size_t)grid * block * 8); cudaMalloc(&in, 4096 * 8); cudaMemset(in, 0, 4096 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; r++) {
        cudaEventRecord(e0); k<MODE, SPLIT><<<grid, block>>>(out, in, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
    }
    const double warps = SPLIT ? 6.0 : 12.0;
    double fma_vec = (MODE & 1) ? (double)iters * 32 * 32 * warps : 0;
    double fma_t = 0;
    if (MODE & 2) fma_t += (double)iters * 16 * 256 * warps;
    if (MODE & 4) fma_t += (double)iters * 4 * 2048 * warps;
    const double sec = best * 1e-3;
    printf("%-46s %8.3f ms   vector %6.2f TFLOP/s   tensor %6.2f TFLOP/s\n", name, best, 2 * fma_vec * grid / sec * 1e-12,
           2 * fma_t * grid / sec * 1e-12);
    cudaFree(out); cudaFree(in);
}
int main() {
    run<1, false>("DFMA only (12 warps)");
    run<2, false>("DMMA m8n8k4 only (12 warps)");
    run<4, false>("DMMA m16n8k16 only (12 warps)");
    run<3, false>("DFMA + DMMA m8n8k4, same warps");
    run<5, false>("DFMA + DMMA m16n8k16, same warps");
    run<1, true>("DFMA only (6 of 12 warps)");
    run<4, true>("DMMA m16n8k16 only (6 of 12 warps)");
    run<5, true>("DFMA odd warps / DMMA m16n8k16 even warps");
    run<3, true>("DFMA odd warps / DMMA m8n8k4 even warps");
    return 0;
}
