// scripts/probes/fp64_probe.cu -- what DFMA rate is reachable with REAL operand patterns?  rs_fp64_peak (the roofline
// denominator) times fma(a, const, const): one register operand.  The blind rotation's FMAs read three register operands.
//   mode 0: a = fma(a, m, c)        m, c compile-time constants          (the peak probe)
//   mode 1: a_k = fma(b_k, c_k, a_k)  three distinct register operands, 8 independent chains
//   mode 2: as 1 with 16 chains (more ILP)
//   mode 3: as 1, 3 warps per SM sub-partition instead of 8+ (the kernel's occupancy)
#include <cuda_runtime.h>
#include <cstdio>
template <int MODE, int CH>
__global__ void k(double* out, const double* in, int iters) {
    double a[CH], b[CH], c[CH];
    for (int i = 0; i < CH; i++) { a[i] = in[threadIdx.x + i]; b[i] = in[threadIdx.x + 32 + i]; c[i] = in[threadIdx.x + 64 + i]; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) {
            if (MODE == 0) a[i] = fma(a[i], 1.0000001, 1e-7);
            else a[i] = fma(b[i], c[(i + 1) % CH], a[i]);
        }
        if (MODE != 0) {
#pragma unroll
            for (int i = 0; i < CH; i++) b[i] = fma(a[i], c[i], b[(i + 3) % CH]);
        }
    }
    double s = 0;
    for (int i = 0; i < CH; i++) s += a[i] + b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE, int CH>
void run(const char* name, int block, int grid_per_sm) {
    int dev = 0; cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    const int grid = p.multiProcessorCount * grid_per_sm, iters = 4000;
    double *out, *in; cudaMalloc(&out, (size_t)grid * block * 8); cudaMalloc(&in, 4096 * 8); cudaMemset(in, 0, 4096 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; r++) {
        cudaEventRecord(e0); k<MODE, CH><<<grid, block>>>(out, in, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
    }
    const double fmas = (double)iters * CH * (MODE == 0 ? 1 : 2) * grid * block;
    printf("%-60s %.2f TFLOP/s\n", name, 2 * fmas / (best * 1e-3) / 1e12);
    cudaFree(out); cudaFree(in);
}
int main() {
    run<0, 8>("mode0 const operands, 8 chains, 8 warps x 8 CTAs/SM", 256, 8);
    run<1, 8>("mode1 3 register operands, 8 chains, 8 warps x 8 CTAs/SM", 256, 8);
    run<1, 16>("mode2 3 register operands, 16 chains, 8 warps x 8 CTAs/SM", 256, 8);
    run<1, 8>("mode3 3 register operands, 8 chains, 12 warps/SM (1 CTA)", 384, 1);
    run<1, 16>("mode3b 3 register operands, 16 chains, 12 warps/SM (1 CTA)", 384, 1);
    run<1, 8>("mode3c 3 register operands, 8 chains, 8 warps/SM (1 CTA)", 256, 1);
    run<1, 8>("mode3d 3 register operands, 8 chains, 4 warps/SM (1 CTA)", 128, 1);
    return 0;
}
