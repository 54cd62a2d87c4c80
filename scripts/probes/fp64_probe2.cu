// scripts/probes/fp64_probe2.cu -- cost model of FP64 instructions by operand pattern (cycles per warp instruction per SM
// sub-partition; the pipe's nominal rate is one per 2 cycles).  12 warps per SM, as in the blind rotation.
#include <cuda_runtime.h>
#include <cstdio>
constexpr int CH = 8;
template <int MODE>
__global__ void __launch_bounds__(384, 1) k(double* out, const double* in, int iters, double uarg) {
    double a[CH], b[CH], c[CH];
    for (int i = 0; i < CH; i++) { a[i] = in[threadIdx.x + i]; b[i] = in[threadIdx.x + 32 + i]; c[i] = in[threadIdx.x + 64 + i]; }
    const double u = uarg;   // warp-uniform (kernel parameter)
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < CH; i++) {
                if (MODE == 0) a[i] = fma(b[i], c[i], a[i]);                 // 3 distinct registers
                if (MODE == 1) a[i] = fma(b[0], c[i], a[i]);                 // one operand shared by consecutive instructions
                if (MODE == 2) a[i] = a[i] + b[i];                           // DADD, 2 registers
                if (MODE == 3) a[i] = fma(b[i], 0.70710678118654752440, a[i]); // constant multiplier
                if (MODE == 4) a[i] = fma(b[i], u, a[i]);                    // uniform multiplier
                if (MODE == 5) a[i] = a[i] * b[i];                           // DMUL, 2 registers
                if (MODE == 6) a[i] = fma(a[i], b[i], a[i]);                 // 3 slots, 2 distinct registers
                if (MODE == 7) a[i] = fma(a[i], 1.0000001, 1e-7);            // the peak probe: 1 register
                if (MODE == 8) a[i] = fma(b[i], c[(i + 1) % CH], a[i]);      // 3 distinct, other pairing
            }
        }
    }
    double s = 0;
    for (int i = 0; i < CH; i++) s += a[i] + b[i] + c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name) {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int grid = p.multiProcessorCount, block = 384, iters = 2000;
    double *out, *in; cudaMalloc(&out, (size_t)grid * block * 8); cudaMalloc(&in, 4096 * 8); cudaMemset(in, 0, 4096 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; r++) {
        cudaEventRecord(e0); k<MODE><<<grid, block>>>(out, in, iters, 1.25); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
    }
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double warp_instr_per_smsp = (double)iters * 4 * CH * 3;          // 12 warps / 4 sub-partitions
    const double cycles = best * 1e-3 * khz * 1e3;
    printf("%-58s %.2f cycles per warp instruction (at %d MHz nominal)\n", name, cycles / warp_instr_per_smsp, khz / 1000);
    cudaFree(out); cudaFree(in);
}
int main() {
    run<7>("fma(a, const, const)            1 register");
    run<2>("a + b                           DADD, 2 registers");
    run<5>("a * b                           DMUL, 2 registers");
    run<3>("fma(b, const, a)                2 registers + immediate");
    run<4>("fma(b, uniform, a)              2 registers + uniform");
    run<6>("fma(a, b, a)                    3 slots, 2 distinct");
    run<1>("fma(b0, c_i, a_i)               3 registers, one reused");
    run<0>("fma(b_i, c_i, a_i)              3 distinct registers");
    run<8>("fma(b_i, c_i+1, a_i)            3 distinct, other pairing");
    return 0;
}
