// scripts/probes/fp64_latency.cu -- FP64 dependent-issue latency and single-warp throughput vs ILP on sm_100a
#include <cuda_runtime.h>
#include <cstdio>
template <int CH, int OP>
__global__ void k(double* out, const double* in, int iters, long long* cyc) {
    double a[CH], b = in[threadIdx.x + 40], c = in[threadIdx.x + 80];
    for (int i = 0; i < CH; i++) a[i] = in[threadIdx.x + i];
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int i = 0; i < CH; i++) {
                if (OP == 0) a[i] = fma(a[i], b, c);        // dependent DFMA chain(s)
                if (OP == 1) a[i] = a[i] + b;               // DADD
                if (OP == 2) a[i] = a[i] * b;               // DMUL
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < CH; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int CH, int OP>
void run(const char* name, int warps) {
    double *out, *in; long long* cyc; cudaMalloc(&out, 4096 * 8); cudaMalloc(&in, 4096 * 8); cudaMemset(in, 0, 4096 * 8); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    k<CH, OP><<<1, 32 * warps>>>(out, in, iters, cyc);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-44s %2d chains, %2d warps on the SM: %.2f cycles per instruction per warp\n", name, CH, warps, (double)h / (iters * 8.0 * CH));
}
int main() {
    run<1, 0>("DFMA dependent chain", 1);
    run<1, 1>("DADD dependent chain", 1);
    run<1, 2>("DMUL dependent chain", 1);
    run<2, 0>("DFMA", 1); run<4, 0>("DFMA", 1); run<8, 0>("DFMA", 1); run<16, 0>("DFMA", 1);
    run<1, 0>("DFMA dependent chain", 4);    // one warp per sub-partition
    run<1, 0>("DFMA dependent chain", 8);    // two per sub-partition
    run<1, 0>("DFMA dependent chain", 12);
    run<4, 0>("DFMA", 12);
    return 0;
}
