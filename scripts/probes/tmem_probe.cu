// scripts/probes/tmem_probe.cu -- microbenchmarks that decide whether the Fourier BSK should be served to the MAC from
// tensor memory (tcgen05.cp smem->TMEM, tcgen05.ld TMEM->registers) instead of ld.shared.
//   A. tcgen05.ld throughput per SM for 4 and 8 reading warps
//   B. source layout of tcgen05.cp.128x256b / 64x128b.warpx2 with a no-swizzle descriptor (pattern dump)
//   C. LDS.128 throughput of the same 8 warps for comparison, and both at once (do the two paths share a pipe?)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_probe tmem_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                 "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;     // sm_100 descriptor version
    return d;                   // swizzle mode 0 (none), base offset 0
}

// mode 0: ld throughput; mode 1: cp layout dump (128x256b); mode 2: cp 64x128b.warpx2::01_23 dump; mode 3: LDS throughput;
// mode 4: LDS + tcgen05.ld interleaved; mode 5: cp 32x128b.warpx4 dump; mode 6: cp 128x128b dump
__global__ void __launch_bounds__(256, 1) probe(int mode, int iters, int lbo, int sbo, uint32_t* out, long long* cyc) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t bar;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* sw = reinterpret_cast<uint32_t*>(smem);
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) sw[i] = i;     // 64 KiB pattern: word index
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the async proxy (tcgen05.cp)
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tbase = tmem_base_s;
    const uint32_t tq = tbase + ((uint32_t)(32 * (warp & 3)) << 16);     // this warp's lane quadrant

    // zero-fill TMEM so dumps of untouched cells are recognisable
    {
        uint32_t z[16];
        for (int k = 0; k < 16; k++) z[k] = 0xDEAD0000u + k;
        if (warp < 4) for (int c = 0; c < 512; c += 16) tmem_st16(tq + c, z);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");

    if (mode == 0 || mode == 3 || mode == 4) {
        uint32_t acc = 0;
        const uint4* s4 = reinterpret_cast<const uint4*>(smem);
        __syncthreads();
        const long long t0 = clock64();
        for (int it = 0; it < iters; it++) {
            if (mode == 0 || mode == 4) {
                uint32_t r[16];
#pragma unroll
                for (int c = 0; c < 64; c += 16) {
                    tmem_ld16(tq + ((it * 64 + c) & 511), r);
                    tmem_ld_wait();
#pragma unroll
                    for (int k = 0; k < 16; k++) acc += r[k];
                }
            }
            if (mode == 3 || mode == 4) {
#pragma unroll
                for (int c = 0; c < 16; c++) {
                    const uint4 v = s4[((it * 16 + c) * 64 + (threadIdx.x & 63)) & 4095];
                    acc += v.x + v.y + v.z + v.w;
                }
            }
        }
        const long long t1 = clock64();
        __syncthreads();
        if (lane == 0) cyc[warp] = t1 - t0;
        out[threadIdx.x] = acc;
    } else if (mode >= 20) {
        // cp throughput with several issuing warps (one lane each): mode 20+w = w+1 warps
        __shared__ __align__(8) uint64_t bars[4];
        const int nw = mode - 19;
        if (threadIdx.x == 0) for (int i = 0; i < 4; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])));
        __syncthreads();
        if (lane == 0 && warp < nw) {
            const uint64_t desc = make_desc(smem_u32(smem) + warp * 16384, 128, 128);
            const long long t0 = clock64();
            for (int it = 0; it < iters; it++) {
#pragma unroll
                for (int k = 0; k < 16; k++)
                    asm volatile("tcgen05.cp.cta_group::1.64x128b.warpx2::02_13 [%0], %1;" ::"r"(tbase + warp * 64 + k * 4), "l"(desc + k * 64) : "memory");
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[warp])) : "memory");
                uint32_t ok = 0;
                while (!ok) {
                    asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}"
                                 : "=r"(ok) : "r"(smem_u32(&bars[warp])), "r"(it & 1) : "memory");
                }
            }
            cyc[warp] = clock64() - t0;
        }
    } else if (mode >= 10) {
        // cp throughput: `iters` rounds of 16 copies + commit + wait, one issuing thread
        if (threadIdx.x == 0) {
            const uint64_t desc = make_desc(smem_u32(smem), 128, 128);
            const long long t0 = clock64();
            for (int it = 0; it < iters; it++) {
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    if (mode == 10) asm volatile("tcgen05.cp.cta_group::1.64x128b.warpx2::02_13 [%0], %1;" ::"r"(tbase + k * 4), "l"(desc + k * 64) : "memory");
                    if (mode == 11) asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tbase + k * 8), "l"(desc + k * 64) : "memory");
                    if (mode == 12) asm volatile("tcgen05.cp.cta_group::1.128x128b [%0], %1;" ::"r"(tbase + k * 4), "l"(desc + k * 64) : "memory");
                    if (mode == 13) asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(tbase + k * 4), "l"(desc + k * 32) : "memory");
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
                uint32_t ok = 0;
                while (!ok) {
                    asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}"
                                 : "=r"(ok) : "r"(smem_u32(&bar)), "r"(it & 1) : "memory");
                }
            }
            cyc[0] = clock64() - t0;
        }
    } else {
        if (threadIdx.x == 0) {
            const uint64_t desc = make_desc(smem_u32(smem), (uint32_t)lbo, (uint32_t)sbo);
            if (mode == 1) asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tbase), "l"(desc) : "memory");
            if (mode == 2) asm volatile("tcgen05.cp.cta_group::1.64x128b.warpx2::01_23 [%0], %1;" ::"r"(tbase), "l"(desc) : "memory");
            if (mode == 5) asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(tbase), "l"(desc) : "memory");
            if (mode == 6) asm volatile("tcgen05.cp.cta_group::1.128x128b [%0], %1;" ::"r"(tbase), "l"(desc) : "memory");
            if (mode == 7) asm volatile("tcgen05.cp.cta_group::1.64x128b.warpx2::02_13 [%0], %1;" ::"r"(tbase), "l"(desc) : "memory");
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        // everyone waits for the copy
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}"
                         : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        }
        asm volatile("tcgen05.fence::after_thread_sync;");
        if (warp < 4) {
            uint32_t r[16];
            tmem_ld16(tq, r);
            tmem_ld_wait();
            for (int k = 0; k < 16; k++) out[(warp * 32 + lane) * 16 + k] = r[k];     // out[tmem lane][column]
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase));
}

int main() {
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 128 * 16 * 4 * 4); cudaMalloc(&cyc, 8 * 8);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    std::vector<uint32_t> h(128 * 16); long long hc[8];
    for (int threads : {128, 256}) {
        for (int mode : {0, 3, 4}) {
            const int iters = 2000;
            probe<<<1, threads, 65536>>>(mode, iters, 0, 0, out, cyc);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
            const double bytes_tm = (mode == 0 || mode == 4) ? (double)iters * 64 * 4 * threads : 0;
            const double bytes_ls = (mode == 3 || mode == 4) ? (double)iters * 16 * 16 * threads : 0;
            printf("mode %d threads %d: %s cycles(warp0) %lld  tmem %.1f B/clk  lds %.1f B/clk\n", mode, threads, cudaGetErrorString(e), hc[0],
                   bytes_tm / hc[0], bytes_ls / hc[0]);
        }
    }
    for (int mode : {10, 11, 12, 13}) {
        const int iters = 1000;
        probe<<<1, 128, 65536>>>(mode, iters, 0, 0, out, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
        const char* nm[] = {"64x128b.warpx2 (1 KiB smem)", "128x256b (4 KiB)", "128x128b (2 KiB)", "32x128b.warpx4 (512 B)"};
        printf("cp mode %d %s: %s  %.1f clk per 16 copies + commit + wait  (%.1f clk/copy)\n", mode, nm[mode - 10], cudaGetErrorString(e),
               (double)hc[0] / iters, (double)hc[0] / iters / 16);
    }
    for (int mode : {20, 21, 23}) {
        const int iters = 1000;
        probe<<<1, 128, 65536>>>(mode, iters, 0, 0, out, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
        printf("cp from %d warps concurrently: %s  warp0 %.1f clk per 16 copies (%.1f clk/copy/warp)\n", mode - 19, cudaGetErrorString(e),
               (double)hc[0] / iters, (double)hc[0] / iters / 16);
    }
    struct Cfg { int mode, lbo, sbo; const char* name; };
    const Cfg cfgs[] = {{1, 128, 256, "128x256b lbo128 sbo256"}, {1, 2048, 128, "128x256b lbo2048 sbo128"}, {6, 128, 128, "128x128b lbo128 sbo128"},
                        {2, 128, 128, "64x128b.warpx2::01_23 lbo128 sbo128"}, {7, 128, 128, "64x128b.warpx2::02_13 lbo128 sbo128"},
                        {5, 128, 128, "32x128b.warpx4 lbo128 sbo128"}};
    for (const Cfg& c : cfgs) {
        cudaMemset(out, 0xff, 128 * 16 * 4);
        probe<<<1, 128, 65536>>>(c.mode, 0, c.lbo, c.sbo, out, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h.data(), out, 128 * 16 * 4, cudaMemcpyDeviceToHost);
        printf("== %s: %s\n", c.name, cudaGetErrorString(e));
        for (int l : {0, 1, 7, 8, 9, 16, 31, 32, 33, 63, 64, 65, 96, 127}) {
            printf("lane %3d:", l);
            for (int k = 0; k < 16; k++) printf(" %5x", h[l * 16 + k]);
            printf("\n");
        }
    }
    return 0;
}
