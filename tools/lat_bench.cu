// Latency microbenchmarks for the building blocks of the invert kernel's panel chain (sm_100a):
// dependent DFMA / DMUL / DADD, MUFU.RCP64H, SHFL, LDS, bar.sync, each alone (1 warp) and with
// other warps streaming DFMAs on the same SM sub-partitions.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/lat_bench tools/lat_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void lat(double *out, long long *clk, int iters, int busy_warps)
{
    __shared__ double sm[1024];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i;
    __syncthreads();
    if (warp == 0) {
        double x = 1.0 + 1e-9 * lane, y = 1.0000001, z = 1e-9;
        int idx = lane;
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                if (MODE == 0) x = fma(x, y, z);
                if (MODE == 1) x = x * y;
                if (MODE == 2) x = x + z;
                if (MODE == 3) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r; }
                if (MODE == 4) x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31);
                if (MODE == 5) { idx = (int) sm[idx & 1023] + (idx & 1023) - 1; }
                if (MODE == 6) { asm volatile("bar.sync 1, 32;" ::: "memory"); }
                if (MODE == 7) { int v = idx; asm volatile("shfl.sync.idx.b32 %0, %0, %1, 0x1f, 0xffffffff;" : "+r"(v) : "r"((lane + 1) & 31)); idx = v; }
            }
        }
        const long long t1 = clock64();
        if (lane == 0) clk[0] = t1 - t0;
        out[threadIdx.x] = x + idx;
    } else if (warp <= busy_warps) {
        // streaming independent DFMAs
        double a0 = 1.0, a1 = 1.1, a2 = 1.2, a3 = 1.3, a4 = 1.4, a5 = 1.5, a6 = 1.6, a7 = 1.7, y = 1.0000001, z = 1e-9;
        for (int i = 0; i < iters * 4; ++i) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                a0 = fma(a0, y, z); a1 = fma(a1, y, z); a2 = fma(a2, y, z); a3 = fma(a3, y, z);
                a4 = fma(a4, y, z); a5 = fma(a5, y, z); a6 = fma(a6, y, z); a7 = fma(a7, y, z);
            }
        }
        out[threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    }
}

template <int MODE>
void run(const char *name, double *out, long long *clk)
{
    const int iters = 2000;
    for (int busy : {0, 3, 7, 15}) {
        lat<MODE><<<1, 32 * (busy + 1)>>>(out, clk, iters, busy);
        cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, clk, sizeof c, cudaMemcpyDeviceToHost);
        printf("%-14s busy warps %2d: %.1f clk per op\n", name, busy, (double) c / (iters * 16.0));
    }
}

int main()
{
    double *out; long long *clk;
    cudaMalloc(&out, 8 * 1024); cudaMalloc(&clk, 64);
    run<0>("DFMA dep", out, clk);
    run<1>("DMUL dep", out, clk);
    run<2>("DADD dep", out, clk);
    run<3>("MUFU.RCP64H", out, clk);
    run<4>("SHFL f64 (x2)", out, clk);
    run<7>("SHFL b32", out, clk);
    run<5>("LDS dep", out, clk);
    run<6>("bar.sync 32", out, clk);
    return 0;
}
