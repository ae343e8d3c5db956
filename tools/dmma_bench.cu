// FP64 tensor-core (DMMA) microbenchmark for sm_100a: dependent-chain latency and sustained throughput of
// mma.sync m8n8k4 / m16n8k4 / m16n8k8 f64 against plain DFMA, at the occupancy of the invert kernel
// (3 CTAs x 8 warps per SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_bench tools/dmma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double (&c)[4], double a0, double a1, double b)
{
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a0), "d"(a1), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], double b0, double b1)
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b0), "d"(b1));
}

// MODE 0: DFMA, 1: m8n8k4, 2: m16n8k4, 3: m16n8k8.  NACC independent accumulator sets per warp.
template <int MODE, int NACC>
__global__ void __launch_bounds__(256, 3) tput(double *out, long long *clk, int iters)
{
    const int lane = threadIdx.x & 31;
    double c[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 1e-3 * (lane + i + j);
    double a[4] = {1.0 + 1e-9 * lane, 1.0 - 1e-9 * lane, 0.5, 0.25}, b0 = 1e-6 * lane, b1 = 2e-6;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (MODE == 0) { c[i][0] = fma(a[0], b0, c[i][0]); c[i][1] = fma(a[1], b0, c[i][1]); }
            if (MODE == 1) dmma884(c[i][0], c[i][1], a[0], b0);
            if (MODE == 2) dmma1684(c[i], a[0], a[1], b0);
            if (MODE == 3) dmma1688(c[i], a, b0, b1);
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}

template <int MODE, int NACC>
void run(const char *name, double flop_per_op, double *out, long long *clk, int ctas_per_sm, int warps)
{
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    tput<MODE, NACC><<<148 * ctas_per_sm, 32 * warps>>>(out, clk, 100);
    cudaEventRecord(e0);
    tput<MODE, NACC><<<148 * ctas_per_sm, 32 * warps>>>(out, clk, iters);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, clk, sizeof c, cudaMemcpyDeviceToHost);
    const double ops = (double) iters * NACC * 148 * ctas_per_sm * warps;
    printf("%-10s acc %d  %d CTA/SM x %d warps: %7.2f TFLOP/s  %6.1f clk per warp-op (issue-to-issue, one warp)  %s\n",
           name, NACC, ctas_per_sm, warps, ops * flop_per_op / (ms * 1e-3) / 1e12, (double) c / ((double) iters * NACC),
           cudaGetErrorString(cudaGetLastError()));
}

// warps with odd index run DFMAs, even ones DMMAs: do the two share a pipe?
__global__ void __launch_bounds__(256, 3) mixed(double *out, int iters, int mode)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = 1e-3 * (lane + i); c[i][1] = 2e-3 * (lane - i); }
    const double a = 1.0 + 1e-9 * lane, b = 1e-6 * lane;
    const bool dm = mode == 1 || (mode == 2 && (warp & 1) == 0);
    if (dm) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);
        }
    } else {
        for (int it = 0; it < iters * 4; ++it) {      // 4 x 16 DFMA = the lane-FMAs of 8 DMMAs
#pragma unroll
            for (int i = 0; i < 8; ++i) { c[i][0] = fma(a, b, c[i][0]); c[i][1] = fma(a, b, c[i][1]); }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

void run_mixed(double *out)
{
    const int iters = 20000;
    for (int mode = 0; mode < 3; ++mode) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        mixed<<<148 * 3, 256>>>(out, 100, mode);
        cudaEventRecord(e0);
        mixed<<<148 * 3, 256>>>(out, iters, mode);
        cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double flop = (double) iters * 8 * 512 * 148 * 3 * 8;       // every warp does the same number of flops
        printf("mixed mode %d (0 all DFMA, 1 all DMMA, 2 half / half): %.3f ms  %.2f TFLOP/s\n", mode, ms, flop / (ms * 1e-3) / 1e12);
    }
}

int main()
{
    double *out; long long *clk;
    cudaMalloc(&out, 8 * 148 * 3 * 256 * 2); cudaMalloc(&clk, 64);
    // latency: one warp, one accumulator
    run<0, 1>("DFMA x2", 2 * 2 * 32, out, clk, 1, 1);
    run<1, 1>("m8n8k4", 2 * 8 * 8 * 4, out, clk, 1, 1);
    run<2, 1>("m16n8k4", 2 * 16 * 8 * 4, out, clk, 1, 1);
    run<3, 1>("m16n8k8", 2 * 16 * 8 * 8, out, clk, 1, 1);
    // one warp, independent accumulators: issue rate of a single warp
    run<1, 8>("m8n8k4", 2 * 8 * 8 * 4, out, clk, 1, 1);
    run<2, 8>("m16n8k4", 2 * 16 * 8 * 4, out, clk, 1, 1);
    run<3, 8>("m16n8k8", 2 * 16 * 8 * 8, out, clk, 1, 1);
    // throughput at the invert kernel's occupancy
    for (int ctas = 1; ctas <= 3; ++ctas) {
        run<0, 8>("DFMA x2", 2 * 2 * 32, out, clk, ctas, 8);
        run<1, 8>("m8n8k4", 2 * 8 * 8 * 4, out, clk, ctas, 8);
        run<2, 8>("m16n8k4", 2 * 16 * 8 * 4, out, clk, ctas, 8);
        run<3, 8>("m16n8k8", 2 * 16 * 8 * 8, out, clk, ctas, 8);
    }
    run<1, 2>("m8n8k4", 2 * 8 * 8 * 4, out, clk, 3, 8);
    run<3, 2>("m16n8k8", 2 * 16 * 8 * 8, out, clk, 3, 8);
    run_mixed(out);
    return 0;
}
