// fp64_peak.cu -- measures the FP64 FMA throughput of the device (the second
// roofline the fused banded factor/solve kernel is quoted against; no FP64 figure is
// in MEASURED_PEAKS.json).  Prints one JSON line.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_kernel(double *out, int iters, double a, double b)
{
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int ILP = 8, threads = 256, blocks = p.multiProcessorCount * 8, iters = 20000;
    double *out; cudaMalloc(&out, sizeof(double) * threads * blocks);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        dfma_kernel<ILP><<<blocks, threads>>>(out, iters, 0.999999, 1e-7);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    const double fma = (double) blocks * threads * ILP * iters;
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"device\": \"%s\", \"sms\": %d, \"fp64_tflops\": %.3f, \"dfma_per_clk_per_sm_at_max_clock\": %.2f, \"ms\": %.3f}\n",
           p.name, p.multiProcessorCount, 2 * fma / (best * 1e-3) / 1e12,
           fma / (best * 1e-3) / ((double) clk * 1e3) / p.multiProcessorCount, best);
    return 0;
}
