// 2-D pitched vs row-by-row 1-D copies between a device mirror and pinned host memory, in the shape of the whole-field
// host entry points (rows of 96 active pencils x 480 complex = 737 KB at a pitch of 145 pencils).
//   nvcc -O2 -o tools/diag_pcie2 tools/diag_pcie2.cu
#include <cstdio>
#include <chrono>
#include <cuda_runtime.h>
int main()
{
    const size_t N = 480, nx = 145, xa = 96, rows = 145;
    const size_t pitch = 16 * N * nx, width = 16 * N * xa, total = pitch * rows;
    char *h, *d;
    cudaHostAlloc(&h, total, cudaHostAllocDefault); cudaMalloc(&d, total);
    memset(h, 1, total);
    cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    auto run = [&](const char *name, int mode, cudaMemcpyKind kind) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaStreamSynchronize(s);
            auto t0 = std::chrono::steady_clock::now();
            char *dst = kind == cudaMemcpyDeviceToHost ? h : d, *src = kind == cudaMemcpyDeviceToHost ? d : h;
            if (mode == 0) cudaMemcpy2DAsync(dst, pitch, src, pitch, width, rows, kind, s);
            else if (mode == 1) for (size_t r = 0; r < rows; ++r) cudaMemcpyAsync(dst + r * pitch, src + r * pitch, width, kind, s);
            else cudaMemcpyAsync(dst, src, width * rows, kind, s);
            cudaStreamSynchronize(s);
            double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            if (rep == 2) printf("%-34s %6.2f ms  %5.1f GB/s\n", name, ms, width * rows / ms / 1e6);
        }
    };
    run("H2D 2-D pitched", 0, cudaMemcpyHostToDevice);
    run("H2D row by row (145 x 1-D)", 1, cudaMemcpyHostToDevice);
    run("H2D one contiguous 1-D", 2, cudaMemcpyHostToDevice);
    run("D2H 2-D pitched", 0, cudaMemcpyDeviceToHost);
    run("D2H row by row (145 x 1-D)", 1, cudaMemcpyDeviceToHost);
    run("D2H one contiguous 1-D", 2, cudaMemcpyDeviceToHost);
    return 0;
}
