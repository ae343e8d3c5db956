// physical.cu -- physical-space sweep that PRODUCES the reference profiles of the linear operator:
// collect_references (apps/perfect/perfect.cpp:1266-1400).  For every wall-normal plane y the 42
// quantities of apps/perfect/references.hpp:83-128 (rho, p, T, u_i u_j, nu u_i u_j, the coefficients
// of the explicit energy terms, rho u_i u_j ...) are summed over the plane's (z, x) points; after the
// caller's all-reduce over ranks and the chi scaling, rows 5..30 of the 42 x Ny result are what
// szb_imexop_set_refs_device hands to the operator.
//
// Five doubles in per point, 42 x Ny doubles out; per point one division, T^beta, one square root and ~90
// multiply-adds, which is what binds it (FP64 pipe), not HBM.  Two deterministic stages:
//   1. a block of 128 threads owns a contiguous slice of one plane; every thread keeps the 42 running
//      sums of its points in registers, the block folds them by warp shuffles and writes one partial
//      row per block;
//   2. one thread per (quantity, y) adds the plane's partial rows in a fixed order.
// No atomics: the same input gives the same bits on every run, whatever the block scheduling.
#include <algorithm>

#include "invert_common.cuh"

namespace szb {
namespace {

constexpr int NQ = 42;              // references::q::count
constexpr int CR_THREADS = 128;

struct CollectArgs {
    double alpha, beta, gamma, Ma;
    const double *e, *mx, *my, *mz, *rho;     // [ny][nzx] each
    size_t nzx;
    int ny, nbx, inviscid_plane;              // local plane index whose mu, lambda are switched off (-1: none)
    double *partial;                          // [ny][nbx][NQ]
};

// The 42 quantities of one point, in row order (perfect.cpp:1296-1376, suzerain/rholut.hpp:769-790,
// suzerain/rholt.hpp:675-709, 1481-1489).  The kernel is bound by the FP64 pipe, not by HBM, so the arithmetic is
// kept short: one division (1 / rho; the reference divides by rho a dozen times), T^beta as exp(beta log T)
// (relative error of a few ulp against pow), products accumulated by FMA.  All within the 1e-12 tolerance of
// the path; the reference's own sums are Kahan-compensated, ours are per-thread partial sums folded in a tree.
__device__ __forceinline__ void accumulate_point(double (&acc)[NQ], double beta, double gamma, double Ma,
                                                 double e, double mx, double my, double mz, double rho, bool viscous)
{
    const double rinv = 1.0 / rho;
    const double ux = rinv * mx, uy = rinv * my, uz = rinv * mz;
    const double m2 = fma(mx, mx, fma(my, my, mz * mz));
    const double p = (gamma - 1.0) * (e - Ma * Ma * rinv * m2 * 0.5);
    const double T = gamma * p * rinv;
    const double mu = viscous ? exp(beta * log(T)) : 0.0;
    const double uxx = ux * ux, uxy = ux * uy, uxz = ux * uz, uyy = uy * uy, uyz = uy * uz, uzz = uz * uz;
    const double u2 = uxx + uyy + uzz;
    const double nu = mu * rinv;
    acc[0] += rho; acc[1] += p; acc[2] = fma(p, p, acc[2]); acc[3] += T; acc[4] += sqrt(T);
    acc[5] += ux; acc[6] += uy; acc[7] += uz; acc[8] += u2;
    acc[9] += uxx; acc[10] += uxy; acc[11] += uxz; acc[12] += uyy; acc[13] += uyz; acc[14] += uzz;
    acc[15] += nu; acc[16] = fma(nu, ux, acc[16]); acc[17] = fma(nu, uy, acc[17]); acc[18] = fma(nu, uz, acc[18]);
    acc[19] = fma(nu, u2, acc[19]);
    acc[20] = fma(nu, uxx, acc[20]); acc[21] = fma(nu, uxy, acc[21]); acc[22] = fma(nu, uxz, acc[22]);
    acc[23] = fma(nu, uyy, acc[23]); acc[24] = fma(nu, uyz, acc[24]); acc[25] = fma(nu, uzz, acc[25]);
    const double r2inv = rinv * rinv;
    const double cg = ((gamma - 2.0) * e - 2.0 * p) * r2inv;
    acc[26] = fma(cg, mx, acc[26]); acc[27] = fma(cg, my, acc[27]); acc[28] = fma(cg, mz, acc[28]);
    acc[29] = fma(e + p, rinv, acc[29]);
    acc[30] = fma(mu * r2inv, (gamma - 1.0) * e - 2.0 * p, acc[30]);
    acc[31] += mx; acc[32] += my; acc[33] += mz; acc[34] += e;
    acc[35] = fma(mx, ux, acc[35]); acc[36] = fma(mx, uy, acc[36]); acc[37] = fma(mx, uz, acc[37]);
    acc[38] = fma(my, uy, acc[38]); acc[39] = fma(my, uz, acc[39]); acc[40] = fma(mz, uz, acc[40]);
    acc[41] = fma(e * e, rinv, acc[41]);
}

__global__ void __launch_bounds__(CR_THREADS, 3)
collect_references_kernel(const CollectArgs A)
{
    __shared__ double s_part[CR_THREADS / 32][NQ];
    const int j = blockIdx.y, bx = blockIdx.x;
    // this block's slice of the plane: whole pairs of points, so that the loads are 16 bytes wide when the
    // plane starts on a 16-byte boundary
    const size_t per = ((A.nzx + A.nbx - 1) / A.nbx + 1) & ~(size_t) 1;
    const size_t lo = per * bx < A.nzx ? per * bx : A.nzx, hi = lo + per < A.nzx ? lo + per : A.nzx;
    const size_t base = (size_t) j * A.nzx;
    const bool viscous = j != A.inviscid_plane;
    double acc[NQ];
#pragma unroll
    for (int i = 0; i < NQ; ++i) acc[i] = 0.0;
    const bool vec = ((base + lo) & 1) == 0
                     && ((reinterpret_cast<size_t>(A.e) | reinterpret_cast<size_t>(A.mx) | reinterpret_cast<size_t>(A.my)
                          | reinterpret_cast<size_t>(A.mz) | reinterpret_cast<size_t>(A.rho)) & 15) == 0;
    if (vec) {
        const size_t npair = (hi - lo) >> 1;
        const double2 *e2 = reinterpret_cast<const double2 *>(A.e + base + lo);
        const double2 *x2 = reinterpret_cast<const double2 *>(A.mx + base + lo);
        const double2 *y2 = reinterpret_cast<const double2 *>(A.my + base + lo);
        const double2 *z2 = reinterpret_cast<const double2 *>(A.mz + base + lo);
        const double2 *r2 = reinterpret_cast<const double2 *>(A.rho + base + lo);
        // the loads of the next pair are issued before the arithmetic of this one (twelve warps per SM do not
        // hide an HBM round trip by themselves)
        size_t k = threadIdx.x;
        double2 e = make_double2(0, 0), mx = e, my = e, mz = e, rho = e;
        if (k < npair) { e = __ldg(e2 + k); mx = __ldg(x2 + k); my = __ldg(y2 + k); mz = __ldg(z2 + k); rho = __ldg(r2 + k); }
        while (k < npair) {
            const size_t kn = k + CR_THREADS;
            double2 en = e, mxn = mx, myn = my, mzn = mz, rhon = rho;
            if (kn < npair) { en = __ldg(e2 + kn); mxn = __ldg(x2 + kn); myn = __ldg(y2 + kn); mzn = __ldg(z2 + kn); rhon = __ldg(r2 + kn); }
            accumulate_point(acc, A.beta, A.gamma, A.Ma, e.x, mx.x, my.x, mz.x, rho.x, viscous);
            accumulate_point(acc, A.beta, A.gamma, A.Ma, e.y, mx.y, my.y, mz.y, rho.y, viscous);
            e = en; mx = mxn; my = myn; mz = mzn; rho = rhon; k = kn;
        }
        if (threadIdx.x == 0 && ((hi - lo) & 1)) {
            const size_t k = base + hi - 1;
            accumulate_point(acc, A.beta, A.gamma, A.Ma, A.e[k], A.mx[k], A.my[k], A.mz[k], A.rho[k], viscous);
        }
    } else {
        for (size_t k = base + lo + threadIdx.x; k < base + hi; k += CR_THREADS)
            accumulate_point(acc, A.beta, A.gamma, A.Ma, __ldg(A.e + k), __ldg(A.mx + k), __ldg(A.my + k),
                             __ldg(A.mz + k), __ldg(A.rho + k), viscous);
    }
    // fold: lanes, then warps (fixed order)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
        double v = acc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) s_part[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < NQ) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < CR_THREADS / 32; ++w) v += s_part[w][threadIdx.x];
        A.partial[((size_t) j * A.nbx + bx) * NQ + threadIdx.x] = v;
    }
}

// refs[i + 42 (y0 + j)] = scale * sum_bx partial[j][bx][i]; the other columns of the 42 x Ny block are zeroed
// (perfect.cpp:1275-1277: "must zero y(j) not present on rank")
__global__ void collect_references_finish_kernel(const double *partial, int ny, int nbx, int y0, int Ny, double scale,
                                                 double *refs)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= NQ * Ny) return;
    const int i = t % NQ, jg = t / NQ, j = jg - y0;
    double v = 0.0;
    if (j >= 0 && j < ny) {
        for (int b = 0; b < nbx; ++b) v += partial[((size_t) j * nbx + b) * NQ + i];
        v *= scale;
    }
    refs[t] = v;
}

}  // namespace
}  // namespace szb

using namespace szb;

extern "C" {

int szb_collect_references_device(const szb_rholut_imexop_scenario *scenario, double beta, int Ny, int y0, int ny,
        size_t nzx, const double *d_sphys, size_t field_stride, int top_is_inviscid, double scale,
        double *d_refs, void *d_workspace, size_t workspace_bytes, size_t *workspace_needed, void *stream)
{
    if (!scenario) return -1;
    if (Ny < 1) return -3;
    if (y0 < 0 || y0 > Ny) return -4;
    if (ny < 0 || y0 + ny > Ny) return -5;
    int dev = 0, sms = 148;
    SZB_CUDA_OK(cudaGetDevice(&dev));
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // blocks per plane: about two waves of three resident blocks per SM over all planes, at least 1 024 points per block
    int nbx = ny > 0 ? std::max(1, 6 * sms / ny) : 1;       // rounded down: ny * nbx blocks fill two waves, not two and a bit
    const size_t maxb = (nzx + 1023) / 1024;
    if ((size_t) nbx > maxb) nbx = (int) std::max<size_t>(1, maxb);
    const size_t need = sizeof(double) * (size_t) std::max(ny, 1) * nbx * NQ;
    if (workspace_needed) *workspace_needed = need;
    if (!d_refs && !d_sphys) return 0;                   // size query
    if (!d_sphys && ny > 0) return -7;
    if (ny > 0 && field_stride < (size_t) ny * nzx) return -8;
    if (!d_refs) return -11;
    if (!d_workspace || workspace_bytes < need) return -12;
    cudaStream_t st = (cudaStream_t) stream;
    CollectArgs A;
    A.alpha = scenario->alpha; A.beta = beta; A.gamma = scenario->gamma; A.Ma = scenario->Ma;
    A.e = d_sphys; A.mx = d_sphys + field_stride; A.my = d_sphys + 2 * field_stride;
    A.mz = d_sphys + 3 * field_stride; A.rho = d_sphys + 4 * field_stride;
    A.nzx = nzx; A.ny = ny; A.nbx = nbx;
    A.inviscid_plane = (top_is_inviscid && y0 + ny == Ny) ? ny - 1 : -1;
    A.partial = static_cast<double *>(d_workspace);
    if (ny > 0 && nzx > 0) {
        collect_references_kernel<<<dim3(nbx, ny), CR_THREADS, 0, st>>>(A);
        count_launch();
    } else if (ny > 0) {
        SZB_CUDA_OK(cudaMemsetAsync(d_workspace, 0, need, st));
    }
    collect_references_finish_kernel<<<(NQ * Ny + 255) / 256, 256, 0, st>>>(A.partial, ny, nbx, y0, Ny, scale, d_refs);
    count_launch();
    SZB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // extern "C"
