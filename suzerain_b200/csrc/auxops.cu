// auxops.cu -- the two wave-space building blocks of the nonlinear operator that share this
// path's data layout, both HBM-bound:
//
//  * szb_bsplineop_accumulate_complex_batch: y <- alpha D^(d) x + beta y over nrhs wall-normal
//    pencils (suzerain_bsplineop_accumulate_complex, suzerain/bsplineop.c:260-297, as batched
//    by operator_tools.hpp:77-116 over all local (kx,kz)).  Persistent CTAs; the pencils of the
//    next three groups are in flight as TMA bulk copies into a zero-halo ring while the current
//    group is computed (about 100 KB per SM on the wire: the latency x bandwidth product of HBM); the operator (ld x n doubles, r-major) stays in L1.
//  * szb_diffwave_{apply,accumulate}_batch: x <- alpha (i kx)^dx (i kz)^dz x  /
//    y <- alpha (i kx)^dx (i kz)^dz x + beta y with dealiased and Nyquist modes zeroed
//    (suzerain_diffwave_apply / _accumulate, suzerain/diffwave.c:65-198).  One warp per pencil:
//    the mode's factor is formed once with the reference's arithmetic, the pencil is streamed.
#include <algorithm>
#include <climits>
#include <mutex>

#include "invert_common.cuh"

namespace szb {

namespace {

// ------------------------------------------------------------------------------------------
// batched banded operator apply
// ------------------------------------------------------------------------------------------
constexpr int BOP_STAGES = 4;

struct BopArgs {
    const double *Dr;       // [ld][n]  Dr[r*n + i] = D[i, i - ku + r]
    int n, kl, ku, ld, nrhs, group, nthr;
    cplx alpha, beta;
    const cplx *x; size_t ldx;
    cplx *y; size_t ldy;
};

__global__ void __launch_bounds__(512)
bop_accumulate_kernel(const BopArgs A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NST = BOP_STAGES;                                // ring depth: enough bytes in flight per SM
    __shared__ __align__(8) unsigned long long s_mbar[NST];
    const int n = A.n, np = n + A.kl + A.ku, G = A.group;
    cplx *s_x = reinterpret_cast<cplx *>(smem_raw);                // [NST][G][ku + n + kl]
    for (int e = threadIdx.x; e < NST * G * np; e += blockDim.x) s_x[e] = cplx(0.0, 0.0);
    if (threadIdx.x == 0) {
        for (int b = 0; b < NST; ++b) fused::mbar_init(&s_mbar[b], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const unsigned bytes = (unsigned) (n * sizeof(cplx));
    const int ngroups = (A.nrhs + G - 1) / G;
    auto fetch = [&](int g, int stage) {
        const int r0 = g * G, cnt = min(G, A.nrhs - r0);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        fused::mbar_expect_tx(&s_mbar[stage], cnt * bytes);
        for (int q = 0; q < cnt; ++q)
            fused::tma_bulk_g2s(s_x + (stage * G + q) * np + A.ku, A.x + (size_t) (r0 + q) * A.ldx, bytes,
                                &s_mbar[stage]);
    };
    if (threadIdx.x == 0)
        for (int b = 0; b < NST - 1; ++b)
            if ((int) (blockIdx.x + b * gridDim.x) < ngroups) fetch(blockIdx.x + b * gridDim.x, b);
    const int q = threadIdx.x / A.nthr, i = threadIdx.x - q * A.nthr;
    const bool beta_zero = is_zero(A.beta);
    int it = 0;
    for (int g = blockIdx.x; g < ngroups; g += gridDim.x, ++it) {
        const int stage = it % NST;
        if (threadIdx.x == 0 && g + (NST - 1) * (int) gridDim.x < ngroups)
            fetch(g + (NST - 1) * gridDim.x, (it + NST - 1) % NST);
        const int rhs = g * G + q;
        cplx yold(0.0, 0.0);
        const bool live = i < n && rhs < A.nrhs;
        cplx *yp = A.y + (size_t) rhs * A.ldy + i;
        if (live && !beta_zero) yold = *yp;                       // in flight while the tile lands
        fused::mbar_wait(&s_mbar[stage], (it / NST) & 1);
        if (live) {
            const cplx *xs = s_x + (stage * G + q) * np + i;       // xs[r] = x[i - ku + r]
            const double *D = A.Dr + i;
            cplx s(0.0, 0.0);
#pragma unroll 4
            for (int r = 0; r < A.ld; ++r) addmul(s, xs[r], __ldg(D + (size_t) r * n));
            *yp = beta_zero ? A.alpha * s : A.alpha * s + A.beta * yold;
        }
        __syncthreads();                                           // the stage may be refilled
    }
}

// The same for real coefficients (suzerain_bsplineop_accumulate / _apply, suzerain/bsplineop.c:222-258, 299-337).
// Pencils of n doubles need not be 16-byte aligned, so the group is staged with plain coalesced loads instead of
// TMA bulk copies; a group is read completely before any of it is written, which makes x == y (in place) safe.
struct BopRealArgs {
    const double *Dr;
    int n, kl, ku, ld, nrhs, group, nthr;
    double alpha, beta;
    const double *x; size_t ldx;
    double *y; size_t ldy;
};

__global__ void __launch_bounds__(512)
bop_real_kernel(const BopRealArgs A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = A.n, np = n + A.kl + A.ku, G = A.group;
    double *s_x = reinterpret_cast<double *>(smem_raw);            // [G][ku + n + kl], zero halo
    for (int e = threadIdx.x; e < G * np; e += blockDim.x) s_x[e] = 0.0;
    __syncthreads();
    const int q = threadIdx.x / A.nthr, i = threadIdx.x - q * A.nthr;
    const int ngroups = (A.nrhs + G - 1) / G;
    for (int g = blockIdx.x; g < ngroups; g += gridDim.x) {
        const int rhs = g * G + q;
        const bool live = i < n && rhs < A.nrhs;
        double yold = 0.0;
        if (live) {
            s_x[q * np + A.ku + i] = A.x[(size_t) rhs * A.ldx + i];
            if (A.beta != 0.0) yold = A.y[(size_t) rhs * A.ldy + i];
        }
        __syncthreads();
        if (live) {
            const double *xs = s_x + q * np + i;
            const double *D = A.Dr + i;
            double s = 0.0;
#pragma unroll 4
            for (int r = 0; r < A.ld; ++r) s = fma(xs[r], __ldg(D + (size_t) r * n), s);
            A.y[(size_t) rhs * A.ldy + i] = A.beta != 0.0 ? A.alpha * s + A.beta * yold : A.alpha * s;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// diffwave
// ------------------------------------------------------------------------------------------
struct DiffwaveArgs {
    int dxcnt, dzcnt;
    cplx alpha, beta;
    const cplx *x; cplx *y;            // y == nullptr: apply (x is written)
    cplx *xw;
    double twopioverLx, twopioverLz;
    int Ny, Nx, dNx, dkbx, dkex, Nz, dNz, dkbz, dkez;
};

__device__ __forceinline__ int wavenumber_abs(int N, int i) { return (i < N / 2 + 1) ? i : N - i; }
__device__ __forceinline__ int wavenumber_diff(int N, int dN, int i)      // inorder.h:282-293
{
    if (i < (N + 1) / 2) return i;
    if (i >= dN - (N - 1) / 2) return -dN + i;
    return 0;
}
// GSL's gsl_sf_pow_int: binary powering (only n >= 0 occurs here)
__device__ __forceinline__ double pow_int(double x, int n)
{
    double value = 1.0;
    do {
        if (n & 1) value = __dmul_rn(value, x);
        n >>= 1;
        x = __dmul_rn(x, x);
    } while (n);
    return value;
}

__global__ void __launch_bounds__(256)
diffwave_kernel(const DiffwaveArgs A)
{
    const int lane = threadIdx.x & 31;
    const int nx = A.dkex - A.dkbx, nz = A.dkez - A.dkbz;
    const long long npencil = (long long) nx * nz;
    const long long warp0 = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarp = ((long long) gridDim.x * blockDim.x) >> 5;
    // alpha * i^(dxcnt + dzcnt)   (diffwave.c:37-48)
    cplx aip;
    switch ((A.dxcnt + A.dzcnt) & 3) {
    case 0:  aip = A.alpha; break;
    case 1:  aip = cplx(-A.alpha.y, A.alpha.x); break;
    case 2:  aip = cplx(-A.alpha.x, -A.alpha.y); break;
    default: aip = cplx(A.alpha.y, -A.alpha.x); break;
    }
    const int absmin_wm = (A.Nx - 1) / 2, absmin_wn = (A.Nz - 1) / 2;
    for (long long pcl = warp0; pcl < npencil; pcl += nwarp) {
        const int nn = (int) (pcl / nx), mm = (int) (pcl - (long long) nn * nx);
        const int n = A.dkbz + nn, m = A.dkbx + mm;
        int nfreq = INT_MAX, mfreq = INT_MAX;
        const bool nkeep = A.dzcnt > 0 ? (nfreq = wavenumber_diff(A.Nz, A.dNz, n)) != 0
                                       : wavenumber_abs(A.dNz, n) <= absmin_wn;
        const bool mkeep = A.dxcnt > 0 ? (mfreq = wavenumber_diff(A.Nx, A.dNx, m)) != 0
                                       : wavenumber_abs(A.dNx, m) <= absmin_wm;
        const size_t off = (size_t) pcl * A.Ny;
        if (nkeep && mkeep) {
            // the reference's association: mscale = nscale * pow(...); malpha = mscale * alpha_ipow
            const double nscale = pow_int(__dmul_rn(A.twopioverLz, (double) nfreq), A.dzcnt);
            const double mscale = __dmul_rn(nscale, pow_int(__dmul_rn(A.twopioverLx, (double) mfreq), A.dxcnt));
            const cplx ma(__dmul_rn(mscale, aip.x), __dmul_rn(mscale, aip.y));
            if (A.y == nullptr) {
                for (int y = lane; y < A.Ny; y += 32) A.xw[off + y] = ma * A.x[off + y];
            } else {
                for (int y = lane; y < A.Ny; y += 32) A.y[off + y] = ma * A.x[off + y] + A.beta * A.y[off + y];
            }
        } else if (A.y == nullptr) {
            for (int y = lane; y < A.Ny; y += 32) A.xw[off + y] = cplx(0.0, 0.0);
        } else {
            for (int y = lane; y < A.Ny; y += 32) A.y[off + y] = A.beta * A.y[off + y];
        }
    }
}

std::mutex g_mutex;

}  // namespace
}  // namespace szb

using namespace szb;

extern "C" {

// device copy of all operators, r-major with the common (max) bandwidths:
// Dr[(d*ld + r)*n + i] = D^(d)[i, i - ku + r]; made once per device
static int bop_upload(const szb_bsplineop *w, int dev)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (w->d_Dr && w->d_dev != dev) { cudaFree(w->d_Dr); w->d_Dr = nullptr; }
    if (!w->d_Dr) {
        const int n = w->n, ld = w->ld;
        std::vector<double> h((size_t) (w->nderiv + 1) * ld * n);
        for (int dd = 0; dd <= w->nderiv; ++dd) {
            const double *blk = w->storage.data() + (size_t) dd * ld * n;
            for (int i = 0; i < n; ++i)
                for (int r = 0; r < ld; ++r) h[((size_t) dd * ld + r) * n + i] = blk[(size_t) i * ld + r];
        }
        SZB_CUDA_OK(cudaMalloc(&w->d_Dr, h.size() * sizeof(double)));
        SZB_CUDA_OK(cudaMemcpy(w->d_Dr, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
        w->d_dev = dev;
    }
    return 0;
}

static int bop_complex_launch(const szb_bsplineop *w, int d, int nrhs, cplx alpha, const szb_complex *d_x, size_t ldx,
                              cplx beta, szb_complex *d_y, size_t ldy, void *stream)
{
    int dev = 0;
    SZB_CUDA_OK(cudaGetDevice(&dev));
    if (int rc = bop_upload(w, dev)) return rc;
    BopArgs A;
    A.Dr = w->d_Dr + (size_t) d * w->ld * w->n;
    A.n = w->n; A.kl = w->max_kl; A.ku = w->max_ku; A.ld = w->ld; A.nrhs = nrhs;
    A.nthr = (w->n + 31) / 32 * 32;
    if (A.nthr > 512) return -1;
    A.group = std::max(1, 384 / A.nthr);
    A.alpha = alpha; A.beta = beta;
    A.x = reinterpret_cast<const cplx *>(d_x); A.ldx = ldx;
    A.y = reinterpret_cast<cplx *>(d_y); A.ldy = ldy;
    const size_t smem = sizeof(cplx) * BOP_STAGES * (size_t) A.group * (A.n + A.kl + A.ku);
    if (smem > 48 * 1024)
        SZB_CUDA_OK(cudaFuncSetAttribute(bop_accumulate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int ngroups = (nrhs + A.group - 1) / A.group;
    bop_accumulate_kernel<<<std::min(ngroups, 4 * sms), A.group * A.nthr, smem, (cudaStream_t) stream>>>(A);
    count_launch();
    SZB_CUDA_OK(cudaGetLastError());
    return 0;
}

static int bop_real_launch(const szb_bsplineop *w, int d, int nrhs, double alpha, const double *d_x, size_t ldx,
                           double beta, double *d_y, size_t ldy, void *stream)
{
    int dev = 0;
    SZB_CUDA_OK(cudaGetDevice(&dev));
    if (int rc = bop_upload(w, dev)) return rc;
    BopRealArgs A;
    A.Dr = w->d_Dr + (size_t) d * w->ld * w->n;
    A.n = w->n; A.kl = w->max_kl; A.ku = w->max_ku; A.ld = w->ld; A.nrhs = nrhs;
    A.nthr = (w->n + 31) / 32 * 32;
    if (A.nthr > 512) return -1;
    A.group = std::max(1, 512 / A.nthr);
    A.alpha = alpha; A.beta = beta;
    A.x = d_x; A.ldx = ldx; A.y = d_y; A.ldy = ldy;
    const size_t smem = sizeof(double) * (size_t) A.group * (A.n + A.kl + A.ku);
    if (smem > 48 * 1024)
        SZB_CUDA_OK(cudaFuncSetAttribute(bop_real_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int ngroups = (nrhs + A.group - 1) / A.group;
    bop_real_kernel<<<std::min(ngroups, 4 * sms), A.group * A.nthr, smem, (cudaStream_t) stream>>>(A);
    count_launch();
    SZB_CUDA_OK(cudaGetLastError());
    return 0;
}

int szb_bsplineop_accumulate_complex_batch(const szb_bsplineop *w, int d, int nrhs,
        const double alpha[2], const szb_complex *d_x, size_t ldx,
        const double beta[2], szb_complex *d_y, size_t ldy, void *stream)
{
    if (!w) return -1;
    if (d < 0 || d > w->nderiv) return -2;
    if (nrhs < 0) return -3;
    if (!alpha) return -4;
    if (!d_x) return -5;
    if (ldx < (size_t) w->n) return -6;
    if (!beta) return -7;
    if (!d_y) return -8;
    if (ldy < (size_t) w->n) return -9;
    if ((const void *) d_x == (const void *) d_y) return -8;        // bsplineop.c:283-286
    if (nrhs == 0) return 0;
    return bop_complex_launch(w, d, nrhs, cplx(alpha[0], alpha[1]), d_x, ldx, cplx(beta[0], beta[1]), d_y, ldy, stream);
}

int szb_bsplineop_apply_complex_batch(const szb_bsplineop *w, int d, int nrhs, double alpha,
        szb_complex *d_x, size_t ldx, void *stream)
{
    if (!w) return -1;
    if (d < 0 || d > w->nderiv) return -2;
    if (nrhs < 0) return -3;
    if (!d_x) return -5;
    if (ldx < (size_t) w->n) return -6;
    if (nrhs == 0) return 0;
    // every pencil is in shared memory before its result is stored, so the scratch copy of the
    // reference (bsplineop.c:356-376) is the staging buffer
    return bop_complex_launch(w, d, nrhs, cplx(alpha, 0.0), d_x, ldx, cplx(0.0, 0.0), d_x, ldx, stream);
}

int szb_bsplineop_accumulate_batch(const szb_bsplineop *w, int d, int nrhs, double alpha,
        const double *d_x, size_t ldx, double beta, double *d_y, size_t ldy, void *stream)
{
    if (!w) return -1;
    if (d < 0 || d > w->nderiv) return -2;
    if (nrhs < 0) return -3;
    if (!d_x) return -5;
    if (ldx < (size_t) w->n) return -6;
    if (!d_y) return -8;
    if (ldy < (size_t) w->n) return -9;
    if ((const void *) d_x == (const void *) d_y) return -8;        // bsplineop.c:244-247
    if (nrhs == 0) return 0;
    return bop_real_launch(w, d, nrhs, alpha, d_x, ldx, beta, d_y, ldy, stream);
}

int szb_bsplineop_apply_batch(const szb_bsplineop *w, int d, int nrhs, double alpha,
        double *d_x, size_t ldx, void *stream)
{
    if (!w) return -1;
    if (d < 0 || d > w->nderiv) return -2;
    if (nrhs < 0) return -3;
    if (!d_x) return -5;
    if (ldx < (size_t) w->n) return -6;
    if (nrhs == 0) return 0;
    return bop_real_launch(w, d, nrhs, alpha, d_x, ldx, 0.0, d_x, ldx, stream);
}

static int diffwave_launch(int dxcnt, int dzcnt, const double alpha[2], const szb_complex *d_x,
                           szb_complex *d_xw, const double beta[2], szb_complex *d_y,
                           const szb_wavegrid *g, int Ny, void *stream)
{
    DiffwaveArgs A;
    A.dxcnt = dxcnt; A.dzcnt = dzcnt;
    A.alpha = cplx(alpha[0], alpha[1]);
    A.beta = beta ? cplx(beta[0], beta[1]) : cplx(0.0, 0.0);
    A.x = reinterpret_cast<const cplx *>(d_x); A.xw = reinterpret_cast<cplx *>(d_xw);
    A.y = reinterpret_cast<cplx *>(d_y);
    // 2 pi / L exactly as diffwave.c:50-61 forms it (no contraction: a product and a quotient)
    volatile double twopi = 2 * 3.1415926535897932384626433832795028841971693993751058209;
    A.twopioverLx = twopi / g->Lx; A.twopioverLz = twopi / g->Lz;
    A.Ny = Ny; A.Nx = g->Nx; A.dNx = g->dNx; A.dkbx = g->dkbx; A.dkex = g->dkex;
    A.Nz = g->Nz; A.dNz = g->dNz; A.dkbz = g->dkbz; A.dkez = g->dkez;
    const long long npencil = (long long) (g->dkex - g->dkbx) * (g->dkez - g->dkbz);
    if (npencil <= 0 || Ny == 0) return 0;
    int dev = 0, sms = 148;
    SZB_CUDA_OK(cudaGetDevice(&dev));
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long blocks = std::min<long long>((npencil + 7) / 8, 8LL * sms);
    diffwave_kernel<<<(unsigned) blocks, 256, 0, (cudaStream_t) stream>>>(A);
    count_launch();
    SZB_CUDA_OK(cudaGetLastError());
    return 0;
}

int szb_diffwave_apply_batch(int dxcnt, int dzcnt, const double alpha[2], szb_complex *d_x,
                             const szb_wavegrid *g, int Ny, void *stream)
{
    if (dxcnt < 0) return -1;
    if (dzcnt < 0) return -2;
    if (!alpha) return -3;
    if (!d_x) return -4;
    if (!g) return -5;
    if (Ny < 0) return -6;
    return diffwave_launch(dxcnt, dzcnt, alpha, d_x, d_x, nullptr, nullptr, g, Ny, stream);
}

int szb_diffwave_accumulate_batch(int dxcnt, int dzcnt, const double alpha[2], const szb_complex *d_x,
                                  const double beta[2], szb_complex *d_y, const szb_wavegrid *g,
                                  int Ny, void *stream)
{
    if (dxcnt < 0) return -1;
    if (dzcnt < 0) return -2;
    if (!alpha) return -3;
    if (!d_x) return -4;
    if (!beta) return -5;
    if (!d_y) return -6;
    if ((const void *) d_x == (const void *) d_y) return -6;        // diffwave.c:145
    if (!g) return -7;
    if (Ny < 0) return -8;
    return diffwave_launch(dxcnt, dzcnt, alpha, d_x, nullptr, beta, d_y, g, Ny, stream);
}

}  // extern "C"
