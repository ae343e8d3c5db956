// rhome_y.cu -- the wavenumber-independent linearisation (linearize::rhome_y,
// apps/perfect/operator_hybrid_isothermal.cpp:691-761): ONE operator P (M + phi L(0,0))^T P^T
// (suzerain_rholut_imexop_packf00 == packf at km = kn = 0, rholut_imexop.h:330-344), one
// zgbtrf, then supply_B / rhs BC / zgbtrs('T') / demand_X for every (kx,kz) pencil.
//
// On the device: assemble + wall BCs + factor once with the pre-assembled batched kernels
// (imexop.cu, gbsv.cu), repack U row-wise with reciprocal diagonal (so that the warps' loads
// are contiguous and the column step multiplies instead of dividing), then one warp per
// pencil: right hand side in shared memory, U^T forward sweep, L^T backward sweep with the
// interchanges undone.  All warps stream the same 1.3 MB of factors: L1 / L2 traffic.
#include <algorithm>
#include <cstring>

#include "szb_internal.hpp"
#include "cplx.cuh"

namespace szb {

namespace {

// Urow[j*(kv+1) + c] = c == 0 ? 1 / U(j,j) : U(j, j+c)   (zero past the matrix)
__global__ void repack_u_kernel(int n, int kl, int ku, const cplx *ab, int ldab, cplx *urow)
{
    const int kv = kl + ku, total = n * (kv + 1);
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int j = e / (kv + 1), c = e - j * (kv + 1);
        cplx v(0.0, 0.0);
        if (j + c < n) v = ab[(size_t) (j + c) * ldab + (kv - c)];
        if (c == 0) v = recip(v);
        urow[e] = v;
    }
}

struct Solve00Args {
    int N, n, kl, ku, ldab;
    const cplx *ab, *urow; const int *ipiv; const int *info1;
    int npencil; const int *index;
    cplx *state; size_t fs, ps;
    int with_bc, wall_begin, wall_end;
    int *ipiv_out, *info_out;
};

__global__ void __launch_bounds__(256)
solve00_kernel(const Solve00Args A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int N = A.N, n = A.n, kv = A.kl + A.ku, kl = A.kl;
    cplx *x = reinterpret_cast<cplx *>(smem_raw) + (size_t) warp * N;
    const int info = *A.info1;
    for (int p = blockIdx.x * nw + warp; p < A.npencil; p += gridDim.x * nw) {
        if (lane == 0) A.info_out[p] = info;
        if (A.ipiv_out)
            for (int k = lane; k < N; k += 32) A.ipiv_out[(size_t) p * N + k] = A.ipiv[k];
        if (info) continue;
        cplx *v = A.state + (A.index ? (size_t) A.index[p] : (size_t) p) * A.ps;
        // b = P state with the wall rows zeroed (bsmbsm_solver.hpp:150-156,
        // operator_hybrid_isothermal.cpp:516-525)
        for (int e = lane; e < N; e += 32) {
            const int f = e / n, y = e - f * n;
            cplx val = v[(size_t) f * A.fs + y];
            if (A.with_bc && f < 4 && ((y == 0 && A.wall_begin == 0) || (y == n - 1 && A.wall_end == 2)))
                val = cplx(0.0, 0.0);
            x[5 * y + f] = val;
        }
        __syncwarp();
        // U^T y = b, column oriented (ztbsv 'U','T','N')
        for (int j = 0; j < N; ++j) {
            const cplx *ur = A.urow + (size_t) j * (kv + 1);
            const cplx xj = x[j] * ur[0];
            const int cmax = min(kv, N - 1 - j);
            for (int c = 1 + lane; c <= cmax; c += 32) {
                cplx w = x[j + c];
                submul(w, ur[c], xj);
                x[j + c] = w;
            }
            __syncwarp();
            if (lane == 0) x[j] = xj;
        }
        __syncwarp();
        // L^T x = y: dot products with the multipliers, interchanges undone in reverse
        for (int j = N - 2; j >= 0; --j) {
            const int lm = min(kl, N - 1 - j);
            const cplx *Lj = A.ab + (size_t) j * A.ldab + kv;
            cplx s(0.0, 0.0);
            for (int i = 1 + lane; i <= lm; i += 32) addmul(s, Lj[i], x[j + i]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                s.x += __shfl_xor_sync(0xffffffffu, s.x, o);
                s.y += __shfl_xor_sync(0xffffffffu, s.y, o);
            }
            if (lane == 0) {
                cplx t = x[j] - s;
                const int l = A.ipiv[j] - 1;
                if (l != j) { const cplx u = x[l]; x[l] = t; t = u; }
                x[j] = t;
            }
            __syncwarp();
        }
        // state = P^T x (bsmbsm_solver.hpp:274-280)
        for (int e = lane; e < N; e += 32) {
            const int f = e / n, y = e - f * n;
            v[(size_t) f * A.fs + y] = x[5 * y + f];
        }
        __syncwarp();
    }
}

}  // namespace

// Returns 0 when done, < 0 on error.
int invert00_dispatch(const szb_imexop *op, const double phi[2], int npencil, const int *d_index,
                      cplx *d_state, size_t fs, size_t ps, int *d_ipiv, int *d_info, cudaStream_t stream)
{
    const int N = op->A.N, KL = op->A.KL, KU = op->A.KU, ldab = op->A.LD + KL, kv = KL + KU;
    // workspace: LU | Urow | ipiv | info | zero wavenumbers
    const size_t b_lu = sizeof(cplx) * (size_t) ldab * N, b_ur = sizeof(cplx) * (size_t) N * (kv + 1);
    const size_t b_ip = ((sizeof(int) * (size_t) N) + 15) & ~(size_t) 15;
    const size_t need = b_lu + b_ur + b_ip + 16 + 16;
    if (need > op->work00_bytes) {
        if (op->d_work00) SZB_CUDA_OK(cudaFree(op->d_work00));
        op->d_work00 = nullptr; op->work00_bytes = 0;
        SZB_CUDA_OK(cudaMalloc(&op->d_work00, need));
        op->work00_bytes = need;
    }
    unsigned char *w = static_cast<unsigned char *>(op->d_work00);
    cplx *LU = reinterpret_cast<cplx *>(w); w += b_lu;
    cplx *Urow = reinterpret_cast<cplx *>(w); w += b_ur;
    int *ipiv = reinterpret_cast<int *>(w); w += b_ip;
    int *info1 = reinterpret_cast<int *>(w); w += 16;
    double *zero = reinterpret_cast<double *>(w);
    SZB_CUDA_OK(cudaMemsetAsync(zero, 0, 16, stream));
    int rc = szb_imexop_pack_batch(op, phi, 1, zero, zero + 1, 1, 1, reinterpret_cast<szb_complex *>(LU), stream);
    if (rc) return rc;
    rc = szb_zgbtrf_batch(N, KL, KU, reinterpret_cast<szb_complex *>(LU), ldab, (size_t) ldab * N, ipiv, info1, 1, stream);
    if (rc) return rc;
    repack_u_kernel<<<64, 256, 0, stream>>>(N, KL, KU, LU, ldab, Urow);
    count_launch();
    Solve00Args A;
    A.N = N; A.n = op->n; A.kl = KL; A.ku = KU; A.ldab = ldab;
    A.ab = LU; A.urow = Urow; A.ipiv = ipiv; A.info1 = info1;
    A.npencil = npencil; A.index = d_index; A.state = d_state; A.fs = fs; A.ps = ps;
    A.with_bc = 1; A.wall_begin = op->iso.enforce_lower ? 0 : 1; A.wall_end = op->iso.enforce_upper ? 2 : 1;
    A.ipiv_out = d_ipiv; A.info_out = d_info;
    int nw = 8;
    size_t smem = sizeof(cplx) * (size_t) nw * N;
    while (smem > 96 * 1024 && nw > 1) { nw /= 2; smem = sizeof(cplx) * (size_t) nw * N; }
    if (smem > 48 * 1024)
        SZB_CUDA_OK(cudaFuncSetAttribute(solve00_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    const int per_sm = (int) std::max<size_t>(1, std::min<size_t>(4, (200 * 1024) / smem));
    const int grid = std::min((npencil + nw - 1) / nw, per_sm * op->sm_count);
    solve00_kernel<<<grid, 32 * nw, smem, stream>>>(A);
    count_launch();
    SZB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace szb
