// gbsv.cu -- batched complex banded LU / solve / iterative refinement and the
// batched invert of (M + phi L) built on them.
//
// Replaces, one launch for all systems instead of one LAPACK call per pencil:
//   suzerain_lapack_zgbtrf / zgbtrs          suzerain/blas_et_al/lapack.c:185-197,261-277
//   suzerain_lapackext_zcgbsvx               suzerain/blas_et_al/dsgbsvx.def:71-318
//   bsmbsm_solver_{zgbsv,zcgbsvx}::solve_hook  suzerain/bsmbsm_solver.cpp:155-182,377-414
//   suzerain_bsmbsm_zaPxpby                  suzerain/bsmbsm_aPxpby_complex.def:37-336
//   the hot loop of invert_mass_plus_scaled_operator
//                                            apps/perfect/operator_hybrid_isothermal.cpp:617-686
#include <cfloat>
#include <cstdio>
#include <cstdlib>

#include "szb_internal.hpp"
#include "cplx.cuh"
#include "kernels.cuh"

namespace szb {

// Carve the LU scratch out of dynamic shared memory.
__device__ __forceinline__ LuScratch lu_scratch(unsigned char *base, int kl, int ku)
{
    LuScratch S;
    S.col = reinterpret_cast<cplx *>(base);
    S.l   = S.col + (kl + 1);
    S.u   = S.l + kl;
    S.ibuf = reinterpret_cast<int *>(S.u + (kl + ku + 1));
    return S;
}
static inline size_t lu_scratch_bytes(int kl, int ku)
{ return sizeof(cplx) * (size_t) (kl + 1 + kl + kl + ku + 1) + 4 * sizeof(int); }

// CTA-wide deterministic sum over threads of a double; result to all threads.
__device__ inline double cta_sum(double v, double *s_red)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[w] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < nw; ++i) t += s_red[i];
    return t;
}

__device__ inline double cta_nrm2(int n, const cplx *v, double *s_red)
{
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += v[i].x * v[i].x + v[i].y * v[i].y;
    return sqrt(cta_sum(s, s_red));
}

// ---------------------------------------------------------------------------
// zgbtrf / zgbtrs batch kernels
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
zgbtrf_batch_kernel(int n, int kl, int ku, cplx *ab, int ldab, size_t stride,
                    int *ipiv, int *info)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const LuScratch S = lu_scratch(smem_raw, kl, ku);
    const size_t b = blockIdx.x;
    const int rc = gbtrf_cta(n, kl, ku, ab + b * stride, ldab, ipiv + b * n, S);
    if (threadIdx.x == 0) info[b] = rc;
}

__global__ void __launch_bounds__(128)
zgbtrs_batch_kernel(char trans, int n, int kl, int ku, int nrhs, const cplx *ab, int ldab,
                    size_t stride, const int *ipiv, cplx *bmat, int ldb, size_t strideb,
                    int nbatch)
{
    const size_t w = (size_t) blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= (size_t) nbatch * nrhs) return;
    const size_t sys = w / nrhs, rhs = w - sys * nrhs;
    cplx *b = bmat + sys * strideb + rhs * (size_t) ldb;
    if (trans == 'T') gbtrs_T_warp(n, kl, ku, ab + sys * stride, ldab, ipiv + sys * n, b);
    else              gbtrs_N_warp(n, kl, ku, ab + sys * stride, ldab, ipiv + sys * n, b);
}

// ---------------------------------------------------------------------------
// Iterative refinement around the double-precision factorisation, CTA-wide.
// Follows dsgbsvx.def:131-318 for fact == 'N' on entry with siter < 0
// (no single-precision attempt): x = 0, r = b; while res > tol and
// diter < dmax: r <- op(LU)^{-1} r; x += r; r = b - op(A) x; stop on
// stagnation (lastres < 2 res once diter >= aiter).  `factored` lets a caller
// reuse the factorisation for further right hand sides (fact == 'D').
// Returns info; *diter_out = refinement counter as the reference reports it.
// ---------------------------------------------------------------------------
__device__ inline int zcgbsvx_cta(char trans, int n, int kl, int ku, int aiter, int dmax,
                                  double tolsc, const cplx *ab, cplx *afb, int *ipiv,
                                  const cplx *b, cplx *x, cplx *r, bool &factored,
                                  int *diter_out, double *res_out,
                                  const LuScratch S, double *s_red)
{
    const int ldab = kl + 1 + ku, ldafb = 2 * kl + 1 + ku;
    const double eps = DBL_EPSILON * 0.5;               // dlamch('E')
    for (int i = threadIdx.x; i < n; i += blockDim.x) { x[i] = cplx(0.0, 0.0); r[i] = b[i]; }
    __syncthreads();
    double res = cta_nrm2(n, r, s_red);
    double lastres = 3.0 * (res + 1.0);
    const bool use_eps = (tolsc == 0.0);
    double tolconst = 1.0, tol = eps;
    if (!use_eps) {
        double s = 0.0;                                   // zlangb('F')
        for (int e = threadIdx.x; e < n * ldab; e += blockDim.x) {
            const int j = e / ldab, rr = e - j * ldab, i = j - ku + rr;
            if (i >= 0 && i < n) { const cplx v = ab[e]; s += v.x * v.x + v.y * v.y; }
        }
        const double afrob = sqrt(cta_sum(s, s_red));
        tolconst = afrob * eps * sqrt((double) n) * tolsc;
        tol = 0.0;
    }
    int diter = -1, info = 0;
    if (dmax >= 0 && res > tol) {
        if (!factored) {
            for (int e = threadIdx.x; e < n * ldab; e += blockDim.x) {    // zlacpy
                const int j = e / ldab, rr = e - j * ldab;
                afb[(size_t) j * ldafb + kl + rr] = ab[e];
            }
            __syncthreads();
            info = gbtrf_cta(n, kl, ku, afb, ldafb, ipiv, S);
            factored = true;
            if (info > 0) { *diter_out = diter; *res_out = res; return info; }
        }
        while (diter < dmax && res > tol) {
            ++diter;
            if (threadIdx.x < 32) {
                if (trans == 'T') gbtrs_T_warp(n, kl, ku, afb, ldafb, ipiv, r);
                else              gbtrs_N_warp(n, kl, ku, afb, ldafb, ipiv, r);
            }
            __syncthreads();
            for (int i = threadIdx.x; i < n; i += blockDim.x) x[i] += r[i];
            __syncthreads();
            if (!use_eps) tol = cta_nrm2(n, x, s_red) * tolconst;
            gb_residual_cta(trans, n, kl, ku, ab, ldab, x, b, r);
            __syncthreads();
            res = cta_nrm2(n, r, s_red);
            if (diter >= aiter && lastres < res * 2.0) break;
            lastres = res;
        }
    } else {
        diter = 0;
    }
    if (res != res) {                                    // NaN right hand side
        for (int i = threadIdx.x; i < n; i += blockDim.x) x[i] = cplx(NAN, NAN);
    }
    __syncthreads();
    *diter_out = diter; *res_out = res;
    return info;
}

// CTA-wide max / min over threads of a double; result to all threads.
__device__ inline double cta_max(double v, double *s_red)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[w] = v;
    __syncthreads();
    double t = s_red[0];
    for (int i = 1; i < nw; ++i) t = fmax(t, s_red[i]);
    return t;
}
__device__ inline double cta_min(double v, double *s_red) { return -cta_max(-v, s_red); }

// ---------------------------------------------------------------------------
// zgbsvx (LAPACK expert driver) for TRANS = 'T' and one right hand side, as
// bsmbsm_solver_zgbsvx::solve_hook calls it (suzerain/bsmbsm_solver.cpp:253-289) with
// FACT = 'E' (spec equil=true) or 'N': optional equilibration (zgbequ + zlaqgb, which
// scales ab in place), zgbtrf, zgbtrs, then zgbrfs: componentwise-backward-error
// refinement, at most ITMAX = 5 corrections, stopping once berr <= eps or berr no longer
// halves.  The solution is unscaled by the row factors afterwards.  rcond / ferr (zgbcon,
// zlacn2) are statistics of the reference's log and are not produced here.
// equed: bit 0 row scaling, bit 1 column scaling (kept across right hand sides).
// ---------------------------------------------------------------------------
__device__ inline int zgbsvx_cta(bool equil, int n, int kl, int ku, cplx *ab, cplx *afb, int *ipiv,
                                 cplx *b, cplx *x, cplx *r, double *rw, double *rs, double *cs,
                                 bool &factored, int &equed, int *count_out, double *berr_out,
                                 const LuScratch S, double *s_red)
{
    const int ldab = kl + 1 + ku, ldafb = 2 * kl + 1 + ku;
    const double eps = DBL_EPSILON * 0.5, safmin = DBL_MIN;          // dlamch('E'), dlamch('S')
    int info = 0;
    if (!factored) {
        equed = 0;
        if (equil) {
            const double smlnum = safmin, bignum = 1.0 / smlnum;
            // zgbequ: r_i = 1 / max_j |a_ij|, c_j = 1 / max_i r_i |a_ij|  (|.| = |re| + |im|)
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                double m = 0.0;
                for (int j = max(0, i - kl); j <= min(n - 1, i + ku); ++j)
                    m = fmax(m, cabs1(ab[(size_t) j * ldab + ku + i - j]));
                rs[i] = m;
            }
            __syncthreads();
            double lmax = 0.0, lmin = bignum;
            for (int i = threadIdx.x; i < n; i += blockDim.x) { lmax = fmax(lmax, rs[i]); lmin = fmin(lmin, rs[i]); }
            const double rcmax = cta_max(lmax, s_red), rcmin = cta_min(lmin, s_red);
            const double amax = rcmax;
            bool ok = rcmin != 0.0;
            double rowcnd = 0.0, colcnd = 0.0;
            if (ok) {
                for (int i = threadIdx.x; i < n; i += blockDim.x) rs[i] = 1.0 / fmin(fmax(rs[i], smlnum), bignum);
                rowcnd = fmax(rcmin, smlnum) / fmin(rcmax, bignum);
                __syncthreads();
                for (int j = threadIdx.x; j < n; j += blockDim.x) {
                    double m = 0.0;
                    for (int i = max(0, j - ku); i <= min(n - 1, j + kl); ++i)
                        m = fmax(m, cabs1(ab[(size_t) j * ldab + ku + i - j]) * rs[i]);
                    cs[j] = m;
                }
                __syncthreads();
                lmax = 0.0; lmin = bignum;
                for (int j = threadIdx.x; j < n; j += blockDim.x) { lmax = fmax(lmax, cs[j]); lmin = fmin(lmin, cs[j]); }
                const double ccmax = cta_max(lmax, s_red), ccmin = cta_min(lmin, s_red);
                ok = ccmin != 0.0;
                if (ok) {
                    for (int j = threadIdx.x; j < n; j += blockDim.x) cs[j] = 1.0 / fmin(fmax(cs[j], smlnum), bignum);
                    colcnd = fmax(ccmin, smlnum) / fmin(ccmax, bignum);
                }
            }
            __syncthreads();
            if (ok) {
                // zlaqgb
                const double small = safmin / DBL_EPSILON, large = 1.0 / small, thresh = 0.1;
                const bool rowsc = !(rowcnd >= thresh && amax >= small && amax <= large);
                const bool colsc = !(colcnd >= thresh);
                equed = (rowsc ? 1 : 0) | (colsc ? 2 : 0);
                if (equed)
                    for (int e = threadIdx.x; e < n * ldab; e += blockDim.x) {
                        const int j = e / ldab, rr = e - j * ldab, i = j - ku + rr;
                        if (i >= 0 && i < n) {
                            const double f = (rowsc ? rs[i] : 1.0) * (colsc ? cs[j] : 1.0);
                            ab[e] = ab[e] * (colsc && rowsc ? cs[j] * rs[i] : f);
                        }
                    }
                __syncthreads();
            }
        }
        for (int e = threadIdx.x; e < n * ldab; e += blockDim.x) {    // zlacpy
            const int j = e / ldab, rr = e - j * ldab;
            afb[(size_t) j * ldafb + kl + rr] = ab[e];
        }
        __syncthreads();
        info = gbtrf_cta(n, kl, ku, afb, ldafb, ipiv, S);
        factored = true;
        if (info > 0) { *count_out = 0; *berr_out = 0.0; return info; }
    }
    // TRANS = 'T': the right hand side takes the column factors, the solution the row factors
    if (equed & 2) for (int i = threadIdx.x; i < n; i += blockDim.x) b[i] = b[i] * cs[i];
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) x[i] = b[i];
    __syncthreads();
    if (threadIdx.x < 32) gbtrs_T_warp(n, kl, ku, afb, ldafb, ipiv, x);
    __syncthreads();
    // zgbrfs
    const int nz = min(kl + ku + 2, n + 1);
    const double safe1 = nz * safmin, safe2 = safe1 / eps;
    int count = 1;
    double lstres = 3.0, berr = 0.0;
    for (;;) {
        gb_residual_cta('T', n, kl, ku, ab, ldab, x, b, r);
        for (int k = threadIdx.x; k < n; k += blockDim.x) {
            double sum = 0.0;
            const cplx *col = ab + (size_t) k * ldab + (ku - k);
            for (int i = max(0, k - ku); i <= min(n - 1, k + kl); ++i) sum += cabs1(col[i]) * cabs1(x[i]);
            rw[k] = cabs1(b[k]) + sum;
        }
        __syncthreads();
        double sm = 0.0;
        for (int i = threadIdx.x; i < n; i += blockDim.x)
            sm = fmax(sm, rw[i] > safe2 ? cabs1(r[i]) / rw[i] : (cabs1(r[i]) + safe1) / (rw[i] + safe1));
        berr = cta_max(sm, s_red);
        if (berr > eps && 2.0 * berr <= lstres && count <= 5) {
            if (threadIdx.x < 32) gbtrs_T_warp(n, kl, ku, afb, ldafb, ipiv, r);
            __syncthreads();
            for (int i = threadIdx.x; i < n; i += blockDim.x) x[i] += r[i];
            __syncthreads();
            lstres = berr;
            ++count;
        } else break;
    }
    if (equed & 1) for (int i = threadIdx.x; i < n; i += blockDim.x) x[i] = x[i] * rs[i];
    __syncthreads();
    *count_out = count - 1; *berr_out = berr;
    return info;
}

__global__ void __launch_bounds__(256)
zcgbsvx_batch_kernel(char trans, int n, int kl, int ku, int aiter, int dmax, double tolsc,
                     const cplx *ab, size_t stride_ab, cplx *afb, size_t stride_afb,
                     int *ipiv, const cplx *b, cplx *x, cplx *rwork, int *iters,
                     double *resv, int *info)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double s_red[32];
    const LuScratch S = lu_scratch(smem_raw, kl, ku);
    const size_t s = blockIdx.x;
    bool factored = false;
    int diter; double res;
    const int rc = zcgbsvx_cta(trans, n, kl, ku, aiter, dmax, tolsc, ab + s * stride_ab,
                               afb + s * stride_afb, ipiv + s * n, b + s * n, x + s * n,
                               rwork + s * n, factored, &diter, &res, S, s_red);
    if (threadIdx.x == 0) {
        info[s] = rc;
        if (iters) iters[s] = diter;
        if (resv) resv[s] = res;
    }
}

// ---------------------------------------------------------------------------
// y <- alpha P x + beta y / alpha P^T x + beta y
// ---------------------------------------------------------------------------
__global__ void zaPxpby_kernel(int transT, int S, int n, cplx alpha, const cplx *x, cplx beta,
                               cplx *y, size_t total)
{
    const size_t N = (size_t) S * n;
    for (size_t e = blockIdx.x * (size_t) blockDim.x + threadIdx.x; e < total;
         e += (size_t) gridDim.x * blockDim.x) {
        const size_t b = e / N; const int k = (int) (e - b * N);
        // 'N': y[k] = x[q(k)], q(k) = (k % S) n + k / S;  'T': y[q(k)] = x[k]
        const int qk = (k % S) * n + k / S;
        const size_t src = transT ? e : b * N + qk;
        const size_t dst = transT ? b * N + qk : e;
        const cplx ax = alpha * x[src];
        y[dst] = is_zero(beta) ? ax : ax + beta * y[dst];
    }
}

// ---------------------------------------------------------------------------
// Batched invert, version 1: persistent CTAs, each owning a slot of global
// scratch holding the LAPACK-layout factor storage (and, for zcgbsvx, the
// unfactored operator plus work vectors).  Per pencil:
//   assemble (+NRBC, +wall BCs) -> b = P state, wall rows zeroed -> factor ->
//   solve 'T' (optionally refined) -> state = P^T x; extra right hand sides
//   reuse the factorisation.
// ---------------------------------------------------------------------------
struct InvertArgs {
    PackArgs pk;
    int method, aiter, diter, equil; double tolsc;
    int npencil; const int *index;
    cplx *state; size_t fs, ps;
    int nextra; cplx *extra;
    int *ipiv_out, *info_out, *iters_out;
    unsigned char *work; size_t slot_bytes;
};

__device__ __forceinline__ size_t align16(size_t v) { return (v + 15) & ~(size_t) 15; }

__global__ void __launch_bounds__(256)
invert_kernel(const InvertArgs A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double s_red[32];
    __shared__ cplx s_x75[75];
    const PackArgs &K = A.pk;
    const int N = K.N, KL = K.KL, KU = K.KU, LD = K.LD, n = K.n;
    const int ldlu = LD + KL;
    cplx *s_alpha = reinterpret_cast<cplx *>(smem_raw);
    const LuScratch S = lu_scratch(smem_raw + sizeof(cplx) * MAXTERMS, KL, KU);

    // slot layout: LU | PAPT (zcgbsvx only) | b | x | r | ipiv
    unsigned char *slot = A.work + (size_t) blockIdx.x * A.slot_bytes;
    cplx *LU = reinterpret_cast<cplx *>(slot);
    size_t off = align16(sizeof(cplx) * (size_t) ldlu * N);
    cplx *PAPT = reinterpret_cast<cplx *>(slot + off);
    if (A.method != SZB_SOLVER_ZGBSV) off += align16(sizeof(cplx) * (size_t) LD * N);
    cplx *vb = reinterpret_cast<cplx *>(slot + off); off += sizeof(cplx) * (size_t) N;
    cplx *vx = reinterpret_cast<cplx *>(slot + off); off += sizeof(cplx) * (size_t) N;
    cplx *vr = reinterpret_cast<cplx *>(slot + off); off += sizeof(cplx) * (size_t) N;
    double *rw = reinterpret_cast<double *>(slot + off); off += 3 * sizeof(double) * (size_t) N;   // zgbsvx: rwork, r, c
    int *ipiv = reinterpret_cast<int *>(slot + off);

    for (int p = blockIdx.x; p < A.npencil; p += gridDim.x) {
        const double km = K.km[p], kn = K.kn[p];
        if (A.method == SZB_SOLVER_ZGBSV) pack_pencil(K, ldlu, km, kn, s_alpha, s_x75, LU + KL);
        else                              pack_pencil(K, LD,   km, kn, s_alpha, s_x75, PAPT);

        int info = 0, diter = 0, equed = 0;
        bool factored = false;
        for (int rhs = 0; rhs <= A.nextra && info == 0; ++rhs) {
            cplx *v = rhs == 0 ? A.state + (A.index ? (size_t) A.index[p] : (size_t) p) * A.ps
                               : A.extra + ((size_t) p * A.nextra + (rhs - 1)) * N;
            const size_t fs = rhs == 0 ? A.fs : (size_t) n;
            // b = P v with the wall rows zeroed (bsmbsm_solver.hpp:150-156,
            // operator_hybrid_isothermal.cpp:516-525)
            for (int k = threadIdx.x; k < N; k += blockDim.x) {
                const int y = k / 5, s = k - 5 * y;
                cplx val = v[(size_t) s * fs + y];
                if (K.with_bc && s < 4) {
                    if ((y == 0 && K.wall_begin == 0) || (y == n - 1 && K.wall_end == 2))
                        val = cplx(0.0, 0.0);
                }
                vb[k] = val;
            }
            __syncthreads();
            const cplx *sol;
            if (A.method == SZB_SOLVER_ZGBSV) {
                if (!factored) { info = gbtrf_cta(N, KL, KU, LU, ldlu, ipiv, S); factored = true; }
                if (info == 0 && threadIdx.x < 32) gbtrs_T_warp(N, KL, KU, LU, ldlu, ipiv, vb);
                sol = vb;
            } else if (A.method == SZB_SOLVER_ZGBSVX) {
                double berr; int it;
                info = zgbsvx_cta(A.equil != 0, N, KL, KU, PAPT, LU, ipiv, vb, vx, vr, rw, rw + N, rw + 2 * N,
                                  factored, equed, &it, &berr, S, s_red);
                if (rhs == 0) diter = it;
                sol = vx;
            } else {
                double res; int it;
                info = zcgbsvx_cta('T', N, KL, KU, A.aiter, A.diter, A.tolsc, PAPT, LU, ipiv,
                                   vb, vx, vr, factored, &it, &res, S, s_red);
                if (rhs == 0) diter = it;
                sol = vx;
            }
            __syncthreads();
            if (info == 0) {
                // v = P^T x (bsmbsm_solver.hpp:274-280)
                for (int k = threadIdx.x; k < N; k += blockDim.x) {
                    const int y = k / 5, s = k - 5 * y;
                    v[(size_t) s * fs + y] = sol[k];
                }
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            A.info_out[p] = info;
            if (A.iters_out) A.iters_out[p] = diter;
        }
        if (A.ipiv_out)
            for (int k = threadIdx.x; k < N; k += blockDim.x) A.ipiv_out[(size_t) p * N + k] = ipiv[k];
        __syncthreads();
    }
}

}  // namespace szb

using namespace szb;

extern "C" {

int szb_zgbtrf_batch(int n, int kl, int ku, szb_complex *d_ab, int ldab, size_t stride,
                     int *d_ipiv, int *d_info, int nbatch, void *stream)
{
    if (n < 0) return -1;
    if (kl < 0) return -2;
    if (ku < 0) return -3;
    if (!d_ab) return -4;
    if (ldab < 2 * kl + ku + 1) return -5;
    if (!d_ipiv) return -7;
    if (!d_info) return -8;
    if (nbatch < 0) return -9;
    if (nbatch == 0 || n == 0) return 0;
    zgbtrf_batch_kernel<<<nbatch, 256, lu_scratch_bytes(kl, ku), (cudaStream_t) stream>>>(
        n, kl, ku, reinterpret_cast<cplx *>(d_ab), ldab, stride, d_ipiv, d_info);
    count_launch();
    SZB_CUDA_OK(cudaGetLastError());
    return 0;
}

int szb_zgbtrs_batch(char trans, int n, int kl, int ku, int nrhs, const szb_complex *d_ab,
                     int ldab, size_t stride, const int *d_ipiv, szb_complex *d_b, int ldb,
                     size_t strideb, int nbatch, void *stream)
{
    if (trans == 't') trans = 'T';
    if (trans == 'n') trans = 'N';
    if (trans != 'N' && trans != 'T') return -1;
    if (n < 0) return -2;
    if (kl < 0) return -3;
    if (ku < 0) return -4;
    if (nrhs < 0) return -5;
    if (!d_ab) return -6;
    if (ldab < 2 * kl + ku + 1) return -7;
    if (!d_ipiv) return -9;
    if (!d_b) return -10;
    if (ldb < n) return -11;
    if (nbatch < 0) return -13;
    if (nbatch == 0 || n == 0 || nrhs == 0) return 0;
    const size_t warps = (size_t) nbatch * nrhs;
    const int wpb = 4;
    zgbtrs_batch_kernel<<<(unsigned) ((warps + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t) stream>>>(
        trans, n, kl, ku, nrhs, reinterpret_cast<const cplx *>(d_ab), ldab, stride, d_ipiv,
        reinterpret_cast<cplx *>(d_b), ldb, strideb, nbatch);
    count_launch();
    SZB_CUDA_OK(cudaGetLastError());
    return 0;
}

int szb_zcgbsvx_batch(char trans, int n, int kl, int ku, int aiter, int diter, double tolsc,
                      const szb_complex *d_ab, size_t stride_ab, szb_complex *d_afb,
                      size_t stride_afb, int *d_ipiv, const szb_complex *d_b,
                      szb_complex *d_x, int *d_iters, double *d_res, int *d_info,
                      int nbatch, void *stream)
{
    if (trans == 't') trans = 'T';
    if (trans == 'n') trans = 'N';
    if (trans != 'N' && trans != 'T') return -1;
    if (n < 0) return -2;
    if (kl < 0) return -3;
    if (ku < 0) return -4;
    if (aiter < 0) return -5;
    if (diter < 0) return -6;
    if (tolsc < 0) return -7;
    if (!d_ab) return -8;
    if (!d_afb) return -10;
    if (!d_ipiv) return -12;
    if (!d_b) return -13;
    if (!d_x) return -14;
    if (!d_info) return -17;
    if (nbatch < 0) return -18;
    if (nbatch == 0 || n == 0) return 0;
    cplx *rwork = nullptr;
    SZB_CUDA_OK(cudaMallocAsync(&rwork, sizeof(cplx) * (size_t) n * nbatch, (cudaStream_t) stream));
    zcgbsvx_batch_kernel<<<nbatch, 256, lu_scratch_bytes(kl, ku), (cudaStream_t) stream>>>(
        trans, n, kl, ku, aiter, diter, tolsc, reinterpret_cast<const cplx *>(d_ab), stride_ab,
        reinterpret_cast<cplx *>(d_afb), stride_afb, d_ipiv, reinterpret_cast<const cplx *>(d_b),
        reinterpret_cast<cplx *>(d_x), rwork, d_iters, d_res, d_info);
    count_launch();
    SZB_CUDA_OK(cudaGetLastError());
    SZB_CUDA_OK(cudaFreeAsync(rwork, (cudaStream_t) stream));
    return 0;
}

int szb_bsmbsm_zaPxpby_batch(char trans, int S, int n, const double alpha[2],
                             const szb_complex *d_x, const double beta[2], szb_complex *d_y,
                             int nbatch, void *stream)
{
    if (trans == 't') trans = 'T';
    if (trans == 'n') trans = 'N';
    if (trans != 'N' && trans != 'T') return -1;
    if (S < 0) return -2;
    if (n < 0) return -3;
    if (!alpha) return -4;
    if (!d_x) return -5;
    if (!beta) return -6;
    if (!d_y) return -7;
    if ((const void *) d_x == (const void *) d_y) return -7;     // aPxpby_complex.def:50
    if (nbatch < 0) return -8;
    const size_t total = (size_t) S * n * nbatch;
    if (total == 0) return 0;
    const unsigned blocks = (unsigned) ((total + 255) / 256 > 65535 ? 65535 : (total + 255) / 256);
    zaPxpby_kernel<<<blocks, 256, 0, (cudaStream_t) stream>>>(trans == 'T', S, n,
        cplx(alpha[0], alpha[1]), reinterpret_cast<const cplx *>(d_x), cplx(beta[0], beta[1]),
        reinterpret_cast<cplx *>(d_y), total);
    count_launch();
    SZB_CUDA_OK(cudaGetLastError());
    return 0;
}

__global__ void zero_pencils_kernel(int npencil, const int *index, int S, int n, cplx *state,
                                    size_t fs, size_t ps)
{
    const int p = blockIdx.x;
    cplx *v = state + (index ? (size_t) index[p] : (size_t) p) * ps;
    for (int e = threadIdx.x; e < S * n; e += blockDim.x) {
        const int f = e / n, y = e - f * n;
        v[(size_t) f * fs + y] = cplx(0.0, 0.0);
    }
}

int szb_zero_pencils(int npencil, const int *d_index, int S, int n, szb_complex *d_state,
                     size_t field_stride, size_t pencil_stride, void *stream)
{
    if (npencil < 0) return -1;
    if (S < 0) return -3;
    if (n < 0) return -4;
    if (!d_state) return -5;
    if (npencil == 0) return 0;
    zero_pencils_kernel<<<npencil, 128, 0, (cudaStream_t) stream>>>(npencil, d_index, S, n,
        reinterpret_cast<cplx *>(d_state), field_stride, pencil_stride);
    count_launch();
    SZB_CUDA_OK(cudaGetLastError());
    return 0;
}

// a <-> b between two state layouts (suzerain/state.hpp:486-520,607-630: exchange of an
// interleaved_state with a contiguous_state as lowstorage::step does after every accumulate,
// suzerain/lowstorage.hpp:1511): one CTA per pencil, whole wall-normal lines of both states.
__global__ void state_exchange_kernel(int npencil, const int *index, int S, int n, cplx *a, size_t afs, size_t aps,
                                      cplx *b, size_t bfs, size_t bps)
{
    for (int q = blockIdx.x; q < npencil; q += gridDim.x) {
        const size_t p = index ? (size_t) index[q] : (size_t) q;
        cplx *va = a + p * aps, *vb = b + p * bps;
        for (int e = threadIdx.x; e < S * n; e += blockDim.x) {
            const int f = e / n, y = e - f * n;
            const cplx x = va[(size_t) f * afs + y], z = vb[(size_t) f * bfs + y];
            va[(size_t) f * afs + y] = z; vb[(size_t) f * bfs + y] = x;
        }
    }
}

int szb_state_exchange(int npencil, const int *d_index, int S, int n,
                       szb_complex *d_a, size_t a_field_stride, size_t a_pencil_stride,
                       szb_complex *d_b, size_t b_field_stride, size_t b_pencil_stride, void *stream)
{
    if (npencil < 0) return -1;
    if (S < 0) return -3;
    if (n < 0) return -4;
    if (!d_a) return -5;
    if (!d_b) return -8;
    if (d_a == d_b) return -8;
    if (npencil == 0) return 0;
    const int grid = npencil < 148 * 16 ? npencil : 148 * 16;
    state_exchange_kernel<<<grid, 128, 0, (cudaStream_t) stream>>>(npencil, d_index, S, n,
        reinterpret_cast<cplx *>(d_a), a_field_stride, a_pencil_stride,
        reinterpret_cast<cplx *>(d_b), b_field_stride, b_pencil_stride);
    count_launch();
    SZB_CUDA_OK(cudaGetLastError());
    return 0;
}

static size_t invert_slot_bytes(const szb_imexop *op, int method)
{
    const size_t N = op->A.N, ldlu = op->A.LD + op->A.KL;
    size_t b = (sizeof(cplx) * ldlu * N + 15) & ~(size_t) 15;
    if (method != SZB_SOLVER_ZGBSV) b += (sizeof(cplx) * (size_t) op->A.LD * N + 15) & ~(size_t) 15;
    b += 3 * sizeof(cplx) * N + 3 * sizeof(double) * N + sizeof(int) * N;
    return (b + 255) & ~(size_t) 255;
}

}  // extern "C"

namespace szb {
int invert_fused_dispatch(const szb_imexop *op, const double phi[2], int npencil,
                          const double *d_km, const double *d_kn, const int *d_index,
                          cplx *d_state, size_t fs, size_t ps, int *d_ipiv, int *d_info,
                          int *d_iters, cudaStream_t stream, int zero_wall_rhs, const int *d_count)
{
    static const int which = [] {
        const char *e = std::getenv("SZB_INVERT");
        return (e && e[0] == 'v' && e[1] == '4') ? 4 : 5;
    }();
    int rc = 1;
    if (which == 5)
        rc = invert_sync_dispatch(op, phi, npencil, d_km, d_kn, d_index, d_state, fs, ps, d_ipiv, d_info, d_iters,
                                  stream, zero_wall_rhs, d_count);
    if (rc == 1)
        rc = invert_pipe_dispatch(op, phi, npencil, d_km, d_kn, d_index, d_state, fs, ps, d_ipiv, d_info, d_iters,
                                  stream, zero_wall_rhs, d_count);
    return rc;
}
}  // namespace szb

extern "C" {

size_t szb_imexop_workspace_bytes(const szb_imexop *op) { return op ? op->work_bytes : 0; }

int szb_imexop_invert_batch(const szb_imexop *op, const szb_zgbsv_spec *spec,
        const double phi[2], int npencil, const double *d_km, const double *d_kn,
        const int *d_index,
        szb_complex *d_state, size_t field_stride, size_t pencil_stride,
        int nextra, szb_complex *d_extra, int *d_ipiv, int *d_info, int *d_iters,
        void *stream)
{
    if (!op) return -1;
    if (!spec || spec->method < SZB_SOLVER_ZGBSV || spec->method > SZB_SOLVER_ZGBSVX) return -2;
    if (!phi) return -3;
    if (npencil < 0) return -4;
    if (!d_km) return -5;
    if (!d_kn) return -6;
    if (!d_state) return -8;
    if (nextra < 0) return -11;
    if (nextra > 0 && !d_extra) return -12;
    if (!d_info) return -14;
    if (npencil == 0) return 0;

    // linearize::rhome_y (operator_hybrid_isothermal.cpp:691-761): the operator does not depend on
    // the wavenumbers: one factorisation, one pair of triangular sweeps per right hand side, and
    // refinement around it for zcgbsvx / zgbsvx (rhome_y.cu); anything else (equilibration, scaled
    // tolerances) runs the general kernels at km = kn = 0.
    if (op->linearization == SZB_LINEARIZE_RHOME_Y) {
        const int mode = spec->method == SZB_SOLVER_ZGBSV ? -1
                       : (spec->method == SZB_SOLVER_ZCGBSVX && spec->tolsc == 0.0) ? 0
                       : (spec->method == SZB_SOLVER_ZGBSVX && !spec->equil) ? 1 : -2;
        if (mode >= -1) {
            const int rc = invert00_dispatch(op, mode, spec->aiter, spec->diter, phi, npencil, d_index,
                                             reinterpret_cast<cplx *>(d_state), field_stride, pencil_stride, nextra,
                                             reinterpret_cast<cplx *>(d_extra), d_ipiv, d_info, d_iters, (cudaStream_t) stream);
            if (rc <= 0) return rc;
        }
        if ((size_t) npencil > op->zero_count) {
            if (op->d_zero) SZB_CUDA_OK(cudaFree(op->d_zero));
            op->d_zero = nullptr; op->zero_count = 0;
            SZB_CUDA_OK(cudaMalloc(&op->d_zero, sizeof(double) * (size_t) npencil));
            SZB_CUDA_OK(cudaMemsetAsync(op->d_zero, 0, sizeof(double) * (size_t) npencil, (cudaStream_t) stream));
            op->zero_count = (size_t) npencil;
        }
        d_km = d_kn = op->d_zero;
    }

    // zgbsv with a single right hand side per pencil: the fused shared-memory-window kernels
    // (invert_sync.cu: v5, default; invert_pipe.cu: v4 with SZB_INVERT=v4 or when v5 has no
    // instantiation).  SZB_INVERT=v1 selects the generic global-memory kernel below.
    if (spec->method == SZB_SOLVER_ZGBSV && nextra == 0) {
        static const bool generic = [] { const char *e = std::getenv("SZB_INVERT"); return e && e[0] == 'v' && e[1] == '1'; }();
        if (!generic) {
            const int rc = invert_fused_dispatch(op, phi, npencil, d_km, d_kn, d_index,
                                                 reinterpret_cast<cplx *>(d_state), field_stride,
                                                 pencil_stride, d_ipiv, d_info, d_iters, (cudaStream_t) stream);
            if (rc <= 0) return rc;
        }
    }
    // zcgbsvx with its default eps tolerance: refinement around the fused kernel (the factors
    // are recomputed per step instead of being stored); SZB_INVERT=v1 keeps the generic kernel
    // Likewise zgbsvx without equilibration: zgbtrs + zgbrfs around the fused kernel.
    const bool refined_zc = spec->method == SZB_SOLVER_ZCGBSVX && spec->tolsc == 0.0 && spec->aiter >= 1;
    const bool refined_zx = spec->method == SZB_SOLVER_ZGBSVX && !spec->equil;
    if (nextra == 0 && (refined_zc || refined_zx)) {
        static const bool generic = [] { const char *e = std::getenv("SZB_INVERT"); return e && e[0] == 'v' && e[1] == '1'; }();
        if (!generic) {
            const int rc = invert_refined_dispatch(op, refined_zx ? 1 : 0, spec->aiter, spec->diter, phi, npencil, d_km, d_kn,
                                                   d_index, reinterpret_cast<cplx *>(d_state), field_stride, pencil_stride,
                                                   d_ipiv, d_info, d_iters, (cudaStream_t) stream);
            if (rc <= 0) return rc;
        }
    }

    InvertArgs A;
    fill_pack_args(op, phi, d_km, d_kn, 0, 1, nullptr, A.pk);
    A.method = spec->method; A.aiter = spec->aiter; A.diter = spec->diter; A.tolsc = spec->tolsc;
    A.equil = spec->equil;
    A.npencil = npencil; A.index = d_index;
    A.state = reinterpret_cast<cplx *>(d_state); A.fs = field_stride; A.ps = pencil_stride;
    A.nextra = nextra; A.extra = reinterpret_cast<cplx *>(d_extra);
    A.ipiv_out = d_ipiv; A.info_out = d_info; A.iters_out = d_iters;

    const size_t smem = sizeof(cplx) * MAXTERMS + lu_scratch_bytes(op->A.KL, op->A.KU);
    int per_sm = 0;
    SZB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, invert_kernel, 256, smem));
    if (per_sm < 1) per_sm = 1;
    int slots = op->sm_count * per_sm;
    if (slots > npencil) slots = npencil;
    A.slot_bytes = invert_slot_bytes(op, spec->method);
    const size_t need = A.slot_bytes * (size_t) slots;
    if (need > op->work_bytes) {
        if (op->d_work) SZB_CUDA_OK(cudaFree(op->d_work));
        op->d_work = nullptr; op->work_bytes = 0;
        SZB_CUDA_OK(cudaMalloc(&op->d_work, need));
        op->work_bytes = need;
    }
    op->work_slots = slots;
    A.work = static_cast<unsigned char *>(op->d_work);
    invert_kernel<<<slots, 256, smem, (cudaStream_t) stream>>>(A);
    count_launch();
    SZB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // extern "C"
