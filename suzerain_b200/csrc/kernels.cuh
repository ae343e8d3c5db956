// kernels.cuh -- CTA-wide device building blocks shared by imexop.cu (pack) and
// gbsv.cu (banded LU / solve / refinement / fused invert).
#pragma once

#include <cstring>

#include "szb_internal.hpp"
#include "cplx.cuh"

namespace szb {

// ---------------------------------------------------------------------------
// wave(km, kn) factors of rholut_terms.def
// ---------------------------------------------------------------------------
__device__ __forceinline__ cplx wave_factor(int w, double km, double kn)
{
    switch (w) {
    case wavid::ONE:  return cplx(1.0, 0.0);
    case wavid::IKM:  return cplx(0.0, km);
    case wavid::IKN:  return cplx(0.0, kn);
    case wavid::KM2:  return cplx(km * km, 0.0);
    case wavid::KN2:  return cplx(kn * kn, 0.0);
    case wavid::KMKN: return cplx(km * kn, 0.0);
    default:          return cplx(km * km + kn * kn, 0.0);
    }
}

// ---------------------------------------------------------------------------
// Assembly of P (M + phi L)^T P^T (+ NRBC corner, + isothermal wall columns)
// ---------------------------------------------------------------------------
struct PackArgs {
    const double *D;        // [3][ld][n]  r-major: D[(d*ld + r)*n + y] = D^(d)[y, y-ku+r]
    const double *refs;     // [27][n]
    const TermTable *terms;
    int n, kl, ku, ld;      // block bandwidths (of the D^T storage)
    int N, KL, KU, LD;
    cplx phi;
    const double *km, *kn;
    cplx *out;              // npencil matrices (pack kernels only)
    int rows;               // column stride: LD (packc) or LD + KL (packf)
    int rowoff;             // 0 or KL
    int with_bc;
    int wall_begin, wall_end;
    double E_factor[2], vel_factor[2][3];
    int nrbc;               // bit0 a, bit1 b, bit2 c
    double a[25], b[25], c[25];
};

inline void fill_pack_args(const szb_imexop *op, const double phi[2],
                           const double *d_km, const double *d_kn,
                           int packf, int with_bc, cplx *out, PackArgs &A)
{
    A.D = op->d_D; A.refs = op->d_refs; A.terms = op->d_terms;
    A.n = op->n; A.kl = op->kl; A.ku = op->ku; A.ld = op->ld;
    A.N = op->A.N; A.KL = op->A.KL; A.KU = op->A.KU; A.LD = op->A.LD;
    A.phi = cplx(phi[0], phi[1]);
    A.km = d_km; A.kn = d_kn; A.out = out;
    A.rows = packf ? A.LD + A.KL : A.LD;
    A.rowoff = packf ? A.KL : 0;
    A.with_bc = with_bc;
    A.wall_begin = op->iso.enforce_lower ? 0 : 1;
    A.wall_end   = op->iso.enforce_upper ? 2 : 1;
    for (int i = 0; i < 2; ++i) {
        A.E_factor[i] = op->E_factor[i];
        for (int j = 0; j < 3; ++j) A.vel_factor[i][j] = op->vel_factor[i][j];
    }
    A.nrbc = (op->have_a ? 1 : 0) | (op->have_b ? 2 : 0) | (op->have_c ? 4 : 0);
    std::memcpy(A.a, op->nrbc_a, sizeof(A.a));
    std::memcpy(A.b, op->nrbc_b, sizeof(A.b));
    std::memcpy(A.c, op->nrbc_c, sizeof(A.c));
}

// alpha_t without phi, one per term, into shared memory (CTA-wide; caller syncs)
__device__ __forceinline__ void term_alphas(const TermTable *tt, double km, double kn,
                                            cplx *s_alpha)
{
    const int nterms = tt->nterms;
    for (int t = threadIdx.x; t < nterms; t += blockDim.x)
        s_alpha[t] = wave_factor(tt->wave[t], km, kn) * tt->sc[t];
}

// Entry (I, J) of the renumbered transpose, I = 5*yI + sI, J = 5*yJ + sJ:
// entry (yJ, yI) of operator block (row = sJ, col = sI), i.e.
//   phi * sum_d c_{sJ,sI,d}[yJ] D^(d)[yJ, yI]   (+ M[yJ, yI] on diagonal blocks)
// accumulated in the reference's order: ops M, D1, D2 into a buffer, then the
// phi scaling, then the mass matrix (rholut_imexop.def:113-137).
__device__ __forceinline__ cplx operator_entry(const PackArgs &A, const cplx *s_alpha,
                                               int I, int J)
{
    const int yI = I / 5, sI = I - 5 * yI;
    const int yJ = J / 5, sJ = J - 5 * yJ;
    const int off = yI - yJ;
    if (off < -A.ku || off > A.kl) return cplx(0.0, 0.0);
    const int r = A.ku + off;
    cplx buf(0.0, 0.0);
    const int blk0 = (sJ * 5 + sI) * 3;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const int tb = A.terms->blk_begin[blk0 + d], te = A.terms->blk_begin[blk0 + d + 1];
        if (tb == te) continue;
        cplx c(0.0, 0.0);
        for (int t = tb; t < te; ++t)
            c += s_alpha[t] * A.refs[(size_t) A.terms->ref[t] * A.n + yJ];
        buf += c * A.D[(size_t) (d * A.ld + r) * A.n + yJ];
    }
    buf = A.phi * buf;
    if (sI == sJ) buf += cplx(A.D[(size_t) (0 * A.ld + r) * A.n + yJ], 0.0);
    return buf;
}

// CTA-wide: assemble one pencil's matrix into M (M points at band row 0 of
// column 0 of the *matrix*, i.e. already offset by rowoff; column stride
// `rows`).  s_alpha: >= nterms cplx of shared memory; s_x: 75 cplx.
__device__ inline void pack_pencil(const PackArgs &A, const int rows, double km, double kn,
                                   cplx *s_alpha, cplx *s_x, cplx *M)
{
    term_alphas(A.terms, km, kn, s_alpha);
    __syncthreads();

    const int total = A.N * A.LD;
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
        const int J = e / A.LD, r = e - J * A.LD;
        const int I = J - A.KU + r;
        if (I < 0 || I >= A.N) continue;
        M[(size_t) J * rows + r] = operator_entry(A, s_alpha, I, J);
    }

    if (A.nrbc) {
        // Lower-right 15 x 5 corner (rholut_imexop.def:505-595):
        //   X <- X - X C^T + [0; 0; C^T - i km phi A^T - i kn phi B^T]
        __syncthreads();
        const int I0 = 5 * (A.n - 3), J0 = 5 * (A.n - 1);
        const int e = threadIdx.x;
        const int i = e % 15, j = e / 15;
        if (e < 75) s_x[e] = M[(size_t) (J0 + j) * rows + (A.KU + (I0 + i) - (J0 + j))];
        __syncthreads();
        if (e < 75) {
            cplx buf(0.0, 0.0);
            if (i >= 10) {
                const cplx ikmphi = cplx(0.0, km) * A.phi, iknphi = cplx(0.0, kn) * A.phi;
                if (A.nrbc & 1) buf -= ikmphi * A.a[5 * (i - 10) + j];
                if (A.nrbc & 2) buf -= iknphi * A.b[5 * (i - 10) + j];
                if (A.nrbc & 4) buf += cplx(A.c[5 * (i - 10) + j], 0.0);
            }
            if (A.nrbc & 4)
                for (int k = 0; k < 5; ++k) buf -= s_x[i + 15 * k] * A.c[j + 5 * k];
            M[(size_t) (J0 + j) * rows + (A.KU + (I0 + i) - (J0 + j))] = s_x[e] + buf;
        }
    }

    if (A.with_bc) {
        // Isothermal wall equations (operator_hybrid_isothermal.cpp:470-510):
        // overwrite column qinv(eq*n + wall) with {+s at self, -s*factor at the
        // wall's rho index, 0 elsewhere}, s = previous diagonal (1 if zero).
        for (int wall = A.wall_begin; wall < A.wall_end; ++wall) {
            const int y = wall == 0 ? 0 : A.n - 1;
            const int irho = 5 * y + 4;
            for (int eq = 0; eq < 4; ++eq) {
                const int J = 5 * y + eq;
                const double factor = eq == 0 ? A.E_factor[wall] : A.vel_factor[wall][eq - 1];
                cplx *col = M + (size_t) J * rows + (A.KU - J);   // col[I]
                __syncthreads();                  // assembly / previous column done
                cplx s = col[J];
                if (is_zero(s)) s = cplx(1.0, 0.0);
                __syncthreads();                  // everyone has read the diagonal
                const int begin = max(0, J - A.KU), end = min(A.N, J + A.KL + 1);
                for (int I = begin + threadIdx.x; I < end; I += blockDim.x)
                    col[I] = I == J ? s : I == irho ? -(s * factor) : cplx(0.0, 0.0);
            }
        }
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------
// Banded LU with partial pivoting, CTA-wide, in place on LAPACK band storage
// (zgbtf2 semantics: ab has 2*kl+ku+1 rows, the matrix occupies rows kl.. ;
// pivot = first maximum of |re|+|im| over the kl+1 candidates; row
// interchanges applied to columns j..ju only; multipliers stored unswapped;
// ipiv 1-based).  Returns info (0, or 1-based index of the first zero pivot).
//
// Shared scratch: s_col[kl+1], s_l[kl], s_u[kl+ku+1] cplx and two ints.
// ---------------------------------------------------------------------------
struct LuScratch {
    cplx *col;   // kl + 1
    cplx *l;     // kl
    cplx *u;     // kl + ku + 1
    int  *ibuf;  // [0] = jp, [1] = info
};

__device__ inline int gbtrf_cta(int n, int kl, int ku, cplx *ab, int ldab, int *ipiv,
                                const LuScratch S)
{
    const int kv = kl + ku;
    const int tid = threadIdx.x, nt = blockDim.x;

    // Zero the fill-in super-diagonals (zgbtf2: columns ku+1 .. min(kv, n)-1,
    // rows kv-j+1 .. kl; then row kl... of each newly reached column).
    for (int j = ku + 1; j < min(kv, n); ++j)
        for (int i = kv - j + tid; i < kl; i += nt) ab[(size_t) j * ldab + i] = cplx(0.0, 0.0);
    if (tid == 0) S.ibuf[1] = 0;
    __syncthreads();

    int ju = 0;
    for (int j = 0; j < n; ++j) {
        // zero fill-in elements in column j + kv
        if (j + kv < n)
            for (int i = tid; i < kl; i += nt) ab[(size_t) (j + kv) * ldab + i] = cplx(0.0, 0.0);

        const int km = min(kl, n - 1 - j);
        cplx *colj = ab + (size_t) j * ldab + kv;          // colj[i] = A(j+i, j)

        // ---- phase A: stage the candidate column; warp 0 finds the pivot ----
        if (tid < 32) {
            double best = -1.0; int bi = 0;
            for (int i = tid; i <= km; i += 32) {
                const cplx v = colj[i];
                S.col[i] = v;
                const double m = cabs1(v);
                if (m > best) { best = m; bi = i; }     // strict >: first maximum per lane
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int    oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            if (tid == 0) {
                S.ibuf[0] = bi;
                ipiv[j] = j + bi + 1;
                if (best == 0.0 && S.ibuf[1] == 0) S.ibuf[1] = j + 1;
            }
        }
        __syncthreads();
        const int jp = S.ibuf[0];
        const cplx piv = S.col[jp];
        const bool nonzero = !is_zero(piv);
        if (nonzero) {
            ju = max(ju, min(j + ku + jp, n - 1));
            // ---- phase B: row interchange + publish pivot row; multipliers ----
            for (int c = tid; c <= ju - j; c += nt) {
                cplx *top = ab + (size_t) (j + c) * ldab + (kv - c);        // A(j, j+c)
                cplx v;
                if (c == 0) { v = piv; *top = piv; }
                else if (jp != 0) { cplx *bot = top + jp; v = *bot; *bot = *top; *top = v; }
                else v = *top;
                S.u[c] = v;
            }
            const cplx rinv = recip(piv);
            for (int i = tid; i < km; i += nt) {
                // row j+1+i after the interchange: the old diagonal if it was swapped here
                const cplx a = (i + 1 == jp) ? S.col[0] : S.col[i + 1];
                const cplx l = a * rinv;
                colj[i + 1] = l;
                S.l[i] = l;
            }
            __syncthreads();
            // ---- phase C: rank-1 update of the trailing window ----
            const int ncol = ju - j;
            const int total = km * ncol;
            for (int e = tid; e < total; e += nt) {
                const int c = e / km, i = e - c * km;                 // column j+1+c, row j+1+i
                cplx *dst = ab + (size_t) (j + 1 + c) * ldab + (kv - (c + 1)) + (i + 1);
                cplx v = *dst;
                submul(v, S.l[i], S.u[c + 1]);
                *dst = v;
            }
        }
        __syncthreads();
    }
    return S.ibuf[1];
}

// ---------------------------------------------------------------------------
// Warp-level triangular solves with the factors from gbtrf_cta (zgbtrs).
// b: one right hand side of length n (global or shared), overwritten by x.
// Must be called by all 32 lanes of one warp.
// ---------------------------------------------------------------------------
__device__ inline void gbtrs_T_warp(int n, int kl, int ku, const cplx *ab, int ldab,
                                    const int *ipiv, cplx *b)
{
    const int kv = kl + ku;
    const int lane = threadIdx.x & 31;
    // U^T y = b: forward, column-oriented (same summation order as ztbsv)
    for (int j = 0; j < n; ++j) {
        const cplx xj = cdiv(b[j], ab[(size_t) j * ldab + kv]);
        __syncwarp();
        if (lane == 0) b[j] = xj;
        const int cmax = min(kv, n - 1 - j);
        for (int c = 1 + lane; c <= cmax; c += 32) {
            cplx v = b[j + c];
            submul(v, ab[(size_t) (j + c) * ldab + (kv - c)], xj);      // U(j, j+c)
            b[j + c] = v;
        }
        __syncwarp();
    }
    // L^T x = y: backward, dot products with the multipliers, undoing pivots
    for (int j = n - 2; j >= 0; --j) {
        const int lm = min(kl, n - 1 - j);
        cplx s(0.0, 0.0);
        for (int i = 1 + lane; i <= lm; i += 32)
            addmul(s, ab[(size_t) j * ldab + kv + i], b[j + i]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s.x += __shfl_xor_sync(0xffffffffu, s.x, o);
            s.y += __shfl_xor_sync(0xffffffffu, s.y, o);
        }
        __syncwarp();
        if (lane == 0) {
            cplx v = b[j] - s;
            const int l = ipiv[j] - 1;
            if (l != j) { const cplx t = b[l]; b[l] = v; v = t; }
            b[j] = v;
        }
        __syncwarp();
    }
}

__device__ inline void gbtrs_N_warp(int n, int kl, int ku, const cplx *ab, int ldab,
                                    const int *ipiv, cplx *b)
{
    const int kv = kl + ku;
    const int lane = threadIdx.x & 31;
    // L y = P b
    for (int j = 0; j < n - 1; ++j) {
        const int lm = min(kl, n - 1 - j);
        const int l = ipiv[j] - 1;
        if (lane == 0 && l != j) { const cplx t = b[l]; b[l] = b[j]; b[j] = t; }
        __syncwarp();
        const cplx bj = b[j];
        for (int i = 1 + lane; i <= lm; i += 32) {
            cplx v = b[j + i];
            submul(v, ab[(size_t) j * ldab + kv + i], bj);
            b[j + i] = v;
        }
        __syncwarp();
    }
    // U x = y
    for (int j = n - 1; j >= 0; --j) {
        const cplx xj = cdiv(b[j], ab[(size_t) j * ldab + kv]);
        __syncwarp();
        if (lane == 0) b[j] = xj;
        const int cmax = min(kv, j);
        for (int c = 1 + lane; c <= cmax; c += 32) {
            cplx v = b[j - c];
            submul(v, ab[(size_t) j * ldab + (kv - c)], xj);            // U(j-c, j)
            b[j - c] = v;
        }
        __syncwarp();
    }
}

// r <- b - op(A) x for the *unfactored* band matrix a (lda rows = kl+1+ku),
// CTA-wide; returns nothing (caller syncs).  trans 'T': op(A) = A^T.
__device__ inline void gb_residual_cta(char trans, int n, int kl, int ku, const cplx *a,
                                       int lda, const cplx *x, const cplx *b, cplx *r)
{
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        cplx s(0.0, 0.0);
        if (trans == 'T') {
            // (A^T x)_j = sum_i A(i, j) x_i : walk column j
            const int i0 = max(0, j - ku), i1 = min(n - 1, j + kl);
            const cplx *col = a + (size_t) j * lda + (ku - j);
            for (int i = i0; i <= i1; ++i) addmul(s, col[i], x[i]);
        } else {
            // (A x)_j = sum_c A(j, c) x_c : walk row j
            const int c0 = max(0, j - kl), c1 = min(n - 1, j + ku);
            for (int c = c0; c <= c1; ++c) addmul(s, a[(size_t) c * lda + (ku + j - c)], x[c]);
        }
        r[j] = b[j] - s;
    }
}

}  // namespace szb
