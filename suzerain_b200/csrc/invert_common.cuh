// invert_common.cuh -- device helpers shared by the fused invert kernels
// (invert_blocked.cu: v3, invert_pipe.cu: v4): named barriers, mbarrier / TMA bulk
// copies, the pivot reciprocal, and the on-the-fly assembly of P (M + phi L)^T P^T
// (rholut_imexop.def:41-597, operator_hybrid_isothermal.cpp:470-510) from a ring of
// per-point block coefficients.  W is the kernel's compile-time configuration
// (KL, KU, KV, CW, CR, NCOEF); SM is its shared-memory carve-up (needs .coef, .alpha,
// .tref, .tblk).
#pragma once

#include "szb_internal.hpp"
#include "cplx.cuh"
#include "kernels.cuh"

namespace szb {
namespace fused {

constexpr int P = 5;

template <int ID> __device__ __forceinline__ void bar_sync_n(int count)
{ asm volatile("bar.sync %0, %1;" :: "n"(ID), "r"(count) : "memory"); }
template <int ID> __device__ __forceinline__ void bar_arrive_n(int count)
{ asm volatile("bar.arrive %0, %1;" :: "n"(ID), "r"(count) : "memory"); }

__device__ __forceinline__ unsigned smem_u32(const void *p)
{ return (unsigned) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *smem, const void *gmem, unsigned bytes,
                                             unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(smem)), "l"(gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// 1/z.  In the comfortable exponent range one division suffices; outside it fall
// back to Smith's algorithm (cplx.cuh recip), which cannot overflow prematurely.
__device__ __forceinline__ cplx recip_fast(cplx z)
{
    const double m = fmax(fabs(z.x), fabs(z.y));
    if (m > 1e-140 && m < 1e140) {
        const double d = 1.0 / fma(z.x, z.x, z.y * z.y);
        return cplx(z.x * d, -z.y * d);
    }
    return recip(z);
}

// Where the B-spline operator entries D^(d)[yJ, yJ - ku + r] come from: straight from global
// memory (read-only path), or -- for the row block yI whose three operator rows a kernel
// staged in shared memory ahead of time, rowblk[d * ld + r] = D^(d) entry with
// yJ = yI - r + ku -- from there.
struct DGlobal {
    const double *D; int n, ld;
    __device__ __forceinline__ DGlobal(const PackArgs &A) : D(A.D), n(A.n), ld(A.ld) {}
    __device__ __forceinline__ double operator()(int d, int r, int yJ, int) const
    { return __ldg(D + (size_t) (d * ld + r) * n + yJ); }
};
struct DStaged {
    const double *D; int n, ld;
    const double *rowblk; int yI;
    __device__ __forceinline__ DStaged(const PackArgs &A, const double *rb, int y)
        : D(A.D), n(A.n), ld(A.ld), rowblk(rb), yI(y) {}
    __device__ __forceinline__ double operator()(int d, int r, int yJ, int y) const
    { return y == yI ? rowblk[d * ld + r] : __ldg(D + (size_t) (d * ld + r) * n + yJ); }
};

// Coefficient ring layout: c_{row sJ, col sI, op d}(y) at coef_index(y, sJ, sI, d).  sJ runs
// fastest, then the ring slot of y: a warp assembling consecutive columns J = 5 yJ + sJ of one
// matrix row (fixed sI) reads consecutive addresses.
template <class W>
__device__ __forceinline__ int coef_index(int y, int sJ, int sI, int d)
{
    // the ring holds CR consecutive collocation points (a power of two in v3 / v4, 2 kb + 1 in v5)
    const int slot = (W::CR & (W::CR - 1)) == 0 ? (y & (W::CR - 1)) : y % W::CR;
    return ((d * 5 + sI) * W::CR + slot) * 5 + sJ;
}

// ---- assembled entries of P (M + phi L)^T P^T from the coefficient ring ----
template <class W, class DS>
__device__ __forceinline__ cplx base_entry(const PackArgs &A, const cplx *s_coef, const DS &ds, int I, int J)
{
    const int yI = I / 5, sI = I - 5 * yI;
    const int yJ = J / 5, sJ = J - 5 * yJ;
    const int off = yI - yJ;
    if (off < -A.ku || off > A.kl) return cplx(0.0, 0.0);
    const int r = A.ku + off;
    const double m0 = ds(0, r, yJ, yI), d1 = ds(1, r, yJ, yI), d2 = ds(2, r, yJ, yI);
    cplx buf = s_coef[coef_index<W>(yJ, sJ, sI, 0)] * m0;
    buf += s_coef[coef_index<W>(yJ, sJ, sI, 1)] * d1;
    buf += s_coef[coef_index<W>(yJ, sJ, sI, 2)] * d2;
    buf = A.phi * buf;
    if (sI == sJ) buf += cplx(m0, 0.0);
    return buf;
}

// + NRBC lower-right corner (rholut_imexop.def:505-595)
template <class W, class DS>
__device__ __forceinline__ cplx nrbc_entry(const PackArgs &A, const cplx *s_coef, const DS &ds, double km,
                                           double kn, int I, int J)
{
    cplx X = base_entry<W>(A, s_coef, ds, I, J);
    if (!A.nrbc) return X;
    const int i = I - 5 * (A.n - 3), J0 = 5 * (A.n - 1), j = J - J0;
    if (i < 0 || i >= 15 || j < 0 || j >= 5) return X;
    cplx buf(0.0, 0.0);
    if (i >= 10) {
        const cplx ikmphi = cplx(0.0, km) * A.phi, iknphi = cplx(0.0, kn) * A.phi;
        if (A.nrbc & 1) buf -= ikmphi * A.a[5 * (i - 10) + j];
        if (A.nrbc & 2) buf -= iknphi * A.b[5 * (i - 10) + j];
        if (A.nrbc & 4) buf += cplx(A.c[5 * (i - 10) + j], 0.0);
    }
    if (A.nrbc & 4)
        for (int k = 0; k < 5; ++k) buf -= base_entry<W>(A, s_coef, ds, I, J0 + k) * A.c[j + 5 * k];
    return X + buf;
}

// + isothermal wall equations (operator_hybrid_isothermal.cpp:470-510)
template <class W, class DS>
__device__ __forceinline__ cplx assembled_entry(const PackArgs &A, const cplx *s_coef, const DS &ds, double km,
                                                double kn, int I, int J)
{
    if (A.with_bc) {
        const int yJ = J / 5, sJ = J - 5 * yJ;
        int wall = -1;
        if (yJ == 0 && A.wall_begin == 0) wall = 0;
        if (yJ == A.n - 1 && A.wall_end == 2) wall = 1;
        if (wall >= 0 && sJ < 4) {
            const int irho = 5 * yJ + 4;
            if (I != J && I != irho) return cplx(0.0, 0.0);
            cplx s = nrbc_entry<W>(A, s_coef, ds, km, kn, J, J);
            if (is_zero(s)) s = cplx(1.0, 0.0);
            if (I == J) return s;
            const double factor = sJ == 0 ? A.E_factor[wall] : A.vel_factor[wall][sJ - 1];
            return -(s * factor);
        }
    }
    return nrbc_entry<W>(A, s_coef, ds, km, kn, I, J);
}

// per-point block coefficients c_{row,col,op}(y) = sum_t alpha_t ref_t(y)
template <class W, class SM>
__device__ __forceinline__ void compute_coef(const PackArgs &A, const SM &S, int y, int t0,
                                             int nt)
{
    if (y < 0 || y >= A.n) return;
    for (int idx = t0; idx < W::NCOEF; idx += nt) {
        const int tb = S.tblk[idx], te = S.tblk[idx + 1];
        cplx c(0.0, 0.0);
        for (int t = tb; t < te; ++t) c += S.alpha[t] * __ldg(A.refs + (size_t) S.tref[t] * A.n + y);
        S.coef[coef_index<W>(y, idx / 15, (idx / 3) % 5, idx % 3)] = c;
    }
}

// the same from a staged column of the reference profiles, refcol[q] = refs[q][y]
template <class W, class SM>
__device__ __forceinline__ void compute_coef_staged(const PackArgs &A, const SM &S, int y,
                                                    const double *refcol, int t0, int nt)
{
    if (y < 0 || y >= A.n) return;
    // the loads of four terms are issued together (a block has at most seven): the plain loop is a chain of
    // three dependent shared-memory loads per term.  Same summation order.
    for (int idx = t0; idx < W::NCOEF; idx += nt) {
        const int tb = S.tblk[idx], te = S.tblk[idx + 1];
        cplx c(0.0, 0.0);
        for (int tq = tb; tq < te; tq += 4) {
            cplx al[4]; double rr[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool ok = tq + u < te;
                const int t = ok ? tq + u : tb;
                const double r = refcol[S.tref[t]];
                const cplx a = S.alpha[t];
                rr[u] = ok ? r : 0.0;
                al[u] = ok ? a : cplx(0.0, 0.0);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) c += al[u] * rr[u];
        }
        S.coef[coef_index<W>(y, idx / 15, (idx / 3) % 5, idx % 3)] = c;
    }
}

// rows 5*yI .. 5*yI+4, all CW column slots (zeros outside the band), into dst[P][CW]
template <class W, class SM, class DS>
__device__ __forceinline__ void assemble_block(const PackArgs &A, const SM &S, const DS &ds, double km,
                                               double kn, int yI, cplx *dst, int t0, int nt)
{
    for (int e = t0; e < P * W::CW; e += nt) {
        const int sI = e / W::CW, ci = e - sI * W::CW;
        const int I = 5 * yI + sI, J = I - W::KL + ci;        // ci in [0, KV]: in band
        int slot = J % W::CW; if (slot < 0) slot += W::CW;
        cplx v(0.0, 0.0);
        if (ci <= W::KV && I < A.N && J >= 0 && J < A.N) v = assembled_entry<W>(A, S.coef, ds, km, kn, I, J);
        dst[sI * W::CW + slot] = v;
    }
}

template <class W, class SM>
__device__ __forceinline__ void assemble_block(const PackArgs &A, const SM &S, double km,
                                               double kn, int yI, cplx *dst, int t0, int nt)
{
    assemble_block<W>(A, S, DGlobal(A), km, kn, yI, dst, t0, nt);
}

// The same for a row block whose columns touch neither an enforced wall point nor the NRBC
// corner (yI - kl >= 1, yI + ku <= n - 2): plain operator entries, same arithmetic as
// base_entry, operator rows from the staged copy rowblk[d * ld + r].
template <class W, class SM>
__device__ __forceinline__ void assemble_block_interior(const PackArgs &A, const SM &S,
                                                        const double *rowblk, int yI, cplx *dst,
                                                        int t0, int nt)
{
#pragma unroll 3
    for (int e = t0; e < P * W::CW; e += nt) {
        const int sI = e / W::CW, ci = e - sI * W::CW;
        const int J = 5 * yI + sI - W::KL + ci;                 // > 0 here
        const int yJ = J / 5, sJ = J - 5 * yJ;
        const int r = A.ku + yI - yJ;
        cplx v(0.0, 0.0);
        if (ci <= W::KV && r >= 0 && r < A.ld) {
            const double m0 = rowblk[r], d1 = rowblk[A.ld + r], d2 = rowblk[2 * A.ld + r];
            cplx buf = S.coef[coef_index<W>(yJ, sJ, sI, 0)] * m0;
            buf += S.coef[coef_index<W>(yJ, sJ, sI, 1)] * d1;
            buf += S.coef[coef_index<W>(yJ, sJ, sI, 2)] * d2;
            buf = A.phi * buf;
            if (sI == sJ) buf += cplx(m0, 0.0);
            v = buf;
        }
        dst[sI * W::CW + J % W::CW] = v;
    }
}

// The same in two steps, for a caller that has to wait for other warps before it may overwrite dst (the rows
// being replaced are still read by them): compute() forms this thread's entries in registers, store() writes them.
template <class W, int NTHR>
struct InteriorBlock {
    static constexpr int NE = (P * W::CW + NTHR - 1) / NTHR;      // entries per thread
    cplx v[NE];
    int off[NE];                                                  // offset in dst, -1: none
    template <class SM>
    __device__ __forceinline__ void compute(const PackArgs &A, const SM &S, const double *rowblk, int yI, int t0)
    {
#pragma unroll
        for (int q = 0; q < NE; ++q) {
            const int e = t0 + q * NTHR;
            v[q] = cplx(0.0, 0.0); off[q] = -1;
            if (e < P * W::CW) {
                const int sI = e / W::CW, ci = e - sI * W::CW;
                const int J = 5 * yI + sI - W::KL + ci;                 // > 0 here
                const int yJ = J / 5, sJ = J - 5 * yJ;
                const int r = A.ku + yI - yJ;
                if (ci <= W::KV && r >= 0 && r < A.ld) {
                    const double m0 = rowblk[r], d1 = rowblk[A.ld + r], d2 = rowblk[2 * A.ld + r];
                    cplx buf = S.coef[coef_index<W>(yJ, sJ, sI, 0)] * m0;
                    buf += S.coef[coef_index<W>(yJ, sJ, sI, 1)] * d1;
                    buf += S.coef[coef_index<W>(yJ, sJ, sI, 2)] * d2;
                    buf = A.phi * buf;
                    if (sI == sJ) buf += cplx(m0, 0.0);
                    v[q] = buf;
                }
                off[q] = sI * W::CW + J % W::CW;
            }
        }
    }
    __device__ __forceinline__ void store(cplx *dst) const
    {
#pragma unroll
        for (int q = 0; q < NE; ++q) if (off[q] >= 0) dst[off[q]] = v[q];
    }
};

}  // namespace fused
}  // namespace szb
