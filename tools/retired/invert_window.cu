// invert_window.cu -- batched invert of (M + phi L), version 2: the "register
// window" kernel.
//
// Replaces the hot loop of invert_mass_plus_scaled_operator
// (apps/perfect/operator_hybrid_isothermal.cpp:617-686) for the zgbsv solver
// specification: suzerain_rholut_imexop_packf (rholut_imexop.def:41-597) +
// IsothermalPATPTEnforcer::op/rhs (:470-525) + bsmbsm_solver::supply_B /
// zgbtrf + zgbtrs('T') / demand_X (bsmbsm_solver.cpp:155-182), one persistent
// CTA per pencil slot, with the matrix never leaving the SM:
//
//  * zgbtf2 touches, at elimination step j, only rows j..j+KL of columns
//    j..j+KL+KU.  That (KL+1) x (KL+KU+1) window lives in REGISTERS, spread
//    cyclically over TR x TC "forward" threads (row i -> slot i mod (KL+1),
//    column c -> slot c mod (KL+KU+1)), so the rank-1 update is pure FP64 FMA
//    on registers; only the pivot column, the old top row and the pivot row
//    pass through shared memory each step (two named barriers per step).
//  * Rows of P (M + phi L)^T P^T (+ NRBC corner, + wall columns) are
//    assembled on the fly, five at a time (one collocation point), from a ring
//    of per-point block coefficients sum_t alpha_t ref_t(y), and enter the
//    window as their predecessor retires.
//  * The right hand side rides along as one more window row: eliminating it
//    yields y^T = b^T U^{-1}, i.e. the U^T forward sweep of zgbtrs('T') is
//    fused with the factorisation and U is never stored.
//  * Only the multipliers (KL per column) go to a per-slot global scratch
//    (L2 resident); a dedicated "solver" warp runs the L^T back substitution
//    with the row interchanges undone, overlapped with the forward threads'
//    factorisation of the slot's next pencil (double-buffered).
//
// Pivot rule, interchange order and update order are those of zgbtf2, so
// ipiv is identical to LAPACK's.  tools/window_lu_model.py is an executable
// model of the index algebra.
#include <cstdio>

#include "szb_internal.hpp"
#include "cplx.cuh"
#include "kernels.cuh"

namespace szb {

namespace {

enum { BAR_FWD = 1, BAR_FULL0 = 2, BAR_EMPTY0 = 4 };   // FULL1 = 3, EMPTY1 = 5

__device__ __forceinline__ void bar_sync(int id, int count)
{ asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int count)
{ asm volatile("bar.arrive %0, %1;" :: "r"(id), "r"(count) : "memory"); }

// ---- mbarrier + TMA bulk copy (global -> shared) used by the solver warp ----
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

template <int KL_, int KU_, int TR_, int TC_, int CR_>
struct WinCfg {
    static constexpr int KL = KL_, KU = KU_, KV = KL_ + KU_, R = KL_ + 1, C = KL_ + KU_ + 1;
    static constexpr int TR = TR_, TC = TC_, NT = TR_ * TC_, NTH = NT + 32;
    static constexpr int RA = (R + 1 + TR - 1) / TR;      // row slots per thread (slot R = RHS)
    static constexpr int CB = (C + TC - 1) / TC;          // column slots per thread
    static constexpr int RP = RA * TR, CP = CB * TC;      // padded slot counts
    static constexpr int A_RHS = R / TR, T_RHS = R % TR;  // owner coordinates of the RHS row
    static constexpr int CR = CR_;                        // coefficient ring (points), power of 2
    static constexpr int NCOEF = 75;
    static constexpr int CH = 8, NB = 4;                  // solver: L columns per TMA chunk, ring depth
    static_assert(NT % 32 == 0, "forward threads must fill whole warps");
    static_assert(R % 5 == 0, "window rows come in groups of five (one collocation point)");
    static_assert((CR & (CR - 1)) == 0, "ring size must be a power of two");
};

struct WindowArgs {
    PackArgs pk;
    int npencil; const int *index;
    cplx *state; size_t fs, ps;
    int *ipiv_out, *info_out, *iters_out;
    cplx *lwork;            // per CTA: 2 buffers of N*KL multipliers
};

// ---- assembled entries of P (M + phi L)^T P^T from the coefficient ring ----
template <class W>
__device__ __forceinline__ cplx base_entry(const PackArgs &A, const cplx *s_coef, int I, int J)
{
    const int yI = I / 5, sI = I - 5 * yI;
    const int yJ = J / 5, sJ = J - 5 * yJ;
    const int off = yI - yJ;
    if (off < -A.ku || off > A.kl) return cplx(0.0, 0.0);
    const int r = A.ku + off;
    const cplx *c = s_coef + (yJ & (W::CR - 1)) * W::NCOEF + (sJ * 5 + sI) * 3;
    const size_t ds = (size_t) A.ld * A.n;
    const double *D = A.D + (size_t) r * A.n + yJ;
    const double m0 = __ldg(D), d1 = __ldg(D + ds), d2 = __ldg(D + 2 * ds);
    cplx buf = c[0] * m0;
    buf += c[1] * d1;
    buf += c[2] * d2;
    buf = A.phi * buf;
    if (sI == sJ) buf += cplx(m0, 0.0);
    return buf;
}

// + NRBC lower-right corner (rholut_imexop.def:505-595)
template <class W>
__device__ __forceinline__ cplx nrbc_entry(const PackArgs &A, const cplx *s_coef, double km,
                                           double kn, int I, int J)
{
    cplx X = base_entry<W>(A, s_coef, I, J);
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
        for (int k = 0; k < 5; ++k) buf -= base_entry<W>(A, s_coef, I, J0 + k) * A.c[j + 5 * k];
    return X + buf;
}

// + isothermal wall equations (operator_hybrid_isothermal.cpp:470-510)
template <class W>
__device__ __forceinline__ cplx assembled_entry(const PackArgs &A, const cplx *s_coef, double km,
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
            cplx s = nrbc_entry<W>(A, s_coef, km, kn, J, J);
            if (is_zero(s)) s = cplx(1.0, 0.0);
            if (I == J) return s;
            const double factor = sJ == 0 ? A.E_factor[wall] : A.vel_factor[wall][sJ - 1];
            return -(s * factor);
        }
    }
    return nrbc_entry<W>(A, s_coef, km, kn, I, J);
}

template <class W>
struct Smem {
    cplx *alpha;      // [MAXTERMS]      wave(km,kn) * scenario factor per term
    cplx *coef;       // [CR][75]        per-point block coefficients
    cplx *stage;      // [2][5][CP]      assembled rows waiting to enter the window
    cplx *col;        // [2][RP]         pivot column (by row slot; slot R = RHS)
    cplx *top;        // [2][CP]         old top row (by column slot)
    cplx *piv;        // [2][CP]         pivot row
    cplx *v;          // [2][N]          b -> y -> x per buffer
    unsigned char *ipiv;   // [2][N]     jp per column
    unsigned char *tref;   // [MAXTERMS] term -> reference profile
    unsigned char *tblk;   // [76]       block-op -> first term
    int *info;        // [2]
    cplx *lring;      // [NB][CH*KL]     multipliers prefetched by TMA for the solver warp
    unsigned long long *mbar;   // [NB]
};

template <class W>
__host__ __device__ inline size_t window_smem_bytes(int N)
{
    size_t b = sizeof(cplx) * ((size_t) MAXTERMS + W::CR * W::NCOEF + 2 * 5 * W::CP + 2 * W::RP
                               + 4 * W::CP + 2 * (size_t) N + W::NB * W::CH * W::KL);
    b += 2 * (size_t) N + MAXTERMS + 80 + 16 + 8 * W::NB;
    return (b + 15) & ~(size_t) 15;
}

template <class W>
__device__ __forceinline__ Smem<W> carve(unsigned char *raw, int N)
{
    Smem<W> S;
    cplx *p = reinterpret_cast<cplx *>(raw);
    S.alpha = p; p += MAXTERMS;
    S.coef = p;  p += W::CR * W::NCOEF;
    S.stage = p; p += 2 * 5 * W::CP;
    S.col = p;   p += 2 * W::RP;
    S.top = p;   p += 2 * W::CP;
    S.piv = p;   p += 2 * W::CP;
    S.v = p;     p += 2 * (size_t) N;
    S.lring = p; p += W::NB * W::CH * W::KL;
    unsigned char *q = reinterpret_cast<unsigned char *>(p);
    S.mbar = reinterpret_cast<unsigned long long *>(q); q += 8 * W::NB;
    S.info = reinterpret_cast<int *>(q); q += 16;
    S.ipiv = q; q += 2 * (size_t) N;
    S.tref = q; q += MAXTERMS;
    S.tblk = q;
    return S;
}

// per-point block coefficients c_{row,col,op}(y) = sum_t alpha_t ref_t(y)
template <class W>
__device__ __forceinline__ void compute_coef(const PackArgs &A, const Smem<W> &S, int y, int tid)
{
    if (y < 0 || y >= A.n) return;
    for (int idx = tid; idx < W::NCOEF; idx += W::NT) {
        const int tb = S.tblk[idx], te = S.tblk[idx + 1];
        cplx c(0.0, 0.0);
        for (int t = tb; t < te; ++t) c += S.alpha[t] * __ldg(A.refs + (size_t) S.tref[t] * A.n + y);
        S.coef[(y & (W::CR - 1)) * W::NCOEF + idx] = c;
    }
}

// rows 5*yI .. 5*yI+4 into stage buffer (yI & 1), by column slot
template <class W>
__device__ __forceinline__ void assemble_block(const PackArgs &A, const Smem<W> &S, double km,
                                               double kn, int yI, int tid)
{
    cplx *dst = S.stage + (size_t) (yI & 1) * 5 * W::CP;
    for (int e = tid; e < 5 * W::C; e += W::NT) {
        const int sI = e / W::C, ci = e - sI * W::C;
        const int I = 5 * yI + sI, J = I - W::KL + ci;
        int slot = J % W::C; if (slot < 0) slot += W::C;
        cplx v(0.0, 0.0);
        if (I < A.N && J >= 0 && J < A.N) v = assembled_entry<W>(A, S.coef, km, kn, I, J);
        dst[sI * W::CP + slot] = v;
    }
}

template <class W>
__global__ void __launch_bounds__(W::NTH)
invert_window_kernel(const WindowArgs A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const PackArgs &K = A.pk;
    const int N = K.N, n = K.n;
    const Smem<W> S = carve<W>(smem_raw, N);
    const int tid = threadIdx.x;
    constexpr int KL = W::KL, KU = W::KU, KV = W::KV, R = W::R, C = W::C;
    constexpr int RA = W::RA, CB = W::CB, TR = W::TR, TC = W::TC, NT = W::NT;
    cplx *lwork = A.lwork + (size_t) blockIdx.x * 2 * (size_t) N * KL;

    for (int t = tid; t < MAXTERMS; t += W::NTH) S.tref[t] = K.terms->ref[t];
    for (int t = tid; t <= NBLOCK; t += W::NTH) S.tblk[t] = K.terms->blk_begin[t];
    if (tid == NT) {
        for (int b = 0; b < W::NB; ++b) mbar_init(S.mbar + b, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (tid >= NT) {
        // =================== solver warp: L^T back substitution ===================
        const int lane = tid - NT;
        int q = 0;
        unsigned chunk_base = 0;        // running chunk count: ring slot and mbarrier phase
        for (int p = blockIdx.x; p < A.npencil; p += gridDim.x, ++q) {
            const int buf = q & 1;
            bar_sync(BAR_FULL0 + buf, W::NTH);
            cplx *x = S.v + (size_t) buf * N;
            const unsigned char *jpv = S.ipiv + (size_t) buf * N;
            const cplx *Lg = lwork + (size_t) buf * N * KL;
            const int info = S.info[buf];
            if (info == 0) {
                // multipliers stream in through a ring of TMA bulk copies, last columns first
                constexpr int CH = W::CH, NB = W::NB;
                const int ncols = N - 1, nchunk = (ncols + CH - 1) / CH;
                asm volatile("fence.proxy.async;" ::: "memory");
                auto issue = [&](int c) {
                    const int jhi = N - 2 - c * CH, jlo = max(jhi - CH + 1, 0);
                    const unsigned bytes = (unsigned) ((jhi - jlo + 1) * KL * sizeof(cplx));
                    const unsigned slot = (chunk_base + c) % NB;
                    mbar_expect_tx(S.mbar + slot, bytes);
                    tma_bulk_g2s(S.lring + (size_t) slot * CH * KL, Lg + (size_t) jlo * KL, bytes,
                                 S.mbar + slot);
                };
                if (lane == 0) for (int c = 0; c < min(NB, nchunk); ++c) issue(c);
                for (int c = 0; c < nchunk; ++c) {
                    const unsigned g = chunk_base + c, slot = g % NB, parity = (g / NB) & 1;
                    mbar_wait(S.mbar + slot, parity);
                    const int jhi = N - 2 - c * CH, jlo = max(jhi - CH + 1, 0);
                    const cplx *Lc = S.lring + (size_t) slot * CH * KL;
                    for (int j = jhi; j >= jlo; --j) {
                        const int lm = min(KL, N - 1 - j);
                        const cplx *Lj = Lc + (size_t) (j - jlo) * KL;
                        cplx s(0.0, 0.0);
                        for (int i = 1 + lane; i <= lm; i += 32) addmul(s, Lj[i - 1], x[j + i]);
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            s.x += __shfl_xor_sync(0xffffffffu, s.x, o);
                            s.y += __shfl_xor_sync(0xffffffffu, s.y, o);
                        }
                        if (lane == 0) {
                            cplx v = x[j] - s;
                            const int l = j + jpv[j];
                            if (l != j) { const cplx t = x[l]; x[l] = v; v = t; }
                            x[j] = v;
                        }
                        __syncwarp();
                    }
                    if (lane == 0 && c + NB < nchunk) {
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        issue(c + NB);
                    }
                }
                chunk_base += nchunk;
                // state = P^T x (bsmbsm_solver.hpp:274-280)
                cplx *v = A.state + (A.index ? (size_t) A.index[p] : (size_t) p) * A.ps;
                for (int e = lane; e < N; e += 32) {
                    const int f = e / n, y = e - f * n;
                    v[(size_t) f * A.fs + y] = x[5 * y + f];
                }
            }
            if (lane == 0) {
                A.info_out[p] = info;
                if (A.iters_out) A.iters_out[p] = 0;
            }
            if (A.ipiv_out)
                for (int k = lane; k < N; k += 32) A.ipiv_out[(size_t) p * N + k] = k + jpv[k] + 1;
            __threadfence_block();
            if (p + 2 * (int) gridDim.x < A.npencil) bar_arrive(BAR_EMPTY0 + buf, W::NTH);
        }
        return;
    }

    // ======================= forward threads: windowed LU =======================
    const int tr = tid % TR, tc = tid / TR;
    const int lane = tid & 31;
    int q = 0;
    for (int p = blockIdx.x; p < A.npencil; p += gridDim.x, ++q) {
        const int buf = q & 1;
        if (q >= 2) bar_sync(BAR_EMPTY0 + buf, W::NTH);
        cplx *sv = S.v + (size_t) buf * N;
        unsigned char *jpv = S.ipiv + (size_t) buf * N;
        cplx *Lg = lwork + (size_t) buf * N * KL;
        const double km = K.km[p], kn = K.kn[p];

        // b = P state with the wall rows zeroed (bsmbsm_solver.hpp:150-156,
        // operator_hybrid_isothermal.cpp:516-525)
        {
            const cplx *v = A.state + (A.index ? (size_t) A.index[p] : (size_t) p) * A.ps;
            for (int e = tid; e < N; e += NT) {
                const int f = e / n, y = e - f * n;
                cplx val = v[(size_t) f * A.fs + y];
                if (K.with_bc && f < 4
                    && ((y == 0 && K.wall_begin == 0) || (y == n - 1 && K.wall_end == 2)))
                    val = cplx(0.0, 0.0);
                sv[5 * y + f] = val;
            }
        }
        for (int t = tid; t < K.terms->nterms; t += NT)
            S.alpha[t] = wave_factor(K.terms->wave[t], km, kn) * K.terms->sc[t];
        if (tid == 0) S.info[buf] = 0;
        bar_sync(BAR_FWD, NT);
        for (int y = 0; y <= K.ku + 1; ++y) compute_coef<W>(K, S, y, tid);
        bar_sync(BAR_FWD, NT);
        assemble_block<W>(K, S, km, kn, 0, tid);
        bar_sync(BAR_FWD, NT);

        // ---- window registers ----
        cplx Wd[RA][CB];
        int relr[RA], relc[CB];
#pragma unroll
        for (int a = 0; a < RA; ++a) {
            const int rs = tr + a * TR;
            relr[a] = rs < R ? rs : (rs == R ? -1 : -1000);     // step 0: jr = 0
#pragma unroll
            for (int b = 0; b < CB; ++b) Wd[a][b] = cplx(0.0, 0.0);
        }
#pragma unroll
        for (int b = 0; b < CB; ++b) {
            const int cs = tc + b * TC;
            relc[b] = cs < C ? cs : -1000;
            if (tr == W::T_RHS && cs < C) Wd[W::A_RHS][b] = cs < N ? sv[cs] : cplx(0.0, 0.0);
        }
        // rows 0..KL enter (virtual steps), block by block
        for (int blk = 0; blk < R / 5; ++blk) {
            compute_coef<W>(K, S, blk + 2 + K.ku, tid);
            assemble_block<W>(K, S, km, kn, blk + 1, tid);
            const cplx *src = S.stage + (size_t) (blk & 1) * 5 * W::CP;
#pragma unroll
            for (int a = 0; a < RA; ++a) {
                const int rs = tr + a * TR;
                if (rs < R && rs / 5 == blk) {
#pragma unroll
                    for (int b = 0; b < CB; ++b) {
                        const int cs = tc + b * TC;
                        if (cs < C) Wd[a][b] = src[(rs - 5 * blk) * W::CP + cs];
                    }
                }
            }
            bar_sync(BAR_FWD, NT);
        }

        int par = 0, ju = 0, info = 0;
        int jr = 0, jc = 0;              // j mod R, j mod C
        for (int j = 0; j < N; ++j) {
            cplx *s_col = S.col + par * W::RP, *s_top = S.top + par * W::CP, *s_piv = S.piv + par * W::CP;
            const int kmj = min(KL, N - 1 - j);
            // ---- A: publish column j and the top row ----
#pragma unroll
            for (int b = 0; b < CB; ++b) {
                if (relc[b] == 0) {
#pragma unroll
                    for (int a = 0; a < RA; ++a)
                        if (relr[a] > -1000) s_col[tr + a * TR] = Wd[a][b];
                }
            }
#pragma unroll
            for (int a = 0; a < RA; ++a) {
                if (relr[a] == 0) {
#pragma unroll
                    for (int b = 0; b < CB; ++b)
                        if (relc[b] > -1000) s_top[tc + b * TC] = Wd[a][b];
                }
            }
            bar_sync(BAR_FWD, NT);
            // ---- C: pivot search, every warp redundantly (izamax: first max of cabs1) ----
            int jp;
            {
                int i0 = lane, i1 = lane + 32;
                int s0 = jr + i0; if (s0 >= R) s0 -= R;
                int s1 = jr + i1; if (s1 >= R) s1 -= R; if (s1 >= R) s1 -= R;
                long long k0 = -1, k1 = -1;
                if (i0 <= kmj) k0 = __double_as_longlong(cabs1(s_col[s0]));
                if (i1 <= kmj) k1 = __double_as_longlong(cabs1(s_col[s1]));
                const bool second = k1 > k0;              // strict: first maximum wins
                const long long kb = second ? k1 : k0;
                const int hi = (int) (kb >> 32);
                const unsigned lo = (unsigned) (kb & 0xffffffffll);
                const int mh = __reduce_max_sync(0xffffffffu, hi);
                const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
                const bool ismax = hi == mh && lo == ml;
                const unsigned b0 = __ballot_sync(0xffffffffu, ismax && !second);
                const unsigned b1 = __ballot_sync(0xffffffffu, ismax);
                jp = b0 ? __ffs(b0) - 1 : 32 + __ffs(b1) - 1;
            }
            int rp = jr + jp; if (rp >= R) rp -= R;
            const cplx pivot = s_col[rp];
            if (is_zero(pivot)) { info = j + 1; break; }          // uniform across the CTA
            ju = max(ju, min(j + KU + jp, N - 1));
            // ---- D: publish the pivot row; the old top row takes its slot ----
#pragma unroll
            for (int a = 0; a < RA; ++a) {
                if (relr[a] == jp) {
#pragma unroll
                    for (int b = 0; b < CB; ++b) {
                        if (relc[b] > -1000) {
                            s_piv[tc + b * TC] = Wd[a][b];
                            if (jp != 0) Wd[a][b] = s_top[tc + b * TC];
                        }
                    }
                }
            }
            const cplx rinv = recip(pivot);
            const cplx topv = s_col[jr];
            cplx l[RA];
#pragma unroll
            for (int a = 0; a < RA; ++a) {
                l[a] = cplx(0.0, 0.0);
                if (relr[a] >= 1 && relr[a] <= kmj) {
                    const cplx val = relr[a] == jp ? topv : s_col[tr + a * TR];
                    l[a] = val * rinv;
                } else if (relr[a] == -1) {
                    l[a] = s_col[R] * rinv;                        // y_j = t_j / U(j,j)
                }
            }
            bar_sync(BAR_FWD, NT);
            // ---- F: rank-1 update of the window (and of the RHS row) ----
            const int width = ju - j;
#pragma unroll
            for (int b = 0; b < CB; ++b) {
                if (relc[b] >= 1 && relc[b] <= width) {
                    const cplx u = s_piv[tc + b * TC];
#pragma unroll
                    for (int a = 0; a < RA; ++a) submul(Wd[a][b], l[a], u);
                }
            }
            // ---- G: retire column j and row j; column j+KV+1 and row j+KL+1 enter ----
            const int cn = j + KV + 1, rn = j + KL + 1;
#pragma unroll
            for (int b = 0; b < CB; ++b) {
                if (relc[b] == 0) {
#pragma unroll
                    for (int a = 0; a < RA; ++a) {
                        if (relr[a] >= 1 && relr[a] <= kmj) Lg[(size_t) j * KL + relr[a] - 1] = l[a];
                        if (relr[a] == -1) {
                            sv[j] = l[a];
                            Wd[a][b] = cn < N ? sv[cn] : cplx(0.0, 0.0);
                        } else {
                            Wd[a][b] = cplx(0.0, 0.0);
                        }
                    }
                }
            }
            if (tid == 0) jpv[j] = (unsigned char) jp;
            {
                const cplx *src = S.stage + (size_t) ((rn / 5) & 1) * 5 * W::CP + (rn % 5) * W::CP;
#pragma unroll
                for (int a = 0; a < RA; ++a) {
                    if (relr[a] == 0) {
#pragma unroll
                        for (int b = 0; b < CB; ++b)
                            if (relc[b] > -1000) Wd[a][b] = rn < N ? src[tc + b * TC] : cplx(0.0, 0.0);
                    }
                }
            }
            // block bookkeeping: a new group of five rows starts entering at steps j % 5 == 0
            if (jr % 5 == 0) {
                const int yI = rn / 5;                    // block whose first row just entered
                compute_coef<W>(K, S, yI + 2 + K.ku, tid);
                assemble_block<W>(K, S, km, kn, yI + 1, tid);
            }
#pragma unroll
            for (int a = 0; a < RA; ++a) if (relr[a] >= 0) relr[a] = relr[a] == 0 ? KL : relr[a] - 1;
#pragma unroll
            for (int b = 0; b < CB; ++b) if (relc[b] >= 0) relc[b] = relc[b] == 0 ? KV : relc[b] - 1;
            if (++jr == R) jr = 0;
            if (++jc == C) jc = 0;
            par ^= 1;
        }
        if (tid == 0) { S.info[buf] = info; if (info) for (int k = 0; k < N; ++k) jpv[k] = 0; }
        __threadfence();
        bar_arrive(BAR_FULL0 + buf, W::NTH);
    }
}

template <class W>
int launch_window(const szb_imexop *op, WindowArgs &A, int npencil, cudaStream_t stream)
{
    const int N = op->A.N;
    const size_t smem = window_smem_bytes<W>(N);
    if (smem > 227 * 1024) return 1;                 // caller falls back to the v1 kernel
    static bool configured = false;
    if (!configured) {
        SZB_CUDA_OK(cudaFuncSetAttribute(invert_window_kernel<W>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured = true;
    }
    int per_sm = 0;
    SZB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, invert_window_kernel<W>,
                                                              W::NTH, smem));
    if (per_sm < 1) return 1;
    int slots = op->sm_count * per_sm;
    if (slots > npencil) slots = npencil;
    const size_t need = (size_t) slots * 2 * N * W::KL * sizeof(cplx);
    if (need > op->work_bytes) {
        if (op->d_work) SZB_CUDA_OK(cudaFree(op->d_work));
        op->d_work = nullptr; op->work_bytes = 0;
        SZB_CUDA_OK(cudaMalloc(&op->d_work, need));
        op->work_bytes = need;
    }
    op->work_slots = slots;
    A.lwork = static_cast<cplx *>(op->d_work);
    invert_window_kernel<W><<<slots, W::NTH, smem, stream>>>(A);
    count_launch();
    SZB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace

// Returns 0 when launched, 1 when this (kl, ku) / size has no window
// instantiation (the caller then uses the generic v1 kernel), <0 on error.
int invert_window_dispatch(const szb_imexop *op, const double phi[2], int npencil,
                           const double *d_km, const double *d_kn, const int *d_index,
                           cplx *d_state, size_t fs, size_t ps, int *d_ipiv, int *d_info,
                           int *d_iters, cudaStream_t stream)
{
    WindowArgs A;
    fill_pack_args(op, phi, d_km, d_kn, 0, 1, nullptr, A.pk);
    A.npencil = npencil; A.index = d_index;
    A.state = d_state; A.fs = fs; A.ps = ps;
    A.ipiv_out = d_ipiv; A.info_out = d_info; A.iters_out = d_iters;
    A.lwork = nullptr;
    if (op->A.KL != op->A.KU) return 1;
    switch (op->A.KL) {
    case 14: return launch_window<WinCfg<14, 14, 8, 8, 8>>(op, A, npencil, stream);     // k = 4
    case 24: return launch_window<WinCfg<24, 24, 8, 8, 16>>(op, A, npencil, stream);    // k = 6
    case 34: return launch_window<WinCfg<34, 34, 12, 8, 16>>(op, A, npencil, stream);   // k = 8
    case 44: return launch_window<WinCfg<44, 44, 12, 16, 32>>(op, A, npencil, stream);  // k = 10
    default: return 1;
    }
}

}  // namespace szb
