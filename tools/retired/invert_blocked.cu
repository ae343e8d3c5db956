// invert_blocked.cu -- batched invert of (M + phi L), version 3: blocked banded LU on
// a sliding shared-memory window.
//
// Replaces the hot loop of invert_mass_plus_scaled_operator
// (apps/perfect/operator_hybrid_isothermal.cpp:617-686) for the zgbsv solver
// specification: suzerain_rholut_imexop_packf (rholut_imexop.def:41-597) +
// IsothermalPATPTEnforcer::op/rhs (:470-525) + bsmbsm_solver::supply_B /
// zgbtrf + zgbtrs('T') / demand_X (bsmbsm_solver.cpp:155-182).  One persistent CTA
// per pencil slot; the matrix is assembled, factored and consumed on the SM and never
// touches HBM.
//
//  * Panel width P = 5 = the five scalars of one collocation point.  At panel j the
//    LAPACK elimination touches rows j..j+P-1+KL of columns j..j+P-1+KL+KU; that
//    window lives in shared memory, cyclic in its columns.
//  * Row interchanges are never performed physically: every window row slot carries
//    the logical row it currently holds, an interchange relabels two slots, and a
//    retired pivot row's slot is handed to the next row entering the window.
//  * One warp factors the P panel columns in registers (izamax pivot search with
//    REDUX, pivot rows broadcast by shuffles) while the other warps assemble the next
//    five rows of P (M + phi L)^T P^T (+ NRBC corner, + wall columns) from a ring of
//    per-point block coefficients.  Then all warps apply the rank-P update to the
//    trailing columns (20 FP64 FMAs per element loaded from shared memory).
//  * The right hand side is one more window row: eliminating it yields
//    y^T = b^T U^-1, the U^T sweep of zgbtrs('T'), so U is never stored.
//  * Multipliers go to a per-slot global scratch in zgbtf2 order; a dedicated solver
//    warp pulls them back with TMA bulk copies and runs the L^T back substitution with
//    the interchanges undone, overlapped with the factorisation of the slot's next
//    pencil.
//
// Arithmetic per element is the same sequence of FMAs as the unblocked zgbtf2 sweep;
// the pivot rule is izamax's (first maximum of |re|+|im|), so ipiv is LAPACK's.
// tools/blocked_window_model.py is an executable model of the index algebra.
#include <climits>
#include <cstdio>

#include "invert_common.cuh"

namespace szb {

namespace {

using namespace fused;


template <int KL_, int KU_, int CR_, int NWK_, int RPG_>
struct BlkCfg {
    static constexpr int KL = KL_, KU = KU_, KV = KL_ + KU_;
    static constexpr int RW = KL_ + P + 1;          // matrix row slots (one spare row keeps blocks aligned)
    static constexpr int NS = RW + 1;               // + the right-hand-side row (slot RW)
    static constexpr int CW = KV + P + 1;           // column slots
    static constexpr int CR = CR_;                  // coefficient ring (collocation points), power of 2
    static constexpr int NWK = NWK_;                // compute warps (warp 0 factors the panels)
    static constexpr int NT = 32 * NWK_, NTH = NT + 32;
    static constexpr int RPG = RPG_;                // rows per trailing-update task
    static constexpr int NG = (NS + RPG_ - 1) / RPG_;
    static constexpr int NCOEF = 75;
    static constexpr int CH = 4, NB = 3;            // solver: L columns per TMA chunk, ring depth
    static_assert(RW % P == 0, "window rows come in groups of five");
    static_assert(NS <= 64, "panel warp holds two row slots per lane");
    static_assert((CR & (CR - 1)) == 0, "ring size must be a power of two");
};

struct BlockedArgs {
    PackArgs pk;
    int npencil; const int *index;
    cplx *state; size_t fs, ps;
    int *ipiv_out, *info_out, *iters_out;
    cplx *lwork;            // per CTA: 2 buffers of N*KL multipliers
};

template <class W>
struct Smem {
    cplx *win;        // [NS][CW]        the window
    cplx *lp;         // [NS][P]         panel multipliers by slot (zeros past a pivot row's own step)
    cplx *lcol;       // [P][KL]         the same multipliers in zgbtf2 (column, row offset) order
    cplx *xch;        // [2][2][1 + P]   per column parity, per panel warp: {key, label, slot}, candidate row
    cplx *stage;      // [2][P][CW]      assembled rows waiting to enter
    cplx *coef;       // [CR][75]        per-point block coefficients
    cplx *alpha;      // [MAXTERMS]
    cplx *v;          // [2][N]          b -> y -> x per buffer
    cplx *lring;      // [NB][CH*KL]     multipliers prefetched by TMA for the solver warp
    unsigned long long *mbar;   // [NB]
    int *pivslot;     // [2][P]
    int *misc;        // [0..1] info per buffer, [2] ju, [3] panel info
    unsigned char *isp;    // [2][64]    slot retired by the panel of that parity
    unsigned char *ipiv;   // [2][N]     jp per column
    unsigned char *tref;   // [MAXTERMS]
    unsigned char *tblk;   // [76]
};

template <class W>
__host__ __device__ inline size_t blocked_smem_bytes(int N)
{
    size_t b = sizeof(cplx) * ((size_t) W::NS * W::CW + W::NS * P + P * W::KL + 4 * (1 + P) + 2 * P * W::CW + W::CR * W::NCOEF
                               + MAXTERMS + 2 * (size_t) N + W::NB * W::CH * W::KL);
    b += 8 * W::NB + 4 * (2 * P + 6) + 2 * 64 + 2 * (size_t) N + MAXTERMS + 80;
    return (b + 15) & ~(size_t) 15;
}

template <class W>
__device__ __forceinline__ Smem<W> carve(unsigned char *raw, int N)
{
    Smem<W> S;
    cplx *p = reinterpret_cast<cplx *>(raw);
    S.win = p;   p += W::NS * W::CW;
    S.lp = p;    p += W::NS * P;
    S.lcol = p;  p += P * W::KL;
    S.xch = p;   p += 4 * (1 + P);
    S.stage = p; p += 2 * P * W::CW;
    S.coef = p;  p += W::CR * W::NCOEF;
    S.alpha = p; p += MAXTERMS;
    S.v = p;     p += 2 * (size_t) N;
    S.lring = p; p += W::NB * W::CH * W::KL;
    unsigned char *q = reinterpret_cast<unsigned char *>(p);
    S.mbar = reinterpret_cast<unsigned long long *>(q); q += 8 * W::NB;
    S.pivslot = reinterpret_cast<int *>(q); q += 4 * 2 * P;
    S.misc = reinterpret_cast<int *>(q); q += 4 * 6;
    S.isp = q;  q += 2 * 64;
    S.ipiv = q; q += 2 * (size_t) N;
    S.tref = q; q += MAXTERMS;
    S.tblk = q;
    return S;
}


template <class W>
__global__ void __launch_bounds__(W::NTH, 2)
invert_blocked_kernel(const BlockedArgs A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const PackArgs &K = A.pk;
    const int N = K.N, n = K.n;
    const Smem<W> S = carve<W>(smem_raw, N);
    const int tid = threadIdx.x;
    constexpr int KL = W::KL, KU = W::KU, RW = W::RW, NS = W::NS, CW = W::CW, NT = W::NT;
    constexpr int BAR_ALL = 1, BAR_FULL0 = 2, BAR_FULL1 = 3, BAR_EMPTY0 = 4, BAR_EMPTY1 = 5, BAR_PANEL = 6;
    const size_t lstride = ((size_t) N * KL + 7) & ~(size_t) 7;     // per buffer, whole 128-byte lines
    cplx *lwork = A.lwork + (size_t) blockIdx.x * 2 * lstride;

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
            if (buf == 0) bar_sync_n<BAR_FULL0>(W::NTH); else bar_sync_n<BAR_FULL1>(W::NTH);
            cplx *x = S.v + (size_t) buf * N;
            const unsigned char *jpv = S.ipiv + (size_t) buf * N;
            const cplx *Lg = lwork + (size_t) buf * lstride;
            const int info = S.misc[buf];
            if (info == 0) {
                // multipliers stream in through a ring of TMA bulk copies, last columns first
                constexpr int CH = W::CH, NB = W::NB;
                // chunk c covers columns [CH (nchunk-1-c), +CH): aligned so that a consumed chunk is a
                // whole number of 128-byte lines
                const int ncols = N - 1, nchunk = (ncols + CH - 1) / CH;
                asm volatile("fence.proxy.async;" ::: "memory");
                auto issue = [&](int c) {
                    const int jlo = CH * (nchunk - 1 - c), jhi = min(jlo + CH - 1, N - 2);
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
                    const int jlo = CH * (nchunk - 1 - c), jhi = min(jlo + CH - 1, N - 2);
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
                    // The multipliers of these columns are dead now: drop their (dirty) L2 lines
                    // instead of letting them be written back to HBM.
                    if ((CH * KL * sizeof(cplx)) % 128 == 0 && jhi - jlo + 1 == CH) {
                        const char *g0 = reinterpret_cast<const char *>(Lg + (size_t) jlo * KL);
                        if ((reinterpret_cast<size_t>(g0) & 127) == 0)
                        for (int ln = lane; ln < (int) (CH * KL * sizeof(cplx) / 128); ln += 32)
                            asm volatile("discard.global.L2 [%0], 128;" :: "l"(g0 + (size_t) ln * 128) : "memory");
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
            if (p + 2 * (int) gridDim.x < A.npencil) {
                if (buf == 0) bar_arrive_n<BAR_EMPTY0>(W::NTH); else bar_arrive_n<BAR_EMPTY1>(W::NTH);
            }
        }
        return;
    }

    // ============================ compute warps ============================
    const int lane = tid & 31, warp = tid >> 5;
    int q = 0;
    for (int p = blockIdx.x; p < A.npencil; p += gridDim.x, ++q) {
        const int buf = q & 1;
        if (q >= 2) { if (buf == 0) bar_sync_n<BAR_EMPTY0>(W::NTH); else bar_sync_n<BAR_EMPTY1>(W::NTH); }
        cplx *sv = S.v + (size_t) buf * N;
        unsigned char *jpv = S.ipiv + (size_t) buf * N;
        cplx *Lg = lwork + (size_t) buf * lstride;
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
        if (tid == 0) { S.misc[buf] = 0; S.misc[3] = 0; }
        if (tid < 2 * 64) S.isp[tid] = 0;
        bar_sync_n<BAR_ALL>(NT);
        for (int y = 0; y <= RW / 5 + K.ku; ++y) compute_coef<W>(K, S, y, tid, NT);
        bar_sync_n<BAR_ALL>(NT);
        // initial window: logical rows 0..RW-1 in slots 0..RW-1; RHS row t_c = b_c
        for (int blk = 0; blk < RW / 5; ++blk)
            assemble_block<W>(K, S, km, kn, blk, S.win + (size_t) blk * P * CW, tid, NT);
        for (int c = tid; c < CW; c += NT) S.win[(size_t) RW * CW + c] = c < N ? sv[c] : cplx(0.0, 0.0);
        bar_sync_n<BAR_ALL>(NT);

        // panel-warp state (warps 0 and 1 factor the panels together): the logical row held by
        // this lane's row slot
        const int fslot = lane + 32 * warp;                  // meaningful for warp < 2
        const bool have = warp < 2 && fslot < NS, rhs = fslot == RW, liveb = have && !rhs;
        int lg = (warp < 2 && fslot < RW) ? fslot : INT_MIN;
        int pk = P;                     // panel step at which the slot's row became a pivot row
        int ju = 0, info = 0, par = 0;
        int jc = 0;                     // j mod CW

        for (int j = 0; j < N; j += P, par ^= 1) {
            int *pivslot = S.pivslot + par * P;
            unsigned char *isp = S.isp + par * 64;
            if (warp < 2) {
                // ---------------- phase 1a: factor the panel in registers ----------------
                // One row slot per lane, two warps; per column the warps exchange their best
                // candidate (key, label, slot, row) through shared memory and one 64-thread
                // barrier.
                const cplx *stg = S.stage + (size_t) (par ^ 1) * P * CW;     // rows that entered after the previous panel
                cplx a[P];
#pragma unroll
                for (int m = 0; m < P; ++m) {
                    int cs = jc + m; if (cs >= CW) cs -= CW;
                    a[m] = cplx(0.0, 0.0);
                    if (have) a[m] = pk < P ? stg[pk * CW + cs] : S.win[(size_t) fslot * CW + cs];
                }
                // a row that retired in the previous panel was replaced by row j+RW-P+k
                if (pk < P) { lg = j + RW - P + pk; pk = P; }
#pragma unroll
                for (int k = 0; k < P; ++k) {
                    const int col = j + k, hi = min(col + KL, N - 1);
                    // izamax over rows col..hi: compare the top 32 bits first; only a near tie
                    // needs the low word and the smallest-row rule
                    const bool cnd = liveb && pk == P && lg <= hi;
                    const long long key = cnd ? __double_as_longlong(cabs1(a[k])) : -1ll;
                    const int hi32 = (int) (key >> 32);
                    const int mh = __reduce_max_sync(0xffffffffu, hi32);
                    bool iswin = hi32 == mh && key >= 0;
                    unsigned bal = __ballot_sync(0xffffffffu, iswin);
                    if (bal & (bal - 1)) {                         // several lanes share the top word
                        const unsigned lo32 = (unsigned) (key & 0xffffffffll);
                        const unsigned ml = __reduce_max_sync(0xffffffffu, iswin ? lo32 : 0u);
                        iswin = iswin && lo32 == ml;
                        const int lmin = __reduce_min_sync(0xffffffffu, iswin ? lg : INT_MAX);
                        iswin = iswin && lg == lmin;
                        bal = __ballot_sync(0xffffffffu, iswin);
                    }
                    cplx *rec = S.xch + ((k & 1) * 2 + warp) * (1 + P);
                    if (bal == 0) {
                        if (lane == 0) *reinterpret_cast<long long *>(rec) = -1ll;
                    } else if (iswin) {
                        *reinterpret_cast<long long *>(rec) = key;
                        reinterpret_cast<int *>(rec)[2] = lg;
                        reinterpret_cast<int *>(rec)[3] = fslot;
#pragma unroll
                        for (int m = k; m < P; ++m) rec[1 + m] = a[m];
                    }
                    bar_sync_n<BAR_PANEL>(64);
                    const cplx *r0 = S.xch + ((k & 1) * 2 + 0) * (1 + P), *r1 = r0 + (1 + P);
                    const long long k0 = *reinterpret_cast<const long long *>(r0);
                    const long long k1 = *reinterpret_cast<const long long *>(r1);
                    const int l0 = reinterpret_cast<const int *>(r0)[2], l1 = reinterpret_cast<const int *>(r1)[2];
                    const bool second = k1 > k0 || (k1 == k0 && k1 >= 0 && l1 < l0);
                    const cplx *rw = second ? r1 : r0;
                    const int lwin = second ? l1 : l0;
                    const int wslot = reinterpret_cast<const int *>(rw)[3];
                    const int jp = lwin - col;
                    cplx pv[P];
#pragma unroll
                    for (int m = k; m < P; ++m) pv[m] = rw[1 + m];
                    // interchange = relabel: the slot holding row `col` takes the winner's label
                    if (liveb && pk == P && lg == col) lg = lwin;
                    if (have && fslot == wslot) { pk = k; lg = col; pivslot[k] = fslot; }
                    if (tid == 0) jpv[col] = (unsigned char) jp;
                    if ((second ? k1 : k0) == 0) { info = col + 1; break; }      // |re|+|im| == 0: zero pivot
                    ju = max(ju, min(col + KU + jp, N - 1));
                    const cplx rinv = recip_fast(pv[k]);
                    if (have && pk == P) {
                        const cplx l = a[k] * rinv; a[k] = l;
#pragma unroll
                        for (int m = k + 1; m < P; ++m) submul(a[m], l, pv[m]);
                        if (rhs) sv[col] = l;
                        else if (lg <= hi) S.lcol[k * KL + lg - (col + 1)] = l;      // L(lg, col)
                    }
                }
                // multipliers by slot; a pivot row keeps only the part below its own diagonal
                if (have) {
#pragma unroll
                    for (int m = 0; m < P; ++m) S.lp[fslot * P + m] = m < pk ? a[m] : cplx(0.0, 0.0);
                    isp[fslot] = pk < P;
                }
                if (tid == 0) { S.misc[2] = ju; S.misc[3] = info; }
            } else {
                // -------- phase 1b: refresh after the previous panel, assemble the next rows --------
                const int t0 = tid - 64, nt = NT - 64;
                if (j > 0) {
                    const int jo = j - P;                         // previous panel
                    const int *opiv = S.pivslot + (par ^ 1) * P;
                    const unsigned char *oisp = S.isp + (par ^ 1) * 64;
                    const cplx *stg = S.stage + (size_t) (par ^ 1) * P * CW;
                    int jco = jc - P; if (jco < 0) jco += CW;
                    // rows jo+RW .. jo+RW+P-1 take the slots of the retired pivot rows
                    for (int e = t0; e < P * CW; e += nt) {
                        const int k = e / CW, cs = e - k * CW;
                        S.win[(size_t) opiv[k] * CW + cs] = stg[e];
                    }
                    // columns jo+CW .. jo+CW+P-1 reuse the retired panel's column slots
                    for (int e = t0; e < NS * P; e += nt) {
                        const int s = e / P, m = e - s * P;
                        int cs = jco + m; if (cs >= CW) cs -= CW;
                        const int cn = jo + CW + m;
                        if (s == RW) S.win[(size_t) RW * CW + cs] = cn < N ? sv[cn] : cplx(0.0, 0.0);
                        else if (!oisp[s]) S.win[(size_t) s * CW + cs] = cplx(0.0, 0.0);
                    }
                }
                const int yI = (j + RW) / 5;                      // block entering after this panel
                compute_coef<W>(K, S, yI + 1 + K.ku, t0, nt);
                assemble_block<W>(K, S, km, kn, yI, S.stage + (size_t) par * P * CW, t0, nt);
            }
            bar_sync_n<BAR_ALL>(NT);
            info = S.misc[3];
            if (info) break;
            ju = S.misc[2];
            // ---------------- phase 2: rank-P update of the trailing columns ----------------
            {
                // this panel's multipliers, already in zgbtf2 order, to the slot's scratch
                for (int e = tid; e < P * KL; e += NT) Lg[(size_t) j * KL + e] = S.lcol[e];
                const int wtrail = ju - (j + P) + 1;
                int ps[P];
#pragma unroll
                for (int m = 0; m < P; ++m) ps[m] = pivslot[m];
                int cb = jc + P; if (cb >= CW) cb -= CW;
                for (int t = tid; t < W::NG * wtrail; t += NT) {
                    const int g = t / wtrail, c = t - g * wtrail;
                    int cs = cb + c; if (cs >= CW) cs -= CW;
                    cplx u[P];
#pragma unroll
                    for (int m = 0; m < P; ++m) u[m] = S.win[(size_t) ps[m] * CW + cs];
#pragma unroll
                    for (int k = 1; k < P; ++k)
#pragma unroll
                        for (int m = 0; m < k; ++m) submul(u[k], S.lp[ps[k] * P + m], u[m]);
#pragma unroll
                    for (int r = 0; r < W::RPG; ++r) {
                        const int s = g * W::RPG + r;
                        if (s < NS && !isp[s]) {
                            cplx w = S.win[(size_t) s * CW + cs];
#pragma unroll
                            for (int m = 0; m < P; ++m) submul(w, S.lp[s * P + m], u[m]);
                            S.win[(size_t) s * CW + cs] = w;
                        }
                    }
                }
            }
            bar_sync_n<BAR_ALL>(NT);
            jc += P; if (jc >= CW) jc -= CW;
        }
        if (tid == 0) { S.misc[buf] = info; if (info) for (int k = 0; k < N; ++k) jpv[k] = 0; }
        __threadfence();
        if (buf == 0) bar_arrive_n<BAR_FULL0>(W::NTH); else bar_arrive_n<BAR_FULL1>(W::NTH);
    }
}

template <class W>
int launch_blocked(const szb_imexop *op, BlockedArgs &A, int npencil, cudaStream_t stream)
{
    const int N = op->A.N;
    const size_t smem = blocked_smem_bytes<W>(N);
    if (smem > 227 * 1024) return 1;                 // caller falls back to the generic kernel
    static bool configured = false;
    if (!configured) {
        SZB_CUDA_OK(cudaFuncSetAttribute(invert_blocked_kernel<W>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured = true;
    }
    int per_sm = 0;
    SZB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, invert_blocked_kernel<W>,
                                                              W::NTH, smem));
    if (per_sm < 1) return 1;
    int slots = op->sm_count * per_sm;
    if (slots > npencil) slots = npencil;
    const size_t need = (size_t) slots * 2 * ((((size_t) N * W::KL) + 7) & ~(size_t) 7) * sizeof(cplx);
    if (need > op->work_bytes) {
        if (op->d_work) SZB_CUDA_OK(cudaFree(op->d_work));
        op->d_work = nullptr; op->work_bytes = 0;
        SZB_CUDA_OK(cudaMalloc(&op->d_work, need));
        op->work_bytes = need;
    }
    op->work_slots = slots;
    A.lwork = static_cast<cplx *>(op->d_work);
    invert_blocked_kernel<W><<<slots, W::NTH, smem, stream>>>(A);
    count_launch();
    SZB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace

// Returns 0 when launched, 1 when this (kl, ku) / size has no instantiation (the
// caller then uses another kernel), <0 on error.
int invert_blocked_dispatch(const szb_imexop *op, const double phi[2], int npencil,
                            const double *d_km, const double *d_kn, const int *d_index,
                            cplx *d_state, size_t fs, size_t ps, int *d_ipiv, int *d_info,
                            int *d_iters, cudaStream_t stream)
{
    BlockedArgs A;
    fill_pack_args(op, phi, d_km, d_kn, 0, 1, nullptr, A.pk);
    A.npencil = npencil; A.index = d_index;
    A.state = d_state; A.fs = fs; A.ps = ps;
    A.ipiv_out = d_ipiv; A.info_out = d_info; A.iters_out = d_iters;
    A.lwork = nullptr;
    if (op->A.KL != op->A.KU) return 1;
    switch (op->A.KL) {
    case 14: return launch_blocked<BlkCfg<14, 14, 8, 4, 6>>(op, A, npencil, stream);     // k = 4
    case 24: return launch_blocked<BlkCfg<24, 24, 16, 6, 6>>(op, A, npencil, stream);    // k = 6
    case 34: return launch_blocked<BlkCfg<34, 34, 16, 8, 6>>(op, A, npencil, stream);    // k = 8
    case 44: return launch_blocked<BlkCfg<44, 44, 32, 8, 8>>(op, A, npencil, stream);    // k = 10
    default: return 1;
    }
}

}  // namespace szb
