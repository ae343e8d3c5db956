// invert_pipe.cu -- batched invert of (M + phi L), version 4: the blocked banded LU of
// invert_blocked.cu with the panel factorisation taken off the critical path of the
// trailing update (two-stage software pipeline per panel).
//
// Replaces the hot loop of invert_mass_plus_scaled_operator
// (apps/perfect/operator_hybrid_isothermal.cpp:617-686) for the zgbsv solver
// specification: suzerain_rholut_imexop_packf (rholut_imexop.def:41-597) +
// IsothermalPATPTEnforcer::op/rhs (:470-525) + bsmbsm_solver::supply_B /
// zgbtrf + zgbtrs('T') / demand_X (bsmbsm_solver.cpp:155-182).  One persistent CTA per
// pencil slot; the matrix is assembled, factored and consumed on the SM and never
// touches HBM.
//
// Window, slot indirection, right-hand side row and solver warp are those of v3.  What
// changes is the schedule.  With j = 5 t the first column of panel t, one iteration is
//
//   panel warp   : L(t-1)  block t (columns j..j+4) of every row -= panel t-1's rank-5
//                          update, in registers (two row slots per lane); rows that
//                          entered after panel t-1 come straight from the stage
//                  F(t)    factor panel t in registers: per column one REDUX pivot
//                          search, the winner's row and its speculatively computed
//                          reciprocal broadcast by shuffles; multipliers to the scratch
//   update warps : U(t-1)  rank-5 update of columns j+5..ju(t-1) of the window
//                  R(t-1)  the rows entering after panel t-1 replace the retired pivot
//                          rows (all column slots but block t's), recycled column slots
//                          are zeroed
//                  A(t)    assemble the five rows entering after panel t into the stage
//   one CTA barrier
//
// so the serial chain per panel is L + F only; tools/pipelined_window_model.py is an
// executable model of the schedule that checks the two sides for shared-memory races.
//
// Arithmetic per element is the same sequence of FMAs as the unblocked zgbtf2 sweep;
// the pivot rule is izamax's (first maximum of |re|+|im|), so ipiv is LAPACK's.
#include <climits>
#include <cstdlib>
#include <cstdio>

#include "invert_fused.cuh"

// SZB_PIPE_SPLITU0 (make SPLITU0=1): the split block update of DESIGN 7.1a' -- correct (all GPU tests, identical
// pivots) but measured slower (23.1 vs 21.9 ms), so it is off by default.

// Optional phase timing (make PROF=1): per-phase clock64() deltas of the panel warp and of
// update warp 0, summed over all pencils and CTAs; read back with szb_debug_pipe_prof().
#ifdef SZB_PIPE_PROF
__device__ unsigned long long g_pipe_prof[16];
#define PROF_DECL long long pt0_ = clock64(); long long pacc_[6] = {0, 0, 0, 0, 0, 0};
#define PROF_MARK(i) do { const long long t_ = clock64(); pacc_[i] += t_ - pt0_; pt0_ = t_; } while (0)
#define PROF_FLUSH(base, who) do { if (who) for (int i_ = 0; i_ < 6; ++i_) atomicAdd(&g_pipe_prof[(base) + i_], (unsigned long long) pacc_[i_]); } while (0)
#else
#define PROF_DECL
#define PROF_MARK(i)
#define PROF_FLUSH(base, who)
#endif

namespace szb {

namespace {

using namespace fused;

template <int KL_, int KU_, int CR_, int NWU_, int NWA_, int RPG_, int MINB_>
struct PipeCfg {
    static constexpr int MINB = MINB_;              // CTAs per SM the register budget is sized for
    static constexpr int NWALLOC = (NWU_ + NWA_ + ((KL_ + 5 + 2) > 32 ? 2 : 1) + 1 + 3) / 4 * 4;      // warps are allocated in fours
    static constexpr int MAXR = (65536 / (MINB_ * 32 * NWALLOC)) / 8 * 8 > 255 ? 255 : (65536 / (MINB_ * 32 * NWALLOC)) / 8 * 8;
    static constexpr int KL = KL_, KU = KU_, KV = KL_ + KU_;
    static constexpr int RW = KL_ + P + 1;          // matrix row slots (one spare row keeps blocks aligned)
    static constexpr int NS = RW + 1;               // + the right-hand-side row (slot RW)
    static constexpr int NWP = NS > 32 ? 2 : 1;     // panel warps: one row slot per lane
    static constexpr int CW = KV + P + 1;           // column slots
    static constexpr int CR = CR_;                  // coefficient ring (collocation points), power of 2
    static constexpr int NWU = NWU_, NWA = NWA_;    // update warps, assembly warps; then the panel warps, the solver warp
    static constexpr int NTU = 32 * NWU_, NTA = 32 * NWA_, NT = NTU + NTA + 32 * NWP, NTH = NT + 32;
    static constexpr int RPG = RPG_;                // rows per trailing-update task
    static constexpr int NG = (NS + RPG_ - 1) / RPG_;
    static constexpr int NCOEF = 75;
    static constexpr int LDMAX = 20;                // >= ld of the B-spline operators (2k - 3 <= 17)
    static constexpr int CH = 4, NB = 3;            // solver: L columns per TMA chunk, ring depth
    // Physical warp w runs on sub-partition w % 4.  With twelve warps (6 update, 3 assembly,
    // 2 panel, 1 solver; logical order U0-5 A0-2 P0-1 S) the layout is
    //   sub-partition 0: P0 U0 U1   1: P1 U2 S   2: A0 A1 U3   3: A2 U4 U5
    // (measured: the assembly warps, not the update warps, slow a panel warp they share a
    // sub-partition with)
    __host__ __device__ static constexpr int logical_warp(int w)
    {
        if (NWU_ == 6 && NWA_ == 3 && NWP == 2) {
            return (int) ((0x53b1472086a9ull >> (4 * w)) & 0xf);      // {9, 10, 6, 8, 0, 2, 7, 4, 1, 11, 3, 5}
        }
        return w;
    }
    static_assert(RW % P == 0, "window rows come in groups of five");
    static_assert(NS <= 64, "at most two panel warps");
    static_assert((CR & (CR - 1)) == 0, "ring size must be a power of two");
};

template <class W>
struct PSmem {
    cplx *win;        // [NS][CW]        the window
    cplx *lp;         // [2][NS][P]      per panel parity: multipliers by slot (zeros past a pivot row's own step)
    cplx *rec;        // [2][2][8]       per column parity and panel warp: its pivot row record {1/pivot, row tail, label}
    cplx *stage;      // [2][P][CW]      assembled rows waiting to enter
    cplx *coef;       // [CR][75]        per-point block coefficients
    cplx *alpha;      // [MAXTERMS]
    cplx *v;          // [2][N]          b -> y -> x per buffer
    cplx *lring;      // [NB][CH*KL]     multipliers prefetched by TMA for the solver warp
    unsigned long long *mbar;   // [NB]
    int *pivslot;     // [2][P]
    int *misc;        // [0..1] info per buffer, [2..3] ju per panel parity, [4..5] panel info per buffer, [6..9] retired-slot mask per panel parity,
                      // [12..19] per column parity and panel warp {top word, ok}, [20..27] exact keys {key, row}
    unsigned char *ipiv;   // [2][N]     jp per column
    unsigned char *tref;   // [MAXTERMS]
    unsigned char *tblk;   // [76]
    double *drow;     // [2][3][LDMAX]   operator rows of the block assembled in iteration parity p
    double *refcol;   // [2][32]         reference-profile column for its new coefficient point
    double *sred;     // [8]             solver warp: the four reduced dot products of a chunk
};

// Everything whose size does not depend on N comes first, so that its shared-memory
// addresses are compile-time offsets; the two N-sized arrays (v, ipiv) close the block.
template <class W>
struct PipeLayout {
    static constexpr size_t C = sizeof(cplx);
    static constexpr size_t win = 0;
    static constexpr size_t lp = win + C * W::NS * W::CW;
    static constexpr size_t rec = lp + C * 2 * W::NS * P;
    static constexpr size_t stage = rec + C * 2 * 2 * 8;
    static constexpr size_t coef = stage + C * 2 * P * W::CW;
    static constexpr size_t alpha = coef + C * W::CR * W::NCOEF;
    static constexpr size_t lring = alpha + C * MAXTERMS;
    static constexpr size_t drow = lring + C * W::NB * W::CH * W::KL;
    static constexpr size_t refcol = drow + 8 * 2 * 3 * W::LDMAX;
    static constexpr size_t sred = refcol + 8 * 2 * 32;
    static constexpr size_t mbar = sred + 8 * 8;
    static constexpr size_t pivslot = mbar + 8 * W::NB;
    static constexpr size_t misc = pivslot + 4 * 2 * P;
    static constexpr size_t tref = misc + 4 * 32;
    static constexpr size_t tblk = tref + MAXTERMS;
    static constexpr size_t v = (tblk + 80 + 15) / 16 * 16;
    // vg: the b -> y -> x vectors live in global memory (large N), only ipiv follows
    __host__ __device__ static constexpr size_t ipiv(int N, bool vg = false) { return vg ? v : v + C * 2 * (size_t) N; }
    __host__ __device__ static constexpr size_t bytes(int N, bool vg = false) { return (ipiv(N, vg) + 2 * (size_t) N + 15) / 16 * 16; }
};

template <class W>
__host__ __device__ inline size_t pipe_smem_bytes(int N, bool vg = false) { return PipeLayout<W>::bytes(N, vg); }

template <class W>
__device__ __forceinline__ PSmem<W> pipe_carve(unsigned char *raw, int N, bool vg)
{
    using Y = PipeLayout<W>;
    PSmem<W> S;
    S.win = reinterpret_cast<cplx *>(raw + Y::win);
    S.lp = reinterpret_cast<cplx *>(raw + Y::lp);
    S.rec = reinterpret_cast<cplx *>(raw + Y::rec);
    S.stage = reinterpret_cast<cplx *>(raw + Y::stage);
    S.coef = reinterpret_cast<cplx *>(raw + Y::coef);
    S.alpha = reinterpret_cast<cplx *>(raw + Y::alpha);
    S.lring = reinterpret_cast<cplx *>(raw + Y::lring);
    S.drow = reinterpret_cast<double *>(raw + Y::drow);
    S.refcol = reinterpret_cast<double *>(raw + Y::refcol);
    S.sred = reinterpret_cast<double *>(raw + Y::sred);
    S.mbar = reinterpret_cast<unsigned long long *>(raw + Y::mbar);
    S.pivslot = reinterpret_cast<int *>(raw + Y::pivslot);
    S.misc = reinterpret_cast<int *>(raw + Y::misc);
    S.tref = raw + Y::tref;
    S.tblk = raw + Y::tblk;
    S.v = reinterpret_cast<cplx *>(raw + Y::v);
    S.ipiv = raw + Y::ipiv(N, vg);
    return S;
}

// Operator rows of row block yI and the reference-profile column of its new coefficient
// point yI + 1 + ku, into the shared-memory buffers of parity `par`.
template <class W, class SM>
__device__ __forceinline__ void stage_rowblock(const PackArgs &K, const SM &S, int yI, int par, int t0, int nt)
{
    const int nd = 3 * K.ld;
    for (int i = t0; i < nd + SZB_NREF + 1; i += nt) {
        if (i < nd) {
            const int d = i / K.ld, r = i - d * K.ld, yJ = yI - r + K.ku;
            S.drow[par * 3 * W::LDMAX + i] = (yJ >= 0 && yJ < K.n) ? __ldg(K.D + (size_t) (d * K.ld + r) * K.n + yJ) : 0.0;
        } else {
            const int q = i - nd, yc = yI + 1 + K.ku;
            S.refcol[par * 32 + q] = yc < K.n ? __ldg(K.refs + (size_t) q * K.n + yc) : 0.0;
        }
    }
}

// SZB_PIPE_SPLITU0 -- U0a(t): block t+1 against pivots 0..3 of panel t, once the panel warps have published
// them (they are in their last column step meanwhile).  Every thread redoes the small unit-lower-triangular
// fix-up of the four pivot rows for its column; rows q0..q3 are read only.  jc: column slot of block t.
template <class W, class SM>
__device__ __forceinline__ void early_block_update(const SM &S, unsigned sbase, int jc, int par, int t0, int nt)
{
    constexpr int NS = W::NS, CW = W::CW;
    const int *npiv = S.pivslot + par * P;
    const cplx *lpn = S.lp + (size_t) par * NS * P;
    const int q0 = npiv[0], q1 = npiv[1], q2 = npiv[2], q3 = npiv[3];
    for (int e = t0; e < NS * P; e += nt) {
        const int s = e / P, m = e - s * P;
        int ccs = jc + P + m; if (ccs >= CW) ccs -= CW;
        const cplx u0 = S.win[(size_t) q0 * CW + ccs];
        cplx u1 = S.win[(size_t) q1 * CW + ccs]; submul(u1, lpn[q1 * P], u0);
        cplx u2 = S.win[(size_t) q2 * CW + ccs]; submul(u2, lpn[q2 * P], u0); submul(u2, lpn[q2 * P + 1], u1);
        cplx u3 = S.win[(size_t) q3 * CW + ccs]; submul(u3, lpn[q3 * P], u0); submul(u3, lpn[q3 * P + 1], u1);
        submul(u3, lpn[q3 * P + 2], u2);
        cplx w = S.win[(size_t) s * CW + ccs];
        submul(w, lpn[s * P], u0); submul(w, lpn[s * P + 1], u1);
        submul(w, lpn[s * P + 2], u2); submul(w, lpn[s * P + 3], u3);
        sts_if(sbase + (unsigned) PipeLayout<W>::win + 16u * (unsigned) (s * CW + ccs), w,
               s != q0 && s != q1 && s != q2 && s != q3);
    }
}

// Start of iteration t > 0, all compute warps (the panel warp has nothing else to do until
// block t is final):
//   X(t-1)   the five pivot rows of panel t-1 become rows of U, in place, for every trailing
//            column j .. ju(t-1): one thread per column (unit lower triangular solve with
//            the pivot rows' own multipliers)
//   U0(t-1)  block t (columns j .. j+4) of every other row -= its multipliers times those
//            rows: one thread per (row slot, column)
// Each ends with a CTA barrier; afterwards the panel warp loads block t and factors it
// while the update warps do the rest of U(t-1).
template <class W, class SM>
__device__ __forceinline__ void lookahead_phases(const SM &S, unsigned sbase, int j, int jc, int par, int tid)
{
#ifdef SZB_PIPE_NOLOOK
    return;
#endif
    constexpr int NS = W::NS, CW = W::CW, NT = W::NT;
    const int *opiv = S.pivslot + (par ^ 1) * P;
    const cplx *lpo = S.lp + (size_t) (par ^ 1) * NS * P;
    int ps[P];
#pragma unroll
    for (int m = 0; m < P; ++m) ps[m] = opiv[m];
    const int wtot = S.misc[2 + (par ^ 1)] - j + 1;               // columns j .. ju(t-1)
#ifdef SZB_PIPE_SPLITU0
    // Block t already carries pivots 0..3 of panel t-1 (early_block_update during F(t-1)'s last column):
    // what is left is a rank-one update of block t with the last pivot row, which needs no fix-up, and the
    // in-place U rows of the columns past block t.  Disjoint columns: one barrier.
    {
        const unsigned long long omask = (unsigned) S.misc[6 + 2 * (par ^ 1)]
            | (unsigned long long) (unsigned) S.misc[7 + 2 * (par ^ 1)] << 32;
        constexpr int XT0 = ((NS * P + 31) / 32) * 32;             // first thread of the X' part (a warp boundary)
        static_assert(XT0 < NT, "threads left for the trailing columns");
        if (tid < XT0) {
            for (int e = tid; e < NS * P; e += XT0) {
                const int s = e / P, m = e - s * P;
                int ccs = jc + m; if (ccs >= CW) ccs -= CW;
                cplx w = S.win[(size_t) s * CW + ccs];
                submul(w, lpo[s * P + P - 1], S.win[(size_t) ps[P - 1] * CW + ccs]);
                sts_if(sbase + (unsigned) PipeLayout<W>::win + 16u * (unsigned) (s * CW + ccs), w, !((omask >> s) & 1));
            }
        } else {
            for (int c = P + tid - XT0; c < wtot; c += NT - XT0) {
                int ccs = jc + c; if (ccs >= CW) ccs -= CW;
                cplx u[P];
#pragma unroll
                for (int k = 0; k < P; ++k) u[k] = S.win[(size_t) ps[k] * CW + ccs];
#pragma unroll
                for (int k = 1; k < P; ++k) {
#pragma unroll
                    for (int i = 0; i < k; ++i) submul(u[k], lpo[ps[k] * P + i], u[i]);
                    S.win[(size_t) ps[k] * CW + ccs] = u[k];
                }
            }
        }
        bar_sync_n<1>(NT);
        return;
    }
#endif
    for (int c = tid; c < wtot; c += NT) {
        int ccs = jc + c; if (ccs >= CW) ccs -= CW;
        cplx u[P];
#pragma unroll
        for (int k = 0; k < P; ++k) u[k] = S.win[(size_t) ps[k] * CW + ccs];
#pragma unroll
        for (int k = 1; k < P; ++k) {
#pragma unroll
            for (int i = 0; i < k; ++i) submul(u[k], lpo[ps[k] * P + i], u[i]);
            S.win[(size_t) ps[k] * CW + ccs] = u[k];
        }
    }
    bar_sync_n<1>(NT);
    const unsigned long long omask = (unsigned) S.misc[6 + 2 * (par ^ 1)]
        | (unsigned long long) (unsigned) S.misc[7 + 2 * (par ^ 1)] << 32;
    for (int e = tid; e < NS * P; e += NT) {
        const int s = e / P, m = e - s * P;
        int ccs = jc + m; if (ccs >= CW) ccs -= CW;
        cplx w = S.win[(size_t) s * CW + ccs];
#pragma unroll
        for (int i = 0; i < P; ++i) submul(w, lpo[s * P + i], S.win[(size_t) ps[i] * CW + ccs]);
        sts_if(sbase + (unsigned) PipeLayout<W>::win + 16u * (unsigned) (s * CW + ccs), w, !((omask >> s) & 1));
    }
    bar_sync_n<1>(NT);
}

template <class W, bool VG>
__global__ void __maxnreg__(W::MAXR)
invert_pipe_kernel(const PipeArgs A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const PackArgs &K = A.pk;
    const int N = K.N, n = K.n;
    const PSmem<W> S = pipe_carve<W>(smem_raw, N, VG);
    cplx *const vbase = VG ? A.vwork + (size_t) blockIdx.x * 2 * N : S.v;
    // logical thread index: warps are dealt to the four SM sub-partitions round-robin, so the
    // roles are permuted to keep the latency-critical panel warps away from the FP64-heavy
    // update warps (see PipeCfg::logical_warp)
    int tid = W::logical_warp(threadIdx.x >> 5) * 32 + (threadIdx.x & 31);
    pin(tid);
    using Y = PipeLayout<W>;
    unsigned sbase = smem_u32(smem_raw);                          // shared-window address of the block
    pin(sbase);
    constexpr int KL = W::KL, KU = W::KU, RW = W::RW, NS = W::NS, CW = W::CW, NT = W::NT, NTU = W::NTU;
    constexpr int BAR_ALL = 1, BAR_UPD = 6, BAR_PP = 7;
    constexpr int BAR_C3 = 8;           // SZB_PIPE_SPLITU0: panel warps arrive once columns 0..3 are published
    constexpr int C3N = NTU + 32 * W::NWP;
    (void) BAR_C3; (void) C3N;
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
        solver_warp_run<W, VG, false>(A, S, vbase, lwork, lstride, S.ipiv, lane);
        return;
    }

    // ============================ compute warps ============================
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int ROLE_UPDATE = 0, ROLE_ASSEMBLE = 1, ROLE_PANEL = 2;
    const int role = warp < W::NWU ? ROLE_UPDATE : warp < W::NWU + W::NWA ? ROLE_ASSEMBLE : ROLE_PANEL;
    int q = 0;
    const int npen = A.count_dev ? min(A.npencil, __ldg(A.count_dev)) : A.npencil;
    for (int p = blockIdx.x; p < npen; p += gridDim.x, ++q) {
        const int buf = q & 1;
        if (q >= 2) { if (buf == 0) bar_sync_n<BAR_EMPTY0>(W::NTH); else bar_sync_n<BAR_EMPTY1>(W::NTH); }
        cplx *sv = vbase + (size_t) buf * N;
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
                if (K.with_bc && A.zero_wall_rhs && f < 4
                    && ((y == 0 && K.wall_begin == 0) || (y == n - 1 && K.wall_end == 2)))
                    val = cplx(0.0, 0.0);
                sv[5 * y + f] = val;
            }
        }
        for (int t = tid; t < K.terms->nterms; t += NT)
            S.alpha[t] = wave_factor(K.terms->wave[t], km, kn) * K.terms->sc[t];
        if (tid == 0) { S.misc[buf] = 0; S.misc[4 + buf] = 0; }
        bar_sync_n<BAR_ALL>(NT);
        for (int y = 0; y <= RW / 5 + K.ku; ++y) compute_coef<W>(K, S, y, tid, NT);
        bar_sync_n<BAR_ALL>(NT);
        // initial window: logical rows 0..RW-1 in slots 0..RW-1; RHS row t_c = b_c
        for (int blk = 0; blk < RW / 5; ++blk)
            assemble_block<W>(K, S, km, kn, blk, S.win + (size_t) blk * P * CW, tid, NT);
        for (int c = tid; c < CW; c += NT) S.win[(size_t) RW * CW + c] = c < N ? ldv<VG>(sv + c) : cplx(0.0, 0.0);
        // operator rows / profile column for the first block assembled inside the panel loop
        stage_rowblock<W>(K, S, RW / 5, 0, tid, NT);
        bar_sync_n<BAR_ALL>(NT);

        int info = 0;
        if (role == ROLE_PANEL) {
            // ====================== panel warps: F(t) ======================
            // Row slot lane + 32 pw lives in lane `lane` of panel warp pw.  lg: the logical row
            // the slot holds (INT_MAX for the right-hand side and for absent slots: never a
            // candidate); pk: P while the row is in play, 0..P-1 once it is this panel's pivot
            // row, -1 for absent slots.  a[0] is the column being eliminated; finished columns
            // are shifted out so that the column loop is one body (instruction-cache footprint).
            constexpr bool TWO = W::NWP == 2;
            const int pw = warp - (W::NWU + W::NWA);
            const int slot = lane + 32 * pw;
            int lg = slot < RW ? slot : INT_MAX;
            int pk = slot < NS ? P : -1;
            cplx a[P];
            unsigned rec_sa = sbase + (unsigned) Y::rec, sv_sa = sbase + (unsigned) (Y::v + sizeof(cplx) * (size_t) buf * N);
            unsigned jpv_sa = sbase + (unsigned) (Y::ipiv(N, VG) + (size_t) buf * N), xmh_sa = sbase + (unsigned) Y::misc + 48;
            pin(rec_sa); pin(sv_sa); pin(jpv_sa); pin(xmh_sa);
            // (b) lane predicates as pinned registers (the compiler otherwise re-reads SR_TID)
            int is_lead = tid == NTU + W::NTA, is_rhs = slot == RW, is_l0 = lane == 0;
            pin(is_lead); pin(is_rhs); pin(is_l0);
            int ju = 0, jc = 0, par = 0;
            PROF_DECL
            for (int j = 0; j < N; j += P, par ^= 1) {
                // block t of every row: entering rows from the stage, the others from the
                // window once panel t-1 has been applied to it
                if (j > 0) lookahead_phases<W>(S, sbase, j, jc, par, tid);
                PROF_MARK(0);
                {
                    const cplx *stg = S.stage + (size_t) (par ^ 1) * P * CW;
                    const bool ret = pk >= 0 && pk < P;
                    const cplx *src = ret ? stg + pk * CW : S.win + (size_t) (pk >= 0 ? slot : 0) * CW;
#pragma unroll
                    for (int m = 0; m < P; ++m) {
                        int c = jc + m; if (c >= CW) c -= CW;
                        a[m] = src[c];
                    }
                    // a row that retired in the previous panel was replaced by row j-P+RW+k
                    // (rows past the end of the matrix are never candidates)
                    if (ret) { lg = j - P + RW + pk; if (lg >= N) lg = INT_MAX; pk = P; }
                }
                PROF_MARK(1);
                unsigned lp_sa = sbase + (unsigned) (Y::lp + sizeof(cplx) * ((size_t) par * NS * P + slot * P));
                unsigned piv_sa = sbase + (unsigned) (Y::pivslot + 4 * par * P);
                pin(lp_sa); pin(piv_sa);
                cplx *Lcol = Lg + (size_t) j * KL - (j + 1);                // L(lg, col) at Lcol[lg]
                // Every lane inverts its candidate of the coming column while the current one is being
                // eliminated (1/z = conj(z) / |z|^2 for z comfortably scaled; anything else takes the
                // exact path): the seed and first residual at the end of a column step, the Newton
                // steps next to the pivot search of the following one.
                double rd_d = fma(a[0].x, a[0].x, a[0].y * a[0].y), rd_r = rcp_seed(rd_d);
                double rd_e = fma(-rd_d, rd_r, 1.0);
#ifdef SZB_PIPE_SPLITU0
                int c3 = 0;
#endif
#pragma unroll 1
                for (int k = 0; k < P; ++k) {
                    const int col = j + k, hi = col + KL;
                    // izamax over rows col..hi on the top 32 bits of |re|+|im|.  In each panel warp
                    // the lane whose candidate carries the warp's maximal top word publishes its row,
                    // the reciprocal pivot and its label straight away, and lane 0 the top word itself
                    // (+ whether it is unique and comfortably scaled: 0x22f00000 ~ 1e-140,
                    // 0x5d000000 ~ 1e+140); the warps then pick the larger one.  Anything else
                    // (ties on the top word, tiny / huge / zero pivots) takes the exact path.
                    const double mag = cabs1(a[0]);
                    cplx rs;
                    {
                        double e = fma(rd_e, rd_e, rd_e), r = fma(rd_r, e, rd_r);
                        e = fma(-rd_d, r, 1.0);
                        r = fma(r, e, r);
                        rs = cplx(a[0].x * r, -a[0].y * r);
                    }
                    const int h = (pk == P && lg <= hi) ? __double2hiint(mag) : -1;
                    const int mh = __reduce_max_sync(0xffffffffu, h);
                    const bool p = h == mh && h >= 0;
                    const unsigned bal = __ballot_sync(0xffffffffu, p);
                    const int okw = mh < 0 || ((bal & (bal - 1)) == 0 && mh >= 0x22f00000 && mh <= 0x5d000000);
                    const unsigned recw = rec_sa + ((k & 1) * 2 + pw) * 8 * (unsigned) sizeof(cplx);
                    sts_if(recw, rs, p);
#pragma unroll
                    for (int m = 1; m < P; ++m) sts_if(recw + m * 16, a[m], p);
                    sts2_if(recw + 5 * 16, lg, slot, p);
                    int g = 0, fast = okw && mh >= 0;
                    if (TWO) {
                        sts2_if(xmh_sa + ((k & 1) * 2 + pw) * 8, mh, okw, is_l0);
                        bar_sync_n<BAR_PP>(64);
                        const int2 x0 = lds_i2(xmh_sa + (k & 1) * 16), x1 = lds_i2(xmh_sa + (k & 1) * 16 + 8);
                        g = x1.x > x0.x;
                        fast = x0.x != x1.x && (g ? x1.y : x0.y);
                    } else {
                        __syncwarp();
                    }
                    bool zp = false;
                    if (!fast) {
                        // exact: first maximum of the full 64-bit |re|+|im|, smallest row on ties
                        const long long key = h >= 0 ? __double_as_longlong(mag) : -1ll;
                        const int e = exact_pivot(key, -1ll, lg, INT_MAX);
                        const int src = e & 0xff;
                        long long kwin = __shfl_sync(0xffffffffu, key, src);
                        int lgw = __shfl_sync(0xffffffffu, lg, src);
                        if ((e >> 16) & 1) { if (kwin < 0) lgw = INT_MAX; }      // no candidate here / zero
                        const unsigned xk_sa = xmh_sa + 32 + pw * 16;
                        if (TWO) {
                            if (lane == 0) {
                                asm volatile("st.shared.v2.b32 [%0], {%1, %2};" :: "r"(xk_sa), "r"((int) (kwin & 0xffffffffll)), "r"((int) (kwin >> 32)) : "memory");
                                asm volatile("st.shared.b32 [%0], %1;" :: "r"(xk_sa + 8), "r"(lgw) : "memory");
                            }
                        }
                        __syncwarp();
                        if (lane == src && kwin >= 0) {
                            const double2 r = exact_recip(a[0].x, a[0].y);
                            sts_if(recw, cplx(r.x, r.y), true);
#pragma unroll
                            for (int m = 1; m < P; ++m) sts_if(recw + m * 16, a[m], true);
                            sts2_if(recw + 5 * 16, lg, slot, true);
                        }
                        if (TWO) {
                            bar_sync_n<BAR_PP>(64);
                            const int2 ka = lds_i2(xmh_sa + 32), kb = lds_i2(xmh_sa + 48);
                            const int la = lds_i2(xmh_sa + 40).x, lb = lds_i2(xmh_sa + 56).x;
                            const long long k0 = (long long) (unsigned) ka.x | (long long) ka.y << 32;
                            const long long k1 = (long long) (unsigned) kb.x | (long long) kb.y << 32;
                            g = k1 > k0 || (k1 == k0 && lb < la);
                            zp = (g ? k1 : k0) <= 0;
                        } else {
                            __syncwarp();
                            zp = kwin <= 0;
                        }
                    }
                    const unsigned recg = rec_sa + ((k & 1) * 2 + g) * 8 * (unsigned) sizeof(cplx);
                    const cplx rinv = lds_c(recg);
                    cplx pv[P];
#pragma unroll
                    for (int m = 1; m < P; ++m) pv[m] = lds_c(recg + 16 * m);
                    const int2 lw = lds_i2(recg + 16 * 5);
                    int lwin = lw.x;
                    const int wslot = lw.y;
                    if (zp) lwin = col;
                    // interchange = relabel: the slot holding row `col` takes the winner's label
                    if (pk == P && lg == col) lg = lwin;
                    if (slot == wslot && !zp) { pk = k; lg = col; }
                    sts8_if(jpv_sa + col, lwin - col, is_lead);
                    sts32_if(piv_sa + 4 * k, wslot, is_lead);
                    if (zp) { info = col + 1; break; }             // |re|+|im| == 0: zero pivot
                    ju = max(ju, lwin + KU);
                    {
                        const bool act = pk == P;
                        cplx l = a[0] * rinv;
                        if (!act) l = cplx(0.0, 0.0);
                        // multipliers by slot (a pivot row keeps only the part below its own
                        // diagonal), in zgbtf2 order to the scratch, y = b^T U^-1 from the RHS row
                        sts_if(lp_sa + 16 * k, l, pk >= 0);
                        st_global_if(Lcol + lg, l, act && lg <= hi);
                        if (VG) st_global_if(sv + col, l, is_rhs); else sts_if(sv_sa + 16 * col, l, is_rhs);
#pragma unroll
                        for (int m = 1; m < P; ++m) {
                            cplx t = a[m];
                            submul(t, l, pv[m]);
                            a[m - 1] = t;
                        }
                        rd_d = fma(a[0].x, a[0].x, a[0].y * a[0].y);
                        rd_r = rcp_seed(rd_d);
                        rd_e = fma(-rd_d, rd_r, 1.0);
                    }
#ifdef SZB_PIPE_SPLITU0
                    if (k == P - 2) { bar_arrive_n<BAR_C3>(C3N); c3 = 1; }
#endif
                    Lcol += KL - 1;
                }
#ifdef SZB_PIPE_SPLITU0
                if (!c3) bar_arrive_n<BAR_C3>(C3N);        // a zero pivot cut the panel short
#endif
                PROF_MARK(2);
                {
                    const unsigned m = __ballot_sync(0xffffffffu, pk >= 0 && pk < P);
                    if (lane == 0) {
                        S.misc[6 + 2 * par + pw] = (int) m;
                        if (!TWO) S.misc[7 + 2 * par] = 0;
                        if (pw == 0) {
                            S.misc[2 + par] = min(ju, N - 1);
                            if (info) S.misc[4 + buf] = info;
                        }
                    }
                }
                PROF_MARK(3);
                bar_sync_n<BAR_ALL>(NT);
                PROF_MARK(4);
                if (info) break;
                jc += P; if (jc >= CW) jc -= CW;
            }
            PROF_FLUSH(0, lane == 0 && pw == 0);
        } else if (role == ROLE_UPDATE) {
            // ============ update warps: X/U0(t-1) with everybody, U(t-1), R(t-1) ============
            int jc = 0, par = 0;
            PROF_DECL
            for (int j = 0; j < N; j += P, par ^= 1) {
                if (j > 0) {
                    lookahead_phases<W>(S, sbase, j, jc, par, tid);
                    PROF_MARK(0);
                    const int jo = j - P;                         // previous panel
                    const int *opiv = S.pivslot + (par ^ 1) * P;
                    const cplx *lpo = S.lp + (size_t) (par ^ 1) * NS * P;
                    const cplx *stg = S.stage + (size_t) (par ^ 1) * P * CW;
                    const unsigned long long omask = (unsigned) S.misc[6 + 2 * (par ^ 1)]
                        | (unsigned long long) (unsigned) S.misc[7 + 2 * (par ^ 1)] << 32;
                    int ps[P];
#pragma unroll
                    for (int m = 0; m < P; ++m) ps[m] = opiv[m];
                    // ---- U(t-1): rank-P update of columns jo+2P .. ju(t-1); the pivot rows already
                    // hold rows of U.  One warp per row group, one lane per column; the next row's
                    // operands are fetched while the current one is updated. ----
                    const int wtrail = S.misc[2 + (par ^ 1)] - (jo + 2 * P) + 1;
                    int cb = jc + P; if (cb >= CW) cb -= CW;
#if !defined(SZB_PIPE_NOUPD) || SZB_PIPE_NOUPD == 2
                    for (int c0 = 0; c0 < wtrail; c0 += 32) {
                        const int c = c0 + lane;
                        if (c < wtrail) {
                            int ccs = cb + c; if (ccs >= CW) ccs -= CW;
                            cplx u[P];
#pragma unroll
                            for (int m = 0; m < P; ++m) u[m] = S.win[(size_t) ps[m] * CW + ccs];
                            for (int g = warp; g < W::NG; g += W::NWU) {
                                const int s0 = g * W::RPG;
                                cplx wn = S.win[(size_t) s0 * CW + ccs], ln[P];
#pragma unroll
                                for (int m = 0; m < P; ++m) ln[m] = lpo[s0 * P + m];
#pragma unroll
                                for (int r = 0; r < W::RPG; ++r) {
                                    const int s = s0 + r;
                                    cplx w = wn, l[P];
#pragma unroll
                                    for (int m = 0; m < P; ++m) l[m] = ln[m];
                                    if (r + 1 < W::RPG) {
                                        const int sn = min(s + 1, NS - 1);
                                        wn = S.win[(size_t) sn * CW + ccs];
#pragma unroll
                                        for (int m = 0; m < P; ++m) ln[m] = lpo[sn * P + m];
                                    }
#pragma unroll
                                    for (int m = 0; m < P; ++m) submul(w, l[m], u[m]);
                                    sts_if(sbase + (unsigned) Y::win + 16u * (unsigned) (min(s, NS - 1) * CW + ccs), w,
                                           s < NS && !((omask >> s) & 1));
                                }
                            }
                        }
                    }
#endif
                    PROF_MARK(1);
                    bar_sync_n<BAR_UPD>(NTU);
                    PROF_MARK(2);
                    // ---- R(t-1): rows jo+RW .. jo+RW+P-1 take the slots of the retired pivot rows;
                    // columns jo+CW .. jo+CW+P-1 reuse the retired panel's column slots ----
                    int jco = jc - P; if (jco < 0) jco += CW;
                    constexpr int RC = (P * CW + NTU - 1) / NTU, RZ = (NS * P + NTU - 1) / NTU;
                    cplx cp[RC];
#pragma unroll
                    for (int i = 0; i < RC; ++i) cp[i] = stg[min(tid + i * NTU, P * CW - 1)];
#pragma unroll
                    for (int i = 0; i < RZ; ++i) {
                        const int e = tid + i * NTU, s = e / P, m = e - s * P;
                        int ccs = jco + m; if (ccs >= CW) ccs -= CW;
                        const int cn = jo + CW + m;
                        if (e < NS * P) {
                            if (s == RW) S.win[(size_t) RW * CW + ccs] = cn < N ? ldv<VG>(sv + cn) : cplx(0.0, 0.0);
                            else if (!((omask >> s) & 1)) S.win[(size_t) s * CW + ccs] = cplx(0.0, 0.0);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < RC; ++i) {
                        const int e = tid + i * NTU, k = e / CW, ccs = e - k * CW;
                        if (e < P * CW) S.win[(size_t) opiv[k] * CW + ccs] = cp[i];
                    }
                    PROF_MARK(3);
                }
#ifdef SZB_PIPE_SPLITU0
                // ---- U0a(t): block t+1 against pivots 0..3 of THIS panel, as soon as the panel warps have
                // published them (they are in their last column step meanwhile).  Every thread redoes the
                // small unit-lower-triangular fix-up of the four pivot rows for its column. ----
                bar_sync_n<BAR_C3>(C3N);
                early_block_update<W>(S, sbase, jc, par, tid, NTU);
#endif
                bar_sync_n<BAR_ALL>(NT);
                PROF_MARK(5);
                info = S.misc[4 + buf];
                if (info) break;
                jc += P; if (jc >= CW) jc -= CW;
            }
            PROF_FLUSH(8, tid == 0);
        } else {
            // ============ assembly warps: X/U0(t-1) with everybody, A(t) ============
            const int ta = tid - NTU;
            int jc = 0, par = 0;
            for (int j = 0; j < N; j += P, par ^= 1) {
                if (j > 0) lookahead_phases<W>(S, sbase, j, jc, par, tid);
                // ---- A(t): the block entering after this panel, from the staged operator rows
                // and profile column; then stage the next iteration's ----
                const int yI = (j + RW) / 5;
                compute_coef_staged<W>(K, S, yI + 1 + K.ku, S.refcol + par * 32, ta, W::NTA);
                cplx *dst = S.stage + (size_t) par * P * CW;
#if !defined(SZB_PIPE_NOUPD) || SZB_PIPE_NOUPD == 1
                if (yI - K.kl >= 1 && yI + K.ku <= n - 2)
                    assemble_block_interior<W>(K, S, S.drow + par * 3 * W::LDMAX, yI, dst, ta, W::NTA);
                else
                    assemble_block<W>(K, S, DStaged(K, S.drow + par * 3 * W::LDMAX, yI), km, kn, yI, dst, ta, W::NTA);
#endif
                stage_rowblock<W>(K, S, yI + 1, par ^ 1, ta, W::NTA);
                bar_sync_n<BAR_ALL>(NT);
                info = S.misc[4 + buf];
                if (info) break;
                jc += P; if (jc >= CW) jc -= CW;
            }
        }
        if (tid == 0) { S.misc[buf] = info; if (info) for (int k = 0; k < N; ++k) jpv[k] = 0; }
        __threadfence();
        if (buf == 0) bar_arrive_n<BAR_FULL0>(W::NTH); else bar_arrive_n<BAR_FULL1>(W::NTH);
    }
}

template <class W, bool VG>
int launch_pipe_vg(const szb_imexop *op, PipeArgs &A, int npencil, cudaStream_t stream)
{
    const int N = op->A.N;
    const size_t smem = pipe_smem_bytes<W>(N, VG);
    if (smem > 227 * 1024) return 1;                 // caller falls back to another kernel
    static bool configured = false;
    if (!configured) {
        SZB_CUDA_OK(cudaFuncSetAttribute(invert_pipe_kernel<W, VG>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured = true;
    }
    int per_sm = 0;
    SZB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, invert_pipe_kernel<W, VG>,
                                                              W::NTH, smem));
    if (per_sm < 1) return 1;
    static const bool debug = std::getenv("SZB_PIPE_DEBUG") != nullptr;
    if (debug)
        std::fprintf(stderr, "invert_pipe: KL=%d N=%d vg=%d threads=%d smem=%zu CTAs/SM=%d\n", W::KL, N, (int) VG, W::NTH, smem, per_sm);
    int slots = op->sm_count * per_sm;
    if (slots > npencil) slots = npencil;
    const size_t lbytes = (size_t) slots * 2 * ((((size_t) N * W::KL) + 7) & ~(size_t) 7) * sizeof(cplx);
    const size_t vbytes = VG ? (size_t) slots * 2 * N * sizeof(cplx) : 0;
    const size_t need = lbytes + vbytes;
    if (need > op->work_bytes) {
        if (op->d_work) SZB_CUDA_OK(cudaFree(op->d_work));
        op->d_work = nullptr; op->work_bytes = 0;
        SZB_CUDA_OK(cudaMalloc(&op->d_work, need));
        op->work_bytes = need;
    }
    op->work_slots = slots;
    A.lwork = static_cast<cplx *>(op->d_work);
    A.vwork = VG ? reinterpret_cast<cplx *>(static_cast<unsigned char *>(op->d_work) + lbytes) : nullptr;
    invert_pipe_kernel<W, VG><<<slots, W::NTH, smem, stream>>>(A);
    count_launch();
    SZB_CUDA_OK(cudaGetLastError());
    return 0;
}

// The b -> y -> x vectors (2 N complex per CTA) stay in shared memory while MINB CTAs still fit
// an SM with them; beyond that (Ny >= 256 at k = 8, any Ny >= 96 at k = 6 with three CTAs per SM)
// they move to a global scratch read at L2.
template <class W>
int launch_pipe(const szb_imexop *op, PipeArgs &A, int npencil, cudaStream_t stream)
{
    // shared memory available to each of the MINB CTAs the register budget is sized for
    const size_t per_cta = (233472 - (size_t) W::MINB * 1024) / W::MINB;
    static const bool allow = [] { const char *e = std::getenv("SZB_PIPE_VG"); return !(e && e[0] == '0'); }();
    if (allow && W::MINB >= 2 && pipe_smem_bytes<W>(op->A.N, false) > per_cta
        && pipe_smem_bytes<W>(op->A.N, true) <= per_cta)
        return launch_pipe_vg<W, true>(op, A, npencil, stream);
    return launch_pipe_vg<W, false>(op, A, npencil, stream);
}

// ---------------------------------------------------------------------------
// Iterative refinement around the fused solve (the zcgbsvx specification with its default
// eps tolerance, suzerain/blas_et_al/dsgbsvx.def:131-318 for siter < 0): the factors are
// never stored, so a refinement step is one more fused factor + solve with the residual as
// right hand side.  residual_kernel forms x += d, r = b - (P A^T P^T)^T x and |r|_2 per pencil
// with the operator re-assembled on the fly, and applies the reference's stopping rules.
// ---------------------------------------------------------------------------
struct ResidualArgs {
    PackArgs pk;                    // km / kn indexed by pencil
    int nlist; const int *pos;      // pencils to process (null: all npencil)
    const int *count_dev;           // optional device-side length of the list (<= nlist)
    const int *index;               // pencil -> slot of the state
    cplx *x; size_t fs, ps;         // solution, state layout
    const cplx *b;                  // [npencil][5][n] right hand sides (wall rows still to be zeroed)
    cplx *r;                        // [npencil][5][n] in: correction d (if add), out: residual
    int add, it, aiter, dmax;
    int mode;                       // 0: zcgbsvx (2-norm, stagnation), 1: zgbrfs (componentwise backward error)
    double tol;
    double *res, *lastres; int *diter, *cont;
};

template <class W>
struct RSmem { cplx *coef, *alpha; unsigned char *tref, *tblk; };

template <class W>
__global__ void __launch_bounds__(256)
residual_kernel(const ResidualArgs A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double s_red[8];
    const PackArgs &K = A.pk;
    const int N = K.N, n = K.n, tid = threadIdx.x;
    constexpr int CW = W::CW, KL = W::KL;
    RSmem<W> S;
    cplx *q = reinterpret_cast<cplx *>(smem_raw);
    S.coef = q; q += W::CR * W::NCOEF;
    S.alpha = q; q += MAXTERMS;
    cplx *stage = q; q += P * CW;
    cplx *xs = q; q += N;
    cplx *rs = q; q += N;
    double *ra = reinterpret_cast<double *>(q); q += (N + 1) / 2;      // zgbrfs: |b| + |A^T| |x|
    S.tref = reinterpret_cast<unsigned char *>(q);
    S.tblk = S.tref + MAXTERMS;
    for (int t = tid; t < MAXTERMS; t += 256) S.tref[t] = K.terms->ref[t];
    for (int t = tid; t <= NBLOCK; t += 256) S.tblk[t] = K.terms->blk_begin[t];
    __syncthreads();
    const int nlist = A.count_dev ? min(A.nlist, __ldg(A.count_dev)) : A.nlist;
    for (int e = blockIdx.x; e < nlist; e += gridDim.x) {
        const int p = A.pos ? A.pos[e] : e;
        const double km = K.km[p], kn = K.kn[p];
        cplx *x = A.x + (A.index ? (size_t) A.index[p] : (size_t) p) * A.ps;
        const cplx *b = A.b + (size_t) p * N;
        cplx *r = A.r + (size_t) p * N;
        for (int t = tid; t < K.terms->nterms; t += 256)
            S.alpha[t] = wave_factor(K.terms->wave[t], km, kn) * K.terms->sc[t];
        for (int k = tid; k < N; k += 256) {
            const int f = k / n, y = k - f * n;
            cplx xv = x[(size_t) f * A.fs + y];
            if (A.add) { xv += r[k]; x[(size_t) f * A.fs + y] = xv; }
            xs[5 * y + f] = xv;
            cplx bv = b[k];
            if (K.with_bc && f < 4 && ((y == 0 && K.wall_begin == 0) || (y == n - 1 && K.wall_end == 2)))
                bv = cplx(0.0, 0.0);
            rs[5 * y + f] = bv;
            ra[5 * y + f] = cabs1(bv);
        }
        __syncthreads();
        for (int y = 0; y <= K.ku; ++y) compute_coef<W>(K, S, y, tid, 256);
        __syncthreads();
        for (int yI = 0; yI < n; ++yI) {
            compute_coef<W>(K, S, yI + K.ku + 1, tid, 256);
            assemble_block<W>(K, S, km, kn, yI, stage, tid, 256);
            __syncthreads();
            // r_J -= sum_I (P A^T P^T)[I, J] x_I over the five rows I of this block; column J is
            // always handled by thread J mod CW
            if (tid < CW) {
#pragma unroll
                for (int sI = 0; sI < P; ++sI) {
                    const int I = 5 * yI + sI, J0 = I - KL;
                    int ci = (tid - J0) % CW; if (ci < 0) ci += CW;
                    const int J = J0 + ci;
                    if (ci <= W::KV && J >= 0 && J < N) {
                        const cplx a = stage[sI * CW + tid], xv = xs[I];
                        submul(rs[J], a, xv);
                        if (A.mode) ra[J] += cabs1(a) * cabs1(xv);
                    }
                }
            }
            __syncthreads();
        }
        double s2 = 0.0;
        // zgbrfs.f: safe1 = nz safmin, safe2 = safe1 / eps guard tiny denominators
        const double safe1 = min(W::KL + W::KU + 2, N + 1) * 2.2250738585072014e-308, safe2 = safe1 / 1.1102230246251565e-16;
        for (int k = tid; k < N; k += 256) {
            const int f = k / n, y = k - f * n;
            const cplx v = rs[5 * y + f];
            r[k] = v;
            if (A.mode) {
                const double den = ra[5 * y + f], num = cabs1(v);
                s2 = fmax(s2, den > safe2 ? num / den : (num + safe1) / (den + safe1));
            } else s2 += v.x * v.x + v.y * v.y;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double other = __shfl_xor_sync(0xffffffffu, s2, o);
            s2 = A.mode ? fmax(s2, other) : s2 + other;
        }
        if ((tid & 31) == 0) s_red[tid >> 5] = s2;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int w = 0; w < 8; ++w) t = A.mode ? fmax(t, s_red[w]) : t + s_red[w];
            if (A.mode) {
                // zgbrfs.f: go on while berr > eps, berr at least halved, at most ITMAX = 5 corrections
                const double berr = t;
                const bool go = berr > 1.1102230246251565e-16 && 2.0 * berr <= A.lastres[p] && A.it < 5;
                A.diter[p] = A.it;
                A.res[p] = berr;
                if (go) A.lastres[p] = berr;
                A.cont[p] = go;
            } else {
                const double res = sqrt(t);
                // dsgbsvx.def:271-284: stagnation once diter >= aiter, else keep the residual
                const bool stop = A.it >= A.aiter && A.lastres[p] < 2.0 * res;
                A.diter[p] = A.it;
                A.res[p] = res;
                if (!stop) A.lastres[p] = res;
                A.cont[p] = !stop && A.it < A.dmax && res > A.tol;
            }
        }
        __syncthreads();
    }
}

// ---- the same residual through accumulate_kernel (imexop.cu): r = b - (M + phi L) x is the operator
// applied to the solution.  accumulate writes (M + phi L) x - b into R (beta = -1 on a copy of b); this
// kernel then replaces the <= 8 wall rows by the boundary equations of IsothermalPATPTEnforcer
// (s x_i - s factor x_rho = 0 with s the assembled diagonal, operator_hybrid_isothermal.cpp:448-506),
// flips the sign, takes |r|_2 and applies the stopping rules of dsgbsvx.def:271-284.
template <class W>
__global__ void __launch_bounds__(128)
residual_fix_kernel(const ResidualArgs A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double s_red[4];
    __shared__ cplx s_wall[8];
    const PackArgs &K = A.pk;
    const int N = K.N, n = K.n, tid = threadIdx.x;
    RSmem<W> S;
    cplx *q = reinterpret_cast<cplx *>(smem_raw);
    S.coef = q; q += W::CR * W::NCOEF;
    S.alpha = q; q += MAXTERMS;
    S.tref = reinterpret_cast<unsigned char *>(q);
    S.tblk = S.tref + MAXTERMS;
    for (int t = tid; t < MAXTERMS; t += 128) S.tref[t] = K.terms->ref[t];
    for (int t = tid; t <= NBLOCK; t += 128) S.tblk[t] = K.terms->blk_begin[t];
    __syncthreads();
    const int nlist = A.count_dev ? min(A.nlist, __ldg(A.count_dev)) : A.nlist;
    for (int e = blockIdx.x; e < nlist; e += gridDim.x) {
        const int p = A.pos ? A.pos[e] : e;
        const double km = K.km[p], kn = K.kn[p];
        const cplx *x = A.x + (A.index ? (size_t) A.index[p] : (size_t) p) * A.ps;
        cplx *r = A.r + (size_t) p * N;
        if (K.with_bc) {
            for (int t = tid; t < K.terms->nterms; t += 128)
                S.alpha[t] = wave_factor(K.terms->wave[t], km, kn) * K.terms->sc[t];
            __syncthreads();
            for (int wall = 0; wall < 2; ++wall) {
                const bool on = wall == 0 ? K.wall_begin == 0 : K.wall_end == 2;
                const int yw = wall == 0 ? 0 : n - 1;
                if (on) {
                    // the diagonal entries of the wall point's four equations need its coefficients and its
                    // neighbours' only through D^(d)[yw, yw]: base_entry reads coef at yJ = yw
                    compute_coef<W>(K, S, yw, tid, 128);
                    __syncthreads();
                    if (tid < 4) {
                        const int J = 5 * yw + tid;
                        cplx sd = nrbc_entry<W>(K, S.coef, DGlobal(K), km, kn, J, J);
                        if (is_zero(sd)) sd = cplx(1.0, 0.0);
                        const double factor = tid == 0 ? K.E_factor[wall] : K.vel_factor[wall][tid - 1];
                        const cplx xi = x[(size_t) tid * A.fs + yw], xr = x[(size_t) 4 * A.fs + yw];
                        // (M + phi L) x - b on this row of the modified system, b = 0
                        s_wall[4 * wall + tid] = sd * (xi - xr * factor);
                    }
                    __syncthreads();
                    if (tid < 4) r[(size_t) tid * n + yw] = s_wall[4 * wall + tid];
                }
            }
            __syncthreads();
        }
        double s2 = 0.0;
        for (int k = tid; k < N; k += 128) {
            const cplx v = -r[k];
            r[k] = v;
            s2 += v.x * v.x + v.y * v.y;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        if ((tid & 31) == 0) s_red[tid >> 5] = s2;
        __syncthreads();
        if (tid == 0) {
            const double res = sqrt(s_red[0] + s_red[1] + s_red[2] + s_red[3]);
            const bool stop = A.it >= A.aiter && A.lastres[p] < 2.0 * res;
            A.diter[p] = A.it;
            A.res[p] = res;
            if (!stop) A.lastres[p] = res;
            A.cont[p] = !stop && A.it < A.dmax && res > A.tol;
        }
        __syncthreads();
    }
}

// x += d for the pencils that go on, and their R <- b for the next accumulate
__global__ void refine_update_kernel(int nlist, const int *count_dev, const int *pos, int N, int n, const int *index, cplx *x, size_t fs, size_t ps,
                                     const cplx *b, cplx *r, int add)
{
    if (count_dev && (int) blockIdx.x >= __ldg(count_dev)) return;
    const int p = pos ? pos[blockIdx.x] : blockIdx.x;
    cplx *xv = x + (index ? (size_t) index[p] : (size_t) p) * ps;
    for (int k = threadIdx.x; k < N; k += blockDim.x) {
        const int f = k / n, y = k - f * n;
        if (add) xv[(size_t) f * fs + y] += r[(size_t) p * N + k];
        r[(size_t) p * N + k] = b[(size_t) p * N + k];
    }
}

__global__ void refine_gather_kernel(int npencil, int N, int n, const int *index, const cplx *state,
                                     size_t fs, size_t ps, cplx *b, double *lastres, double lastres0)
{
    const int p = blockIdx.x;
    const cplx *v = state + (index ? (size_t) index[p] : (size_t) p) * ps;
    double s2 = 0.0;
    for (int k = threadIdx.x; k < N; k += blockDim.x) {
        const int f = k / n, y = k - f * n;
        const cplx val = v[(size_t) f * fs + y];
        b[(size_t) p * N + k] = val;
        s2 += val.x * val.x + val.y * val.y;
    }
    (void) s2; (void) npencil;
    if (threadIdx.x == 0) lastres[p] = lastres0;
}

__global__ void refine_compact_kernel(int npencil, const int *cont, const int *info, const double *km,
                                      const double *kn, int *pos, double *kma, double *kna, int *count,
                                      const int *index, int *slot)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npencil || !cont[p] || info[p] != 0) return;
    const int i = atomicAdd(count, 1);
    pos[i] = p; kma[i] = km[p]; kna[i] = kn[p];
    if (slot) slot[i] = index ? index[p] : p;
}

__global__ void refine_finish_kernel(int npencil, const int *info, const int *diter, int *iters)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < npencil && iters) iters[p] = info[p] ? -1 : diter[p];
}

template <class W>
int launch_residual(const szb_imexop *op, ResidualArgs &A, cudaStream_t stream)
{
    const size_t smem = sizeof(cplx) * ((size_t) W::CR * W::NCOEF + MAXTERMS + P * W::CW + 2 * (size_t) op->A.N
                                       + ((size_t) op->A.N + 1) / 2) + MAXTERMS + 96;
    if (smem > 200 * 1024) return 1;
    static size_t configured = 0;
    if (smem > configured) {
        SZB_CUDA_OK(cudaFuncSetAttribute(residual_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        configured = smem;
    }
    int grid = A.nlist < op->sm_count * 4 ? A.nlist : op->sm_count * 4;
    if (grid < 1) return 0;
    residual_kernel<W><<<grid, 256, smem, stream>>>(A);
    count_launch();
    SZB_CUDA_OK(cudaGetLastError());
    return 0;
}

int dispatch_residual(const szb_imexop *op, ResidualArgs &A, cudaStream_t stream)
{
    switch (op->A.KL) {
    case 14: return launch_residual<PipeCfg<14, 14, 8, 3, 1, 7, 3>>(op, A, stream);
    case 24: return launch_residual<PipeCfg<24, 24, 16, 4, 2, 8, 3>>(op, A, stream);
    case 34: return launch_residual<PipeCfg<34, 34, 16, 6, 3, 7, 2>>(op, A, stream);
    case 44: return launch_residual<PipeCfg<44, 44, 32, 6, 2, 9, 1>>(op, A, stream);
    default: return 1;
    }
}

template <class W>
int launch_residual_fix(const szb_imexop *op, ResidualArgs &A, cudaStream_t stream)
{
    const size_t smem = sizeof(cplx) * ((size_t) W::CR * W::NCOEF + MAXTERMS) + MAXTERMS + 96;
    int grid = A.nlist < op->sm_count * 16 ? A.nlist : op->sm_count * 16;
    if (grid < 1) return 0;
    residual_fix_kernel<W><<<grid, 128, smem, stream>>>(A);
    count_launch();
    SZB_CUDA_OK(cudaGetLastError());
    return 0;
}

int dispatch_residual_fix(const szb_imexop *op, ResidualArgs &A, cudaStream_t stream)
{
    switch (op->A.KL) {
    case 14: return launch_residual_fix<PipeCfg<14, 14, 8, 3, 1, 7, 3>>(op, A, stream);
    case 24: return launch_residual_fix<PipeCfg<24, 24, 16, 4, 2, 8, 3>>(op, A, stream);
    case 34: return launch_residual_fix<PipeCfg<34, 34, 16, 6, 3, 7, 2>>(op, A, stream);
    case 44: return launch_residual_fix<PipeCfg<44, 44, 32, 6, 2, 9, 1>>(op, A, stream);
    default: return 1;
    }
}

}  // namespace

// Returns 0 when launched, 1 when this (kl, ku) / size has no instantiation (the
// caller then uses another kernel), <0 on error.
int invert_pipe_dispatch(const szb_imexop *op, const double phi[2], int npencil,
                         const double *d_km, const double *d_kn, const int *d_index,
                         cplx *d_state, size_t fs, size_t ps, int *d_ipiv, int *d_info,
                         int *d_iters, cudaStream_t stream, int zero_wall_rhs, const int *d_count)
{
    PipeArgs A;
    fill_pack_args(op, phi, d_km, d_kn, 0, 1, nullptr, A.pk);
    A.npencil = npencil; A.index = d_index; A.count_dev = d_count;
    A.state = d_state; A.fs = fs; A.ps = ps;
    A.ipiv_out = d_ipiv; A.info_out = d_info; A.iters_out = d_iters;
    A.lwork = nullptr; A.vwork = nullptr; A.ipwork = nullptr; A.xwork = nullptr;
    A.zero_wall_rhs = zero_wall_rhs;
    if (op->A.KL != op->A.KU) return 1;
    switch (op->A.KL) {
    case 14: return launch_pipe<PipeCfg<14, 14, 8, 3, 1, 7, 3>>(op, A, npencil, stream);     // k = 4
    case 24: return launch_pipe<PipeCfg<24, 24, 16, 4, 2, 8, 3>>(op, A, npencil, stream);    // k = 6
    case 34: return launch_pipe<PipeCfg<34, 34, 16, 6, 3, 7, 2>>(op, A, npencil, stream);    // k = 8
    case 44: return launch_pipe<PipeCfg<44, 44, 32, 6, 2, 9, 1>>(op, A, npencil, stream);    // k = 10
    default: return 1;
    }
}


// zcgbsvx with the default eps tolerance (tolsc == 0, no single-precision attempt; mode 0) or
// zgbsvx without equilibration (zgbtrs + zgbrfs refinement; mode 1) on top of the fused kernel.  Returns 0 when done, 1 when not applicable (the caller then uses the generic
// kernel), < 0 on error.  Synchronises the stream once per refinement step (active-list count).
int invert_refined_dispatch(const szb_imexop *op, int mode, int aiter, int dmax, const double phi[2], int npencil,
                            const double *d_km, const double *d_kn, const int *d_index,
                            cplx *d_state, size_t fs, size_t ps, int *d_ipiv, int *d_info,
                            int *d_iters, cudaStream_t stream)
{
    return invert_refined_stage(op, mode, aiter, dmax, phi, npencil, d_km, d_kn, d_index, d_state, fs, ps, d_ipiv, d_info,
                                d_iters, stream, 3, 0, npencil);
}

// The same in two stages, for a caller that feeds the pencils in pieces (the whole-field host entry point uploads
// wave space chunk by chunk): stage 1 = right-hand sides saved + first solve of `npencil` pencils that occupy
// positions pos0 .. pos0 + npencil - 1 of a list of `capacity` pencils; stage 2 = the refinement loop over positions
// 0 .. npencil - 1 (pos0 = 0), whose first solves have all been issued.  The pointer arguments always describe the
// pencils of THIS call.  One refinement over the union of several chunks costs what one over a single chunk does
// (its passes over the few pencils that go on are latency-bound).
int invert_refined_stage(const szb_imexop *op, int mode, int aiter, int dmax, const double phi[2], int npencil,
                         const double *d_km, const double *d_kn, const int *d_index,
                         cplx *d_state, size_t fs, size_t ps, int *d_ipiv, int *d_info,
                         int *d_iters, cudaStream_t stream, int stages, int pos0, int capacity)
{
    // applicability first: nothing may have been launched when the caller is told to fall back
    if (op->A.KL != op->A.KU || dmax < 0) return 1;
    if (!(op->A.KL == 14 || op->A.KL == 24 || op->A.KL == 34 || op->A.KL == 44)) return 1;   // the fused kernels' orders
    if (!mode && aiter < 1) return 1;       // lastres starts at 3: exact for aiter >= 1 (dsgbsvx.def:271-284)
    if (pos0 < 0 || pos0 + npencil > capacity || ((stages & 2) && pos0 != 0)) return -4;
    const int N = op->A.N, n = op->n;
    // workspace: b, r | res, lastres, kma, kna | diter, cont, pos, info2, count
    const size_t nb = (size_t) capacity * N * sizeof(cplx);
    const size_t nd = (((size_t) capacity * sizeof(double)) + 15) & ~(size_t) 15;
    const size_t ni = (((size_t) capacity * sizeof(int)) + 15) & ~(size_t) 15;
    const size_t need = 2 * nb + 4 * nd + 5 * ni + 16;
    if (need > op->refine_bytes) {
        if (op->d_refine) SZB_CUDA_OK(cudaFree(op->d_refine));
        op->d_refine = nullptr; op->refine_bytes = 0;
        SZB_CUDA_OK(cudaMalloc(&op->d_refine, need));
        op->refine_bytes = need;
    }
    unsigned char *w = static_cast<unsigned char *>(op->d_refine);
    cplx *B = reinterpret_cast<cplx *>(w); w += nb;
    cplx *R = reinterpret_cast<cplx *>(w); w += nb;
    double *res = reinterpret_cast<double *>(w); w += nd;
    double *lastres = reinterpret_cast<double *>(w); w += nd;
    double *kma = reinterpret_cast<double *>(w); w += nd;
    double *kna = reinterpret_cast<double *>(w); w += nd;
    int *diter = reinterpret_cast<int *>(w); w += ni;
    int *cont = reinterpret_cast<int *>(w); w += ni;
    int *pos = reinterpret_cast<int *>(w); w += ni;
    int *info2 = reinterpret_cast<int *>(w); w += ni;
    int *slot2 = reinterpret_cast<int *>(w); w += ni;     // state slot of each pencil of the active list
    int *count = reinterpret_cast<int *>(w);

    if (mode) { aiter = 1; dmax = 5; }                    // zgbrfs: ITMAX
    int rc = 0;
    if (stages & 1) {
        refine_gather_kernel<<<npencil, 128, 0, stream>>>(npencil, N, n, d_index, d_state, fs, ps, B + (size_t) pos0 * N,
                                                          lastres + pos0, mode ? 3.0 : 0.0);
        count_launch();
        // first pass: x = 0 + (LU)^-T b, in place in the state
        rc = invert_fused_dispatch(op, phi, npencil, d_km, d_kn, d_index, d_state, fs, ps, d_ipiv, d_info,
                                   nullptr, stream, 1);
        if (rc) return rc;
    }
    if (!(stages & 2)) { SZB_CUDA_OK(cudaGetLastError()); return 0; }
    ResidualArgs A;
    fill_pack_args(op, phi, d_km, d_kn, 0, 1, nullptr, A.pk);
    A.nlist = npencil; A.pos = nullptr; A.index = d_index;
    A.x = d_state; A.fs = fs; A.ps = ps; A.b = B; A.r = R;
    A.add = 0; A.it = 0; A.aiter = aiter; A.dmax = dmax; A.mode = mode;
    A.tol = 2.220446049250313e-16 * 0.5;                  // dlamch('E')
    A.res = res; A.lastres = lastres; A.diter = diter; A.cont = cont;
    // the reference starts from lastres = 3 (|b| + 1): never a stagnation at it = 0 unless aiter = 0;
    // with aiter = 0 the test lastres < 2 res needs |b|: keep it simple and exact for aiter >= 1
    // zcgbsvx: the residual is the operator applied to the solution -- accumulate_kernel (0.25 ms for 18 336
    // pencils) plus a small fix-up kernel instead of re-assembling every row block (7.4 ms).  SZB_REFINE_ACC=0
    // keeps the assembling kernel, which zgbsvx (mode 1: it needs |A^T| |x| too) always uses.
    static const bool via_accumulate = [] { const char *e = std::getenv("SZB_REFINE_ACC"); return !(e && e[0] == '0'); }();
    const bool acc = via_accumulate && mode == 0 && op->A.KL == op->A.KU
                     && (op->A.KL == 14 || op->A.KL == 24 || op->A.KL == 34 || op->A.KL == 44);
    // No host synchronisation: the list of pencils that go on lives on the device together with its length
    // (count), which the kernels of a refinement step read themselves; the host enqueues all dmax steps, the
    // ones past the last active pencil find an empty list and return at once (~5 us per launch).
    auto residual = [&](int nlist, const int *cnt, const int *list, const double *kml, const double *knl,
                        const int *slot_in, int add, int it) -> int {
        A.nlist = nlist; A.count_dev = cnt; A.pos = list; A.add = add; A.it = it;
        if (!acc) return dispatch_residual(op, A, stream);
        if (nlist < 1) return 0;
        // x += d and R <- b for the listed pencils, R <- (M + phi L) x - R, then walls / sign / norm / stopping rule
        refine_update_kernel<<<nlist, 128, 0, stream>>>(nlist, cnt, list, N, n, d_index, d_state, fs, ps, B, R, add);
        count_launch();
        const double minus_one[2] = { -1.0, 0.0 };
        int rc2 = accumulate_launch(op, phi, nlist, kml, knl, slot_in, list, 1, reinterpret_cast<const szb_complex *>(d_state),
                                    fs, ps, minus_one, reinterpret_cast<szb_complex *>(R), (size_t) n, (size_t) N, stream, cnt);
        if (rc2) return rc2;
        return dispatch_residual_fix(op, A, stream);
    };
    if ((rc = residual(npencil, nullptr, nullptr, d_km, d_kn, d_index, 0, 0))) return rc;
    for (int it = 1; it <= dmax; ++it) {
        SZB_CUDA_OK(cudaMemsetAsync(count, 0, sizeof(int), stream));
        refine_compact_kernel<<<(npencil + 255) / 256, 256, 0, stream>>>(npencil, cont, d_info, d_km, d_kn,
                                                                         pos, kma, kna, count, d_index, acc ? slot2 : nullptr);
        count_launch();
        // d = (LU)^-T r, in place in R (compact layout: field stride n, pencil stride N)
        rc = invert_fused_dispatch(op, phi, npencil, kma, kna, pos, R, (size_t) n, (size_t) N, nullptr, info2,
                                   nullptr, stream, 0, count);
        if (rc) return rc;
        if ((rc = residual(npencil, count, pos, kma, kna, slot2, 1, it))) return rc;
    }
    refine_finish_kernel<<<(npencil + 255) / 256, 256, 0, stream>>>(npencil, d_info, diter, d_iters);
    count_launch();
    SZB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace szb

// debug hook (not part of the C ABI): phase clocks accumulated by a PROF=1 build
extern "C" int szb_debug_pipe_prof(unsigned long long out[16], int reset)
{
#ifdef SZB_PIPE_PROF
    if (cudaMemcpyFromSymbol(out, g_pipe_prof, sizeof(unsigned long long) * 16) != cudaSuccess) return -1;
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_pipe_prof, z, sizeof z); }
    return 1;
#else
    for (int i = 0; i < 16; ++i) out[i] = 0;
    (void) reset;
    return 0;
#endif
}
