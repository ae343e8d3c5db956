// rhome_y.cu -- the wavenumber-independent linearisation (linearize::rhome_y,
// apps/perfect/operator_hybrid_isothermal.cpp:691-761): ONE operator P (M + phi L(0,0))^T P^T
// (suzerain_rholut_imexop_packf00 == packf at km = kn = 0, rholut_imexop.h:330-344), one
// zgbtrf, then supply_B / rhs BC / zgbtrs('T') / demand_X for every (kx,kz) pencil.
//
// On the device: assemble + wall BCs + factor once with the pre-assembled batched kernels
// (imexop.cu), factor it once on a shared-memory column ring, regroup the factors by blocks of
// four columns (reciprocal diagonal, contiguous per-lane loads), then one warp per pencil:
// right hand side in shared memory, U^T forward sweep, L^T backward sweep with the
// interchanges undone, four columns per step, the factors of the next step prefetched into
// registers.  All warps stream the same ~1.3 MB of factors: L2 traffic.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "szb_internal.hpp"
#include "cplx.cuh"

namespace szb {

namespace {

// zgbtf2 of ONE band matrix by one CTA, on a ring of kv + 1 + AHEAD columns in shared memory: column
// j + kv + AHEAD is requested while column j is eliminated, column j leaves once its multipliers are
// scaled.  Same arithmetic as the reference's LAPACK path (bsmbsm_solver.cpp:171-173):
// first maximum of |re| + |im|, reciprocal pivot times column, rank-one update; a zero pivot
// sets info and the factorisation carries on.
constexpr int FACTOR00_AHEAD = 4;
constexpr int FACTOR00_TASKS = 8;      // update entries per thread: kl (kl + ku) <= 8 * 512

// 1/z with one division in the comfortable exponent range, Smith's algorithm outside it
__device__ __forceinline__ cplx fast_recip(cplx z)
{
    const double m = fmax(fabs(z.x), fabs(z.y));
    if (m > 1e-140 && m < 1e140) {
        const double d = 1.0 / fma(z.x, z.x, z.y * z.y);
        return cplx(z.x * d, -z.y * d);
    }
    return recip(z);
}

__global__ void __launch_bounds__(512)
factor00_kernel(int n, int kl, int ku, cplx *ab, int ldab, int *ipiv, int *info)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx *ring = reinterpret_cast<cplx *>(smem_raw);
    __shared__ cplx s_piv, s_d0, s_rinv;
    __shared__ int s_jp, s_ok;
    const int kv = kl + ku, RC = kv + 1 + FACTOR00_AHEAD, tid = threadIdx.x, nt = blockDim.x;
    int sj = 0;                                  // ring slot of column j
    // column j + c, 0 <= c < RC, without a division
    auto colr = [&](int c) { int sl = sj + c; sl -= sl >= RC ? RC : 0; return ring + (size_t) sl * ldab; };
    // columns arrive by cp.async, FACTOR00_AHEAD column steps before they are first touched
    auto load_col = [&](int c, int slot) {
        if (c < n) {
            cplx *d = ring + (size_t) slot * ldab; const cplx *g = ab + (size_t) c * ldab;
            for (int r = tid; r < ldab; r += nt) {
                if (r < kl) d[r] = cplx(0.0, 0.0);
                else asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((unsigned) __cvta_generic_to_shared(d + r)), "l"(g + r) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    for (int c = 0; c < kv + FACTOR00_AHEAD; ++c) load_col(c, c);
    int ju = 0, inf = 0;
    for (int j = 0; j < n; ++j) {
        const int km = min(kl, n - 1 - j);
        cplx *cj = ring + (size_t) sj * ldab;
        // column j + kv (the last one this step can touch) has landed for every thread, and
        // column j - 1 has left: its slot takes the next request
        asm volatile("cp.async.wait_group %0;" :: "n"(FACTOR00_AHEAD - 1) : "memory");
        __syncthreads();
        load_col(j + kv + FACTOR00_AHEAD, sj == 0 ? RC - 1 : sj - 1);
        if (tid < 32) {
            double best = -1.0; int bi = 0;
            for (int i = tid; i <= km; i += 32) {
                const double m = cabs1(cj[kv + i]);
                if (m > best) { best = m; bi = i; }
            }
            // the top word of the magnitudes decides almost every column with one REDUX; lanes
            // that share the largest top word settle it exactly
            const unsigned key = best < 0.0 ? 0u : (unsigned) __double2hiint(best) + 1u;
            const unsigned kmax = __reduce_max_sync(0xffffffffu, key);
            unsigned tie = __ballot_sync(0xffffffffu, key == kmax);
            if (tie & (tie - 1)) {
                if (key != kmax) best = -1.0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
                }
                tie = (tid == (bi & 31)) ? 1u << tid : 0u;
                tie = __ballot_sync(0xffffffffu, tie != 0);
            }
            if (tid == __ffs(tie) - 1) {
                const cplx pv = cj[kv + bi];
                s_jp = bi; s_piv = pv; s_d0 = cj[kv];
                s_ok = (pv.x != 0.0 || pv.y != 0.0);
                if (s_ok) s_rinv = fast_recip(pv);
                ipiv[j] = j + bi + 1;
            }
        }
        __syncthreads();
        const int jp = s_jp, ok = s_ok;
        if (ok) {
            ju = max(ju, min(j + ku + jp, n - 1));
            const int nc = ju - j;
            // interchange over columns j .. ju, multipliers of column j
            if (tid <= nc) {
                if (jp) {
                    if (tid == 0) cj[kv] = s_piv;
                    else {
                        cplx *cc = colr(tid);
                        const cplx a = cc[kv - tid], b = cc[kv + jp - tid];
                        cc[kv - tid] = b; cc[kv + jp - tid] = a;
                    }
                }
            } else if (tid >= 128 && tid - 128 < km) {
                const int i = tid - 128 + 1;
                const cplx v = (i == jp) ? s_d0 : cj[kv + i];
                cj[kv + i] = v * s_rinv;
            }
            __syncthreads();
            // rank-one update of columns j+1 .. ju: loads first, stores last
            const int total = km * nc;
            const float rkm = 1.0f / (float) km;
            cplx *pw[FACTOR00_TASKS]; cplx w[FACTOR00_TASKS];
#pragma unroll
            for (int t = 0; t < FACTOR00_TASKS; ++t) {
                const int e = tid + t * 512;
                pw[t] = nullptr;
                if (e < total) {
                    const int c0 = __float2int_rz(((float) e + 0.5f) * rkm), c = c0 + 1, i = e - c0 * km + 1;
                    cplx *cc = colr(c);
                    pw[t] = cc + kv + i - c;
                    w[t] = *pw[t];
                    submul(w[t], cj[kv + i], cc[kv - c]);
                }
            }
#pragma unroll
            for (int t = 0; t < FACTOR00_TASKS; ++t) if (pw[t]) *pw[t] = w[t];
        } else if (!inf) inf = j + 1;
        // column j is final: write it back; the column that enters next takes a free slot
        {
            cplx *g = ab + (size_t) j * ldab;
            for (int r = tid; r < ldab; r += nt) g[r] = cj[r];
        }
        sj = sj + 1 == RC ? 0 : sj + 1;
    }
    if (tid == 0) *info = inf;
}

// The factors, regrouped for warps that take four columns per step (block b = columns
// 4b .. 4b+3; the matrix is extended by identity rows up to a multiple of four):
//   fw_tri[b][10]     1/U(j,j) for the four columns, then U(j0+m, j0+m') for m < m'
//   fw_upd[b][96][4]  U(j0+m, j0+4+r): what row j0+4+r of U^T takes from the block's columns
//   bw_tri[b][6]      L(j0+m', j0+m) for m < m'
//   bw_upd[b][64][4]  L(j0+4+t, j0+m): what column j0+m's dot product takes from final rows
//   plain[b]          no interchange in the block (then the four columns go together)
constexpr int FWR = 96, BWR = 64;

__global__ void regroup00_kernel(int N, int kl, int ku, const cplx *ab, int ldab, const int *ipiv, int nblk,
                                 cplx *fw_tri, cplx *fw_upd, cplx *bw_tri, cplx *bw_upd, unsigned char *plain)
{
    const int kv = kl + ku;
    auto U = [&](int i, int j) {        // i <= j
        if (i >= N || j >= N) return i == j ? cplx(1.0, 0.0) : cplx(0.0, 0.0);
        return j - i <= kv ? ab[(size_t) j * ldab + kv + i - j] : cplx(0.0, 0.0);
    };
    auto L = [&](int i, int j) {        // i > j
        if (i >= N || j >= N) return cplx(0.0, 0.0);
        return i - j <= kl ? ab[(size_t) j * ldab + kv + i - j] : cplx(0.0, 0.0);
    };
    const int per = 10 + 4 * FWR + 6 + 4 * BWR + 1, total = nblk * per;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int b = e / per, j0 = 4 * b;
        int q = e - b * per;
        if (q < 10) {
            cplx v;
            if (q < 4) v = recip(U(j0 + q, j0 + q));
            else {
                const int m = q < 7 ? 0 : q < 9 ? 1 : 2, mp = q < 7 ? q - 3 : q < 9 ? q - 5 : 3;
                v = U(j0 + m, j0 + mp);
            }
            fw_tri[(size_t) b * 10 + q] = v;
            continue;
        }
        q -= 10;
        if (q < 4 * FWR) { fw_upd[(size_t) b * 4 * FWR + q] = U(j0 + (q & 3), j0 + 4 + (q >> 2)); continue; }
        q -= 4 * FWR;
        if (q < 6) {
            const int m = q < 3 ? 0 : q < 5 ? 1 : 2, mp = q < 3 ? q + 1 : q < 5 ? q - 1 : 3;
            bw_tri[(size_t) b * 6 + q] = L(j0 + mp, j0 + m);
            continue;
        }
        q -= 6;
        if (q < 4 * BWR) { bw_upd[(size_t) b * 4 * BWR + q] = L(j0 + 4 + (q >> 2), j0 + (q & 3)); continue; }
        bool pl = true;
        for (int m = 0; m < 4; ++m) if (j0 + m < N && ipiv[j0 + m] != j0 + m + 1) pl = false;
        plain[b] = pl;
    }
}

struct Solve00Args {
    int N, n, kl, ku, ldab, nblk;
    const cplx *ab, *fw_tri, *fw_upd, *bw_tri, *bw_upd; const unsigned char *plain;
    const int *ipiv; const int *info1;
    int npencil; const int *index;
    cplx *state; size_t fs, ps;
    int with_bc, wall_begin, wall_end;
    int *ipiv_out, *info_out;
};

__device__ __forceinline__ cplx ldg_c(const cplx *p)
{
    const double2 v = __ldg(reinterpret_cast<const double2 *>(p));
    return cplx(v.x, v.y);
}

constexpr int SOLVE00_WARPS = 4;

// One warp per pencil, right hand side in shared memory (padded with zeros past N), the
// regrouped factors streamed through registers one block ahead of their use.
template <int FWP, int BWP>
__global__ void __launch_bounds__(32 * SOLVE00_WARPS, 4)
solve00_kernel(const Solve00Args A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int N = A.N, n = A.n, kv = A.kl + A.ku, kl = A.kl, nblk = A.nblk;
    const int XP = 4 * nblk + FWR;
    cplx *x = reinterpret_cast<cplx *>(smem_raw) + (size_t) warp * (XP + 4);
    double *sred = reinterpret_cast<double *>(x + XP);
    const int info = *A.info1;
    for (int p = blockIdx.x * nw + warp; p < A.npencil; p += gridDim.x * nw) {
        if (lane == 0) A.info_out[p] = info;
        if (A.ipiv_out)
            for (int k = lane; k < N; k += 32) A.ipiv_out[(size_t) p * N + k] = A.ipiv[k];
        if (info) continue;
        cplx *v = A.state + (A.index ? (size_t) A.index[p] : (size_t) p) * A.ps;
        // b = P state with the wall rows zeroed (bsmbsm_solver.hpp:150-156,
        // operator_hybrid_isothermal.cpp:516-525)
        for (int e = lane; e < XP; e += 32) {
            cplx val(0.0, 0.0);
            if (e < N) {
                const int f = e / n, y = e - f * n;
                val = v[(size_t) f * A.fs + y];
                if (A.with_bc && f < 4 && ((y == 0 && A.wall_begin == 0) || (y == n - 1 && A.wall_end == 2)))
                    val = cplx(0.0, 0.0);
                x[5 * y + f] = val;
            } else x[e] = val;
        }
        __syncwarp();
        // ---- U^T y = b (ztbsv 'U','T','N'), four columns per step ----
        {
            cplx up[FWP][4], tr;
            auto fetch = [&](int b, int ps) {
                const cplx *src = A.fw_upd + ((size_t) b * FWR + lane) * 4 + ps * 128;
#pragma unroll
                for (int m = 0; m < 4; ++m) up[ps][m] = ldg_c(src + m);
            };
#pragma unroll
            for (int ps = 0; ps < FWP; ++ps) fetch(0, ps);
            tr = ldg_c(A.fw_tri + (lane < 10 ? lane : 0));
            for (int b = 0; b < nblk; ++b) {
                const int j0 = 4 * b, bn = b + 1 < nblk ? b + 1 : b;
                auto T = [&](int i) { return cplx(__shfl_sync(0xffffffffu, tr.x, i), __shfl_sync(0xffffffffu, tr.y, i)); };
                const cplx y0 = x[j0] * T(0);
                cplx y1 = x[j0 + 1]; submul(y1, T(4), y0); y1 = y1 * T(1);
                cplx y2 = x[j0 + 2]; submul(y2, T(5), y0); submul(y2, T(7), y1); y2 = y2 * T(2);
                cplx y3 = x[j0 + 3]; submul(y3, T(6), y0); submul(y3, T(8), y1); submul(y3, T(9), y2); y3 = y3 * T(3);
                tr = ldg_c(A.fw_tri + (size_t) bn * 10 + (lane < 10 ? lane : 0));
                // every register block is refilled for the next step right after its last use
#pragma unroll
                for (int ps = 0; ps < FWP; ++ps) {
                    cplx w = x[j0 + 4 + lane + 32 * ps];
                    submul(w, up[ps][0], y0); submul(w, up[ps][1], y1);
                    submul(w, up[ps][2], y2); submul(w, up[ps][3], y3);
                    x[j0 + 4 + lane + 32 * ps] = w;
                    fetch(bn, ps);
                }
                if (lane < 4) x[j0 + lane] = lane == 0 ? y0 : lane == 1 ? y1 : lane == 2 ? y2 : y3;
                __syncwarp();
            }
        }
        // ---- L^T x = y: dot products with the multipliers, interchanges undone in reverse ----
        {
            cplx lw[BWP][4], tr;
            auto fetch = [&](int b, int ps) {
                const cplx *src = A.bw_upd + ((size_t) b * BWR + lane) * 4 + ps * 128;
#pragma unroll
                for (int m = 0; m < 4; ++m) lw[ps][m] = ldg_c(src + m);
            };
#pragma unroll
            for (int ps = 0; ps < BWP; ++ps) fetch(nblk - 1, ps);
            tr = ldg_c(A.bw_tri + (size_t) (nblk - 1) * 6 + (lane < 6 ? lane : 0));
            for (int b = nblk - 1; b >= 0; --b) {
                const int j0 = 4 * b, bn = b > 0 ? b - 1 : 0;
                if (A.plain[b]) {
                    double s[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) s[i] = 0.0;
#pragma unroll
                    for (int ps = 0; ps < BWP; ++ps) {
                        const cplx xv = x[j0 + 4 + lane + 32 * ps];
#pragma unroll
                        for (int m = 0; m < 4; ++m) {
                            cplx sq(s[2 * m], s[2 * m + 1]);
                            addmul(sq, lw[ps][m], xv);
                            s[2 * m] = sq.x; s[2 * m + 1] = sq.y;
                        }
                        fetch(bn, ps);
                    }
                    // packed butterfly: 8 -> 4 -> 2 -> 1 doubles per lane; lane 4 i ends with component i
                    double w4[4], w2[2], w1;
                    {
                        const bool hi = lane & 16;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const double send = hi ? s[i] : s[i + 4], keep = hi ? s[i + 4] : s[i];
                            w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                        }
                    }
                    {
                        const bool hi = lane & 8;
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const double send = hi ? w4[i] : w4[i + 2], keep = hi ? w4[i + 2] : w4[i];
                            w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                        }
                    }
                    {
                        const bool hi = lane & 4;
                        const double send = hi ? w2[0] : w2[1], keep = hi ? w2[1] : w2[0];
                        w1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                    }
                    w1 += __shfl_xor_sync(0xffffffffu, w1, 2);
                    w1 += __shfl_xor_sync(0xffffffffu, w1, 1);
                    if ((lane & 3) == 0) sred[lane >> 2] = w1;
                    auto T = [&](int i) { return cplx(__shfl_sync(0xffffffffu, tr.x, i), __shfl_sync(0xffffffffu, tr.y, i)); };
                    const cplx l10 = T(0), l20 = T(1), l30 = T(2), l21 = T(3), l31 = T(4), l32 = T(5);
                    __syncwarp();
                    if (lane == 0) {
                        const cplx s0(sred[0], sred[1]), s1(sred[2], sred[3]), s2(sred[4], sred[5]), s3(sred[6], sred[7]);
                        const cplx x3 = x[j0 + 3] - s3;
                        cplx x2 = x[j0 + 2] - s2; submul(x2, l32, x3);
                        cplx x1 = x[j0 + 1] - s1; submul(x1, l21, x2); submul(x1, l31, x3);
                        cplx x0 = x[j0] - s0; submul(x0, l10, x1); submul(x0, l20, x2); submul(x0, l30, x3);
                        x[j0 + 3] = x3; x[j0 + 2] = x2; x[j0 + 1] = x1; x[j0] = x0;
                    }
                    __syncwarp();
                } else {
                    for (int j = min(j0 + 3, N - 2); j >= j0; --j) {
                        const int lm = min(kl, N - 1 - j);
                        const cplx *Lj = A.ab + (size_t) j * A.ldab + kv;
                        cplx s(0.0, 0.0);
                        for (int i = 1 + lane; i <= lm; i += 32) addmul(s, ldg_c(Lj + i), x[j + i]);
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
#pragma unroll
                    for (int ps = 0; ps < BWP; ++ps) fetch(bn, ps);
                }
                tr = ldg_c(A.bw_tri + (size_t) bn * 6 + (lane < 6 ? lane : 0));
            }
        }
        // state = P^T x (bsmbsm_solver.hpp:274-280)
        for (int e = lane; e < N; e += 32) {
            const int f = e / n, y = e - f * n;
            v[(size_t) f * A.fs + y] = x[5 * y + f];
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
// Thread-per-pencil sweeps (k = 4, 6, 8, 10: KL = KU = 5k - 6).  Every lane owns one
// right hand side, so an entry of the factors is ONE uniform load for 32 pencils, the
// interchanges are uniform control flow, and there is nothing to reduce across lanes.
// Blocks are the five rows of one collocation point (N = 5 n).  Each lane keeps a ring of
// XRY collocation points of its vector in shared memory ([row][lane]: conflict free); the
// points enter by cp.async two steps ahead and leave with plain stores, so the state makes
// two round trips through HBM (b -> y, y -> x) and nothing else does.
//   t5_fw_tri[b][15]     1/U(j,j) (5), then U(j0+m, j0+m') for m < m' (10)
//   t5_fw_upd[b][KV][5]  U(j0+m, j0+5+i)
//   t5_bw_tri[b][10]     L(j0+m', j0+m) for m < m'
//   t5_bw_upd[b][KL][5]  L(j0+5+i, j0+m)
// ---------------------------------------------------------------------------
__global__ void regroup5_kernel(int N, int kl, int ku, const cplx *ab, int ldab, const int *ipiv, int nblk,
                                cplx *fw_tri, cplx *fw_upd, cplx *bw_tri, cplx *bw_upd, unsigned char *plain)
{
    const int kv = kl + ku;
    auto U = [&](int i, int j) {        // i <= j
        return (j < N && j - i <= kv) ? ab[(size_t) j * ldab + kv + i - j] : cplx(0.0, 0.0);
    };
    auto L = [&](int i, int j) {        // i > j
        return (i < N && i - j <= kl) ? ab[(size_t) j * ldab + kv + i - j] : cplx(0.0, 0.0);
    };
    const int per = 15 + 5 * kv + 10 + 5 * kl + 1, total = nblk * per;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int b = e / per, j0 = 5 * b;
        int q = e - b * per;
        // pairs (m, m') with m < m' in the order (0,1) (0,2) (0,3) (0,4) (1,2) (1,3) (1,4) (2,3) (2,4) (3,4)
        auto pair = [](int t, int &m, int &mp) {
            if (t < 4) { m = 0; mp = t + 1; } else if (t < 7) { m = 1; mp = t - 2; }
            else if (t < 9) { m = 2; mp = t - 4; } else { m = 3; mp = 4; }
        };
        if (q < 15) {
            cplx v;
            if (q < 5) v = recip(U(j0 + q, j0 + q));
            else { int m, mp; pair(q - 5, m, mp); v = U(j0 + m, j0 + mp); }
            fw_tri[(size_t) b * 15 + q] = v;
            continue;
        }
        q -= 15;
        if (q < 5 * kv) { fw_upd[(size_t) b * 5 * kv + q] = U(j0 + q % 5, j0 + 5 + q / 5); continue; }
        q -= 5 * kv;
        if (q < 10) { int m, mp; pair(q, m, mp); bw_tri[(size_t) b * 10 + q] = L(j0 + mp, j0 + m); continue; }
        q -= 10;
        if (q < 5 * kl) { bw_upd[(size_t) b * 5 * kl + q] = L(j0 + 5 + q / 5, j0 + q % 5); continue; }
        bool pl = true;
        for (int m = 0; m < 5; ++m) if (ipiv[j0 + m] != j0 + m + 1) pl = false;
        plain[b] = pl;
    }
}

template <int KV, int KL>
struct Tpp {
    static constexpr int PD = 1;                       // collocation points (and factor blocks) requested ahead
    static constexpr int YA = (KV + 4) / 5;            // forward: points past b that a step touches
    static constexpr int YB = (KL + 4) / 5;            // backward: points past b an interchange can still reach
    static constexpr int XRY = YA + PD + 2;            // ring, in collocation points
    static constexpr int XR = 5 * XRY;                 // ring, in rows
    static constexpr int TS = 5 * KV + 15;             // one block of the factors (forward; backward is smaller)
    // NW groups of 32 pencils per CTA, two warps per group (they split the rows of every step)
    static constexpr int NW = (sizeof(cplx) * (4 * 32 * (XR + 5) + 2 * TS) + 1024 <= 227 * 1024) ? 4 : 3;
    static constexpr size_t smem = sizeof(cplx) * ((size_t) NW * 32 * (XR + 5) + 2 * TS);
};

template <int KV, int KL>
__global__ void __launch_bounds__(64 * Tpp<KV, KL>::NW)
solve00_tpp_kernel(const Solve00Args A)
{
    using T = Tpp<KV, KL>;
    constexpr int YA = T::YA, YB = T::YB, XRY = T::XRY, XR = T::XR, TS = T::TS, NW = T::NW, NT = 64 * NW, NP = 32 * NW;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n = A.n, N = A.N;
    const int group = warp % NW, half = warp / NW;               // the two warps of a group share its ring
    const int p = blockIdx.x * NP + group * 32 + lane;
    const bool active = p < A.npencil, lead = half == 0;
    const int info = *A.info1;
    if (active && lead) A.info_out[p] = info;
    if (A.ipiv_out) {
        const int base = blockIdx.x * NP, cnt = max(0, min(NP, A.npencil - base)) * N;
        for (int e = tid; e < cnt; e += NT) A.ipiv_out[(size_t) base * N + e] = A.ipiv[e % N];
    }
    if (info) return;
    cplx *tab = reinterpret_cast<cplx *>(smem_raw);                               // [2][TS]
    cplx *ring = tab + 2 * TS + (size_t) group * 32 * (XR + 5) + lane;            // row r of this lane: ring[(r % XR) * 32]
    cplx *part = ring + (size_t) XR * 32;                                         // [5][32] partial dot products (backward)
    cplx *v = A.state + (A.index ? (size_t) A.index[active ? p : 0] : (size_t) (active ? p : 0)) * A.ps;
    const size_t fs = A.fs;
    const bool bc_lo = A.with_bc && A.wall_begin == 0, bc_hi = A.with_bc && A.wall_end == 2;

    auto cp16 = [](cplx *dst, const cplx *src) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;"
                     :: "r"((unsigned) __cvta_generic_to_shared(dst)), "l"(src) : "memory");
    };
    // collocation point y -> its five ring rows (XR is a multiple of five: a point never wraps)
    auto request = [&](int y, bool forward) {
        if (y < 0 || !lead) return;
        cplx *d = ring + (size_t) (y % XRY) * 5 * 32;
        const bool wall = forward && ((y == 0 && bc_lo) || (y == n - 1 && bc_hi));
#pragma unroll
        for (int f = 0; f < 5; ++f) {
            if (y >= n || !active || (wall && f < 4)) d[f * 32] = cplx(0.0, 0.0);
            else cp16(d + f * 32, v + (size_t) f * fs + y);
        }
    };
    auto release = [&](int y) {        // ring -> state
        if (y >= 0 && y < n && active && lead) {
            const cplx *d = ring + (size_t) (y % XRY) * 5 * 32;
#pragma unroll
            for (int f = 0; f < 5; ++f) v[(size_t) f * fs + y] = d[f * 32];
        }
    };
    // block b of the factors -> stage b & 1, by the whole CTA
    auto request_fw = [&](int b) {
        if (b >= n) return;
        cplx *d = tab + (b & 1) * TS;
        const cplx *u = A.fw_upd + (size_t) b * 5 * KV, *t = A.fw_tri + (size_t) b * 15;
        for (int e = tid; e < 5 * KV; e += NT) cp16(d + 15 + e, u + e);
        if (tid < 15) cp16(d + tid, t + tid);
    };
    auto request_bw = [&](int b) {
        if (b < 0) return;
        cplx *d = tab + (b & 1) * TS;
        const cplx *u = A.bw_upd + (size_t) b * 5 * KL, *t = A.bw_tri + (size_t) b * 10;
        for (int e = tid; e < 5 * KL; e += NT) cp16(d + 15 + e, u + e);
        if (tid < 10) cp16(d + tid, t + tid);
    };
    auto commit = [] { asm volatile("cp.async.commit_group;" ::: "memory"); };
    auto land = [] { asm volatile("cp.async.wait_group 0;" ::: "memory"); __syncthreads(); };

    // ---- U^T y = b ----
    for (int y = 0; y <= YA; ++y) request(y, true);
    request_fw(0);
    commit();
    int s0 = 0;                                   // ring row of j0 = 5 b
    for (int b = 0; b < n; ++b) {
        land();                                   // block b and point b + YA are here; step b - 1 is over for every warp
        request(b + YA + 1, true); request_fw(b + 1); commit();
        const cplx *tri = tab + (b & 1) * TS, *up = tri + 15;
        cplx *r0 = ring + (size_t) s0 * 32;
        const cplx y0 = r0[0] * tri[0];
        cplx y1 = r0[32]; submul(y1, tri[5], y0); y1 = y1 * tri[1];
        cplx y2 = r0[64]; submul(y2, tri[6], y0); submul(y2, tri[9], y1); y2 = y2 * tri[2];
        cplx y3 = r0[96]; submul(y3, tri[7], y0); submul(y3, tri[10], y1); submul(y3, tri[12], y2);
        y3 = y3 * tri[3];
        cplx y4 = r0[128]; submul(y4, tri[8], y0); submul(y4, tri[11], y1); submul(y4, tri[13], y2);
        submul(y4, tri[14], y3); y4 = y4 * tri[4];
        // four rows at a time: loads first, the twenty products interleaved over the rows, stores last
        static_assert(KV % 4 == 0, "rows are taken in fours");
#pragma unroll 2
        for (int i = 4 * half; i < KV; i += 8) {
            cplx *px[4], w[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                int sl = s0 + 5 + i + g; sl -= sl >= XR ? XR : 0;
                px[g] = ring + (size_t) sl * 32;
                w[g] = *px[g];
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) submul(w[g], up[5 * (i + g)], y0);
#pragma unroll
            for (int g = 0; g < 4; ++g) submul(w[g], up[5 * (i + g) + 1], y1);
#pragma unroll
            for (int g = 0; g < 4; ++g) submul(w[g], up[5 * (i + g) + 2], y2);
#pragma unroll
            for (int g = 0; g < 4; ++g) submul(w[g], up[5 * (i + g) + 3], y3);
#pragma unroll
            for (int g = 0; g < 4; ++g) submul(w[g], up[5 * (i + g) + 4], y4);
#pragma unroll
            for (int g = 0; g < 4; ++g) *px[g] = w[g];
        }
        if (active && lead) {
            v[b] = y0; v[fs + b] = y1; v[2 * fs + b] = y2; v[3 * fs + b] = y3; v[4 * fs + b] = y4;
        }
        s0 += 5; s0 -= s0 >= XR ? XR : 0;
    }
    land();
    __threadfence();

    // ---- L^T x = y, interchanges undone in reverse ----
    for (int y = n; y <= n + YB; ++y) request(y, false);           // zeros past the matrix
    request(n - 1, false); request_bw(n - 1); commit();
    for (int b = n - 1; b >= 0; --b) {
        land();
        request(b - 1, false); request_bw(b - 1); commit();
        const int j0 = 5 * b;
        s0 = (b % XRY) * 5;
        cplx *r0 = ring + (size_t) s0 * 32;
        const cplx *tri = tab + (b & 1) * TS, *lp = tri + 15;
        const bool plain = A.plain[b];
        if (plain) {
            cplx a0(0.0, 0.0), a1(0.0, 0.0), a2(0.0, 0.0), a3(0.0, 0.0), a4(0.0, 0.0);
#pragma unroll 2
            for (int i = half; i < KL; i += 2) {
                int sl = s0 + 5 + i; sl -= sl >= XR ? XR : 0;
                const cplx xv = ring[(size_t) sl * 32];
                addmul(a0, lp[5 * i], xv); addmul(a1, lp[5 * i + 1], xv); addmul(a2, lp[5 * i + 2], xv);
                addmul(a3, lp[5 * i + 3], xv); addmul(a4, lp[5 * i + 4], xv);
            }
            if (!lead) { part[0] = a0; part[32] = a1; part[64] = a2; part[96] = a3; part[128] = a4; }
            __syncthreads();
            if (lead) {
                a0 += part[0]; a1 += part[32]; a2 += part[64]; a3 += part[96]; a4 += part[128];
                const cplx x4 = r0[128] - a4;
                cplx x3 = r0[96] - a3; submul(x3, tri[9], x4);
                cplx x2 = r0[64] - a2; submul(x2, tri[7], x3); submul(x2, tri[8], x4);
                cplx x1 = r0[32] - a1; submul(x1, tri[4], x2); submul(x1, tri[5], x3); submul(x1, tri[6], x4);
                cplx x0 = r0[0] - a0; submul(x0, tri[0], x1); submul(x0, tri[1], x2); submul(x0, tri[2], x3);
                submul(x0, tri[3], x4);
                r0[0] = x0; r0[32] = x1; r0[64] = x2; r0[96] = x3; r0[128] = x4;
            }
        } else if (lead) {
            for (int m = 4; m >= 0; --m) {
                const int j = j0 + m;
                if (j > N - 2) continue;
                const int lm = min(KL, N - 1 - j);
                cplx s(0.0, 0.0);
                // L(j + i, j) = bw_upd[b][i - 5 + m][m] past the block, bw_tri inside it
                for (int i = 1; i <= lm; ++i) {
                    const int mp = m + i;
                    cplx l;
                    if (mp >= 5) l = lp[5 * (mp - 5) + m];
                    else l = tri[m == 0 ? mp - 1 : m == 1 ? mp + 2 : m == 2 ? mp + 4 : 9];
                    addmul(s, l, ring[(size_t) ((j + i) % XR) * 32]);
                }
                cplx t = r0[m * 32] - s;
                const int l = A.ipiv[j] - 1;
                if (l != j) { cplx *pl = ring + (size_t) (l % XR) * 32; const cplx u = *pl; *pl = t; t = u; }
                r0[m * 32] = t;
            }
        }
        release(b + YB);
    }
    for (int y = YB - 1; y >= 0; --y) release(y);
}

// ---------------------------------------------------------------------------
// Refinement around the shared factorisation (zcgbsvx with the eps tolerance, dsgbsvx.def:131-318;
// zgbsvx without equilibration, zgbrfs.f): r = b - A^T x with the UNFACTORED operator, one thread
// per pencil, same ring of collocation points as the sweeps; the five columns of P A^T P^T that
// belong to step b are staged by the CTA (zero padded so that no band test is needed).  Norms and
// stopping rules are thread-local.
// ---------------------------------------------------------------------------
__global__ void extract_papt_kernel(int N, int kl, int ku, const cplx *lu, int ldlu, cplx *papt)
{
    const int ld = kl + 1 + ku, total = N * ld;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int j = e / ld, r = e - j * ld, i = j - ku + r;
        papt[e] = (i >= 0 && i < N) ? lu[(size_t) j * ldlu + kl + r] : cplx(0.0, 0.0);
    }
}

struct Residual00Args {
    int N, n, npencil;
    const cplx *papt;               // [N][LD], zero outside the matrix
    const int *index; cplx *x; size_t fs, ps;       // solution (state layout)
    const cplx *b;                  // [npencil][5][n] right hand sides (wall rows still to be zeroed)
    cplx *r;                        // [npencil][5][n] in: correction d (if add), out: residual
    int with_bc, wall_begin, wall_end;
    int add, it, aiter, dmax, mode;
    double tol;
    double *res, *lastres; int *diter, *cont, *count;
};

template <int KL>
struct Res00 {
    static constexpr int KU = KL, LD = KL + 1 + KU, PADL = 5, LDP = LD + 10;
    static constexpr int YB = (KL + 4) / 5;            // points either side of b that column block b touches
    static constexpr int XRY = 2 * YB + 3, XR = 5 * XRY;
    static constexpr int TS = 5 * LDP;
    static constexpr int NW = (sizeof(cplx) * (4 * 32 * XR + 2 * TS) + 1024 <= 227 * 1024) ? 4 : 3;
    static constexpr size_t smem = sizeof(cplx) * ((size_t) NW * 32 * XR + 2 * TS);
    static_assert(5 * YB <= KU + 1, "padding covers the rows of the outermost points");
};

template <int KL>
__global__ void __launch_bounds__(32 * Res00<KL>::NW)
residual00_kernel(const Residual00Args A)
{
    using T = Res00<KL>;
    constexpr int KU = T::KU, LD = T::LD, PADL = T::PADL, LDP = T::LDP, YB = T::YB, XRY = T::XRY, TS = T::TS, NT = 32 * T::NW;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n = A.n, N = A.N;
    const int p = blockIdx.x * NT + tid;
    const bool inrange = p < A.npencil;
    // pencils that stopped earlier sit this pass out
    const bool active = inrange && (A.it == 0 || A.cont[p]);
    cplx *tab = reinterpret_cast<cplx *>(smem_raw);                               // [2][5][LDP]
    cplx *ring = tab + 2 * TS + (size_t) warp * 32 * T::XR + lane;
    const size_t pp = active ? (size_t) p : 0;
    cplx *xg = A.x + (A.index ? (size_t) A.index[pp] : pp) * A.ps;
    const cplx *bg = A.b + pp * N;
    cplx *rg = A.r + pp * N;
    const size_t fs = A.fs;
    const bool bc_lo = A.with_bc && A.wall_begin == 0, bc_hi = A.with_bc && A.wall_end == 2;
    for (int e = tid; e < 2 * TS; e += NT) tab[e] = cplx(0.0, 0.0);
    __syncthreads();

    auto slot = [&](int y) { return ring + (size_t) ((y + XRY) % XRY) * 5 * 32; };
    // x (+ d) of collocation point y into registers; written back to the state when corrected
    auto fetch = [&](int y, cplx (&v)[5]) {
#pragma unroll
        for (int f = 0; f < 5; ++f) v[f] = cplx(0.0, 0.0);
        if (active && y >= 0 && y < n) {
#pragma unroll
            for (int f = 0; f < 5; ++f) v[f] = xg[(size_t) f * fs + y];
            if (A.add) {
#pragma unroll
                for (int f = 0; f < 5; ++f) { v[f] += rg[(size_t) f * n + y]; xg[(size_t) f * fs + y] = v[f]; }
            }
        }
    };
    auto put = [&](int y, const cplx (&v)[5]) {
        cplx *d = slot(y);
#pragma unroll
        for (int f = 0; f < 5; ++f) d[f * 32] = v[f];
    };
    auto request_tab = [&](int b) {
        if (b < n) {
            cplx *d = tab + (b & 1) * TS;
            const cplx *src = A.papt + (size_t) 5 * b * LD;
            for (int e = tid; e < 5 * LD; e += NT) {
                const int m = e / LD, rr = e - m * LD;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;"
                             :: "r"((unsigned) __cvta_generic_to_shared(d + m * LDP + PADL + rr)), "l"(src + e) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    cplx nx[5];
    for (int y = -YB; y <= YB; ++y) { fetch(y, nx); put(y, nx); }
    fetch(YB + 1, nx);
    request_tab(0);
    double s2 = 0.0;
    const double safe1 = min(KL + KU + 2, N + 1) * 2.2250738585072014e-308, safe2 = safe1 / 1.1102230246251565e-16;
    for (int b = 0; b < n; ++b) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                            // block b staged; every warp is done with block b - 1
        request_tab(b + 1);
        if (b > 0) { put(b + YB, nx); fetch(b + YB + 1, nx); }
        const cplx *tb = tab + (b & 1) * TS;
        cplx acc[5]; double ra[5];
#pragma unroll
        for (int m = 0; m < 5; ++m) {
            cplx bv(0.0, 0.0);
            if (active) bv = bg[(size_t) m * n + b];
            if (m < 4 && ((b == 0 && bc_lo) || (b == n - 1 && bc_hi))) bv = cplx(0.0, 0.0);
            acc[m] = bv; ra[m] = cabs1(bv);
        }
        for (int dy = -YB; dy <= YB; ++dy) {
            const cplx *xr = slot(b + dy);
#pragma unroll
            for (int f = 0; f < 5; ++f) {
                const cplx xv = xr[f * 32];
                const int base = PADL + KU + 5 * dy + f;            // + m * LDP - m
#pragma unroll
                for (int m = 0; m < 5; ++m) {
                    const cplx a = tb[m * LDP + base - m];
                    submul(acc[m], a, xv);
                    if (A.mode) ra[m] += cabs1(a) * cabs1(xv);
                }
            }
        }
        if (active) {
#pragma unroll
            for (int m = 0; m < 5; ++m) {
                rg[(size_t) m * n + b] = acc[m];
                if (A.mode) {
                    const double num = cabs1(acc[m]);
                    s2 = fmax(s2, ra[m] > safe2 ? num / ra[m] : (num + safe1) / (ra[m] + safe1));
                } else s2 += acc[m].x * acc[m].x + acc[m].y * acc[m].y;
            }
        }
    }
    bool go = false;
    if (active) {
        if (A.mode) {
            const double berr = s2;
            go = berr > 1.1102230246251565e-16 && 2.0 * berr <= A.lastres[p] && A.it < 5;
            A.diter[p] = A.it; A.res[p] = berr;
            if (go) A.lastres[p] = berr;
        } else {
            const double res = sqrt(s2);
            const bool stop = A.it >= A.aiter && A.lastres[p] < 2.0 * res;
            A.diter[p] = A.it; A.res[p] = res;
            if (!stop) A.lastres[p] = res;
            go = !stop && A.it < A.dmax && res > A.tol;
        }
        A.cont[p] = go;
    }
    const unsigned any = __ballot_sync(0xffffffffu, go);
    if (lane == 0 && any) atomicAdd(A.count, __popc(any));
}

__global__ void gather00_kernel(int npencil, int N, int n, const int *index, const cplx *state, size_t fs, size_t ps,
                                cplx *b, double *lastres, double lastres0)
{
    const int p = blockIdx.x;
    const cplx *v = state + (index ? (size_t) index[p] : (size_t) p) * ps;
    for (int k = threadIdx.x; k < N; k += blockDim.x) {
        const int f = k / n, y = k - f * n;
        b[(size_t) p * N + k] = v[(size_t) f * fs + y];
    }
    if (threadIdx.x == 0) lastres[p] = lastres0;
}

__global__ void finish00_kernel(int npencil, const int *info, const int *diter, int *iters)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < npencil && iters) iters[p] = info[p] ? -1 : diter[p];
}

// The factorisation of the km = kn = 0 operator and everything derived from it, in op->d_work00.
struct Ctx00 {
    int N, n, KL, KU, ldab, kv, nblk;
    bool tpp;
    cplx *LU, *fw_tri, *fw_upd, *bw_tri, *bw_upd, *papt;
    int *ipiv, *info1; double *zero; unsigned char *plain;
};

int prepare00(const szb_imexop *op, const double phi[2], bool want_papt, Ctx00 &C, cudaStream_t stream)
{
    const int N = op->A.N, KL = op->A.KL, KU = op->A.KU, ldab = op->A.LD + KL, kv = KL + KU;
    if (kv + 3 > FWR || KL > BWR || kv >= 128 || KL * kv > FACTOR00_TASKS * 512) return -1;
    const int nblk = (N + 3) / 4;
    // workspace: LU | regrouped factors | ipiv | info | zero wavenumbers | plain flags | unfactored operator
    const size_t b_lu = sizeof(cplx) * (size_t) ldab * N;
    // (sized for either grouping: blocks of four columns, or of five = one collocation point)
    const size_t n5 = (size_t) op->n;
    const size_t b_ft = sizeof(cplx) * std::max((size_t) nblk * 10, n5 * 15);
    const size_t b_fu = sizeof(cplx) * std::max((size_t) nblk * 4 * FWR, n5 * 5 * kv);
    const size_t b_bt = sizeof(cplx) * std::max((size_t) nblk * 6, n5 * 10);
    const size_t b_bu = sizeof(cplx) * std::max((size_t) nblk * 4 * BWR, n5 * 5 * KL);
    const size_t b_ip = ((sizeof(int) * (size_t) N) + 15) & ~(size_t) 15;
    const size_t b_pl = ((size_t) std::max(nblk, op->n) + 15) & ~(size_t) 15;
    const size_t b_pa = sizeof(cplx) * (size_t) op->A.LD * N;
    const size_t need = b_lu + b_ft + b_fu + b_bt + b_bu + b_ip + 16 + 16 + b_pl + b_pa;
    if (need > op->work00_bytes) {
        if (op->d_work00) SZB_CUDA_OK(cudaFree(op->d_work00));
        op->d_work00 = nullptr; op->work00_bytes = 0;
        SZB_CUDA_OK(cudaMalloc(&op->d_work00, need));
        op->work00_bytes = need;
    }
    unsigned char *w = static_cast<unsigned char *>(op->d_work00);
    C.N = N; C.n = op->n; C.KL = KL; C.KU = KU; C.ldab = ldab; C.kv = kv; C.nblk = nblk;
    C.LU = reinterpret_cast<cplx *>(w); w += b_lu;
    C.fw_tri = reinterpret_cast<cplx *>(w); w += b_ft;
    C.fw_upd = reinterpret_cast<cplx *>(w); w += b_fu;
    C.bw_tri = reinterpret_cast<cplx *>(w); w += b_bt;
    C.bw_upd = reinterpret_cast<cplx *>(w); w += b_bu;
    C.ipiv = reinterpret_cast<int *>(w); w += b_ip;
    C.info1 = reinterpret_cast<int *>(w); w += 16;
    C.zero = reinterpret_cast<double *>(w); w += 16;
    C.plain = w; w += b_pl;
    C.papt = reinterpret_cast<cplx *>(w);
    SZB_CUDA_OK(cudaMemsetAsync(C.zero, 0, 16, stream));
    int rc = szb_imexop_pack_batch(op, phi, 1, C.zero, C.zero + 1, 1, 1, reinterpret_cast<szb_complex *>(C.LU), stream);
    if (rc) return rc;
    if (want_papt) {
        extract_papt_kernel<<<op->sm_count, 256, 0, stream>>>(N, KL, KU, C.LU, ldab, C.papt);
        count_launch();
    }
    const size_t fsm = sizeof(cplx) * (size_t) (kv + 1 + FACTOR00_AHEAD) * ldab;
    if (fsm > 48 * 1024)
        SZB_CUDA_OK(cudaFuncSetAttribute(factor00_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) fsm));
    factor00_kernel<<<1, 512, fsm, stream>>>(N, KL, KU, C.LU, ldab, C.ipiv, C.info1);
    count_launch();
    static const bool use_warp = std::getenv("SZB_RHOME_Y_WARP") != nullptr;
    C.tpp = !use_warp && N == 5 * op->n && KL == KU && (KL == 14 || KL == 24 || KL == 34 || KL == 44);
    if (C.tpp)
        regroup5_kernel<<<op->sm_count, 256, 0, stream>>>(N, KL, KU, C.LU, ldab, C.ipiv, op->n, C.fw_tri, C.fw_upd,
                                                         C.bw_tri, C.bw_upd, C.plain);
    else
        regroup00_kernel<<<op->sm_count, 256, 0, stream>>>(N, KL, KU, C.LU, ldab, C.ipiv, nblk, C.fw_tri, C.fw_upd,
                                                          C.bw_tri, C.bw_upd, C.plain);
    count_launch();
    SZB_CUDA_OK(cudaGetLastError());
    return 0;
}

// state <- (LU)^-T P state for npencil right hand sides against the prepared factorisation
int solve00(const szb_imexop *op, const Ctx00 &C, int npencil, const int *d_index, cplx *d_state, size_t fs, size_t ps,
            int with_bc, int *d_ipiv, int *d_info, cudaStream_t stream)
{
    Solve00Args A;
    A.N = C.N; A.n = C.n; A.kl = C.KL; A.ku = C.KU; A.ldab = C.ldab; A.nblk = C.tpp ? C.n : C.nblk;
    A.ab = C.LU; A.fw_tri = C.fw_tri; A.fw_upd = C.fw_upd; A.bw_tri = C.bw_tri; A.bw_upd = C.bw_upd; A.plain = C.plain;
    A.ipiv = C.ipiv; A.info1 = C.info1;
    A.npencil = npencil; A.index = d_index; A.state = d_state; A.fs = fs; A.ps = ps;
    A.with_bc = with_bc; A.wall_begin = op->iso.enforce_lower ? 0 : 1; A.wall_end = op->iso.enforce_upper ? 2 : 1;
    A.ipiv_out = d_ipiv; A.info_out = d_info;
    int rc = 0;
    if (C.tpp) {
        // nt pencils per CTA, two warps per group of 32
        auto go = [&](auto kern, size_t smem, int nt) -> int {
            if (smem > 48 * 1024)
                SZB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
            kern<<<(npencil + nt - 1) / nt, 2 * nt, smem, stream>>>(A);
            return 0;
        };
        rc = C.KL == 14 ? go(solve00_tpp_kernel<28, 14>, Tpp<28, 14>::smem, 32 * Tpp<28, 14>::NW)
           : C.KL == 24 ? go(solve00_tpp_kernel<48, 24>, Tpp<48, 24>::smem, 32 * Tpp<48, 24>::NW)
           : C.KL == 34 ? go(solve00_tpp_kernel<68, 34>, Tpp<68, 34>::smem, 32 * Tpp<68, 34>::NW)
                        : go(solve00_tpp_kernel<88, 44>, Tpp<88, 44>::smem, 32 * Tpp<88, 44>::NW);
    } else {
        const int nw = SOLVE00_WARPS;
        const size_t smem = sizeof(cplx) * (size_t) nw * (4 * C.nblk + FWR + 4);
        const int per_sm = (int) std::max<size_t>(1, std::min<size_t>(4, (224 * 1024) / (smem + 1024)));
        const int grid = std::min((npencil + nw - 1) / nw, per_sm * op->sm_count);
        const int fwp = (C.kv + 3 + 31) / 32, bwp = (C.KL + 31) / 32;
        auto go = [&](auto kern) -> int {
            if (smem > 48 * 1024)
                SZB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
            kern<<<grid, 32 * nw, smem, stream>>>(A);
            return 0;
        };
        rc = fwp == 1 ? go(solve00_kernel<1, 1>) : fwp == 2 ? go(solve00_kernel<2, 1>)
           : bwp == 1 ? go(solve00_kernel<3, 1>) : go(solve00_kernel<3, 2>);
    }
    if (rc) return rc;
    count_launch();
    SZB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace

// linearize::rhome_y invert.  mode < 0: zgbsv; 0: zcgbsvx (eps tolerance); 1: zgbsvx without
// equilibration.  Extra right hand sides (the integral-constraint columns) are just more pencils
// here: every pencil has the same operator.  Returns 0 when done, 1 when this shape / mode has
// no kernel (the caller then uses the general kernels at km = kn = 0), < 0 on error.
int invert00_dispatch(const szb_imexop *op, int mode, int aiter, int dmax, const double phi[2], int npencil,
                      const int *d_index, cplx *d_state, size_t fs, size_t ps, int nextra, cplx *d_extra,
                      int *d_ipiv, int *d_info, int *d_iters, cudaStream_t stream)
{
    Ctx00 C;
    int rc = prepare00(op, phi, mode >= 0, C, stream);
    if (rc == -1) return 1;
    if (rc) return rc;
    if (mode >= 0 && (!C.tpp || nextra > 0 || aiter < 1 || dmax < 0)) return 1;
    const int N = C.N, n = C.n;
    if (mode == 1) { aiter = 1; dmax = 5; }
    // workspace of the refinement: b, r | res, lastres | diter, cont, info2 | count; zgbsv with extra
    // right hand sides only needs an info array for them
    const size_t nb = mode >= 0 ? (size_t) npencil * N * sizeof(cplx) : 0;
    const size_t nd = mode >= 0 ? (((size_t) npencil * sizeof(double)) + 15) & ~(size_t) 15 : 0;
    const size_t ni = (((size_t) npencil * std::max(1, nextra) * sizeof(int)) + 15) & ~(size_t) 15;
    const size_t need = 2 * nb + 2 * nd + 3 * ni + 16;
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
    int *diter = reinterpret_cast<int *>(w); w += ni;
    int *cont = reinterpret_cast<int *>(w); w += ni;
    int *info2 = reinterpret_cast<int *>(w); w += ni;
    int *count = reinterpret_cast<int *>(w);

    if (mode >= 0) {
        gather00_kernel<<<npencil, 128, 0, stream>>>(npencil, N, n, d_index, d_state, fs, ps, B, lastres, mode ? 3.0 : 0.0);
        count_launch();
    }
    // x = (LU)^-T b, in place in the state
    if ((rc = solve00(op, C, npencil, d_index, d_state, fs, ps, 1, d_ipiv, d_info, stream))) return rc;
    if (mode < 0) {
        if (nextra > 0 &&
            (rc = solve00(op, C, npencil * nextra, nullptr, d_extra, (size_t) n, (size_t) N, 1, nullptr, info2, stream)))
            return rc;
        if (d_iters) SZB_CUDA_OK(cudaMemsetAsync(d_iters, 0, sizeof(int) * (size_t) npencil, stream));
        return 0;
    }
    Residual00Args A;
    A.N = N; A.n = n; A.npencil = npencil; A.papt = C.papt;
    A.index = d_index; A.x = d_state; A.fs = fs; A.ps = ps; A.b = B; A.r = R;
    A.with_bc = 1; A.wall_begin = op->iso.enforce_lower ? 0 : 1; A.wall_end = op->iso.enforce_upper ? 2 : 1;
    A.add = 0; A.it = 0; A.aiter = aiter; A.dmax = dmax; A.mode = mode;
    A.tol = 2.220446049250313e-16 * 0.5;                  // dlamch('E')
    A.res = res; A.lastres = lastres; A.diter = diter; A.cont = cont; A.count = count;
    auto residual = [&]() -> int {
        SZB_CUDA_OK(cudaMemsetAsync(count, 0, sizeof(int), stream));
        auto go = [&](auto kern, size_t smem, int nt) -> int {
            if (smem > 48 * 1024)
                SZB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
            kern<<<(npencil + nt - 1) / nt, nt, smem, stream>>>(A);
            return 0;
        };
        const int r2 = C.KL == 14 ? go(residual00_kernel<14>, Res00<14>::smem, 32 * Res00<14>::NW)
                     : C.KL == 24 ? go(residual00_kernel<24>, Res00<24>::smem, 32 * Res00<24>::NW)
                     : C.KL == 34 ? go(residual00_kernel<34>, Res00<34>::smem, 32 * Res00<34>::NW)
                                  : go(residual00_kernel<44>, Res00<44>::smem, 32 * Res00<44>::NW);
        if (r2) return r2;
        count_launch();
        SZB_CUDA_OK(cudaGetLastError());
        return 0;
    };
    if ((rc = residual())) return rc;
    for (int it = 1; it <= dmax; ++it) {
        int nact = 0;
        SZB_CUDA_OK(cudaMemcpyAsync(&nact, count, sizeof(int), cudaMemcpyDeviceToHost, stream));
        SZB_CUDA_OK(cudaStreamSynchronize(stream));
        if (nact == 0) break;
        // d = (LU)^-T r, in place in R (field stride n, pencil stride N), every pencil; the residual
        // kernel only takes the corrections of the pencils that go on
        if ((rc = solve00(op, C, npencil, nullptr, R, (size_t) n, (size_t) N, 0, nullptr, info2, stream))) return rc;
        A.add = 1; A.it = it;
        if ((rc = residual())) return rc;
    }
    finish00_kernel<<<(npencil + 255) / 256, 256, 0, stream>>>(npencil, d_info, diter, d_iters);
    count_launch();
    SZB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace szb
