// bspline.cpp -- host-side construction of B-spline Greville-collocation
// operators, plus band/BSMBSM index algebra and grid stretching.
//
// Replaces, for SUZERAIN_BSPLINEOP_COLLOCATION_GREVILLE only:
//   suzerain_bsplineop_alloc                       suzerain/bsplineop.c:101-201
//   ..._determine_collocation_bandwidths           suzerain/bsplineop.c:403-517
//   ..._create_collocation_operator_transposes     suzerain/bsplineop.c:553-652
// and the GSL calls they lean on (gsl_bspline_alloc/knots,
// gsl_bspline_greville_abscissa, gsl_bspline_basis_deriv).  The basis
// functions and their derivatives are evaluated with the triangular
// Cox-de Boor scheme (de Boor, "A Practical Guide to Splines", BSPLVD; the
// tabular form is Piegl & Tiller algorithm A2.3).  Setup-time only.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>

#include "szb_internal.hpp"

namespace {

// Nonzero basis functions of order k (degree p = k-1) and their derivatives
// up to nd at x, where t[span] <= x < t[span+1] (or x == right end).
// ders[d*(p+1) + r] = d-th derivative of B_{span-p+r}.
void basis_derivs(const double *t, int p, int span, double x, int nd, double *ders)
{
    const int P1 = p + 1;
    std::vector<double> ndu(P1 * P1), left(P1), right(P1), a(2 * P1);
    ndu[0] = 1.0;
    for (int j = 1; j <= p; ++j) {
        left[j]  = x - t[span + 1 - j];
        right[j] = t[span + j] - x;
        double saved = 0.0;
        for (int r = 0; r < j; ++r) {
            ndu[j * P1 + r] = right[r + 1] + left[j - r];          // knot differences
            const double temp = ndu[r * P1 + (j - 1)] / ndu[j * P1 + r];
            ndu[r * P1 + j] = saved + right[r + 1] * temp;
            saved = left[j - r] * temp;
        }
        ndu[j * P1 + j] = saved;
    }
    for (int j = 0; j <= p; ++j) ders[j] = ndu[j * P1 + p];
    for (int d = 1; d <= nd; ++d)
        for (int j = 0; j <= p; ++j) ders[d * P1 + j] = 0.0;

    const int ndmax = std::min(nd, p);
    for (int r = 0; r <= p; ++r) {
        int s1 = 0, s2 = 1;
        a[0] = 1.0;
        for (int d = 1; d <= ndmax; ++d) {
            double acc = 0.0;
            const int rk = r - d, pk = p - d;
            if (r >= d) {
                a[s2 * P1 + 0] = a[s1 * P1 + 0] / ndu[(pk + 1) * P1 + rk];
                acc = a[s2 * P1 + 0] * ndu[rk * P1 + pk];
            }
            const int j1 = (rk >= -1) ? 1 : -rk;
            const int j2 = (r - 1 <= pk) ? d - 1 : p - r;
            for (int j = j1; j <= j2; ++j) {
                a[s2 * P1 + j] = (a[s1 * P1 + j] - a[s1 * P1 + j - 1])
                               / ndu[(pk + 1) * P1 + rk + j];
                acc += a[s2 * P1 + j] * ndu[(rk + j) * P1 + pk];
            }
            if (r <= pk) {
                a[s2 * P1 + d] = -a[s1 * P1 + d - 1] / ndu[(pk + 1) * P1 + r];
                acc += a[s2 * P1 + d] * ndu[r * P1 + pk];
            }
            ders[d * P1 + r] = acc;
            std::swap(s1, s2);
        }
    }
    double fac = p;
    for (int d = 1; d <= ndmax; ++d) {
        for (int j = 0; j <= p; ++j) ders[d * P1 + j] *= fac;
        fac *= (p - d);
    }
}

int find_span(const std::vector<double> &t, int n, int p, double x)
{
    // last span [t[i], t[i+1]) with t[i] <= x, restricted to i in [p, n-1]
    const double *lo = t.data() + p, *hi = t.data() + n + 1;
    int i = (int) (std::upper_bound(lo, hi, x) - t.data()) - 1;
    if (i < p) i = p;
    if (i > n - 1) i = n - 1;
    while (i > p && t[i] == t[i + 1]) --i;      // skip empty spans
    return i;
}

// Dense evaluation D[d][i][j] = B_j^(d)(xi_i) restricted to its k nonzeros;
// calls visit(d, i, j, value) for every (possibly zero) evaluated entry.
template <class F>
void for_each_collocation_entry(const szb_bsplineop &w, int i0, int i1, F visit)
{
    const int p = w.k - 1, P1 = w.k;
    std::vector<double> ders((w.nderiv + 1) * P1);
    for (int i = i0; i < i1; ++i) {
        const double x = w.greville[i];
        const int span = find_span(w.knots, w.n, p, x);
        basis_derivs(w.knots.data(), p, span, x, w.nderiv, ders.data());
        for (int d = 0; d <= w.nderiv; ++d)
            for (int r = 0; r <= p; ++r)
                visit(d, i, span - p + r, ders[d * P1 + r]);
    }
}

}  // namespace

extern "C" {

int szb_gbmatrix_offset(int ld, int kl, int ku, int i, int j)
{ (void) kl; return j * ld + (ku + i - j); }

int szb_gbmatrix_in_band(int kl, int ku, int i, int j)
{ return j - ku <= i && i <= j + kl; }

szb_bsmbsm szb_bsmbsm_construct(int S, int n, int kl, int ku)
{
    szb_bsmbsm t;
    t.S = S; t.n = n; t.kl = kl; t.ku = ku;
    t.ld = kl + 1 + ku;
    t.N  = S * n;
    t.KL = S * (kl + 1) - 1;
    t.KU = S * (ku + 1) - 1;
    t.LD = t.KL + 1 + t.KU;
    return t;
}

int szb_bsmbsm_q(int S, int n, int i)    { return (i % S) * n + i / S; }
int szb_bsmbsm_qinv(int S, int n, int i) { return (i % n) * S + i / n; }

double szb_htstretch1(double delta, double L, double x)
{
    if (delta == 0.0) return x / L;
    return 1 + std::tanh(delta * (x / L - 1)) / std::tanh(delta);
}

double szb_htstretch2(double delta, double L, double x)
{
    if (delta == 0.0) return x / L;
    return 0.5 * (1 + std::tanh(delta * (x / L - 0.5)) / std::tanh(delta / 2));
}

int szb_bsplineop_alloc(int k, int nbreak, const double *breakpoints,
                        int nderiv, szb_bsplineop **out)
{
    if (k < 1) return -1;
    if (nbreak < 2) return -2;
    if (!breakpoints) return -3;
    if (nderiv < 0) return -4;
    if (!out) return -5;
    for (int i = 1; i < nbreak; ++i)
        if (!(breakpoints[i] > breakpoints[i - 1])) return -3;

    szb_bsplineop *w = new (std::nothrow) szb_bsplineop();
    if (!w) return -5;
    w->k = k;
    w->n = nbreak + k - 2;
    w->nderiv = nderiv;
    const int n = w->n;

    // Open knot vector: end breakpoints repeated k times.
    w->knots.resize(n + k);
    for (int i = 0; i < k; ++i) w->knots[i] = breakpoints[0];
    for (int i = 1; i < nbreak - 1; ++i) w->knots[k - 1 + i] = breakpoints[i];
    for (int i = 0; i < k; ++i) w->knots[n + i] = breakpoints[nbreak - 1];

    // Greville abscissae: running mean of k-1 consecutive knots (order 1: the
    // span midpoints).
    w->greville.resize(n);
    for (int i = 0; i < n; ++i) {
        if (k == 1) {
            w->greville[i] = 0.5 * (breakpoints[i] + breakpoints[i + 1]);
        } else {
            double mean = 0;
            for (int j = 0; j < k - 1; ++j)
                mean += (w->knots[i + 1 + j] - mean) / (j + 1);
            w->greville[i] = mean;
        }
    }

    // Bandwidths: start from k-1 and trim each all-zero outer diagonal found
    // in the upper-left and lower-right k x k corners (exact-zero test).
    // kl/ku describe the *transposed* operator's storage, as in the reference.
    w->kl.assign(nderiv + 1, k - 1);
    w->ku.assign(nderiv + 1, k - 1);
    {
        const int kk = std::min(k, n);
        // asum[d][k-1 + (j - i)] over corner entries, j = basis, i = point
        std::vector<double> asum((size_t) (nderiv + 1) * (2 * k - 1), 0.0);
        auto corner = [&](int i0, int i1, int lo, int hi) {
            for_each_collocation_entry(*w, i0, i1, [&](int d, int i, int j, double v) {
                if (j < lo || j >= hi) return;
                const int off = j - i;                   // D^T[j, i]: row j, col i
                if (off <= -k || off >= k) return;
                asum[(size_t) d * (2 * k - 1) + (k - 1 + off)] += std::fabs(v);
            });
        };
        corner(0, kk, 0, kk);
        corner(n - kk, n, n - kk, n);
        for (int d = 0; d <= nderiv; ++d) {
            const double *s = &asum[(size_t) d * (2 * k - 1) + (k - 1)];
            // super-diagonals of D^T: row j < col i, i.e. off = j - i < 0
            for (int m = k - 1; m >= 1 && s[-m] == 0.0; --m) --w->ku[d];
            for (int m = k - 1; m >= 1 && s[+m] == 0.0; --m) --w->kl[d];
        }
    }
    w->max_kl = *std::max_element(w->kl.begin(), w->kl.end());
    w->max_ku = *std::max_element(w->ku.begin(), w->ku.end());
    w->ld = w->max_ku + 1 + w->max_kl;

    // One zeroed block for all operators; D_T[d] = block d stepped past the
    // unused super-diagonals.  Entry D^T[j, i] = B_j^(d)(xi_i) sits at
    // i*ld + (ku[d] + j - i) from D_T[d].
    w->storage.assign((size_t) (nderiv + 1) * w->ld * n, 0.0);
    int bad = 0;
    for_each_collocation_entry(*w, 0, n, [&](int d, int i, int j, double v) {
        if (j < 0 || j >= n) return;
        if (i - w->ku[d] <= j && j <= i + w->kl[d]) {
            double *DT = w->storage.data() + (size_t) d * w->ld * n + (w->max_ku - w->ku[d]);
            DT[(size_t) i * w->ld + (w->ku[d] + j - i)] = v;
        } else if (v != 0.0) {
            bad = 1;                                      // nonzero outside band
        }
    });
    if (bad) { delete w; return 1; }
    *out = w;
    return 0;
}

int szb_bsplineop_from_storage(int k, int n, int nderiv, const int *kl,
                               const int *ku, const double *storage,
                               szb_bsplineop **out)
{
    if (k < 1) return -1;
    if (n < 1) return -2;
    if (nderiv < 0) return -3;
    if (!kl) return -4;
    if (!ku) return -5;
    if (!storage) return -6;
    if (!out) return -7;
    szb_bsplineop *w = new (std::nothrow) szb_bsplineop();
    if (!w) return -7;
    w->k = k; w->n = n; w->nderiv = nderiv;
    w->kl.assign(kl, kl + nderiv + 1);
    w->ku.assign(ku, ku + nderiv + 1);
    w->max_kl = *std::max_element(w->kl.begin(), w->kl.end());
    w->max_ku = *std::max_element(w->ku.begin(), w->ku.end());
    w->ld = w->max_ku + 1 + w->max_kl;
    w->storage.assign(storage, storage + (size_t) (nderiv + 1) * w->ld * n);
    w->greville.assign(n, 0.0);
    *out = w;
    return 0;
}

void szb_bsplineop_free(szb_bsplineop *w) { if (w && w->d_Dr) cudaFree(w->d_Dr); delete w; }
int szb_bsplineop_k     (const szb_bsplineop *w) { return w->k; }
int szb_bsplineop_n     (const szb_bsplineop *w) { return w->n; }
int szb_bsplineop_nderiv(const szb_bsplineop *w) { return w->nderiv; }
int szb_bsplineop_kl    (const szb_bsplineop *w, int d) { return w->kl[d]; }
int szb_bsplineop_ku    (const szb_bsplineop *w, int d) { return w->ku[d]; }
int szb_bsplineop_max_kl(const szb_bsplineop *w) { return w->max_kl; }
int szb_bsplineop_max_ku(const szb_bsplineop *w) { return w->max_ku; }
int szb_bsplineop_ld    (const szb_bsplineop *w) { return w->ld; }
const double *szb_bsplineop_D_T(const szb_bsplineop *w, int d)
{ return (d < 0 || d > w->nderiv) ? nullptr : w->D_T(d); }
int szb_bsplineop_greville(const szb_bsplineop *w, double *xi)
{
    if (!xi) return -2;
    std::memcpy(xi, w->greville.data(), sizeof(double) * w->n);
    return 0;
}

/* ---- wave-space bookkeeping (suzerain/inorder.h:92-96,146-150) ---- */
static inline int wavenumber(int N, int i) { return (i < N / 2 + 1) ? i : -N + i; }
static inline int wavenumber_absmin(int N) { return (N - 1) / 2; }

int szb_wavegrid_npencils(const szb_wavegrid *g)
{ return (g->dkex - g->dkbx) * (g->dkez - g->dkbz); }

int szb_wavegrid_wavenumbers(const szb_wavegrid *g, double *km, double *kn, int *active)
{
    // 2*pi/L evaluated on its own and then multiplied by the integer
    // wavenumber, as the reference does under fp_contract(off)
    // (operator_hybrid_isothermal.cpp:53-65,133-134); volatile blocks
    // contraction/reassociation here.
    volatile double twopioverLx = 2 * M_PI / g->Lx;
    volatile double twopioverLz = 2 * M_PI / g->Lz;
    const int nx = g->dkex - g->dkbx;
    int nact = 0;
    for (int n = g->dkbz; n < g->dkez; ++n) {
        const int wn = wavenumber(g->dNz, n);
        for (int m = g->dkbx; m < g->dkex; ++m) {
            const int wm = wavenumber(g->dNx, m);
            const size_t p = (size_t) (n - g->dkbz) * nx + (m - g->dkbx);
            const int act = !(std::abs(wn) > wavenumber_absmin(g->Nz)
                           || std::abs(wm) > wavenumber_absmin(g->Nx));
            if (km) km[p] = twopioverLx * wm;
            if (kn) kn[p] = twopioverLz * wn;
            if (active) active[p] = act;
            nact += act;
        }
    }
    return nact;
}

int szb_wavegrid_nactive(const szb_wavegrid *g)
{ return szb_wavegrid_wavenumbers(g, nullptr, nullptr, nullptr); }

}  // extern "C"
