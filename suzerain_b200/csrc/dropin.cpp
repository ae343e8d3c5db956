// dropin.cpp -- libsuzerain_b200_dropin.so: the reference's own per-pencil entry points
// (suzerain/rholut_imexop.h:186-469), same names, same argument lists, served by the kernels of
// libsuzerain_b200.so.  See include/suzerain_b200_dropin.h.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/suzerain_b200_dropin.h"

namespace {

// the reference's error convention for void functions: message, then abort (suzerain/error.c:40-56)
[[noreturn]] void fail(const char *fn, const char *what, int rc)
{
    std::fprintf(stderr, "suzerain_b200 drop-in: %s: %s (code %d)\n", fn, what, rc);
    std::abort();
}

struct Workspace {
    szb_bsplineop *w = nullptr;
    ~Workspace() { szb_bsplineop_free(w); }
};

// the reference's workspace (suzerain/bsplineop.h:125-180) -> ours: the band storage of derivative d
// viewed with the common (max_kl, max_ku, ld) starts at D_T[d] - (max_ku - ku[d]) (bsplineop.h:163-174)
void adopt(const char *fn, const szb_dropin_bsplineop_workspace *rw, Workspace &out)
{
    if (!rw) fail(fn, "null B-spline workspace", -7);
    if (rw->nderiv < 2) fail(fn, "workspace must hold derivatives 0..2 (rholut_imexop.c:77)", -7);
    const size_t blk = (size_t) rw->ld * rw->n;
    std::vector<double> storage((size_t) (rw->nderiv + 1) * blk);
    for (int d = 0; d <= rw->nderiv; ++d)
        std::memcpy(&storage[d * blk], rw->D_T[d] - (rw->max_ku - rw->ku[d]), sizeof(double) * blk);
    const int rc = szb_bsplineop_from_storage(rw->k, rw->n, rw->nderiv, rw->kl, rw->ku, storage.data(), &out.w);
    if (rc) fail(fn, "szb_bsplineop_from_storage", rc);
}

void check_order(const char *fn, int rho_E, int rho_u, int rho_v, int rho_w, int rho)
{
    if (!(rho_E == 0 && rho_u == 1 && rho_v == 2 && rho_w == 3 && rho == 4))
        fail(fn, "only the application's scalar ordering rho_E=0, rho_u=1, rho_v=2, rho_w=3, rho=4 "
                 "(operator_hybrid_isothermal.cpp:644-653) is assembled on the device", -8);
}

void check_structure(const char *fn, const szb_dropin_bsmbsm *A_T, const szb_dropin_bsplineop_workspace *w)
{
    if (!A_T) fail(fn, "null A_T", -14);
    if (A_T->S != 5 || A_T->n != w->n) fail(fn, "A_T must describe S = 5 scalars of w->n points (rholut_imexop.def:80-83)", -14);
    if (A_T->kl < w->max_kl || A_T->ku < w->max_ku) fail(fn, "A_T bandwidths below the operators'", -14);
}

inline const szb_complex *in(const szb_dropin_complex *p) { return reinterpret_cast<const szb_complex *>(p); }
inline szb_complex *out(szb_dropin_complex *p) { return reinterpret_cast<szb_complex *>(p); }
inline void split(const szb_dropin_complex &z, double v[2]) { std::memcpy(v, &z, 2 * sizeof(double)); }

void pack(const char *fn, int packf, const szb_dropin_complex phi, double km, double kn,
          const szb_dropin_scenario *s, const szb_dropin_ref *r, const szb_dropin_refld *ld,
          const szb_dropin_bsplineop_workspace *w, int rho_E, int rho_u, int rho_v, int rho_w, int rho,
          szb_dropin_bsmbsm *A_T, szb_dropin_complex *patpt, const double *a, const double *b, const double *c)
{
    check_order(fn, rho_E, rho_u, rho_v, rho_w, rho);
    Workspace W;
    adopt(fn, w, W);
    check_structure(fn, A_T, w);
    double p2[2];
    split(phi, p2);
    szb_bsmbsm mine;
    const int rc = (packf ? szb_rholut_imexop_packf : szb_rholut_imexop_packc)(
        p2, km, kn, reinterpret_cast<const szb_rholut_imexop_scenario *>(s),
        reinterpret_cast<const szb_rholut_imexop_ref *>(r), reinterpret_cast<const szb_rholut_imexop_refld *>(ld),
        W.w, &mine, out(patpt), a, b, c);
    if (rc) fail(fn, "device assembly failed", rc);
    if (mine.N != A_T->N || mine.KL != A_T->KL || mine.KU != A_T->KU || mine.LD != A_T->LD)
        fail(fn, "A_T does not match suzerain_bsmbsm_construct(5, n, max_kl, max_ku)", -14);
}

}  // namespace

extern "C" {

void suzerain_rholut_imexop_accumulate(
        const szb_dropin_complex phi, const double km, const double kn,
        const szb_dropin_scenario *s, const szb_dropin_ref *r, const szb_dropin_refld *ld,
        const szb_dropin_bsplineop_workspace *w,
        const szb_dropin_complex *in_rho_E, const szb_dropin_complex *in_rho_u,
        const szb_dropin_complex *in_rho_v, const szb_dropin_complex *in_rho_w,
        const szb_dropin_complex *in_rho, const szb_dropin_complex beta,
        szb_dropin_complex *out_rho_E, szb_dropin_complex *out_rho_u, szb_dropin_complex *out_rho_v,
        szb_dropin_complex *out_rho_w, szb_dropin_complex *out_rho,
        const double *a, const double *b, const double *c)
{
    static const char fn[] = "suzerain_rholut_imexop_accumulate";
    Workspace W;
    adopt(fn, w, W);
    double p2[2], b2[2];
    split(phi, p2); split(beta, b2);
    const int rc = szb_rholut_imexop_accumulate(
        p2, km, kn, reinterpret_cast<const szb_rholut_imexop_scenario *>(s),
        reinterpret_cast<const szb_rholut_imexop_ref *>(r), reinterpret_cast<const szb_rholut_imexop_refld *>(ld),
        W.w, in(in_rho_E), in(in_rho_u), in(in_rho_v), in(in_rho_w), in(in_rho), b2,
        out(out_rho_E), out(out_rho_u), out(out_rho_v), out(out_rho_w), out(out_rho), a, b, c);
    if (rc) fail(fn, "device apply failed", rc);
}

void suzerain_rholut_imexop_accumulate00(
        const szb_dropin_complex phi,
        const szb_dropin_scenario *s, const szb_dropin_ref *r, const szb_dropin_refld *ld,
        const szb_dropin_bsplineop_workspace *w,
        const szb_dropin_complex *in_rho_E, const szb_dropin_complex *in_rho_u,
        const szb_dropin_complex *in_rho_v, const szb_dropin_complex *in_rho_w,
        const szb_dropin_complex *in_rho, const szb_dropin_complex beta,
        szb_dropin_complex *out_rho_E, szb_dropin_complex *out_rho_u, szb_dropin_complex *out_rho_v,
        szb_dropin_complex *out_rho_w, szb_dropin_complex *out_rho,
        const double *c)
{
    suzerain_rholut_imexop_accumulate(phi, 0.0, 0.0, s, r, ld, w, in_rho_E, in_rho_u, in_rho_v, in_rho_w, in_rho, beta,
                                      out_rho_E, out_rho_u, out_rho_v, out_rho_w, out_rho, nullptr, nullptr, c);
}

void suzerain_rholut_imexop_packc(
        const szb_dropin_complex phi, const double km, const double kn,
        const szb_dropin_scenario *s, const szb_dropin_ref *r, const szb_dropin_refld *ld,
        const szb_dropin_bsplineop_workspace *w,
        const int rho_E, const int rho_u, const int rho_v, const int rho_w, const int rho,
        szb_dropin_complex *buf, szb_dropin_bsmbsm *A_T, szb_dropin_complex *patpt,
        const double *a, const double *b, const double *c)
{
    (void) buf;
    pack("suzerain_rholut_imexop_packc", 0, phi, km, kn, s, r, ld, w, rho_E, rho_u, rho_v, rho_w, rho, A_T, patpt, a, b, c);
}

void suzerain_rholut_imexop_packc00(
        const szb_dropin_complex phi,
        const szb_dropin_scenario *s, const szb_dropin_ref *r, const szb_dropin_refld *ld,
        const szb_dropin_bsplineop_workspace *w,
        const int rho_E, const int rho_u, const int rho_v, const int rho_w, const int rho,
        szb_dropin_complex *buf, szb_dropin_bsmbsm *A_T, szb_dropin_complex *patpt,
        const double *c)
{
    (void) buf;
    pack("suzerain_rholut_imexop_packc00", 0, phi, 0.0, 0.0, s, r, ld, w, rho_E, rho_u, rho_v, rho_w, rho, A_T, patpt,
         nullptr, nullptr, c);
}

void suzerain_rholut_imexop_packf(
        const szb_dropin_complex phi, const double km, const double kn,
        const szb_dropin_scenario *s, const szb_dropin_ref *r, const szb_dropin_refld *ld,
        const szb_dropin_bsplineop_workspace *w,
        const int rho_E, const int rho_u, const int rho_v, const int rho_w, const int rho,
        szb_dropin_complex *buf, szb_dropin_bsmbsm *A_T, szb_dropin_complex *patpt,
        const double *a, const double *b, const double *c)
{
    (void) buf;
    pack("suzerain_rholut_imexop_packf", 1, phi, km, kn, s, r, ld, w, rho_E, rho_u, rho_v, rho_w, rho, A_T, patpt, a, b, c);
}

void suzerain_rholut_imexop_packf00(
        const szb_dropin_complex phi,
        const szb_dropin_scenario *s, const szb_dropin_ref *r, const szb_dropin_refld *ld,
        const szb_dropin_bsplineop_workspace *w,
        const int rho_E, const int rho_u, const int rho_v, const int rho_w, const int rho,
        szb_dropin_complex *buf, szb_dropin_bsmbsm *A_T, szb_dropin_complex *patpt,
        const double *c)
{
    (void) buf;
    pack("suzerain_rholut_imexop_packf00", 1, phi, 0.0, 0.0, s, r, ld, w, rho_E, rho_u, rho_v, rho_w, rho, A_T, patpt,
         nullptr, nullptr, c);
}

}  // extern "C"
