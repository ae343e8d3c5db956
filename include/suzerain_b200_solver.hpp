// suzerain_b200_solver.hpp -- header-only C++ drop-in for suzerain::bsmbsm_solver
// (suzerain/bsmbsm_solver.hpp:70-330, suzerain/bsmbsm_solver.cpp:36-414) on top of the C ABI.
//
// Same protocol, same public members, same storage conventions as the reference class, for callers
// that drive the solver object directly (apps/reacting, tests/test_bsmbsm.cpp-style code,
// apps/perfect/operator_hybrid_isothermal.cpp:646-674):
//
//     std::unique_ptr<bsmbsm_solver_b200> s(bsmbsm_solver_b200::build(A, spec, nrhs));
//     s->supply_B(b);                       // PB = P b          (suzerain_bsmbsm_zaPxpby 'N')
//     ... fill s->PAPT (or s->LU, in place) with P A^T P^T ...
//     s->supplied_PAPT();
//     s->solve('T');                        // zgbtrf + zgbtrs, or zcgbsvx, on the B200
//     s->demand_X(x);                       // x = P^T PX        (suzerain_bsmbsm_zaPxpby 'T')
//
// LU is (KL + LD) x N column-major; with an in-place specification (zgbsv) PAPT aliases LU + KL with column
// stride LD + KL and PX aliases PB (bsmbsm_solver.cpp:64-67); otherwise PAPT (LD x N) and PX are separate
// (:202-203, :310-311).  ipiv is LAPACK's (1-based).  Only the C ABI is needed: no CUDA, Eigen or Boost
// headers.  zgbsvx has no pre-assembled device entry point; build() refuses it instead of mapping it silently.
#ifndef SUZERAIN_B200_SOLVER_HPP
#define SUZERAIN_B200_SOLVER_HPP

#include <complex>
#include <cstddef>
#include <stdexcept>
#include <string>
#include <vector>

#include "suzerain_b200.h"

namespace suzerain_b200 {

typedef std::complex<double> complex_double;

/** Column-major matrix view with Eigen's spelling of the few members callers touch. */
template <class T>
struct matrix_view {
    T *p; int r, c, stride;
    matrix_view() : p(0), r(0), c(0), stride(0) {}
    matrix_view(T *p_, int r_, int c_, int s_) : p(p_), r(r_), c(c_), stride(s_) {}
    T *data() { return p; }
    const T *data() const { return p; }
    int rows() const { return r; }
    int cols() const { return c; }
    int colStride() const { return stride; }
    int outerStride() const { return stride; }
    T &operator()(int i, int j) { return p[i + (std::ptrdiff_t) j * stride]; }
    const T &operator()(int i, int j) const { return p[i + (std::ptrdiff_t) j * stride]; }
    void setZero() { for (int j = 0; j < c; ++j) for (int i = 0; i < r; ++i) (*this)(i, j) = T(); }
};

class bsmbsm_solver_b200 : public szb_bsmbsm
{
public:
    typedef matrix_view<complex_double> LU_type, PB_type, PAPT_type, PX_type;

    static bsmbsm_solver_b200 *build(const szb_bsmbsm &bsmbsm, const szb_zgbsv_spec &spec, const int nrhs)
    {
        if (spec.method != SZB_SOLVER_ZGBSV && spec.method != SZB_SOLVER_ZCGBSVX)
            throw std::invalid_argument("bsmbsm_solver_b200: zgbsvx is served only by the fused device entry points "
                                        "(szb_imexop_invert_batch); use zgbsv or zcgbsvx here");
        return new bsmbsm_solver_b200(bsmbsm, spec, nrhs);
    }
    virtual ~bsmbsm_solver_b200() {}

    szb_zgbsv_spec spec;
    LU_type   LU;       ///< (KL + LD) x N factorisation storage
    PB_type   PB;       ///< N x nrhs right hand sides, renumbered
    PAPT_type PAPT;     ///< LD x N operator: a view into LU (in place) or separate storage
    PX_type   PX;       ///< N x nrhs solutions: PB itself (in place) or separate storage
    std::vector<int> ipiv;

    bool in_place() const { return spec.method == SZB_SOLVER_ZGBSV; }          // specification_zgbsv.cpp:116-124

    int supply_b(const complex_double *b, const int j, const int incb = 1)
    {
        complex_double *y = &PB(0, j);
        for (int k = 0; k < N; ++k) y[k] = b[(std::ptrdiff_t) szb_bsmbsm_q(S, n, k) * incb];     // y = P x
        return 0;
    }
    int supply_B(const complex_double *B, const int ldB, const int incB = 1)
    {
        for (int j = 0; j < PB.cols(); ++j) supply_b(B + (std::ptrdiff_t) j * ldB, j, incB);
        return 0;
    }
    int supply_B(const complex_double *B) { return supply_B(B, N, 1); }

    char fact() const { return fact_; }
    char default_fact() const { return spec.equil ? 'E' : 'N'; }
    bool apprx() const { return apprx_ != 0; }
    virtual bsmbsm_solver_b200 &supplied_PAPT() { fact_ = default_fact(); return *this; }   // bsmbsm_solver.cpp:80-89
    virtual bool apprx(const bool acceptable) { apprx_ = acceptable && spec.reuse; return apprx(); }

    int solve(const char trans, const int nrhs)
    {
        if (nrhs == 0) return 0;
        iters.assign(nrhs, 0); res.assign(nrhs, 0.0);
        const int info = szb_bsmbsm_solver_solve(this, &spec, trans, nrhs,
            reinterpret_cast<szb_complex *>(LU.data()),
            in_place() ? 0 : reinterpret_cast<const szb_complex *>(PAPT.data()), ipiv.data(),
            reinterpret_cast<szb_complex *>(PB.data()),
            in_place() ? 0 : reinterpret_cast<szb_complex *>(PX.data()), iters.data(), res.data());
        if (info > 0) {
            // bsmbsm_solver.cpp:123-141: name the singular row and the state scalar it belongs to
            const int row = info - 1, qrow = szb_bsmbsm_q(S, n, row);
            throw std::runtime_error("bsmbsm_solver_b200: singularity in PAP^T row " + std::to_string(row)
                                     + " corresponding to A row " + std::to_string(qrow) + " for state scalar "
                                     + std::to_string(qrow / n));
        }
        if (info == 0) fact_ = 'F';
        return info;
    }
    int solve(const char trans) { return solve(trans, PB.cols()); }

    int demand_x(complex_double *x, const int j, const int incx = 1) const
    {
        const complex_double *y = &PX(0, j);
        for (int k = 0; k < N; ++k) x[(std::ptrdiff_t) szb_bsmbsm_q(S, n, k) * incx] = y[k];     // x = P^T y
        return 0;
    }
    int demand_X(complex_double *X, const int ldX, const int incX = 1) const
    {
        for (int j = 0; j < PB.cols(); ++j) demand_x(X + (std::ptrdiff_t) j * ldX, j, incX);
        return 0;
    }
    int demand_X(complex_double *X) const { return demand_X(X, N, 1); }

    std::vector<int>    iters;   ///< zcgbsvx: refinement steps of the last solve, per right hand side
    std::vector<double> res;     ///< zcgbsvx: final residual 2-norms

protected:
    bsmbsm_solver_b200(const szb_bsmbsm &bsmbsm, const szb_zgbsv_spec &spec_, const int nrhs)
        : szb_bsmbsm(bsmbsm), spec(spec_), ipiv(bsmbsm.N, 0), fact_(spec_.equil ? 'E' : 'N'), apprx_(0),
          lu_((std::size_t) (bsmbsm.KL + bsmbsm.LD) * bsmbsm.N), pb_((std::size_t) bsmbsm.N * nrhs)
    {
        LU = LU_type(lu_.data(), KL + LD, N, KL + LD);
        PB = PB_type(pb_.data(), N, nrhs, N);
        if (in_place()) {
            PAPT = PAPT_type(lu_.data() + KL, LD, N, KL + LD);
            PX = PB;
        } else {
            papt_.resize((std::size_t) LD * N);
            px_.resize((std::size_t) N * nrhs);
            PAPT = PAPT_type(papt_.data(), LD, N, LD);
            PX = PX_type(px_.data(), N, nrhs, N);
        }
    }
    char fact_;
    int apprx_;

private:
    std::vector<complex_double> lu_, pb_, papt_, px_;
};

}  // namespace suzerain_b200

#endif
