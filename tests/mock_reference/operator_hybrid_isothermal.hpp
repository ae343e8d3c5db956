// Stand-ins for the reference types include/suzerain_b200_shim.hpp names, with the member names and
// signatures of the real ones (apps/perfect/operator_hybrid_isothermal.hpp:101-180,
// suzerain/specification_zgbsv.hpp:43-58, apps/perfect/definition_scenario.hpp:117-164,
// suzerain/specification_isothermal.hpp:186-266, suzerain/specification_grid.hpp, suzerain/pencil_grid.hpp,
// suzerain/bspline.hpp:348-590, apps/perfect/operator_common_block.hpp:59, apps/perfect/references.hpp,
// suzerain/bsplineop.h:125-180, suzerain/rholut_imexop.h:67-133).  Test infrastructure only: it lets the shim
// be compiled (and linked against libsuzerain_b200.so) without Boost / Eigen / MPI.
#pragma once
#include <complex>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <string>

extern "C" {
typedef struct suzerain_bsplineop_workspace {
    int method; int k, n, nderiv; int *kl, *ku; int max_kl, max_ku, ld; double **D_T;
} suzerain_bsplineop_workspace;
typedef struct suzerain_rholut_imexop_ref {
    double *ux, *uy, *uz, *u2, *uxux, *uxuy, *uxuz, *uyuy, *uyuz, *uzuz, *nu, *nuux, *nuuy, *nuuz, *nuu2,
           *nuuxux, *nuuxuy, *nuuxuz, *nuuyuy, *nuuyuz, *nuuzuz, *ex_gradrho, *ey_gradrho, *ez_gradrho, *e_divm, *e_deltarho;
} suzerain_rholut_imexop_ref;
typedef struct suzerain_rholut_imexop_refld {
    int ux, uy, uz, u2, uxux, uxuy, uxuz, uyuy, uyuz, uzuz, nu, nuux, nuuy, nuuz, nuu2,
        nuuxux, nuuxuy, nuuxuz, nuuyuy, nuuyuz, nuuzuz, ex_gradrho, ey_gradrho, ez_gradrho, e_divm, e_deltarho;
} suzerain_rholut_imexop_refld;
}

#define SUZERAIN_ESANITY 7
#define SUZERAIN_EFAILED 5
#define SUZERAIN_ENSURE(expr) do { if (!(expr)) { std::fprintf(stderr, "ENSURE failed: %s\n", #expr); std::abort(); } } while (0)
#define SUZERAIN_ERROR_VOID(msg, code) do { std::fprintf(stderr, "error %d: %s\n", (int) (code), msg); std::abort(); } while (0)

namespace suzerain {

typedef double real_t;
typedef std::complex<double> complex_t;

struct Array3i { int v[3]; int x() const { return v[0]; } int y() const { return v[1]; } int z() const { return v[2]; } };
struct Array3r { double v[3]; double x() const { return v[0]; } double y() const { return v[1]; } double z() const { return v[2]; } };
struct Matrix5r { double m[25]; const double *data() const { return m; } };

class specification_zgbsv {
public:
    enum method_type { zgbsvx = 1, zgbsv, zcgbsvx };
    method_type method() const { return method_; }
    bool equil() const { return equil_; }  bool reuse() const { return reuse_; }
    int aiter() const { return aiter_; }   int siter() const { return siter_; }   int diter() const { return diter_; }
    double tolsc() const { return tolsc_; }
    method_type method_ = zcgbsvx; bool equil_ = false, reuse_ = false; int aiter_ = 1, siter_ = -1, diter_ = 5; double tolsc_ = 0;
};
struct specification_isothermal { real_t lower_T, lower_u, lower_v, lower_w, lower_rho, upper_T, upper_u, upper_v, upper_w, upper_rho; };
struct specification_grid {
    Array3r L; Array3i N, dN; double htdelta;
    bool two_sided() const { return htdelta >= 0; }
    bool one_sided() const { return !two_sided(); }
};
struct pencil_grid { Array3i local_wave_start, local_wave_end; };
struct bsplineop { const suzerain_bsplineop_workspace *w; const suzerain_bsplineop_workspace *get() const { return w; } };
struct bspline {};

namespace multi_array {
template <class T, int D> struct ref {
    T *p; std::size_t shp[D];
    T *data() { return p; } const T *data() const { return p; }
    const std::size_t *shape() const { return shp; }
};
}
template <int D, class T> struct contiguous_state {
    T *p; std::ptrdiff_t str[D];
    T *data() { return p; } const std::ptrdiff_t *strides() const { return str; }
};
namespace lowstorage { template <class T> struct method_interface {}; }

namespace perfect {

struct definition_scenario { real_t Re, Ma, Pr, bulk_rho, bulk_rho_u, bulk_rho_E, alpha, beta, gamma; };
namespace linearize { enum type { rhome_xyz = 1, rhome_y, none }; }
struct references {
    void rholut_imexop(suzerain_rholut_imexop_ref &r, suzerain_rholut_imexop_refld &ld) { (void) r; (void) ld; }
};
struct operator_common_block { linearize::type linearization; references ref; };

class operator_hybrid_isothermal {
public:
    struct linear { typedef real_t component; };
    operator_hybrid_isothermal(const specification_zgbsv&, const definition_scenario&, const specification_isothermal&,
                               const specification_grid&, const pencil_grid&, const bsplineop&, bspline&,
                               operator_common_block& common_) : common(common_) {}
    virtual ~operator_hybrid_isothermal() {}
    virtual void apply_mass_plus_scaled_operator(const complex_t&, multi_array::ref<complex_t,4>&, const std::size_t) const = 0;
    virtual void accumulate_mass_plus_scaled_operator(const complex_t&, const multi_array::ref<complex_t,4>&, const complex_t&,
                                                      contiguous_state<4,complex_t>&, const std::size_t) const = 0;
    virtual void invert_mass_plus_scaled_operator(const complex_t&, multi_array::ref<complex_t,4>&,
                                                  const lowstorage::method_interface<complex_t>&, const linear::component,
                                                  const std::size_t, multi_array::ref<complex_t,4>* ic0 = NULL) const = 0;
protected:
    operator_common_block& common;
    Matrix5r upper_nrbc_a, upper_nrbc_b, upper_nrbc_c;
};

}  // namespace perfect
}  // namespace suzerain
