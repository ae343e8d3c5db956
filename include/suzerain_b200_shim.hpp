// suzerain_b200_shim.hpp -- the translation unit a maintainer adds to apps/perfect/:
// operator_hybrid_isothermal with its three linear-operator virtuals
// (apps/perfect/operator_hybrid_isothermal.hpp:112-140) executed on a B200 through the C ABI.
//
// Include it AFTER the reference's own "operator_hybrid_isothermal.hpp" (it names the reference's
// types and nothing else); instantiate operator_hybrid_isothermal_b200 where apps/perfect/main_advance.cpp:647-653
// builds operator_hybrid_isothermal; link -lsuzerain_b200 -lcudart.  tests/test_shim_compiles.py compiles
// this header against stand-ins for the reference types (tests/mock_reference/).
#ifndef SUZERAIN_B200_SHIM_HPP
#define SUZERAIN_B200_SHIM_HPP

#include <cstddef>
#include <string>

#include "suzerain_b200.h"

namespace suzerain { namespace perfect {

/** operator_hybrid_isothermal with the three linear-operator virtuals executed on a B200. */
class operator_hybrid_isothermal_b200 : public operator_hybrid_isothermal
{
public:
    operator_hybrid_isothermal_b200(const specification_zgbsv& spec, const definition_scenario& scenario,
            const specification_isothermal& isothermal, const specification_grid& grid,
            const pencil_grid& dgrid, const bsplineop& cop, bspline& b, operator_common_block& common)
        : operator_hybrid_isothermal(spec, scenario, isothermal, grid, dgrid, cop, b, common),
          bop_(NULL), op_(NULL), one_sided_(grid.one_sided())
    {
        // operators: hand the reference's own D_T storage over (bsplineop.c:163-187: one block, derivative d
        // viewed with the common bandwidths starts at D_T[d] - (max_ku - ku[d]))
        const suzerain_bsplineop_workspace *w = cop.get();
        SUZERAIN_ENSURE(0 == szb_bsplineop_from_storage(w->k, w->n, w->nderiv, w->kl, w->ku,
                               w->D_T[0] - (w->max_ku - w->ku[0]), &bop_));
        SUZERAIN_ENSURE(0 == szb_imexop_create(bop_, &op_));
        const szb_rholut_imexop_scenario s = { scenario.Re, scenario.Pr, scenario.Ma, scenario.alpha, scenario.gamma };
        SUZERAIN_ENSURE(0 == szb_imexop_set_scenario(op_, &s));
        // the lower wall is always enforced, the upper one only on two-sided grids (operator_hybrid_isothermal.cpp:599-601)
        const szb_isothermal iso = { 1, grid.two_sided() ? 1 : 0,
            isothermal.lower_T, isothermal.lower_u, isothermal.lower_v, isothermal.lower_w,
            isothermal.upper_T, isothermal.upper_u, isothermal.upper_v, isothermal.upper_w };
        SUZERAIN_ENSURE(0 == szb_imexop_set_isothermal(op_, &iso));
        wg_.Nx = grid.N.x();  wg_.dNx = grid.dN.x();
        wg_.dkbx = dgrid.local_wave_start.x();  wg_.dkex = dgrid.local_wave_end.x();
        wg_.Nz = grid.N.z();  wg_.dNz = grid.dN.z();
        wg_.dkbz = dgrid.local_wave_start.z();  wg_.dkez = dgrid.local_wave_end.z();
        wg_.Lx = grid.L.x();  wg_.Lz = grid.L.z();
        // specification_zgbsv -> szb_zgbsv_spec, every field (specification_zgbsv.hpp:47-58)
        spec_ = szb_zgbsv_spec_default();
        switch (spec.method()) {
        case specification_zgbsv::zgbsv:   spec_.method = SZB_SOLVER_ZGBSV;   break;
        case specification_zgbsv::zgbsvx:  spec_.method = SZB_SOLVER_ZGBSVX;  break;
        case specification_zgbsv::zcgbsvx: spec_.method = SZB_SOLVER_ZCGBSVX; break;
        }
        spec_.equil = spec.equil();  spec_.reuse = spec.reuse();
        spec_.aiter = spec.aiter();  spec_.siter = spec.siter();  spec_.diter = spec.diter();
        spec_.tolsc = spec.tolsc();
    }
    ~operator_hybrid_isothermal_b200() { szb_imexop_destroy(op_); szb_bsplineop_free(bop_); }

    void apply_mass_plus_scaled_operator(const complex_t& phi, multi_array::ref<complex_t,4>& state,
                                         const std::size_t) const
    {
        refresh();
        check(szb_operator_apply_mass_plus_scaled_operator(op_, &wg_, d2(phi), c(state.data())));
    }
    void accumulate_mass_plus_scaled_operator(const complex_t& phi, const multi_array::ref<complex_t,4>& input,
            const complex_t& beta, contiguous_state<4,complex_t>& output, const std::size_t) const
    {
        refresh();
        check(szb_operator_accumulate_mass_plus_scaled_operator(op_, &wg_, d2(phi), c(input.data()),
                                                               d2(beta), c(output.data()), output.strides()[0]));
    }
    void invert_mass_plus_scaled_operator(const complex_t& phi, multi_array::ref<complex_t,4>& state,
            const lowstorage::method_interface<complex_t>&, const linear::component, const std::size_t,
            multi_array::ref<complex_t,4>* ic0 = NULL) const
    {
        refresh();
        int bad = -1;
        const int nc = ic0 ? (int) (ic0->shape()[2] * ic0->shape()[3]) : 0;      // :580-587
        const int info = szb_operator_invert_mass_plus_scaled_operator(op_, &spec_, &wg_, d2(phi),
                c(state.data()), nc, ic0 ? c(ic0->data()) : NULL, &bad);
        if (info > 0) {   // same decoding as bsmbsm_solver.cpp:123-141
            const int n = szb_bsplineop_n(bop_), q = szb_bsmbsm_q(5, n, info - 1);
            const std::string msg = std::string("pencil ") + std::to_string(bad) + ": singularity in PAP^T row "
                + std::to_string(info - 1) + " corresponding to A row " + std::to_string(q) + " for state scalar "
                + std::to_string(q / n);
            SUZERAIN_ERROR_VOID(msg.c_str(), SUZERAIN_ESANITY);
        }
        check(info);
    }

private:
    /** Per call: linearisation, reference profiles (they change once per step, perfect.cpp:1397), Giles matrices. */
    void refresh() const
    {
        check(szb_imexop_set_linearization(op_, common.linearization == linearize::rhome_y
                                                ? SZB_LINEARIZE_RHOME_Y : SZB_LINEARIZE_RHOME_XYZ));
        suzerain_rholut_imexop_ref r; suzerain_rholut_imexop_refld ld;
        common.ref.rholut_imexop(r, ld);
        check(szb_imexop_set_refs(op_, reinterpret_cast<szb_rholut_imexop_ref*>(&r),
                                       reinterpret_cast<szb_rholut_imexop_refld*>(&ld)));
        if (one_sided_)
            check(szb_imexop_set_nrbc(op_, upper_nrbc_a.data(), upper_nrbc_b.data(), upper_nrbc_c.data()));
    }
    static const double* d2(const complex_t& z) { return reinterpret_cast<const double*>(&z); }
    template<class T> static szb_complex* c(T* p) { return reinterpret_cast<szb_complex*>(const_cast<complex_t*>(p)); }
    static void check(int rc) { if (rc) SUZERAIN_ERROR_VOID("suzerain_b200 failure", SUZERAIN_EFAILED); }

    szb_bsplineop* bop_; szb_imexop* op_; szb_wavegrid wg_; szb_zgbsv_spec spec_; bool one_sided_;
};

}}

#endif
