// Instantiates include/suzerain_b200_shim.hpp against the stand-in reference types: every member
// function is compiled and the program links against libsuzerain_b200.so (tests/test_shim_compiles.py).
#include "operator_hybrid_isothermal.hpp"
#include "suzerain_b200_shim.hpp"

int main(int argc, char **)
{
    using namespace suzerain;
    using namespace suzerain::perfect;
    if (argc < 100) return 0;          // compiled and linked, never run on a machine without a GPU
    specification_zgbsv spec; definition_scenario sc = {}; specification_isothermal iso = {}; specification_grid grid = {};
    pencil_grid dgrid = {}; bsplineop cop = { 0 }; bspline b; operator_common_block common = { linearize::rhome_xyz, references() };
    operator_hybrid_isothermal_b200 L(spec, sc, iso, grid, dgrid, cop, b, common);
    multi_array::ref<complex_t,4> a = { 0, { 5, 1, 1, 1 } };
    contiguous_state<4,complex_t> out = { 0, { 1, 1, 1, 1 } };
    lowstorage::method_interface<complex_t> m;
    L.apply_mass_plus_scaled_operator(complex_t(1, 0), a, 0);
    L.accumulate_mass_plus_scaled_operator(complex_t(1, 0), a, complex_t(0, 0), out, 0);
    L.invert_mass_plus_scaled_operator(complex_t(1, 0), a, m, 1.0, 0);
    return 0;
}
