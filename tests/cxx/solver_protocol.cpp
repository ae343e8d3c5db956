// Drives suzerain_b200::bsmbsm_solver_b200 through the reference's protocol
// (supply_B -> fill PAPT -> supplied_PAPT -> solve('T') -> demand_X; apps/perfect/operator_hybrid_isothermal.cpp:646-674)
// on a system read from a file; writes the solution and the pivots.  Built and run by tests/test_gpu_round2.py.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#include "suzerain_b200_solver.hpp"

using namespace suzerain_b200;

int main(int argc, char **argv)
{
    if (argc < 4) return 2;
    FILE *f = std::fopen(argv[1], "rb");
    if (!f) return 3;
    int hdr[5];
    if (std::fread(hdr, sizeof(int), 5, f) != 5) return 4;
    const int S = hdr[0], n = hdr[1], kl = hdr[2], ku = hdr[3], nrhs = hdr[4];
    const szb_bsmbsm A = szb_bsmbsm_construct(S, n, kl, ku);
    std::vector<complex_double> b((size_t) A.N * nrhs), papt((size_t) A.N * A.LD), x((size_t) A.N * nrhs);
    if (std::fread(b.data(), sizeof(complex_double), b.size(), f) != b.size()) return 4;
    if (std::fread(papt.data(), sizeof(complex_double), papt.size(), f) != papt.size()) return 4;
    std::fclose(f);
    szb_zgbsv_spec spec = szb_zgbsv_spec_default();
    spec.method = std::strcmp(argv[3], "zgbsv") == 0 ? SZB_SOLVER_ZGBSV : SZB_SOLVER_ZCGBSVX;
    std::unique_ptr<bsmbsm_solver_b200> s(bsmbsm_solver_b200::build(A, spec, nrhs));
    if (s->PAPT.colStride() != (s->in_place() ? A.LD + A.KL : A.LD)) return 5;
    s->supply_B(b.data());
    for (int j = 0; j < A.N; ++j)
        for (int i = 0; i < A.LD; ++i) s->PAPT(i, j) = papt[(size_t) j * A.LD + i];
    s->supplied_PAPT();
    const int info = s->solve('T');
    s->demand_X(x.data());
    f = std::fopen(argv[2], "wb");
    if (!f) return 3;
    std::fwrite(&info, sizeof(int), 1, f);
    std::fwrite(s->ipiv.data(), sizeof(int), A.N, f);
    std::fwrite(x.data(), sizeof(complex_double), x.size(), f);
    std::fclose(f);
    return 0;
}
