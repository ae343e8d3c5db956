// szb_internal.hpp -- private definitions shared by the translation units of
// libsuzerain_b200.so.  Nothing here is part of the C ABI (include/suzerain_b200.h).
#pragma once

#include <cstddef>
#include <cstdint>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/suzerain_b200.h"

#define SZB_CUDA_OK(expr)                                                   \
    do {                                                                    \
        cudaError_t szb_err__ = (expr);                                     \
        if (szb_err__ != cudaSuccess) {                                     \
            szb::report_cuda(szb_err__, #expr, __FILE__, __LINE__);         \
            return SZB_ECUDA_BASE - (int) szb_err__;                        \
        }                                                                   \
    } while (0)

namespace szb {

void report_cuda(cudaError_t e, const char *what, const char *file, int line);
void count_launch(unsigned n = 1);
void field_ctx_free(void *p);      // capi.cu: cached whole-field plan of an operator context
struct cplx;
int accumulate_launch(const szb_imexop *op, const double phi[2], int npencil, const double *d_km, const double *d_kn,
                      const int *d_index, const int *d_index_out, int out_plain, const szb_complex *d_in, size_t in_fs, size_t in_ps,
                      const double beta[2], szb_complex *d_out, size_t out_fs, size_t out_ps, void *stream,
                      const int *d_count = nullptr);
int invert_pipe_dispatch(const szb_imexop *op, const double phi[2], int npencil,
                         const double *d_km, const double *d_kn, const int *d_index,
                         cplx *d_state, size_t fs, size_t ps, int *d_ipiv, int *d_info,
                         int *d_iters, cudaStream_t stream, int zero_wall_rhs = 1, const int *d_count = nullptr);
int invert_sync_dispatch(const szb_imexop *op, const double phi[2], int npencil,
                         const double *d_km, const double *d_kn, const int *d_index,
                         cplx *d_state, size_t fs, size_t ps, int *d_ipiv, int *d_info,
                         int *d_iters, cudaStream_t stream, int zero_wall_rhs = 1, const int *d_count = nullptr);
// the fused zgbsv kernel in use: v5 (invert_sync.cu) unless SZB_INVERT=v4 or v5 has no instantiation
// d_count: optional device-side pencil count (<= npencil), read by the kernel instead of a host sync
int invert_fused_dispatch(const szb_imexop *op, const double phi[2], int npencil,
                          const double *d_km, const double *d_kn, const int *d_index,
                          cplx *d_state, size_t fs, size_t ps, int *d_ipiv, int *d_info,
                          int *d_iters, cudaStream_t stream, int zero_wall_rhs = 1, const int *d_count = nullptr);
int invert_refined_dispatch(const szb_imexop *op, int mode, int aiter, int dmax, const double phi[2], int npencil,
                            const double *d_km, const double *d_kn, const int *d_index,
                            cplx *d_state, size_t fs, size_t ps, int *d_ipiv, int *d_info,
                            int *d_iters, cudaStream_t stream);
// stage 1: save b + first solve of pencils at positions pos0.. of a list of `capacity`; stage 2: the refinement
// loop over positions 0 .. npencil - 1 (stages = 1 | 2: both, as invert_refined_dispatch)
int invert_refined_stage(const szb_imexop *op, int mode, int aiter, int dmax, const double phi[2], int npencil,
                         const double *d_km, const double *d_kn, const int *d_index,
                         cplx *d_state, size_t fs, size_t ps, int *d_ipiv, int *d_info,
                         int *d_iters, cudaStream_t stream, int stages, int pos0, int capacity);
int invert00_dispatch(const szb_imexop *op, int mode, int aiter, int dmax, const double phi[2], int npencil,
                      const int *d_index, cplx *d_state, size_t fs, size_t ps, int nextra, cplx *d_extra,
                      int *d_ipiv, int *d_info, int *d_iters, cudaStream_t stream);
// Term table dimensions (rholut_terms.def)
enum Field { E = 0, U = 1, V = 2, W = 3, R = 4, NFIELD = 5 };
enum Oper  { M = 0, D1 = 1, D2 = 2, NOPER = 3 };
// wave(km,kn) factors and reference-profile ids named as in rholut_terms.def
namespace wavid { enum { ONE = 0, IKM, IKN, KM2, KN2, KMKN, K2, NWAVE }; }
namespace refid {
enum { ux, uy, uz, u2, uxux, uxuy, uxuz, uyuy, uyuz, uzuz, nu, nuux, nuuy,
       nuuz, nuu2, nuuxux, nuuxuy, nuuxuz, nuuyuy, nuuyuz, nuuzuz,
       ex_gradrho, ey_gradrho, ez_gradrho, e_divm, e_deltarho,
       ONE /* pseudo-profile of ones appended to the 26 */ };
}
enum { REF_ONE = refid::ONE };
static_assert(REF_ONE == SZB_NREF, "reference profile count");
enum { MAXTERMS = 128, NBLOCK = NFIELD * NFIELD * NOPER };

// Flattened term table living in __constant__ memory (one copy per context
// generation; scenario constants are folded in on the host).
struct TermTable {
    int     nterms;
    double  sc[MAXTERMS];              // scenario factor of each term
    uint8_t ref[MAXTERMS];             // 0..25 or REF_ONE
    uint8_t wave[MAXTERMS];            // Wave
    uint8_t blk_begin[NBLOCK + 1];     // block b = (row*5 + col)*3 + op owns
                                       // terms [blk_begin[b], blk_begin[b+1])
};

}  // namespace szb

struct szb_bsplineop {
    int k, n, nderiv;
    std::vector<int> kl, ku;
    int max_kl, max_ku, ld;
    std::vector<double> knots;         // n + k
    std::vector<double> greville;      // n
    std::vector<double> storage;       // (nderiv+1) * ld * n, reference layout
    const double *D_T(int d) const { return storage.data() + (size_t) d * ld * n + (max_ku - ku[d]); }
    mutable double *d_Dr = nullptr;    // device copy, r-major (auxops.cu), made on first batched apply
    mutable int d_dev = -1;
};

struct szb_imexop {
    // geometry
    int n, k, kl, ku, ld;              // per-block bandwidths are the max over D_T[0..2]
    szb_bsmbsm A;                      // S = 5
    // device tables
    double *d_D;                       // [3][n][ld]: d_D[(d*n + i)*ld + ku + j - i] = D^(d)[i, j]
    double *d_refs;                    // [27][n]; row 26 is all ones
    szb::TermTable *d_terms;           // device copy of the term table
    szb::TermTable  h_terms;
    // scenario & boundary data
    szb_rholut_imexop_scenario scen;
    szb_isothermal iso;
    double E_factor[2], vel_factor[2][3];
    bool   have_a, have_b, have_c;
    double nrbc_a[25], nrbc_b[25], nrbc_c[25];
    // scratch for the batched invert: persistent per-CTA LU workspaces
    mutable void  *d_work;
    mutable size_t work_bytes;
    mutable int    work_slots;
    mutable void  *field_ctx;          // cached whole-field plan (capi.cu)
    mutable void  *d_refine;           // workspace of the refined fused invert (invert_pipe.cu)
    mutable size_t refine_bytes;
    mutable void  *d_work00;           // rhome_y: the one factorisation, U rows, pivots (rhome_y.cu)
    mutable size_t work00_bytes;
    mutable double *d_zero;            // zero wavenumbers for the general kernels under rhome_y
    mutable size_t zero_count;
    int linearization;                 // SZB_LINEARIZE_*
    int sm_count;
};
