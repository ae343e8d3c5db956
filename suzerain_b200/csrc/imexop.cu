// imexop.cu -- device-resident (M + phi L) operator context and the batched
// accumulate / apply and assembly kernels.
//
// Replaces suzerain_rholut_imexop_accumulate (suzerain/rholut_imexop.c:43-547),
// suzerain_rholut_imexop_pack{c,f} (suzerain/rholut_imexop.def:41-597),
// suzerain_bsmbsm_z{,d}pack (suzerain/bsmbsm_pack.def:37-123) and
// IsothermalPATPTEnforcer::op (apps/perfect/operator_hybrid_isothermal.cpp:
// 470-510) with kernels that work on every local (kx,kz) pencil in one launch.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <atomic>

#include "szb_internal.hpp"
#include "cplx.cuh"
#include <algorithm>
#include "kernels.cuh"
#include "invert_common.cuh"

namespace szb {

static std::atomic<unsigned long long> g_launches{0};
void count_launch(unsigned n) { g_launches += n; }

void report_cuda(cudaError_t e, const char *what, const char *file, int line)
{
    std::fprintf(stderr, "suzerain_b200: CUDA error %d (%s) at %s:%d: %s\n",
                 (int) e, cudaGetErrorString(e), file, line, what);
}

// ---------------------------------------------------------------------------
// Term table construction (host): fold the scenario into per-term constants.
// ---------------------------------------------------------------------------
static void build_terms(const szb_rholut_imexop_scenario &s, TermTable &tt)
{
    // Shorthands as in rholut_imexop.c:85-97
    const double g        = s.gamma;
    const double gm1      = s.gamma - 1;
    const double gm3      = s.gamma - 3;
    const double ap43     = s.alpha + 4.0 / 3.0;
    const double ap13     = s.alpha + 1.0 / 3.0;
    const double Ma2      = s.Ma * s.Ma;
    const double invRe    = 1 / s.Re;
    const double invMa2   = 1 / Ma2;
    const double ginvPr   = s.gamma / s.Pr;
    const double ginvRePr = s.gamma / (s.Re * s.Pr);
    (void) g; (void) gm3; (void) invMa2;

    std::memset(&tt, 0, sizeof(tt));
    int t = 0, last_blk = -1;
    for (int b = 0; b <= NBLOCK; ++b) tt.blk_begin[b] = 0;
    auto add = [&](int row, int col, int op, int ref, int wave, double sc) {
        const int blk = (row * NFIELD + col) * NOPER + op;
        // blocks arrive in increasing order; close all blocks up to this one
        for (int b = last_blk + 1; b <= blk; ++b) tt.blk_begin[b] = (uint8_t) t;
        last_blk = blk;
        tt.sc[t] = sc; tt.ref[t] = (uint8_t) ref; tt.wave[t] = (uint8_t) wave;
        ++t;
    };
#define SZB_TERM(row, col, op, ref, wave, scen) \
    add(szb::row, szb::col, szb::op, refid::ref, wavid::wave, (scen));
#include "rholut_terms.def"
#undef SZB_TERM
    for (int b = last_blk + 1; b <= NBLOCK; ++b) tt.blk_begin[b] = (uint8_t) t;
    tt.nterms = t;
}

// ---------------------------------------------------------------------------
// accumulate: out <- (M + phi L) in + beta out, one CTA per pencil.
//
// Shared memory holds the five input pencils (zero halo) and the per-term coefficients
// alpha_t = phi * sc_t * wave_t(km, kn); one thread per collocation point y:
//   P[d][j](y) = sum_r D^(d)[y, y - ku + r] * in_j[y - ku + r]      fifteen products in registers
//   out_i(y)   = beta out_i + sum_{blocks (i,j,d)} (sum_t alpha_t ref_t[y]) P[d][j]
//                + P[M][i]                                (mass added last)
// The block structure of each row is unrolled at compile time from rholut_terms.def.
// HBM traffic is the algorithmic minimum: every state element is read once and written
// once (plus one read of the output when beta != 0); operators and profiles stay in L1/L2.
// ---------------------------------------------------------------------------
struct AccumulateArgs {
    const double *D;        // [3][ld][n]  (r-major: D[(d*ld + r)*n + y])
    const double *refs;     // [27][n]
    const TermTable *terms;
    int n, kl, ku, ld;
    cplx phi, beta;
    const double *km, *kn; const int *index;
    const cplx *in;  size_t in_fs,  in_ps;
    cplx       *out; size_t out_fs, out_ps;
    int nrbc;               // bit0 a, bit1 b, bit2 c
    double a[25], b[25], c[25];
    int npencil;
    const int *count_dev;   // optional device-side pencil count (<= npencil)
    int zero_wave;          // linearize::rhome_y: the operator at km = kn = 0 for every pencil
    const int *index_out;   // slot of the output pencil when it differs from the input's (null: the same ...
    int out_plain;          // ... or, with this flag, the pencil's own number)
};

// One output row of phi L at collocation point y, statically specialised on the equation
// ROW: only that row's terms of rholut_terms.def survive constant folding.  Pm[op][col] are
// this point's fifteen banded products, held in registers.
template <int ROW>
__device__ __forceinline__ cplx accumulate_row(const double *refs, int n, int y, const cplx *s_alpha,
                                               const cplx (&Pm)[3][5])
{
    cplx acc(0.0, 0.0), c(0.0, 0.0);
    int t = 0, cur = -1, ccol = 0, cop = 0;
#define SZB_FLUSH() do { if (cur >= 0) acc += c * Pm[cop][ccol]; } while (0)
#define SZB_TERM(row, col, op, ref, wave, scen)                                  \
    if (szb::row == ROW) {                                                       \
        if ((szb::col) * 3 + szb::op != cur) {                                   \
            SZB_FLUSH();                                                         \
            cur = (szb::col) * 3 + szb::op; ccol = szb::col; cop = szb::op;      \
            c = cplx(0.0, 0.0);                                                  \
        }                                                                        \
        c += s_alpha[t] * __ldg(refs + (size_t) refid::ref * n + y);             \
    }                                                                            \
    ++t;
#include "rholut_terms.def"
#undef SZB_TERM
    SZB_FLUSH();
#undef SZB_FLUSH
    return acc;
}

// One CTA per pencil, one thread per collocation point y.  The thread forms the fifteen
// banded products D^(0..2) in_j at y in registers (every operator entry is loaded once and
// used for the five fields; the fields sit in shared memory with a zero halo of ku / kl
// points, so there are no bounds tests) and combines them into the five output rows
// straight away: no second pass, no products in shared memory.
template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
accumulate_kernel(const AccumulateArgs A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long s_mbar[2];
    const int n = A.n, ku = A.ku, kl = A.kl, np = n + ku + kl;
    cplx *s_inbuf = reinterpret_cast<cplx *>(smem_raw);            // [2][5][ku + n + kl]
    cplx *s_alpha = s_inbuf + 2 * 5 * np;                          // [MAXTERMS]
    cplx *s_top   = s_alpha + MAXTERMS;                            // [5] phi L in at the upper boundary

    // Persistent CTA: the five field pencils of the next (kx,kz) arrive by TMA bulk copies
    // (one per field, interior of a zero-halo buffer) while the current one is computed.
    for (int e = threadIdx.x; e < 2 * 5 * np; e += blockDim.x) s_inbuf[e] = cplx(0.0, 0.0);
    if (threadIdx.x == 0) {
        fused::mbar_init(&s_mbar[0], 1); fused::mbar_init(&s_mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const unsigned bytes = (unsigned) (n * sizeof(cplx));
    auto fetch = [&](int p, int stage) {
        const size_t slot = A.index ? (size_t) A.index[p] : (size_t) p;
        const cplx *src = A.in + slot * A.in_ps;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        fused::mbar_expect_tx(&s_mbar[stage], 5 * bytes);
        for (int f = 0; f < 5; ++f)
            fused::tma_bulk_g2s(s_inbuf + (stage * 5 + f) * np + ku, src + (size_t) f * A.in_fs, bytes, &s_mbar[stage]);
    };
    const int npen = A.count_dev ? min(A.npencil, __ldg(A.count_dev)) : A.npencil;
    if (threadIdx.x == 0 && (int) blockIdx.x < npen) fetch(blockIdx.x, 0);
    const int nterms = A.terms->nterms;
    int it = 0;
    for (int p = blockIdx.x; p < npen; p += gridDim.x, ++it) {
    const int stage = it & 1;
    const cplx *s_in = s_inbuf + stage * 5 * np;
    const double km = A.zero_wave ? 0.0 : A.km[p], kn = A.zero_wave ? 0.0 : A.kn[p];
    const size_t slot = A.index ? (size_t) A.index[p] : (size_t) p;
    cplx *out = A.out + (A.index_out ? (size_t) A.index_out[p] : A.out_plain ? (size_t) p : slot) * A.out_ps;
    if (threadIdx.x == 0 && p + (int) gridDim.x < npen) fetch(p + gridDim.x, stage ^ 1);
    for (int t = threadIdx.x; t < nterms; t += blockDim.x)
        s_alpha[t] = A.phi * (wave_factor(A.terms->wave[t], km, kn) * A.terms->sc[t]);
    fused::mbar_wait(&s_mbar[stage], (it >> 1) & 1);
    __syncthreads();

    const bool beta_zero = is_zero(A.beta);
    const bool nrbc = A.nrbc != 0;
    for (int y = threadIdx.x; y < n; y += blockDim.x) {
        // ---- banded products: P[d][j](y) = sum_r D^(d)[y, y-ku+r] in_j[y-ku+r] ----
        cplx Pm[3][5];
#pragma unroll
        for (int d = 0; d < 3; ++d)
#pragma unroll
            for (int j = 0; j < 5; ++j) Pm[d][j] = cplx(0.0, 0.0);
        const double *D0 = A.D + y;
        const size_t ds = (size_t) A.ld * n;
#pragma unroll 2
        for (int r = 0; r < A.ld; ++r) {
            // entries outside the matrix are zero in the band storage and meet the zero halo
            const double d0 = __ldg(D0 + (size_t) r * n), d1 = __ldg(D0 + ds + (size_t) r * n),
                         d2 = __ldg(D0 + 2 * ds + (size_t) r * n);
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const cplx v = s_in[j * np + y + r];
                addmul(Pm[0][j], v, d0);
                addmul(Pm[1][j], v, d1);
                addmul(Pm[2][j], v, d2);
            }
        }
        // ---- rows: out_i = beta out_i + phi L in (+ NRBC) + M in_i, mass added last ----
        cplx phiL[5];
        phiL[0] = accumulate_row<0>(A.refs, n, y, s_alpha, Pm);
        phiL[1] = accumulate_row<1>(A.refs, n, y, s_alpha, Pm);
        phiL[2] = accumulate_row<2>(A.refs, n, y, s_alpha, Pm);
        phiL[3] = accumulate_row<3>(A.refs, n, y, s_alpha, Pm);
        phiL[4] = accumulate_row<4>(A.refs, n, y, s_alpha, Pm);
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            cplx *o = out + (size_t) i * A.out_fs + y;
            cplx acc = beta_zero ? phiL[i] : A.beta * (*o) + phiL[i];
            acc += Pm[0][i];                                         // M in_i (d = 0, j = i)
            if (nrbc && y == n - 1) s_top[i] = phiL[i];
            *o = acc;
        }
    }
    if (nrbc) {
        // upper-boundary correction (rholut_imexop.c:510-545):
        //   t = -i phi (km a + kn b) in(top) - c (phi L in)(top)
        __syncthreads();
        const int i = threadIdx.x;
        if (i < 5) {
            cplx tt(0.0, 0.0);
            const cplx ikmphi = cplx(0.0, km) * A.phi, iknphi = cplx(0.0, kn) * A.phi;
            if (A.nrbc & 1) {
                cplx s(0.0, 0.0);
                for (int j = 0; j < 5; ++j) s += s_in[j * np + ku + (n - 1)] * A.a[i + 5 * j];
                tt -= ikmphi * s;
            }
            if (A.nrbc & 2) {
                cplx s(0.0, 0.0);
                for (int j = 0; j < 5; ++j) s += s_in[j * np + ku + (n - 1)] * A.b[i + 5 * j];
                tt -= iknphi * s;
            }
            if (A.nrbc & 4) {
                cplx s(0.0, 0.0);
                for (int j = 0; j < 5; ++j) s += s_top[j] * A.c[i + 5 * j];
                tt -= s;
            }
            cplx *o = out + (size_t) i * A.out_fs + (n - 1);
            *o = *o + tt;
        }
    }
    __syncthreads();          // alpha, s_top and this stage's buffer are free again
    }
}

// ---------------------------------------------------------------------------
// pack: assemble P (M + phi L)^T P^T into LAPACK band storage, one CTA per
// pencil (see pack_pencil in kernels.cuh).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pack_kernel(const PackArgs A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx *s_alpha = reinterpret_cast<cplx *>(smem_raw);
    __shared__ cplx s_x[75];
    const int p = blockIdx.x;
    cplx *M = A.out + (size_t) p * A.N * A.rows + A.rowoff;
    pack_pencil(A, A.rows, A.km[p], A.kn[p], s_alpha, s_x, M);
}

// accumulate with separate pencil slots for input and output (the refinement residual writes a compact buffer)
int accumulate_launch(const szb_imexop *op, const double phi[2],
        int npencil, const double *d_km, const double *d_kn, const int *d_index, const int *d_index_out, int out_plain,
        const szb_complex *d_in, size_t in_fs, size_t in_ps,
        const double beta[2],
        szb_complex *d_out, size_t out_fs, size_t out_ps, void *stream, const int *d_count)
{
    if (!op) return -1;
    if (!phi) return -2;
    if (npencil < 0) return -3;
    if (!d_km) return -4;
    if (!d_kn) return -5;
    if (!d_in) return -7;
    if (!beta) return -10;
    if (!d_out) return -11;
    if (npencil == 0) return 0;
    if ((const void *) d_in == (const void *) d_out
        && !(beta[0] == 0.0 && beta[1] == 0.0 && in_fs == out_fs && in_ps == out_ps)) return -11;
    AccumulateArgs A;
    A.D = op->d_D; A.refs = op->d_refs; A.terms = op->d_terms;
    A.n = op->n; A.kl = op->kl; A.ku = op->ku; A.ld = op->ld;
    A.phi = cplx(phi[0], phi[1]); A.beta = cplx(beta[0], beta[1]);
    A.km = d_km; A.kn = d_kn; A.index = d_index; A.index_out = d_index_out; A.out_plain = out_plain;
    A.in = reinterpret_cast<const cplx *>(d_in); A.in_fs = in_fs; A.in_ps = in_ps;
    A.out = reinterpret_cast<cplx *>(d_out); A.out_fs = out_fs; A.out_ps = out_ps;
    A.nrbc = (op->have_a ? 1 : 0) | (op->have_b ? 2 : 0) | (op->have_c ? 4 : 0);
    std::memcpy(A.a, op->nrbc_a, sizeof(A.a));
    std::memcpy(A.b, op->nrbc_b, sizeof(A.b));
    std::memcpy(A.c, op->nrbc_c, sizeof(A.c));
    A.npencil = npencil; A.count_dev = d_count;
    A.zero_wave = op->linearization == SZB_LINEARIZE_RHOME_Y;
    const size_t smem = sizeof(cplx) * (10 * (size_t) (op->n + op->kl + op->ku) + MAXTERMS + 8);
    if (smem > 227 * 1024) return -1;
    int threads = (op->n + 31) / 32 * 32;
    if (threads > 512) threads = 512;
    if (threads <= 128) {
        if (smem > 48 * 1024)
            SZB_CUDA_OK(cudaFuncSetAttribute(accumulate_kernel<128, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        accumulate_kernel<128, 4><<<std::min(npencil, 5 * op->sm_count), threads, smem, (cudaStream_t) stream>>>(A);
    } else {
        if (smem > 48 * 1024)
            SZB_CUDA_OK(cudaFuncSetAttribute(accumulate_kernel<512, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        accumulate_kernel<512, 1><<<std::min(npencil, 2 * op->sm_count), threads, smem, (cudaStream_t) stream>>>(A);
    }
    count_launch();
    SZB_CUDA_OK(cudaGetLastError());
    return 0;
}


}  // namespace szb

using namespace szb;

// ---------------------------------------------------------------------------
// C ABI: context management
// ---------------------------------------------------------------------------
namespace szb {
// references -> dense profile table: rows q::u .. q::e_deltarho of the reference's 42 x Ny column-major
// `references` block (apps/perfect/references.hpp:82-125) are exactly the 26 profiles of
// references::rholut_imexop (references.cpp:50-108), in that order.
__global__ void gather_references_kernel(int n, const double *src, int ld, double *dst)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= SZB_NREF * n) return;
    const int q = e / n, y = e - q * n;
    dst[e] = src[(size_t) y * ld + SZB_REFERENCES_FIRST + q];
}
}  // namespace szb

extern "C" {

const char *szb_version(void) { return "suzerain_b200 0.1 (sm_100a, FP64)"; }
unsigned long long szb_launch_count(void) { return szb::g_launches.load(); }

int szb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

szb_zgbsv_spec szb_zgbsv_spec_default(void)
{
    szb_zgbsv_spec s;
    s.method = SZB_SOLVER_ZCGBSVX; s.aiter = 1; s.diter = 5; s.tolsc = 0.0;
    s.equil = 0; s.reuse = 0; s.siter = -1;
    return s;
}

int szb_imexop_create(const szb_bsplineop *w, szb_imexop **out)
{
    if (!w) return -1;
    if (!out) return -2;
    if (w->nderiv < 2) return -1;                   // rholut_imexop.c:77
    szb_imexop *op = new (std::nothrow) szb_imexop();
    if (!op) return -2;
    std::memset(&op->scen, 0, sizeof(op->scen));
    op->n = w->n; op->k = w->k; op->kl = w->max_kl; op->ku = w->max_ku; op->ld = w->ld;
    op->A = szb_bsmbsm_construct(5, w->n, w->max_kl, w->max_ku);
    op->d_D = nullptr; op->d_refs = nullptr; op->d_terms = nullptr;
    op->d_work = nullptr; op->work_bytes = 0; op->work_slots = 0; op->field_ctx = nullptr;
    op->d_refine = nullptr; op->refine_bytes = 0;
    op->d_work00 = nullptr; op->work00_bytes = 0; op->d_zero = nullptr; op->zero_count = 0;
    op->linearization = SZB_LINEARIZE_RHOME_XYZ;
    op->have_a = op->have_b = op->have_c = false;
    std::memset(&op->iso, 0, sizeof(op->iso));
    op->iso.enforce_lower = 1; op->iso.enforce_upper = 1;
    for (int i = 0; i < 2; ++i) { op->E_factor[i] = 0; for (int j = 0; j < 3; ++j) op->vel_factor[i][j] = 0; }

    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { delete op; report_cuda(e, "cudaGetDevice", __FILE__, __LINE__); return SZB_ECUDA_BASE - (int) e; }
    cudaDeviceGetAttribute(&op->sm_count, cudaDevAttrMultiProcessorCount, dev);

    // Operators in r-major order with the common (max) bandwidth view:
    // h[(d*ld + r)*n + i] = D^(d)[i, i - ku + r] = (D_T[d] - (max_ku - ku[d]))[i*ld + r]
    const int n = op->n, ld = op->ld;
    std::vector<double> h((size_t) 3 * ld * n, 0.0);
    for (int d = 0; d < 3; ++d) {
        const double *blk = w->storage.data() + (size_t) d * ld * n;
        for (int i = 0; i < n; ++i)
            for (int r = 0; r < ld; ++r)
                h[((size_t) d * ld + r) * n + i] = blk[(size_t) i * ld + r];
    }
    std::vector<double> ones((size_t) (SZB_NREF + 1) * n, 0.0);
    for (int i = 0; i < n; ++i) ones[(size_t) SZB_NREF * n + i] = 1.0;
    if ((e = cudaMalloc(&op->d_D, h.size() * sizeof(double))) != cudaSuccess ||
        (e = cudaMalloc(&op->d_refs, ones.size() * sizeof(double))) != cudaSuccess ||
        (e = cudaMalloc(&op->d_terms, sizeof(TermTable))) != cudaSuccess ||
        (e = cudaMemcpy(op->d_D, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(op->d_refs, ones.data(), ones.size() * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess) {
        report_cuda(e, "imexop_create", __FILE__, __LINE__);
        szb_imexop_destroy(op);
        return SZB_ECUDA_BASE - (int) e;
    }
    *out = op;
    return 0;
}

void szb_imexop_destroy(szb_imexop *op)
{
    if (!op) return;
    if (op->field_ctx) szb::field_ctx_free(op->field_ctx);
    cudaFree(op->d_D); cudaFree(op->d_refs); cudaFree(op->d_terms); cudaFree(op->d_work); cudaFree(op->d_refine); cudaFree(op->d_work00); cudaFree(op->d_zero);
    delete op;
}

szb_bsmbsm szb_imexop_bsmbsm(const szb_imexop *op) { return op->A; }

int szb_imexop_set_linearization(szb_imexop *op, int linearization)
{
    if (!op) return -1;
    if (linearization != SZB_LINEARIZE_RHOME_XYZ && linearization != SZB_LINEARIZE_RHOME_Y) return -2;
    op->linearization = linearization;
    return 0;
}

int szb_imexop_set_scenario(szb_imexop *op, const szb_rholut_imexop_scenario *s)
{
    if (!op) return -1;
    if (!s) return -2;
    op->scen = *s;
    build_terms(op->scen, op->h_terms);
    SZB_CUDA_OK(cudaMemcpy(op->d_terms, &op->h_terms, sizeof(TermTable), cudaMemcpyHostToDevice));
    // E_factor depends on gamma and Ma as well as the wall data
    return szb_imexop_set_isothermal(op, &op->iso);
}

int szb_imexop_set_refs(szb_imexop *op, const szb_rholut_imexop_ref *r,
                        const szb_rholut_imexop_refld *ld)
{
    if (!op) return -1;
    if (!r) return -2;
    if (!ld) return -3;
    const int n = op->n;
    double *const *ptrs = reinterpret_cast<double *const *>(r);
    const int *lds = reinterpret_cast<const int *>(ld);
    std::vector<double> h((size_t) SZB_NREF * n);
    for (int q = 0; q < SZB_NREF; ++q) {
        if (!ptrs[q]) return -2;
        for (int i = 0; i < n; ++i) h[(size_t) q * n + i] = ptrs[q][(size_t) i * lds[q]];
    }
    SZB_CUDA_OK(cudaMemcpy(op->d_refs, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
    return 0;
}

int szb_imexop_set_refs_device(szb_imexop *op, const double *d_references, int ld, void *stream)
{
    if (!op) return -1;
    if (!d_references) return -2;
    if (ld < SZB_REFERENCES_FIRST + SZB_NREF) return -3;
    const int total = SZB_NREF * op->n;
    szb::gather_references_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t) stream>>>(op->n, d_references, ld, op->d_refs);
    szb::count_launch();
    SZB_CUDA_OK(cudaGetLastError());
    return 0;
}

int szb_imexop_set_isothermal(szb_imexop *op, const szb_isothermal *iso)
{
    if (!op) return -1;
    if (!iso) return -2;
    op->iso = *iso;
    const double g = op->scen.gamma, Ma = op->scen.Ma;
    // operator_hybrid_isothermal.cpp:448-461
    op->E_factor[0] = iso->lower_T / (g * (g - 1))
                    + Ma * Ma / 2 * (iso->lower_u * iso->lower_u + iso->lower_v * iso->lower_v + iso->lower_w * iso->lower_w);
    op->E_factor[1] = iso->upper_T / (g * (g - 1))
                    + Ma * Ma / 2 * (iso->upper_u * iso->upper_u + iso->upper_v * iso->upper_v + iso->upper_w * iso->upper_w);
    op->vel_factor[0][0] = iso->lower_u; op->vel_factor[0][1] = iso->lower_v; op->vel_factor[0][2] = iso->lower_w;
    op->vel_factor[1][0] = iso->upper_u; op->vel_factor[1][1] = iso->upper_v; op->vel_factor[1][2] = iso->upper_w;
    return 0;
}

int szb_imexop_set_nrbc(szb_imexop *op, const double *a, const double *b, const double *c)
{
    if (!op) return -1;
    op->have_a = a != nullptr; op->have_b = b != nullptr; op->have_c = c != nullptr;
    if (a) std::memcpy(op->nrbc_a, a, sizeof(op->nrbc_a));
    if (b) std::memcpy(op->nrbc_b, b, sizeof(op->nrbc_b));
    if (c) std::memcpy(op->nrbc_c, c, sizeof(op->nrbc_c));
    return 0;
}

// ---------------------------------------------------------------------------
// C ABI: batched accumulate / pack
// ---------------------------------------------------------------------------
int szb_imexop_accumulate_batch(const szb_imexop *op, const double phi[2],
        int npencil, const double *d_km, const double *d_kn, const int *d_index,
        const szb_complex *d_in, size_t in_fs, size_t in_ps,
        const double beta[2],
        szb_complex *d_out, size_t out_fs, size_t out_ps, void *stream)
{
    return szb::accumulate_launch(op, phi, npencil, d_km, d_kn, d_index, nullptr, 0, d_in, in_fs, in_ps, beta,
                                  d_out, out_fs, out_ps, stream);
}

int szb_imexop_pack_batch(const szb_imexop *op, const double phi[2],
        int npencil, const double *d_km, const double *d_kn,
        int packf, int with_bc, szb_complex *d_patpt, void *stream)
{
    if (!op) return -1;
    if (!phi) return -2;
    if (npencil < 0) return -3;
    if (!d_km) return -4;
    if (!d_kn) return -5;
    if (!d_patpt) return -8;
    if (npencil == 0) return 0;
    PackArgs A;
    szb::fill_pack_args(op, phi, d_km, d_kn, packf, with_bc, reinterpret_cast<cplx *>(d_patpt), A);
    pack_kernel<<<npencil, 256, sizeof(cplx) * MAXTERMS, (cudaStream_t) stream>>>(A);
    count_launch();
    SZB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // extern "C"
