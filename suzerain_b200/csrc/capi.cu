// capi.cu -- HOST-pointer entry points of the C ABI: per-pencil wrappers with the
// reference's signatures, the three whole-field virtuals of
// operator_hybrid_isothermal, and the batched B-spline operator apply.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <thread>
#include <vector>

#include "szb_internal.hpp"
#include "cplx.cuh"

using namespace szb;

namespace {

// RAII device buffer
template <class T> struct DevBuf {
    T *p = nullptr;
    size_t count = 0;
    cudaError_t alloc(size_t n) { count = n; return cudaMalloc(&p, sizeof(T) * (n ? n : 1)); }
    ~DevBuf() { if (p) cudaFree(p); }
};

struct TempOp {
    szb_imexop *op = nullptr;
    ~TempOp() { szb_imexop_destroy(op); }
};

int make_temp_op(const szb_rholut_imexop_scenario *s, const szb_rholut_imexop_ref *r,
                 const szb_rholut_imexop_refld *ld, const szb_bsplineop *w,
                 const double *a, const double *b, const double *c, TempOp &t)
{
    int rc = szb_imexop_create(w, &t.op);
    if (rc) return rc;
    if ((rc = szb_imexop_set_scenario(t.op, s))) return rc;
    if ((rc = szb_imexop_set_refs(t.op, r, ld))) return rc;
    return szb_imexop_set_nrbc(t.op, a, b, c);
}

}  // namespace

extern "C" {

int szb_rholut_imexop_accumulate(const double phi[2], double km, double kn,
        const szb_rholut_imexop_scenario *s, const szb_rholut_imexop_ref *r,
        const szb_rholut_imexop_refld *ld, const szb_bsplineop *w,
        const szb_complex *in_rho_E, const szb_complex *in_rho_u,
        const szb_complex *in_rho_v, const szb_complex *in_rho_w,
        const szb_complex *in_rho, const double beta[2],
        szb_complex *out_rho_E, szb_complex *out_rho_u, szb_complex *out_rho_v,
        szb_complex *out_rho_w, szb_complex *out_rho,
        const double *a, const double *b, const double *c)
{
    if (!phi) return -1;
    if (!s) return -4;
    if (!r) return -5;
    if (!ld) return -6;
    if (!w) return -7;
    const szb_complex *in[5] = { in_rho_E, in_rho_u, in_rho_v, in_rho_w, in_rho };
    szb_complex *out[5] = { out_rho_E, out_rho_u, out_rho_v, out_rho_w, out_rho };
    for (int f = 0; f < 5; ++f) { if (!in[f]) return -(8 + f); if (!out[f]) return -(14 + f); }
    if (!beta) return -13;
    TempOp t;
    int rc = make_temp_op(s, r, ld, w, a, b, c, t);
    if (rc) return rc;
    const int n = w->n;
    std::vector<szb_complex> hin(5 * (size_t) n), hout(5 * (size_t) n);
    for (int f = 0; f < 5; ++f) {
        std::memcpy(&hin[(size_t) f * n], in[f], sizeof(szb_complex) * n);
        std::memcpy(&hout[(size_t) f * n], out[f], sizeof(szb_complex) * n);
    }
    DevBuf<szb_complex> din, dout; DevBuf<double> dk;
    SZB_CUDA_OK(din.alloc(5 * (size_t) n)); SZB_CUDA_OK(dout.alloc(5 * (size_t) n)); SZB_CUDA_OK(dk.alloc(2));
    const double k2[2] = { km, kn };
    SZB_CUDA_OK(cudaMemcpy(din.p, hin.data(), sizeof(szb_complex) * hin.size(), cudaMemcpyHostToDevice));
    SZB_CUDA_OK(cudaMemcpy(dout.p, hout.data(), sizeof(szb_complex) * hout.size(), cudaMemcpyHostToDevice));
    SZB_CUDA_OK(cudaMemcpy(dk.p, k2, sizeof(k2), cudaMemcpyHostToDevice));
    rc = szb_imexop_accumulate_batch(t.op, phi, 1, dk.p, dk.p + 1, nullptr, din.p, n, 5 * (size_t) n,
                                     beta, dout.p, n, 5 * (size_t) n, nullptr);
    if (rc) return rc;
    SZB_CUDA_OK(cudaMemcpy(hout.data(), dout.p, sizeof(szb_complex) * hout.size(), cudaMemcpyDeviceToHost));
    for (int f = 0; f < 5; ++f) std::memcpy(out[f], &hout[(size_t) f * n], sizeof(szb_complex) * n);
    return 0;
}

static int pack_host(int packf, const double phi[2], double km, double kn,
        const szb_rholut_imexop_scenario *s, const szb_rholut_imexop_ref *r,
        const szb_rholut_imexop_refld *ld, const szb_bsplineop *w,
        szb_bsmbsm *A_T, szb_complex *patpt,
        const double *a, const double *b, const double *c)
{
    if (!phi) return -1;
    if (!s) return -4;
    if (!r) return -5;
    if (!ld) return -6;
    if (!w) return -7;
    if (!A_T) return -8;
    if (!patpt) return -9;
    TempOp t;
    int rc = make_temp_op(s, r, ld, w, a, b, c, t);
    if (rc) return rc;
    *A_T = t.op->A;
    const size_t rows = packf ? A_T->LD + A_T->KL : A_T->LD;
    const size_t cnt = rows * A_T->N;
    DevBuf<szb_complex> dm; DevBuf<double> dk;
    SZB_CUDA_OK(dm.alloc(cnt)); SZB_CUDA_OK(dk.alloc(2));
    const double k2[2] = { km, kn };
    // the caller's storage may be NaN-poisoned outside the matrix: keep it
    SZB_CUDA_OK(cudaMemcpy(dm.p, patpt, sizeof(szb_complex) * cnt, cudaMemcpyHostToDevice));
    SZB_CUDA_OK(cudaMemcpy(dk.p, k2, sizeof(k2), cudaMemcpyHostToDevice));
    rc = szb_imexop_pack_batch(t.op, phi, 1, dk.p, dk.p + 1, packf, 0, dm.p, nullptr);
    if (rc) return rc;
    SZB_CUDA_OK(cudaMemcpy(patpt, dm.p, sizeof(szb_complex) * cnt, cudaMemcpyDeviceToHost));
    return 0;
}

int szb_rholut_imexop_packc(const double phi[2], double km, double kn,
        const szb_rholut_imexop_scenario *s, const szb_rholut_imexop_ref *r,
        const szb_rholut_imexop_refld *ld, const szb_bsplineop *w,
        szb_bsmbsm *A_T, szb_complex *patpt,
        const double *a, const double *b, const double *c)
{ return pack_host(0, phi, km, kn, s, r, ld, w, A_T, patpt, a, b, c); }

int szb_rholut_imexop_packf(const double phi[2], double km, double kn,
        const szb_rholut_imexop_scenario *s, const szb_rholut_imexop_ref *r,
        const szb_rholut_imexop_refld *ld, const szb_bsplineop *w,
        szb_bsmbsm *A_T, szb_complex *patpt,
        const double *a, const double *b, const double *c)
{ return pack_host(1, phi, km, kn, s, r, ld, w, A_T, patpt, a, b, c); }

}  // extern "C"

// ---------------------------------------------------------------------------
// Whole-field entry points
//
// The state arrives in (ideally pinned) HOST memory.  Only active pencils cross the
// bus: wave space is cut into chunks of consecutive active kz rows (one strided 2-D
// copy each), and chunk c+1 is uploaded while chunk c is computed and chunk c-1 is
// downloaded (three streams, one event pair per chunk).  Dealiased pencils never
// travel: invert zero-fills them on the host, apply/accumulate leave them alone.
// Plans, device mirrors, streams and events are cached on the operator context.
// ---------------------------------------------------------------------------
namespace {

struct Chunk {
    int row0, nrows;          // local kz rows [row0, row0 + nrows)
    int a0, a1;               // range of the active-pencil list
};

struct FieldCtx {
    szb_wavegrid g;
    int nx = 0, nz = 0, xa = 0, npencil = 0, nact = 0;
    int zero_zero = -1;                     // position of the (0,0) pencil in the active list
    std::vector<int> active_idx, inactive_idx;
    std::vector<Chunk> chunks;              // transfer-bound calls (apply, accumulate): eight equal chunks
    std::vector<Chunk> chunks_inv;          // invert: short first and last chunk, two long ones
    DevBuf<double> d_km, d_kn;
    DevBuf<int> d_act, d_info;
    DevBuf<szb_complex> d_a, d_b;           // interleaved mirror, contiguous-state mirror
    cudaStream_t s_in = nullptr, s_comp = nullptr, s_out = nullptr;
    std::vector<cudaEvent_t> ev_in, ev_done;
    cudaEvent_t ev_prev = nullptr;
    ~FieldCtx() {
        for (auto e : ev_in) cudaEventDestroy(e);
        for (auto e : ev_done) cudaEventDestroy(e);
        if (ev_prev) cudaEventDestroy(ev_prev);
        if (s_in) cudaStreamDestroy(s_in);
        if (s_comp) cudaStreamDestroy(s_comp);
        if (s_out) cudaStreamDestroy(s_out);
    }
};

int build_ctx(const szb_wavegrid *g, FieldCtx &F)
{
    F.g = *g;
    F.nx = g->dkex - g->dkbx; F.nz = g->dkez - g->dkbz;
    F.npencil = szb_wavegrid_npencils(g);
    std::vector<double> km(F.npencil), kn(F.npencil), akm, akn;
    std::vector<int> act(F.npencil);
    szb_wavegrid_wavenumbers(g, km.data(), kn.data(), act.data());
    // active kx form a prefix of every row (non-negative wavenumbers only); check it
    F.xa = 0;
    std::vector<int> row_active(F.nz, 0);
    for (int r = 0; r < F.nz; ++r) {
        int cnt = 0; bool prefix = true;
        for (int m = 0; m < F.nx; ++m) {
            if (act[(size_t) r * F.nx + m]) { if (cnt != m) prefix = false; ++cnt; }
        }
        if (!prefix) return -2;
        if (cnt) { if (F.xa && cnt != F.xa) return -2; F.xa = cnt; row_active[r] = 1; }
    }
    std::vector<int> row_a0(F.nz + 1, 0);
    for (int r = 0; r < F.nz; ++r) {
        row_a0[r] = (int) F.active_idx.size();
        for (int m = 0; m < F.nx; ++m) {
            const int p = r * F.nx + m;
            if (act[p]) {
                const int mm = g->dkbx + m, nn = g->dkbz + r;
                if (mm == 0 && nn == 0) F.zero_zero = (int) F.active_idx.size();
                F.active_idx.push_back(p); akm.push_back(km[p]); akn.push_back(kn[p]);
            } else {
                F.inactive_idx.push_back(p);
            }
        }
    }
    row_a0[F.nz] = (int) F.active_idx.size();
    F.nact = (int) F.active_idx.size();
    // Chunks of consecutive active rows.  apply / accumulate are transfer-bound: eight equal
    // chunks keep upload, compute and download overlapped.  For invert every chunk is one
    // launch of the persistent kernel and pays its tail (slots idling until the slowest pencil
    // of the chunk is done) while only the first upload and the last download are exposed: a
    // short first and last chunk (1/8 of the rows each) around one long one.
    int nrows_active = 0;
    for (int r = 0; r < F.nz; ++r) nrows_active += row_active[r];
    auto cut = [&](std::vector<int> sizes, std::vector<Chunk> &out) {
        size_t si = 0;
        int left = sizes.empty() ? nrows_active : std::max(sizes[0], 1);
        for (int r = 0; r < F.nz;) {
            if (!row_active[r]) { ++r; continue; }
            int e = r;
            while (e < F.nz && row_active[e] && e - r < left) ++e;
            // an inactive row inside the range (the dealiased gap between the positive and the negative kz) ends
            // the chunk early; what is left of its size becomes a chunk of its own, so that the planned LAST chunk
            // stays as short as planned (its download is the exposed one)
            out.push_back(Chunk{ r, e - r, row_a0[r], row_a0[e] });
            left -= e - r;
            if (left <= 0) { ++si; left = si < sizes.size() ? std::max(sizes[si], 1) : nrows_active; }
            r = e;
        }
    };
    cut(std::vector<int>(8, nrows_active ? (nrows_active + 7) / 8 : 1), F.chunks);
    if (nrows_active >= 32) {
        const int edge = nrows_active / 8, mid = nrows_active - 2 * edge;
        cut({ edge, mid, edge }, F.chunks_inv);
    } else {
        F.chunks_inv = F.chunks;
    }
    SZB_CUDA_OK(F.d_km.alloc(akm.size())); SZB_CUDA_OK(F.d_kn.alloc(akn.size()));
    SZB_CUDA_OK(F.d_act.alloc(F.active_idx.size())); SZB_CUDA_OK(F.d_info.alloc(F.nact + 1));
    if (F.nact) {
        SZB_CUDA_OK(cudaMemcpy(F.d_km.p, akm.data(), sizeof(double) * akm.size(), cudaMemcpyHostToDevice));
        SZB_CUDA_OK(cudaMemcpy(F.d_kn.p, akn.data(), sizeof(double) * akn.size(), cudaMemcpyHostToDevice));
        SZB_CUDA_OK(cudaMemcpy(F.d_act.p, F.active_idx.data(), sizeof(int) * F.nact, cudaMemcpyHostToDevice));
    }
    SZB_CUDA_OK(cudaStreamCreateWithFlags(&F.s_in, cudaStreamNonBlocking));
    SZB_CUDA_OK(cudaStreamCreateWithFlags(&F.s_comp, cudaStreamNonBlocking));
    SZB_CUDA_OK(cudaStreamCreateWithFlags(&F.s_out, cudaStreamNonBlocking));
    const size_t nev = std::max(F.chunks.size(), F.chunks_inv.size());
    F.ev_in.resize(nev); F.ev_done.resize(nev);
    for (auto &e : F.ev_in) SZB_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto &e : F.ev_done) SZB_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    SZB_CUDA_OK(cudaEventCreateWithFlags(&F.ev_prev, cudaEventDisableTiming));
    return 0;
}

int get_ctx(const szb_imexop *op, const szb_wavegrid *g, FieldCtx **out)
{
    FieldCtx *F = static_cast<FieldCtx *>(op->field_ctx);
    if (F && std::memcmp(&F->g, g, sizeof(*g)) != 0) { delete F; F = nullptr; op->field_ctx = nullptr; }
    if (!F) {
        F = new (std::nothrow) FieldCtx();
        if (!F) return -1;
        const int rc = build_ctx(g, *F);
        if (rc) { delete F; return rc; }
        op->field_ctx = F;
    }
    *out = F;
    return 0;
}

template <class T> int ensure(DevBuf<T> &b, size_t n)
{
    if (b.count >= n && b.p) return 0;
    if (b.p) { cudaFree(b.p); b.p = nullptr; }
    SZB_CUDA_OK(b.alloc(n));
    return 0;
}

// interleaved state rows [row0, row0+nrows), active prefix only: one strided copy
int copy_interleaved(const FieldCtx &F, const Chunk &c, size_t N, szb_complex *dst, const szb_complex *src,
                     cudaMemcpyKind kind, cudaStream_t s)
{
    const size_t pitch = sizeof(szb_complex) * N * F.nx, width = sizeof(szb_complex) * N * F.xa;
    const size_t off = (size_t) c.row0 * F.nx * N;
    SZB_CUDA_OK(cudaMemcpy2DAsync(dst + off, pitch, src + off, pitch, width, c.nrows, kind, s));
    return 0;
}

// contiguous state (field slowest): five strided copies per chunk
int copy_contiguous(const FieldCtx &F, const Chunk &c, size_t n, size_t fs, szb_complex *dst,
                    const szb_complex *src, cudaMemcpyKind kind, cudaStream_t s)
{
    const size_t pitch = sizeof(szb_complex) * n * F.nx, width = sizeof(szb_complex) * n * F.xa;
    for (int f = 0; f < 5; ++f) {
        const size_t off = (size_t) f * fs + (size_t) c.row0 * F.nx * n;
        SZB_CUDA_OK(cudaMemcpy2DAsync(dst + off, pitch, src + off, pitch, width, c.nrows, kind, s));
    }
    return 0;
}

}  // namespace

void szb::field_ctx_free(void *p) { delete static_cast<FieldCtx *>(p); }

extern "C" {

int szb_operator_apply_mass_plus_scaled_operator(const szb_imexop *op,
        const szb_wavegrid *g, const double phi[2], szb_complex *state)
{
    if (!op) return -1;
    if (!g) return -2;
    if (!phi) return -3;
    if (!state) return -4;
    FieldCtx *F;
    int rc = get_ctx(op, g, &F);
    if (rc) return rc;
    if (F->nact == 0) return 0;
    const size_t N = op->A.N;
    if ((rc = ensure(F->d_a, N * F->npencil))) return rc;
    const double zero[2] = { 0.0, 0.0 };
    for (size_t c = 0; c < F->chunks.size(); ++c) {
        const Chunk &ch = F->chunks[c];
        if ((rc = copy_interleaved(*F, ch, N, F->d_a.p, state, cudaMemcpyHostToDevice, F->s_in))) return rc;
        SZB_CUDA_OK(cudaEventRecord(F->ev_in[c], F->s_in));
        SZB_CUDA_OK(cudaStreamWaitEvent(F->s_comp, F->ev_in[c], 0));
        rc = szb_imexop_accumulate_batch(op, phi, ch.a1 - ch.a0, F->d_km.p + ch.a0, F->d_kn.p + ch.a0,
                                         F->d_act.p + ch.a0, F->d_a.p, op->n, N, zero, F->d_a.p, op->n, N,
                                         F->s_comp);
        if (rc) return rc;
        SZB_CUDA_OK(cudaEventRecord(F->ev_done[c], F->s_comp));
        SZB_CUDA_OK(cudaStreamWaitEvent(F->s_out, F->ev_done[c], 0));
        if ((rc = copy_interleaved(*F, ch, N, state, F->d_a.p, cudaMemcpyDeviceToHost, F->s_out))) return rc;
    }
    SZB_CUDA_OK(cudaStreamSynchronize(F->s_out));
    return 0;
}

int szb_operator_accumulate_mass_plus_scaled_operator(const szb_imexop *op,
        const szb_wavegrid *g, const double phi[2], const szb_complex *input,
        const double beta[2], szb_complex *output, size_t out_field_stride)
{
    if (!op) return -1;
    if (!g) return -2;
    if (!phi) return -3;
    if (!input) return -4;
    if (!beta) return -5;
    if (!output) return -6;
    FieldCtx *F;
    int rc = get_ctx(op, g, &F);
    if (rc) return rc;
    if (F->nact == 0) return 0;
    const size_t N = op->A.N, n = op->n;
    if (out_field_stride < n * (size_t) F->npencil) return -7;
    const size_t out_total = 4 * out_field_stride + n * (size_t) F->npencil;
    if ((rc = ensure(F->d_a, N * F->npencil))) return rc;
    if ((rc = ensure(F->d_b, out_total))) return rc;
    const bool need_out = !(beta[0] == 0.0 && beta[1] == 0.0);
    for (size_t c = 0; c < F->chunks.size(); ++c) {
        const Chunk &ch = F->chunks[c];
        if ((rc = copy_interleaved(*F, ch, N, F->d_a.p, input, cudaMemcpyHostToDevice, F->s_in))) return rc;
        if (need_out &&
            (rc = copy_contiguous(*F, ch, n, out_field_stride, F->d_b.p, output, cudaMemcpyHostToDevice, F->s_in)))
            return rc;
        SZB_CUDA_OK(cudaEventRecord(F->ev_in[c], F->s_in));
        SZB_CUDA_OK(cudaStreamWaitEvent(F->s_comp, F->ev_in[c], 0));
        rc = szb_imexop_accumulate_batch(op, phi, ch.a1 - ch.a0, F->d_km.p + ch.a0, F->d_kn.p + ch.a0,
                                         F->d_act.p + ch.a0, F->d_a.p, n, N, beta, F->d_b.p,
                                         out_field_stride, n, F->s_comp);
        if (rc) return rc;
        SZB_CUDA_OK(cudaEventRecord(F->ev_done[c], F->s_comp));
        SZB_CUDA_OK(cudaStreamWaitEvent(F->s_out, F->ev_done[c], 0));
        if ((rc = copy_contiguous(*F, ch, n, out_field_stride, output, F->d_b.p, cudaMemcpyDeviceToHost, F->s_out)))
            return rc;
    }
    SZB_CUDA_OK(cudaStreamSynchronize(F->s_out));
    return 0;
}

int szb_operator_invert_mass_plus_scaled_operator(const szb_imexop *op,
        const szb_zgbsv_spec *spec, const szb_wavegrid *g, const double phi[2],
        szb_complex *state, int nconstraints, szb_complex *ic0, int *first_bad_pencil)
{
    if (!op) return -1;
    if (!spec) return -2;
    if (!g) return -3;
    if (!phi) return -4;
    if (!state) return -5;
    if (nconstraints < 0) return -6;
    if (nconstraints > 0 && !ic0) return -7;
    if (first_bad_pencil) *first_bad_pencil = -1;
    static const bool trace = std::getenv("SZB_HOST_TRACE") != nullptr;
    const auto tr0 = std::chrono::steady_clock::now();
    auto since = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tr0).count(); };
    double tr[6] = {0, 0, 0, 0, 0, 0};
    FieldCtx *F;
    int rc = get_ctx(op, g, &F);
    if (rc) return rc;
    if (F->npencil == 0) return 0;
    if (nconstraints > 0 && F->zero_zero < 0) return -6;     // must own the (0,0) mode (:580-587)
    const size_t N = op->A.N;
    if ((rc = ensure(F->d_a, N * F->npencil))) return rc;
    std::vector<int> info(F->nact + 1, 0);

    DevBuf<szb_complex> scratch, dic;
    if (nconstraints > 0) {
        // Constraint right hand sides ride on the (0,0) pencil's operator: that one pencil is
        // solved on a scratch copy with the constraints attached (same factorisation
        // arithmetic; the scratch state is discarded).  Issued first so that it overlaps
        // the bulk upload.
        SZB_CUDA_OK(scratch.alloc(N));
        SZB_CUDA_OK(dic.alloc(N * (size_t) nconstraints));
        SZB_CUDA_OK(cudaMemcpyAsync(scratch.p, state + N * (size_t) F->active_idx[F->zero_zero],
                                    sizeof(szb_complex) * N, cudaMemcpyHostToDevice, F->s_comp));
        SZB_CUDA_OK(cudaMemcpyAsync(dic.p, ic0, sizeof(szb_complex) * N * nconstraints,
                                    cudaMemcpyHostToDevice, F->s_comp));
        rc = szb_imexop_invert_batch(op, spec, phi, 1, F->d_km.p + F->zero_zero, F->d_kn.p + F->zero_zero,
                                     nullptr, scratch.p, op->n, N, nconstraints, dic.p, nullptr,
                                     F->d_info.p + F->nact, nullptr, F->s_comp);
        if (rc) return rc;
        SZB_CUDA_OK(cudaMemcpyAsync(ic0, dic.p, sizeof(szb_complex) * N * nconstraints,
                                    cudaMemcpyDeviceToHost, F->s_comp));
    }
    // zcgbsvx / zgbsvx around the fused kernel: the refinement passes run over the few pencils that go on and are
    // latency-bound, so one refinement per chunk would cost as much as one over the whole field, three times.  All
    // chunks but the last get their first solve as they arrive and ONE refinement together (their downloads then
    // overlap the last chunk's solve); the last chunk is solved and refined on its own.
    const size_t nch = F->chunks_inv.size();
    std::vector<cudaEvent_t> tev;                              // SZB_HOST_TRACE: device-side timeline
    std::vector<const char *> tname;
    auto mark = [&](const char *name, cudaStream_t st) {
        if (!trace) return;
        cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); tev.push_back(e); tname.push_back(name);
    };
    mark("start", F->s_in);
    const bool refined_zc = spec->method == SZB_SOLVER_ZCGBSVX && spec->tolsc == 0.0 && spec->aiter >= 1;
    const bool refined_zx = spec->method == SZB_SOLVER_ZGBSVX && !spec->equil;
    // (off by default, SZB_HOST_STAGED=1: on one GPU it is a wash, 24.78 vs 24.85 ms, and with several ranks sharing the host's
    // PCIe / memory system the late burst of downloads is no longer hidden: 4 GPUs 304 vs 276 ms per substep)
    static const bool staging = [] { const char *e = std::getenv("SZB_HOST_STAGED"); return e && e[0] == '1'; }();
    bool staged = staging && nch >= 3 && (refined_zc || refined_zx) && op->linearization == SZB_LINEARIZE_RHOME_XYZ;
    const int g0 = nch >= 1 ? F->chunks_inv[0].a0 : 0, g1 = nch >= 2 ? F->chunks_inv[nch - 2].a1 : 0;    // union of all chunks but the last
    for (size_t c = 0; c < nch; ++c) {
        const Chunk &ch = F->chunks_inv[c];
        if ((rc = copy_interleaved(*F, ch, N, F->d_a.p, state, cudaMemcpyHostToDevice, F->s_in))) return rc;
        mark("upload", F->s_in);
        SZB_CUDA_OK(cudaEventRecord(F->ev_in[c], F->s_in));
        SZB_CUDA_OK(cudaStreamWaitEvent(F->s_comp, F->ev_in[c], 0));
        if (staged && c + 1 < nch) {
            rc = szb::invert_refined_stage(op, refined_zx ? 1 : 0, spec->aiter, spec->diter, phi, ch.a1 - ch.a0,
                                           F->d_km.p + ch.a0, F->d_kn.p + ch.a0, F->d_act.p + ch.a0,
                                           reinterpret_cast<szb::cplx *>(F->d_a.p), op->n, N, nullptr, F->d_info.p + ch.a0,
                                           nullptr, F->s_comp, 1, ch.a0 - g0, g1 - g0);
            if (rc == 1 && c == 0) staged = false;               // no fused kernel for this operator: plain path below
            else if (rc) return rc;
            if (staged) {
                mark("first solve", F->s_comp);
                if (c + 2 < nch) continue;
                // the last of the group: refine the union, then release every download of the group
                rc = szb::invert_refined_stage(op, refined_zx ? 1 : 0, spec->aiter, spec->diter, phi, g1 - g0,
                                               F->d_km.p + g0, F->d_kn.p + g0, F->d_act.p + g0,
                                               reinterpret_cast<szb::cplx *>(F->d_a.p), op->n, N, nullptr, F->d_info.p + g0,
                                               nullptr, F->s_comp, 2, 0, g1 - g0);
                if (rc) return rc;
                mark("refinement", F->s_comp);
                SZB_CUDA_OK(cudaEventRecord(F->ev_done[c], F->s_comp));
                SZB_CUDA_OK(cudaStreamWaitEvent(F->s_out, F->ev_done[c], 0));
                for (size_t d = 0; d <= c; ++d) {
                    if ((rc = copy_interleaved(*F, F->chunks_inv[d], N, state, F->d_a.p, cudaMemcpyDeviceToHost, F->s_out)))
                        return rc;
                    mark("download", F->s_out);
                }
                continue;
            }
        }
        rc = szb_imexop_invert_batch(op, spec, phi, ch.a1 - ch.a0, F->d_km.p + ch.a0, F->d_kn.p + ch.a0,
                                     F->d_act.p + ch.a0, F->d_a.p, op->n, N, 0, nullptr, nullptr,
                                     F->d_info.p + ch.a0, nullptr, F->s_comp);
        if (rc) return rc;
        mark("solve", F->s_comp);
        SZB_CUDA_OK(cudaEventRecord(F->ev_done[c], F->s_comp));
        SZB_CUDA_OK(cudaStreamWaitEvent(F->s_out, F->ev_done[c], 0));
        if ((rc = copy_interleaved(*F, ch, N, state, F->d_a.p, cudaMemcpyDeviceToHost, F->s_out))) return rc;
        mark("download", F->s_out);
    }
    tr[0] = since();
    // dealiased / Nyquist pencils are zero-filled (:632-637): on the host, they never travel
    // (a few host threads: 180 MB on the bench grid, done while the device works on the active pencils)
    {
        const size_t ni = F->inactive_idx.size();
        const int nthr = ni * N * sizeof(szb_complex) > (size_t) (8u << 20) ? 4 : 1;
        auto zero_range = [&](size_t lo, size_t hi) {
            for (size_t i = lo; i < hi; ++i)
                std::memset(state + N * (size_t) F->inactive_idx[i], 0, sizeof(szb_complex) * N);
        };
        if (nthr == 1) zero_range(0, ni);
        else {
            std::vector<std::thread> pool;
            for (int t = 1; t < nthr; ++t) pool.emplace_back(zero_range, ni * t / nthr, ni * (t + 1) / nthr);
            zero_range(0, ni / nthr);
            for (auto &th : pool) th.join();
        }
    }
    tr[1] = since();
    SZB_CUDA_OK(cudaStreamSynchronize(F->s_comp));
    tr[2] = since();
    SZB_CUDA_OK(cudaMemcpy(info.data(), F->d_info.p, sizeof(int) * (F->nact + (nconstraints > 0)),
                           cudaMemcpyDeviceToHost));
    tr[3] = since();
    SZB_CUDA_OK(cudaStreamSynchronize(F->s_out));
    tr[4] = since();
    if (trace) {
        for (size_t i = 1; i < tev.size(); ++i) {
            float ms = 0; cudaEventElapsedTime(&ms, tev[0], tev[i]);
            std::fprintf(stderr, "  %s @ %.2f", tname[i], ms);
        }
        std::fprintf(stderr, "\n");
        for (auto e : tev) cudaEventDestroy(e);
    }
    if (trace)
        std::fprintf(stderr, "invert host: enqueued %.2f, zero-fill done %.2f, compute done %.2f, info %.2f, downloads done %.2f ms\n",
                     tr[0], tr[1], tr[2], tr[3], tr[4]);
    for (int p = 0; p < F->nact + (nconstraints > 0); ++p) {
        if (info[p]) {
            if (first_bad_pencil) *first_bad_pencil = p < F->nact ? F->active_idx[p] : F->active_idx[F->zero_zero];
            return info[p];
        }
    }
    return 0;
}

/* bsmbsm_solver::solve on HOST storage (suzerain/bsmbsm_solver.cpp:155-182 zgbsv, :377-414 zcgbsvx): the
 * protocol object of include/suzerain_b200_solver.hpp keeps LU / PAPT / PB / PX / ipiv on the host exactly like
 * the reference's; one call moves them to the device, runs the batched LAPACK kernels on this one system and
 * brings the results back. */
int szb_bsmbsm_solver_solve(const szb_bsmbsm *A, const szb_zgbsv_spec *spec, char trans, int nrhs,
                            szb_complex *lu, const szb_complex *papt, int *ipiv,
                            szb_complex *pb, szb_complex *px, int *iters, double *res)
{
    if (!A) return -1;
    if (!spec) return -2;
    if (trans != 'N' && trans != 'T') return -3;
    if (nrhs < 0) return -4;
    if (!lu) return -5;
    if (!ipiv) return -7;
    if (!pb) return -8;
    if (nrhs == 0) return 0;
    const int N = A->N, KL = A->KL, KU = A->KU, LD = A->LD, ldlu = LD + KL;
    DevBuf<szb_complex> dlu, dpapt, db, dx; DevBuf<int> dipiv, dinfo, diters; DevBuf<double> dres;
    SZB_CUDA_OK(dlu.alloc((size_t) ldlu * N)); SZB_CUDA_OK(dipiv.alloc(N)); SZB_CUDA_OK(dinfo.alloc(1));
    SZB_CUDA_OK(db.alloc((size_t) N * nrhs));
    int info = 0;
    if (spec->method == SZB_SOLVER_ZGBSV) {
        // in place: PAPT aliases LU + KL (bsmbsm_solver.cpp:64-67), PB is overwritten by the solution
        SZB_CUDA_OK(cudaMemcpy(dlu.p, lu, sizeof(szb_complex) * ldlu * N, cudaMemcpyHostToDevice));
        SZB_CUDA_OK(cudaMemcpy(db.p, pb, sizeof(szb_complex) * N * nrhs, cudaMemcpyHostToDevice));
        int rc = szb_zgbtrf_batch(N, KL, KU, dlu.p, ldlu, (size_t) ldlu * N, dipiv.p, dinfo.p, 1, nullptr);
        if (rc) return rc;
        SZB_CUDA_OK(cudaMemcpy(&info, dinfo.p, sizeof(int), cudaMemcpyDeviceToHost));
        if (!info) {
            rc = szb_zgbtrs_batch(trans, N, KL, KU, nrhs, dlu.p, ldlu, (size_t) ldlu * N, dipiv.p, db.p, N,
                                  (size_t) N * nrhs, 1, nullptr);
            if (rc) return rc;
            SZB_CUDA_OK(cudaMemcpy(pb, db.p, sizeof(szb_complex) * N * nrhs, cudaMemcpyDeviceToHost));
        }
        SZB_CUDA_OK(cudaMemcpy(lu, dlu.p, sizeof(szb_complex) * ldlu * N, cudaMemcpyDeviceToHost));
        SZB_CUDA_OK(cudaMemcpy(ipiv, dipiv.p, sizeof(int) * N, cudaMemcpyDeviceToHost));
        return info;
    }
    if (spec->method != SZB_SOLVER_ZCGBSVX) return -2;      // zgbsvx has no pre-assembled entry point: say so
    if (!papt) return -6;
    if (!px) return -9;
    SZB_CUDA_OK(dpapt.alloc((size_t) LD * N)); SZB_CUDA_OK(dx.alloc(N)); SZB_CUDA_OK(diters.alloc(1)); SZB_CUDA_OK(dres.alloc(1));
    SZB_CUDA_OK(cudaMemcpy(dpapt.p, papt, sizeof(szb_complex) * LD * N, cudaMemcpyHostToDevice));
    SZB_CUDA_OK(cudaMemcpy(db.p, pb, sizeof(szb_complex) * N * nrhs, cudaMemcpyHostToDevice));
    for (int j = 0; j < nrhs && !info; ++j) {
        // bsmbsm_solver_zcgbsvx::solve_hook refactors per right hand side too (gotcha 8)
        int rc = szb_zcgbsvx_batch(trans, N, KL, KU, spec->aiter, spec->diter, spec->tolsc, dpapt.p, (size_t) LD * N,
                                   dlu.p, (size_t) ldlu * N, dipiv.p, db.p + (size_t) j * N, dx.p, diters.p, dres.p,
                                   dinfo.p, 1, nullptr);
        if (rc) return rc;
        SZB_CUDA_OK(cudaMemcpy(&info, dinfo.p, sizeof(int), cudaMemcpyDeviceToHost));
        SZB_CUDA_OK(cudaMemcpy(px + (size_t) j * N, dx.p, sizeof(szb_complex) * N, cudaMemcpyDeviceToHost));
        if (iters) SZB_CUDA_OK(cudaMemcpy(iters + j, diters.p, sizeof(int), cudaMemcpyDeviceToHost));
        if (res) SZB_CUDA_OK(cudaMemcpy(res + j, dres.p, sizeof(double), cudaMemcpyDeviceToHost));
    }
    SZB_CUDA_OK(cudaMemcpy(lu, dlu.p, sizeof(szb_complex) * ldlu * N, cudaMemcpyDeviceToHost));
    SZB_CUDA_OK(cudaMemcpy(ipiv, dipiv.p, sizeof(int) * N, cudaMemcpyDeviceToHost));
    return info;
}

/* The km = kn = 0 special cases (suzerain/rholut_imexop.h:209-238, 330-368, 447-469: "equivalent
 * to calling ... using km == 0 and kn == 0"), used by linearize::rhome_y. */
int szb_rholut_imexop_accumulate00(const double phi[2],
        const szb_rholut_imexop_scenario *s, const szb_rholut_imexop_ref *r,
        const szb_rholut_imexop_refld *ld, const szb_bsplineop *w,
        const szb_complex *in_rho_E, const szb_complex *in_rho_u,
        const szb_complex *in_rho_v, const szb_complex *in_rho_w,
        const szb_complex *in_rho, const double beta[2],
        szb_complex *out_rho_E, szb_complex *out_rho_u, szb_complex *out_rho_v,
        szb_complex *out_rho_w, szb_complex *out_rho, const double *c)
{
    return szb_rholut_imexop_accumulate(phi, 0.0, 0.0, s, r, ld, w, in_rho_E, in_rho_u, in_rho_v, in_rho_w,
                                        in_rho, beta, out_rho_E, out_rho_u, out_rho_v, out_rho_w, out_rho,
                                        nullptr, nullptr, c);
}

int szb_rholut_imexop_packc00(const double phi[2],
        const szb_rholut_imexop_scenario *s, const szb_rholut_imexop_ref *r,
        const szb_rholut_imexop_refld *ld, const szb_bsplineop *w,
        szb_bsmbsm *A_T, szb_complex *patpt, const double *c)
{
    return szb_rholut_imexop_packc(phi, 0.0, 0.0, s, r, ld, w, A_T, patpt, nullptr, nullptr, c);
}

int szb_rholut_imexop_packf00(const double phi[2],
        const szb_rholut_imexop_scenario *s, const szb_rholut_imexop_ref *r,
        const szb_rholut_imexop_refld *ld, const szb_bsplineop *w,
        szb_bsmbsm *A_T, szb_complex *patpt, const double *c)
{
    return szb_rholut_imexop_packf(phi, 0.0, 0.0, s, r, ld, w, A_T, patpt, nullptr, nullptr, c);
}

}  // extern "C"
