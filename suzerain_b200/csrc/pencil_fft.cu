// pencil_fft.cu -- device replacement of suzerain::pencil_grid: the wave <-> physical
// transforms of the nonlinear operator (suzerain/pencil_grid.hpp:57-246, pencil_grid.cpp:60-194;
// called at apps/perfect/navier_stokes.hpp:320-331,969).  See include/suzerain_b200_fft.h for the
// layouts.  FFTs are cuFFT (batched 2-D Z2D / D2Z over the (z, x) plane of every y owned in
// physical space); what is ours are the layout changes around them:
//
//   wave [Z][X][Y] (Y fastest)  <->  [Y][Z][X] (X fastest)
//
// is a transpose of the (Z*X) x Y matrix, done through 32 x 32 shared-memory tiles so that both
// sides move whole 512-byte rows.  With several ranks (wave space cut in Z, physical space in Y)
// the same kernel writes one block per destination rank -- the send buffer of the all-to-all --
// and the received blocks are placed with strided device copies.
#include <atomic>
#include <cstdio>
#include <new>
#include <vector>

#include <cuda_runtime.h>
#include <cufft.h>

#include "../../include/suzerain_b200_fft.h"

namespace {

struct __align__(16) cplx { double x, y; };

std::atomic<unsigned long long> g_launches{0};

#define FFT_CUDA_OK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    std::fprintf(stderr, "suzerain_b200_fft: %s at %s:%d: %s\n", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); \
    return -100 - (int) e_; } } while (0)
#define FFT_CUFFT_OK(call) do { cufftResult r_ = (call); if (r_ != CUFFT_SUCCESS) { \
    std::fprintf(stderr, "suzerain_b200_fft: %s at %s:%d: cufft error %d\n", #call, __FILE__, __LINE__, (int) r_); \
    return -200 - (int) r_; } } while (0)

// out[c * ld_out + r] = in[r * ld_in + c],  0 <= r < rows, 0 <= c < cols
__global__ void __launch_bounds__(256)
transpose_kernel(const cplx *__restrict__ in, size_t ld_in, cplx *__restrict__ out, size_t ld_out, int rows, int cols)
{
    __shared__ cplx tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        if (r < rows && c < cols) tile[i][threadIdx.x] = in[(size_t) r * ld_in + c];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (r < rows && c < cols) out[(size_t) c * ld_out + r] = tile[threadIdx.x][i];
    }
}

int transpose(const cplx *in, size_t ld_in, cplx *out, size_t ld_out, int rows, int cols, cudaStream_t s)
{
    if (rows <= 0 || cols <= 0) return 0;
    const dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
    transpose_kernel<<<grid, block, 0, s>>>(in, ld_in, out, ld_out, rows, cols);
    ++g_launches;
    FFT_CUDA_OK(cudaGetLastError());
    return 0;
}

// The same tile transpose with the columns cut into blocks that go to different buffers (one per
// rank): column c of block b (bound[b] <= c < bound[b+1]) lands at
//   ptr[b][(c - bound[b]) * ld_out + row_off + r].
// One launch serves every destination; over NVLink peer mappings this is the all-to-all.
constexpr int MAX_PEERS = 16;
struct Targets { cplx *ptr[MAX_PEERS]; int bound[MAX_PEERS + 1]; int n; size_t ld_out, row_off; };

__global__ void __launch_bounds__(256)
transpose_scatter_kernel(const cplx *__restrict__ in, size_t ld_in, int rows, int cols, const Targets T)
{
    __shared__ cplx tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        if (r < rows && c < cols) tile[i][threadIdx.x] = in[(size_t) r * ld_in + c];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (r < rows && c < cols) {
            int b = 0;
            while (b + 1 < T.n && c >= T.bound[b + 1]) ++b;
            T.ptr[b][(size_t) (c - T.bound[b]) * T.ld_out + T.row_off + r] = tile[threadIdx.x][i];
        }
    }
}

int transpose_scatter(const cplx *in, size_t ld_in, int rows, int cols, const Targets &T, cudaStream_t s)
{
    if (rows <= 0 || cols <= 0) return 0;
    const dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
    transpose_scatter_kernel<<<grid, block, 0, s>>>(in, ld_in, rows, cols, T);
    ++g_launches;
    FFT_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace

struct szb_pencil_grid {
    int dNx, Ny, dNz, nxw, nranks, rank;
    std::vector<int> zs, ys;            // rank r owns wave z in [zs[r], zs[r+1]), physical y in [ys[r], ys[r+1])
    cufftHandle c2r, r2c;
    bool have_plans;
    cplx *scratch;                      // [Yloc][dNz][nxw]
    int zloc(int r) const { return zs[r + 1] - zs[r]; }
    int yloc(int r) const { return ys[r + 1] - ys[r]; }
};

extern "C" {

unsigned long long szb_fft_launch_count(void) { return g_launches.load(); }

int szb_pencil_grid_create(int dNx, int Ny, int dNz, int nranks, int rank, szb_pencil_grid **out)
{
    if (dNx < 1) return -1;
    if (Ny < 1) return -2;
    if (dNz < 1) return -3;
    if (nranks < 1) return -4;
    if (rank < 0 || rank >= nranks) return -5;
    if (!out) return -6;
    szb_pencil_grid *g = new (std::nothrow) szb_pencil_grid();
    if (!g) return -6;
    g->dNx = dNx; g->Ny = Ny; g->dNz = dNz; g->nxw = dNx / 2 + 1; g->nranks = nranks; g->rank = rank;
    g->zs.resize(nranks + 1); g->ys.resize(nranks + 1);
    for (int r = 0; r <= nranks; ++r) {
        g->zs[r] = (int) ((long long) r * dNz / nranks);
        g->ys[r] = (int) ((long long) r * Ny / nranks);
    }
    g->have_plans = false; g->scratch = nullptr;
    const int yl = g->yloc(rank);
    if (yl > 0) {
        const size_t elems = (size_t) yl * dNz * g->nxw;
        cudaError_t e = cudaMalloc(&g->scratch, elems * sizeof(cplx));
        if (e != cudaSuccess) { delete g; return -100 - (int) e; }
        int n[2] = { dNz, dNx }, cemb[2] = { dNz, g->nxw }, remb[2] = { dNz, dNx };
        cufftResult r1 = cufftPlanMany(&g->c2r, 2, n, cemb, 1, dNz * g->nxw, remb, 1, dNz * dNx, CUFFT_Z2D, yl);
        cufftResult r2 = r1 == CUFFT_SUCCESS
                       ? cufftPlanMany(&g->r2c, 2, n, remb, 1, dNz * dNx, cemb, 1, dNz * g->nxw, CUFFT_D2Z, yl)
                       : r1;
        if (r1 != CUFFT_SUCCESS || r2 != CUFFT_SUCCESS) {
            if (r1 == CUFFT_SUCCESS) cufftDestroy(g->c2r);
            cudaFree(g->scratch); delete g;
            return -200 - (int) (r1 != CUFFT_SUCCESS ? r1 : r2);
        }
        g->have_plans = true;
    }
    *out = g;
    return 0;
}

void szb_pencil_grid_destroy(szb_pencil_grid *g)
{
    if (!g) return;
    if (g->have_plans) { cufftDestroy(g->c2r); cufftDestroy(g->r2c); }
    if (g->scratch) cudaFree(g->scratch);
    delete g;
}

int szb_pencil_grid_extents(const szb_pencil_grid *g, int which, int r, int start[3], int end[3])
{
    if (!g) return -1;
    if (which != 0 && which != 1) return -2;
    if (r < 0 || r >= g->nranks) return -3;
    if (!start) return -4;
    if (!end) return -5;
    if (which == 0) {          // physical: X and Z complete, Y cut
        start[0] = 0; end[0] = g->dNx; start[1] = g->ys[r]; end[1] = g->ys[r + 1]; start[2] = 0; end[2] = g->dNz;
    } else {                   // wave: X (kept half) and Y complete, Z cut
        start[0] = 0; end[0] = g->nxw; start[1] = 0; end[1] = g->Ny; start[2] = g->zs[r]; end[2] = g->zs[r + 1];
    }
    return 0;
}

size_t szb_pencil_grid_local_wave_storage(const szb_pencil_grid *g)
{
    if (!g) return 0;
    const size_t wave = (size_t) g->zloc(g->rank) * g->nxw * g->Ny, phys = (size_t) g->yloc(g->rank) * g->dNz * g->dNx;
    return wave > phys / 2 + phys % 2 ? wave : phys / 2 + phys % 2;
}

size_t szb_pencil_grid_local_physical_storage(const szb_pencil_grid *g)
{
    if (!g) return 0;
    const size_t wave = (size_t) g->zloc(g->rank) * g->nxw * g->Ny, phys = (size_t) g->yloc(g->rank) * g->dNz * g->dNx;
    return 2 * wave > phys ? 2 * wave : phys;
}

int szb_pencil_grid_has_zero_zero_modes(const szb_pencil_grid *g)
{
    return g && g->zs[g->rank] == 0 && g->zloc(g->rank) > 0 && g->Ny > 0;
}

int szb_pencil_grid_exchange_counts(const szb_pencil_grid *g, int dir, long long *send, long long *recv)
{
    if (!g) return -1;
    if (dir != 0 && dir != 1) return -2;
    if (!send) return -3;
    if (!recv) return -4;
    const int me = g->rank;
    for (int p = 0; p < g->nranks; ++p) {
        const long long a = (long long) g->yloc(p) * g->zloc(me) * g->nxw;      // [Yloc_p][Zloc_me][X]
        const long long b = (long long) g->yloc(me) * g->zloc(p) * g->nxw;      // [Yloc_me][Zloc_p][X]
        send[p] = dir == 0 ? a : b;
        recv[p] = dir == 0 ? b : a;
    }
    return 0;
}

int szb_pencil_grid_w2p_pack(szb_pencil_grid *g, const double *d_wave, double *d_send, void *stream)
{
    if (!g) return -1;
    if (!d_wave) return -2;
    if (!d_send) return -3;
    const cplx *wave = reinterpret_cast<const cplx *>(d_wave);
    cplx *send = reinterpret_cast<cplx *>(d_send);
    const int Q = g->zloc(g->rank) * g->nxw;
    size_t off = 0;
    for (int s = 0; s < g->nranks; ++s) {
        // block s [y - ys[s]][q] = wave[q][y]
        const int rc = transpose(wave + g->ys[s], (size_t) g->Ny, send + off, (size_t) Q, Q, g->yloc(s), (cudaStream_t) stream);
        if (rc) return rc;
        off += (size_t) g->yloc(s) * Q;
    }
    return 0;
}

int szb_pencil_grid_w2p_finish(szb_pencil_grid *g, const double *d_recv, double *d_phys, void *stream)
{
    if (!g) return -1;
    if (!d_recv) return -2;
    if (!d_phys) return -3;
    const int yl = g->yloc(g->rank);
    if (yl == 0) return 0;
    const cplx *recv = reinterpret_cast<const cplx *>(d_recv);
    const cplx *fft_in = recv;
    if (g->nranks > 1) {
        // block r [y][zl][x] -> scratch [y][zs[r] + zl][x]
        size_t off = 0;
        for (int r = 0; r < g->nranks; ++r) {
            const size_t row = sizeof(cplx) * (size_t) g->zloc(r) * g->nxw;
            if (row)
                FFT_CUDA_OK(cudaMemcpy2DAsync(g->scratch + (size_t) g->zs[r] * g->nxw, sizeof(cplx) * (size_t) g->dNz * g->nxw,
                                              recv + off, row, row, yl, cudaMemcpyDeviceToDevice, (cudaStream_t) stream));
            off += (size_t) yl * g->zloc(r) * g->nxw;
        }
        fft_in = g->scratch;
    }
    FFT_CUFFT_OK(cufftSetStream(g->c2r, (cudaStream_t) stream));
    FFT_CUFFT_OK(cufftExecZ2D(g->c2r, reinterpret_cast<cufftDoubleComplex *>(const_cast<cplx *>(fft_in)), d_phys));
    return 0;
}

int szb_pencil_grid_p2w_start(szb_pencil_grid *g, const double *d_phys, double *d_send, void *stream)
{
    if (!g) return -1;
    if (!d_phys) return -2;
    if (!d_send) return -3;
    const int yl = g->yloc(g->rank);
    if (yl == 0) return 0;
    cplx *send = reinterpret_cast<cplx *>(d_send);
    cplx *fft_out = g->nranks > 1 ? g->scratch : send;
    FFT_CUFFT_OK(cufftSetStream(g->r2c, (cudaStream_t) stream));
    FFT_CUFFT_OK(cufftExecD2Z(g->r2c, const_cast<double *>(d_phys), reinterpret_cast<cufftDoubleComplex *>(fft_out)));
    if (g->nranks > 1) {
        size_t off = 0;
        for (int r = 0; r < g->nranks; ++r) {
            const size_t row = sizeof(cplx) * (size_t) g->zloc(r) * g->nxw;
            if (row)
                FFT_CUDA_OK(cudaMemcpy2DAsync(send + off, row, g->scratch + (size_t) g->zs[r] * g->nxw,
                                              sizeof(cplx) * (size_t) g->dNz * g->nxw, row, yl, cudaMemcpyDeviceToDevice,
                                              (cudaStream_t) stream));
            off += (size_t) yl * g->zloc(r) * g->nxw;
        }
    }
    return 0;
}

int szb_pencil_grid_p2w_unpack(szb_pencil_grid *g, const double *d_recv, double *d_wave, void *stream)
{
    if (!g) return -1;
    if (!d_recv) return -2;
    if (!d_wave) return -3;
    const cplx *recv = reinterpret_cast<const cplx *>(d_recv);
    cplx *wave = reinterpret_cast<cplx *>(d_wave);
    const int Q = g->zloc(g->rank) * g->nxw;
    size_t off = 0;
    for (int s = 0; s < g->nranks; ++s) {
        // wave[q][y] = block s [y - ys[s]][q]
        const int rc = transpose(recv + off, (size_t) Q, wave + g->ys[s], (size_t) g->Ny, g->yloc(s), Q, (cudaStream_t) stream);
        if (rc) return rc;
        off += (size_t) g->yloc(s) * Q;
    }
    return 0;
}

// ---- peer-memory variants: the layout change IS the exchange ----
// Every rank exposes two buffers to its peers (NVLink peer mappings obtained by the caller, e.g.
// torch symmetric memory): fft = [Yloc][dNz][X], the input of its backward FFT / output of its
// forward FFT, and wave = [Zloc][X][Y].  The transposing kernels store straight into the peers'
// buffers at the final position, so there is no send buffer, no all-to-all and no unpacking pass;
// the caller brackets them with device-side barriers.
int szb_pencil_grid_w2p_pack_peers(szb_pencil_grid *g, const double *d_wave, const unsigned long long *peer_fft, void *stream)
{
    if (!g) return -1;
    if (!d_wave) return -2;
    if (!peer_fft) return -3;
    if (g->nranks > MAX_PEERS) return -1;
    const cplx *wave = reinterpret_cast<const cplx *>(d_wave);
    const int me = g->rank, Q = g->zloc(me) * g->nxw;
    // peer s: fft[y - ys[s]][zs[me] * X + q] = wave[q][y], every peer in one launch
    Targets T;
    T.n = g->nranks; T.ld_out = (size_t) g->dNz * g->nxw; T.row_off = (size_t) g->zs[me] * g->nxw;
    for (int s = 0; s < g->nranks; ++s) { T.ptr[s] = reinterpret_cast<cplx *>(peer_fft[s]); T.bound[s] = g->ys[s]; }
    T.bound[g->nranks] = g->ys[g->nranks];
    return transpose_scatter(wave, (size_t) g->Ny, Q, g->Ny, T, (cudaStream_t) stream);
}

int szb_pencil_grid_w2p_fft(szb_pencil_grid *g, const double *d_fft, double *d_phys, void *stream)
{
    if (!g) return -1;
    if (!d_fft) return -2;
    if (!d_phys) return -3;
    if (g->yloc(g->rank) == 0) return 0;
    FFT_CUFFT_OK(cufftSetStream(g->c2r, (cudaStream_t) stream));
    FFT_CUFFT_OK(cufftExecZ2D(g->c2r, reinterpret_cast<cufftDoubleComplex *>(const_cast<double *>(d_fft)), d_phys));
    return 0;
}

int szb_pencil_grid_p2w_fft(szb_pencil_grid *g, const double *d_phys, double *d_fft, void *stream)
{
    if (!g) return -1;
    if (!d_phys) return -2;
    if (!d_fft) return -3;
    if (g->yloc(g->rank) == 0) return 0;
    FFT_CUFFT_OK(cufftSetStream(g->r2c, (cudaStream_t) stream));
    FFT_CUFFT_OK(cufftExecD2Z(g->r2c, const_cast<double *>(d_phys), reinterpret_cast<cufftDoubleComplex *>(d_fft)));
    return 0;
}

int szb_pencil_grid_p2w_scatter_peers(szb_pencil_grid *g, const double *d_fft, const unsigned long long *peer_wave, void *stream)
{
    if (!g) return -1;
    if (!d_fft) return -2;
    if (!peer_wave) return -3;
    if (g->nranks > MAX_PEERS) return -1;
    const cplx *fft = reinterpret_cast<const cplx *>(d_fft);
    const int me = g->rank;
    // peer r: wave[q][ys[me] + y] = fft[y][zs[r] * X + q], every peer in one launch
    Targets T;
    T.n = g->nranks; T.ld_out = (size_t) g->Ny; T.row_off = (size_t) g->ys[me];
    for (int r = 0; r < g->nranks; ++r) { T.ptr[r] = reinterpret_cast<cplx *>(peer_wave[r]); T.bound[r] = g->zs[r] * g->nxw; }
    T.bound[g->nranks] = g->dNz * g->nxw;
    return transpose_scatter(fft, (size_t) g->dNz * g->nxw, g->yloc(me), g->dNz * g->nxw, T, (cudaStream_t) stream);
}

int szb_pencil_grid_transform_wave_to_physical(szb_pencil_grid *g, double *d_inout, void *stream)
{
    if (!g) return -1;
    if (g->nranks != 1) return -1;
    if (!d_inout) return -2;
    // one rank: the single send block [Y][Z][X] is the FFT input
    int rc = szb_pencil_grid_w2p_pack(g, d_inout, reinterpret_cast<double *>(g->scratch), stream);
    if (rc) return rc;
    return szb_pencil_grid_w2p_finish(g, reinterpret_cast<const double *>(g->scratch), d_inout, stream);
}

int szb_pencil_grid_transform_physical_to_wave(szb_pencil_grid *g, double *d_inout, void *stream)
{
    if (!g) return -1;
    if (g->nranks != 1) return -1;
    if (!d_inout) return -2;
    int rc = szb_pencil_grid_p2w_start(g, d_inout, reinterpret_cast<double *>(g->scratch), stream);
    if (rc) return rc;
    return szb_pencil_grid_p2w_unpack(g, reinterpret_cast<const double *>(g->scratch), d_inout, stream);
}

}  // extern "C"
