/* suzerain_b200_fft.h -- device replacement of suzerain::pencil_grid.
 *
 * C ABI of libsuzerain_b200_fft.so (separate from libsuzerain_b200.so so that the implicit
 * operator does not depend on cuFFT).  Replaces
 *   suzerain::pencil_grid / pencil_grid_p3dfft      suzerain/pencil_grid.hpp:57-246,
 *                                                   suzerain/pencil_grid.cpp:60-194
 *   p3dfft_btran_c2r / p3dfft_ftran_r2c             lib/suzerain-p3dfft (STRIDE1 build)
 * as they are used by apps/perfect/navier_stokes.hpp:320-331,969 and perfect.cpp:136-175:
 *
 *   wave space      complex [Z][X][Y], Y fastest ("stride one in Y"), X = dNx/2 + 1 kept modes of the
 *                   real-to-complex direction, Z = dNz; Y complete on every rank
 *   physical space  real    [Y][Z][X], X fastest (physical_view.hpp:43-47); X and Z complete
 *   transforms      unnormalised: wave -> physical is a backward complex FFT in z followed by
 *                   complex-to-real in x, physical -> wave the forward pair; a round trip scales
 *                   by dNx * dNz (= 1 / pencil_grid::chi(), pencil_grid.hpp:159-163)
 *
 * Decomposition: slabs.  Wave space is cut in Z, physical space in Y (the reference's P0 x P1
 * processor grid with P0 = 1), contiguous and balanced: rank r owns [r n / R, (r + 1) n / R).
 * With more than one rank a transform is  pack -> all-to-all -> finish; the exchange itself is the
 * caller's (NCCL all_to_all_single in suzerain_b200/pencil.py), the two local phases are here.
 * All pointers are DEVICE pointers; extents use the reference's (X, Y, Z) index order.
 */
#ifndef SUZERAIN_B200_FFT_H
#define SUZERAIN_B200_FFT_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct szb_pencil_grid szb_pencil_grid;

/* global_physical_extent = (dNx, Ny, dNz) (pencil_grid.hpp:80); nranks >= 1, 0 <= rank < nranks */
int  szb_pencil_grid_create(int dNx, int Ny, int dNz, int nranks, int rank, szb_pencil_grid **out);
void szb_pencil_grid_destroy(szb_pencil_grid *g);

/* local_{physical,wave}_{start,end} (pencil_grid.hpp:103-139): which = 0 physical, 1 wave; of rank r */
int  szb_pencil_grid_extents(const szb_pencil_grid *g, int which, int r, int start[3], int end[3]);
/* local_wave_storage() in complex scalars / local_physical_storage() in real scalars (pencil_grid.cpp:166-178) */
size_t szb_pencil_grid_local_wave_storage(const szb_pencil_grid *g);
size_t szb_pencil_grid_local_physical_storage(const szb_pencil_grid *g);
/* does this rank hold the (0,0) Fourier modes (pencil_grid.hpp:173-178) */
int  szb_pencil_grid_has_zero_zero_modes(const szb_pencil_grid *g);

/* One rank: transform_wave_to_physical / transform_physical_to_wave, in place (pencil_grid.hpp:200-221) */
int  szb_pencil_grid_transform_wave_to_physical(szb_pencil_grid *g, double *d_inout, void *stream);
int  szb_pencil_grid_transform_physical_to_wave(szb_pencil_grid *g, double *d_inout, void *stream);

/* Several ranks.  Exchange buffers hold one block per peer, in rank order:
 *   wave -> physical   send block s = [Yloc_s][Zloc_me][X],  recv block r = [Yloc_me][Zloc_r][X]
 *   physical -> wave   send block r = [Yloc_me][Zloc_r][X],  recv block s = [Yloc_s][Zloc_me][X]
 * counts[] (complex scalars per peer) from szb_pencil_grid_exchange_counts: dir 0 = wave -> physical. */
int  szb_pencil_grid_exchange_counts(const szb_pencil_grid *g, int dir, long long *send, long long *recv);
int  szb_pencil_grid_w2p_pack(szb_pencil_grid *g, const double *d_wave, double *d_send, void *stream);
int  szb_pencil_grid_w2p_finish(szb_pencil_grid *g, const double *d_recv, double *d_phys, void *stream);
int  szb_pencil_grid_p2w_start(szb_pencil_grid *g, const double *d_phys, double *d_send, void *stream);
int  szb_pencil_grid_p2w_unpack(szb_pencil_grid *g, const double *d_recv, double *d_wave, void *stream);

/* Peer-memory variants (NVLink P2P): peer_fft[r] / peer_wave[r] are device addresses, valid on this
 * GPU, of rank r's [Yloc_r][dNz][X] FFT buffer / [Zloc_r][X][Y] wave buffer.  The transposing kernels
 * store directly into the peers' buffers at the final position -- layout change and exchange are one
 * pass; the caller separates the phases with cross-GPU barriers. */
int  szb_pencil_grid_w2p_pack_peers(szb_pencil_grid *g, const double *d_wave, const unsigned long long *peer_fft, void *stream);
int  szb_pencil_grid_w2p_fft(szb_pencil_grid *g, const double *d_fft, double *d_phys, void *stream);
int  szb_pencil_grid_p2w_fft(szb_pencil_grid *g, const double *d_phys, double *d_fft, void *stream);
int  szb_pencil_grid_p2w_scatter_peers(szb_pencil_grid *g, const double *d_fft, const unsigned long long *peer_wave, void *stream);

unsigned long long szb_fft_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
