"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the wave <-> physical transforms behind
suzerain::pencil_grid (suzerain/pencil_grid.hpp:200-221, pencil_grid.cpp:186-194).

The arithmetic lives in a third-party dependency that is absent here: P3DFFT (the reference's
patched copy under lib/suzerain-p3dfft is Fortran; no gfortran, FFTW or MPI in this image), built
with STRIDE1 so that wave space is stride one in Y (pencil_grid.cpp:124-131).  Its published
algorithm: backward = complex FFT in z with exp(+i...) then complex-to-real in x; forward =
real-to-complex in x then complex FFT in z with exp(-i...); neither is normalised
(tests/test_diffwave_p3dfft.cpp:131-134 rescales by dNx*dNz).  PARITY UNPINNED against P3DFFT
itself; anchored instead on the reference's own use of it: tests/test_diffwave_p3dfft.cpp
(analytic derivatives through physical -> wave -> diffwave -> physical) is replayed in
tests/test_pencil_grid.py, and numpy.fft has FFTW's sign convention.

Also the slab protocol (pack / all-to-all / finish) in numpy, used by the CPU gloo test to pin the
block order and counts the CUDA pack kernels must produce."""
import numpy as np


def wave_to_physical(wave, dNx):
    """wave: complex [dNz][dNx/2+1][Ny]  ->  real [Ny][dNz][dNx]."""
    dNz = wave.shape[0]
    w = np.transpose(wave, (2, 0, 1))
    t = np.fft.ifft(w, axis=1) * dNz
    return np.fft.irfft(t, n=dNx, axis=2) * dNx


def physical_to_wave(phys):
    """real [Ny][dNz][dNx]  ->  complex [dNz][dNx/2+1][Ny]."""
    t = np.fft.fft(np.fft.rfft(phys, axis=2), axis=1)
    return np.ascontiguousarray(np.transpose(t, (1, 2, 0)))


def slab_bounds(n, nranks):
    return [r * n // nranks for r in range(nranks + 1)]


def w2p_pack(wave_local, ys):
    """wave_local [Zloc][X][Y] -> list of blocks, block s = [Yloc_s][Zloc][X]."""
    return [np.ascontiguousarray(np.transpose(wave_local[:, :, ys[s]:ys[s + 1]], (2, 0, 1))) for s in range(len(ys) - 1)]


def w2p_finish(blocks, dNx):
    """blocks r = [Yloc][Zloc_r][X] -> physical [Yloc][dNz][dNx]."""
    w = np.concatenate(blocks, axis=1)
    t = np.fft.ifft(w, axis=1) * w.shape[1]
    return np.fft.irfft(t, n=dNx, axis=2) * dNx


def p2w_start(phys_local, zs):
    """physical [Yloc][dNz][dNx] -> list of blocks, block r = [Yloc][Zloc_r][X]."""
    t = np.fft.fft(np.fft.rfft(phys_local, axis=2), axis=1)
    return [np.ascontiguousarray(t[:, zs[r]:zs[r + 1], :]) for r in range(len(zs) - 1)]


def p2w_unpack(blocks):
    """blocks s = [Yloc_s][Zloc][X] -> wave_local [Zloc][X][Y]."""
    return np.ascontiguousarray(np.transpose(np.concatenate(blocks, axis=0), (1, 2, 0)))
