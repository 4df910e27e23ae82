# -*- coding: utf-8 -*-
"""
The decomposition behind csrc/qmb_envelope.cu, restated in numpy with the kernels' index arithmetic
(CPU, no GPU needed): Stockham auto-sort passes of radix <= 16 for power-of-two lengths and
Bluestein's chirp-z convolution for arbitrary trace lengths, against numpy's FFT and
scipy.signal.hilbert (what stalta.py:520-523 calls).  The GPU parity test of the kernels themselves
is tests/test_gpu_preprocess.py::test_envelope_matches_scipy_hilbert.
"""

import numpy as np
import pytest
from scipy.signal import hilbert


def _bitrev(r, lr):
    v = 0
    for b in range(lr):
        v |= ((r >> b) & 1) << (lr - 1 - b)
    return v


def _fft_pass(x, M, p, LR, sign):
    """k_fft_pass<LR>: thread i, k = i mod p, u[r] = x[i + r T] w^(r k), DFT_R in registers, y[(i-k) R + k + s p]."""
    R, T = 1 << LR, M >> LR
    y = np.empty_like(x)
    i = np.arange(T)
    k = i & (p - 1)
    j = ((i - k) << LR) + k
    u = [x[i + r * T] * np.exp(sign * 2j * np.pi * r * k / (p * R)) for r in range(R)]
    span = R >> 1
    while span >= 1:                       # decimation in frequency -> bit-reversed positions
        for base in range(0, R, 2 * span):
            for t in range(span):
                a, b = u[base + t], u[base + t + span]
                u[base + t] = a + b
                u[base + t + span] = (a - b) * np.exp(sign * 2j * np.pi * t / (2 * span))
        span >>= 1
    for pos in range(R):
        y[j + _bitrev(pos, LR) * p] = u[pos]
    return y


def _fft_pow2(x, sign=-1):
    M = len(x)
    m = M.bit_length() - 1
    p, rem = 1, m
    while rem > 0:
        LR = min(4, rem)
        x = _fft_pass(x, M, p, LR, sign)
        p <<= LR
        rem -= LR
    return x


def _envelope(x):
    """qmb_envelope_device: chirp-in, FFT_M, filter, FFT_M^-1, spectrum weights, the same again, modulus."""
    N = len(x)
    M = 1
    while M < 2 * N - 1:
        M <<= 1
    n = np.arange(N)
    w = np.exp(1j * np.pi * (((n.astype(object) ** 2) % (2 * N)).astype(float) / N))   # k_chirp
    b = np.zeros(M, complex)                                                           # k_chirp_filter
    b[:N] = w
    b[M - np.arange(1, N)] = w[1:]
    Bf = _fft_pow2(b)
    c = np.zeros(M, complex)                                                           # k_chirp_in
    c[:N] = x * np.conj(w)
    conv = _fft_pow2(_fft_pow2(c) * Bf, +1)
    h = np.zeros(N)                                                                    # k_analytic_mid
    if N % 2 == 0:
        h[0] = h[N // 2] = 1
        h[1:N // 2] = 2
    else:
        h[0] = 1
        h[1:(N + 1) // 2] = 2
    A = np.conj(w) * conv[:N] / M
    c2 = np.zeros(M, complex)
    c2[:N] = np.conj(h * A) * np.conj(w)
    conv2 = _fft_pow2(_fft_pow2(c2) * Bf, +1)
    return np.abs(conv2[:N]) / (M * N), A                                              # k_envelope


@pytest.mark.parametrize("m", [1, 2, 3, 4, 5, 7, 8, 9, 12])
def test_stockham_passes_match_numpy_fft(m):
    rng = np.random.default_rng(m)
    M = 1 << m
    x = rng.normal(size=M) + 1j * rng.normal(size=M)
    np.testing.assert_allclose(_fft_pow2(x), np.fft.fft(x), rtol=0, atol=1e-12 * M)
    np.testing.assert_allclose(_fft_pow2(x, +1), np.fft.ifft(x) * M, rtol=0, atol=1e-12 * M)


@pytest.mark.parametrize("n", [1, 2, 3, 5, 8, 100, 101, 1000, 4099])
def test_bluestein_envelope_matches_scipy_hilbert(n):
    rng = np.random.default_rng(n)
    x = rng.normal(size=n)
    env, spectrum = _envelope(x)
    np.testing.assert_allclose(spectrum, np.fft.fft(x), rtol=0, atol=1e-11 * max(1.0, np.abs(x).sum()))
    np.testing.assert_allclose(env, np.abs(hilbert(x)), rtol=0, atol=1e-12 * np.abs(x).max() * max(1, n) ** 0.5)
