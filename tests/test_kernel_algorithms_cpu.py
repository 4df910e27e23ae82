# -*- coding: utf-8 -*-
"""
The two non-obvious decompositions the CUDA kernels rely on, restated in numpy and checked on the
CPU (no GPU needed; the kernels themselves are checked against the oracle in the -m gpu suite):

* the 16-bit scan's exact integer stack (csrc/qmb_scan.cu, MODE 4+): whole-word sums modulo 2^32
  plus upper-half sums recover both per-sample sums, at the largest row count and values allowed;
  and the quantisation error bound that feeds the float64 candidate re-evaluation;
* the chunk-parallel IIR of csrc/qmb_preprocess.cu: zero-state chunk responses + Hillis-Steele scan
  with the powers of the chunk transition matrix + tile carry + re-run from the true states equals
  scipy's serial sosfilt.
"""
import numpy as np
import pytest

from oracle import preprocess_oracle as PO


def test_q16_word_sums_recover_both_halves_exactly():
    rng = np.random.default_rng(0)
    for S in (1, 7, 32, 128, 256):
        for fill in ("random", "max"):
            lo = rng.integers(0, 65536, S, dtype=np.uint32) if fill == "random" else np.full(S, 65535, np.uint32)
            hi = rng.integers(0, 65536, S, dtype=np.uint32) if fill == "random" else np.full(S, 65535, np.uint32)
            words = (hi << np.uint32(16)) | lo
            A = np.uint32(0)
            H = np.uint32(0)
            with np.errstate(over="ignore"):
                for w in words:
                    A = np.uint32(A + w)               # IMAD / IADD3: wraps modulo 2^32
                    H = np.uint32(H + (w >> np.uint32(16)))  # LEA.HI
                got_lo = np.uint32(A - np.uint32(H << np.uint32(16)))
            assert int(H) == int(hi.astype(np.uint64).sum())
            assert int(got_lo) == int(lo.astype(np.uint64).sum())
            assert int(H) < 2 ** 24 and int(got_lo) < 2 ** 24  # exact in float32 (S <= 256)


def test_q16_error_band_is_rigorous():
    """|quantised stack - exact| <= S*step/2 + the float32 evaluation of c0 + step*sum(u)."""
    rng = np.random.default_rng(1)
    for S in (4, 32, 128):
        x = np.log(np.clip(rng.gamma(2.0, 0.8, size=(S, 4000)), 0.4, np.inf))
        x[rng.random(x.shape) < 0.3] = np.log(0.4)
        qmin, qmax = x.min(), x.max()
        step = (qmax - qmin) / 65535.0
        u = np.clip(np.rint((x - qmin) * (1.0 / step)), 0, 65535)
        assert (u[x == qmin] == 0).all()  # the clip floor sits on the grid: no error at all there
        idx = rng.integers(0, x.shape[1], size=(3000, S))
        exact = x[np.arange(S), idx].sum(axis=1)
        usum = u[np.arange(S), idx].sum(axis=1)
        v = (np.float32(usum) * np.float32(step) + np.float32(S * qmin)).astype(np.float32)  # kernel: one fma
        delta = 0.5 * S * step + 4.0 * 2.0 ** -24 * S * (abs(qmin) + 65535.0 * step)
        assert np.abs(v.astype(np.float64) - exact).max() <= delta
        # typical error is far below the bound: ~ step * sqrt(S / 12)
        assert np.abs(v.astype(np.float64) - exact).std() < 2.0 * step * np.sqrt(S / 12.0) + 1e-6


L, TPB = 32, 256


def _step(coef, s, x):
    for k, (b0, b1, b2, a1, a2) in enumerate(coef):
        y = b0 * x + s[2 * k]
        s[2 * k] = b1 * x - a1 * y + s[2 * k + 1]
        s[2 * k + 1] = b2 * x - a2 * y
        x = y
    return x


def _chunked_pass(x, coef):
    """k_sos_state + k_sos_carry + k_sos_apply for one trace in processing order."""
    n, D, tile = len(x), 2 * len(coef), L * TPB
    ntiles = (n + tile - 1) // tile
    M = np.zeros((D, D))
    for j in range(D):
        s = np.zeros(D)
        s[j] = 1.0
        for _ in range(L):
            _step(coef, s, 0.0)
        M[:, j] = s
    mp = [np.eye(D)]
    for _ in range(TPB):
        mp.append(mp[-1] @ M)
    xp = np.zeros(ntiles * tile)
    xp[:n] = x
    chunk_state, tile_state = np.zeros((ntiles, TPB, D)), np.zeros((ntiles, D))
    for T in range(ntiles):
        st = np.zeros((TPB, D))
        for t in range(TPB):
            s = np.zeros(D)
            for k in range(L):
                _step(coef, s, xp[T * tile + t * L + k])
            st[t] = s
        o = 1
        while o < TPB:  # Hillis-Steele with M^o
            new = st.copy()
            new[o:] += st[:-o] @ mp[o].T
            st, o = new, o << 1
        chunk_state[T, 1:], tile_state[T] = st[:-1], st[-1]
    tile_in, cur = np.zeros((ntiles, D)), np.zeros(D)
    for T in range(ntiles):
        tile_in[T] = cur
        cur = tile_state[T] + mp[TPB] @ cur
    y = np.zeros(ntiles * tile)
    for T in range(ntiles):
        for t in range(TPB):
            s = chunk_state[T, t] + mp[t] @ tile_in[T]
            for k in range(L):
                q = T * tile + t * L + k
                y[q] = _step(coef, s, xp[q])
    return y[:n]


@pytest.mark.parametrize("band,corners", [((2.0, 16.0), 2), ((0.1, 0.5), 2), ((1.0, 3.0), 3)])
def test_chunk_parallel_iir_equals_serial_sosfilt(band, corners):
    rng = np.random.default_rng(corners)
    n = 9000  # two tiles, the second one partial
    x = PO.detrend_taper(rng.standard_normal(n) * 100.0 + np.linspace(0.0, 500.0, n))
    sos = PO.butter_sos("bandpass", band, 100.0, corners)
    coef = [(r[0], r[1], r[2], r[4], r[5]) for r in sos]
    fwd = _chunked_pass(x, coef)
    both = _chunked_pass(fwd[::-1], coef)[::-1]
    want = PO.sosfilt_zerophase(sos, x)
    assert np.abs(both - want).max() <= 1e-9 * np.abs(want).max()
