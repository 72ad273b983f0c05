#!/usr/bin/env python3
"""NumPy model of the index arithmetic used by the CUDA FFT kernels (development aid).

* in-place radix-R DIF forward / DIT inverse pair with digit-reversed spectrum (csrc/pc_kernels.cu)
* out-of-place Stockham stage with one thread per output element (csrc/mtd_kernels.cu, generic P)
Run: python tools/fft_index_model.py
"""
import numpy as np


def dif_forward(x, R, S):
    NT = R ** S
    d = x.astype(complex).copy()
    for s in range(S):
        stride = NT // R ** (s + 1)
        new = d.copy()
        for u in range(NT // R):
            q, blk = u % stride, u // stride
            base = blk * stride * R + q
            v = d[base + stride * np.arange(R)]
            V = np.array([sum(v[j] * np.exp(-2j * np.pi * j * k / R) for j in range(R)) for k in range(R)])
            if stride > 1:
                V = V * np.exp(-2j * np.pi * (q * np.arange(R) * R ** s) / NT)
            new[base + stride * np.arange(R)] = V
        d = new
    return d


def dit_inverse(d, R, S):
    NT = R ** S
    d = d.copy()
    for s in range(S - 1, -1, -1):
        stride = NT // R ** (s + 1)
        new = d.copy()
        for u in range(NT // R):
            q, blk = u % stride, u // stride
            base = blk * stride * R + q
            v = d[base + stride * np.arange(R)]
            if stride > 1:
                v = v * np.exp(+2j * np.pi * (q * np.arange(R) * R ** s) / NT)
            y = np.array([sum(v[k] * np.exp(+2j * np.pi * j * k / R) for k in range(R)) for j in range(R)])
            new[base + stride * np.arange(R)] = y
        d = new
    return d


def digit_reverse_perm(R, S):
    """freq[pos]: frequency index held at in-place position pos after the DIF forward."""
    NT = R ** S
    freq = np.zeros(NT, dtype=int)
    for pos in range(NT):
        p, f = pos, 0
        for i in range(S):
            digit = p // (NT // R ** (i + 1)) % R
            f += digit * R ** i
        freq[pos] = f
    return freq


def stockham(x, radices):
    P = len(x)
    a = x.astype(complex).copy()
    w = np.exp(-2j * np.pi * np.arange(P) / P)
    Ns = 1
    for Rr in radices:
        b = np.zeros(P, dtype=complex)
        span = Ns * Rr
        for o in range(P):
            m = o % span
            high = o // span
            j = high * Ns + (m % Ns)
            step = (m * (P // span)) % P
            idx = 0
            acc = 0
            for t in range(Rr):
                acc += a[j + t * (P // Rr)] * w[idx]
                idx += step
                if idx >= P:
                    idx -= P
            b[o] = acc
        a = b
        Ns = span
    return a


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for R, S in ((8, 3), (16, 2), (4, 3), (8, 2)):
        NT = R ** S
        x = rng.normal(size=NT) + 1j * rng.normal(size=NT)
        X = dif_forward(x, R, S)
        fr = digit_reverse_perm(R, S)
        assert np.allclose(X, np.fft.fft(x)[fr]), (R, S)
        y = dit_inverse(X, R, S) / NT
        assert np.allclose(y, x), (R, S)
        print("DIF/DIT ok", R, S)
    for P, rad in ((8, [8]), (64, [8, 8]), (1536, [16, 16, 6]), (1536, [3, 8, 8, 8]), (155, [5, 31]), (7, [7]), (12, [4, 3])):
        x = rng.normal(size=P) + 1j * rng.normal(size=P)
        assert np.allclose(stockham(x, rad), np.fft.fft(x)), (P, rad)
        print("Stockham ok", P, rad)
