"""CPU emulation of k_band_solve_stream (hybridsbp_b200/csrc/api_band.cuh): panels of PB columns of the banded Cholesky
factor, inverted diagonal blocks, a circular window for the active part of the right-hand side / solution.  Checks the
index arithmetic of the two sweeps (window wrap-around, rows entering late, band edge d <= kd) against a dense solve."""
import numpy as np
import pytest


def emulate(L, kd, g, PB=16, WIN=64):
    n = L.shape[0]
    assert n % PB == 0 and kd + 2 * PB <= WIN and kd >= PB
    npan = n // PB
    mask = WIN - 1
    inv = [np.linalg.inv(L[k * PB:(k + 1) * PB, k * PB:(k + 1) * PB]) for k in range(npan)]
    col = lambda c, d: L[c + d, c] if c + d < n else 0.0          # band storage AB[c * ld + d], zero beyond the matrix
    # ---- L y = g
    win = np.zeros(WIN)
    for i in range(min(PB + kd, WIN)):
        win[i] = g[i] if i < n else 0.0
    y = np.zeros(n)
    for k in range(npan):
        k0 = k * PB
        gnew = [(g[k0 + PB + kd + m] if k0 + PB + kd + m < n else 0.0) for m in range(PB)]
        pv = inv[k] @ np.array([win[(k0 + j) & mask] for j in range(PB)])
        y[k0:k0 + PB] = pv
        for t in range(kd):
            s = 0.0
            for j in range(PB):
                d = t + PB - j
                if d <= kd:
                    s += col(k0 + j, d) * pv[j]
            win[(k0 + PB + t) & mask] -= s
        for m in range(PB):
            win[(k0 + PB + kd + m) & mask] = gnew[m]
    # ---- L^T x = y
    win[:] = 0.0
    x = np.zeros(n)
    for k in reversed(range(npan)):
        k0 = k * PB
        s = np.zeros(PB)
        for j in range(PB):
            for t in range(kd - PB + j + 1):
                s[j] += col(k0 + j, t + PB - j) * win[(k0 + PB + t) & mask]
        xv = inv[k].T @ (y[k0:k0 + PB] - s)
        for i in range(PB):
            win[(k0 + i) & mask] = xv[i]
        x[k0:k0 + PB] = xv
    return x


@pytest.mark.parametrize("n,kd", [(96, 16), (96, 20), (160, 31), (128, 32)])
def test_streamed_band_solve_emulation(n, kd):
    rng = np.random.default_rng(n + kd)
    A = np.zeros((n, n))
    for d in range(kd + 1):
        v = rng.uniform(-1, 1, n - d)
        A += np.diag(v, -d) + (np.diag(v, d) if d else 0)
    A += np.eye(n) * (2 * kd + 4)                                  # diagonally dominant: SPD, banded
    L = np.linalg.cholesky(A)
    assert np.abs(np.tril(L, -kd - 1)).max() == 0.0                # no fill outside the band
    g = rng.uniform(-1, 1, n)
    x = emulate(L, kd, g)
    assert np.linalg.norm(x - np.linalg.solve(A, g)) <= 1e-12 * np.linalg.norm(x)
