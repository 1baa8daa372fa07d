"""CPU emulation of the batched eigensolver of the FDM setup (hybridsbp_b200/csrc/k_eig.cuh): one-sided (Hestenes) Jacobi on
G = A V, a sweep = all pairs of 16-column blocks, inside a block pair 31 round-robin steps of 16 disjoint column pairs.  Checked
here: the round-robin schedule meets every pair exactly once, the rotation orthogonalises a pair, and on symmetric positive
definite matrices the loop converges to the eigen-decomposition (ascending order after the rank sort)."""
import numpy as np

CB = 16


def round_robin(step, k, RR=2 * CB - 1):
    p = RR if k == 0 else (step + k) % RR
    q = (step + RR - k) % RR
    return p, q


def test_round_robin_meets_every_pair_once():
    seen = set()
    for step in range(2 * CB - 1):
        cols = set()
        for k in range(CB):
            p, q = round_robin(step, k)
            assert p != q and p not in cols and q not in cols        # disjoint pairs within a step
            cols.update((p, q))
            seen.add((min(p, q), max(p, q)))
        assert len(cols) == 2 * CB
    assert len(seen) == 2 * CB * (2 * CB - 1) // 2


def rotate(gp, gq, vp, vq, tol):
    a, b, c = gp @ gp, gq @ gq, gp @ gq
    if not abs(c) > tol * np.sqrt(a * b):
        return False
    d = b - a
    t = np.copysign(2.0 * c, d * c) / (abs(d) + np.sqrt(d * d + 4.0 * c * c))
    cs = 1.0 / np.sqrt(1.0 + t * t)
    sn = cs * t
    gp[:], gq[:] = cs * gp - sn * gq, sn * gp + cs * gq
    vp[:], vq[:] = cs * vp - sn * vq, sn * vp + cs * vq
    return True


def test_rotation_orthogonalises_the_pair():
    rng = np.random.default_rng(1)
    for _ in range(20):
        gp, gq = rng.normal(size=40), rng.normal(size=40)
        vp, vq = rng.normal(size=40), rng.normal(size=40)
        n0 = gp @ gp + gq @ gq
        assert rotate(gp, gq, vp, vq, 1e-15)
        assert abs(gp @ gq) <= 1e-14 * n0 and abs(gp @ gp + gq @ gq - n0) <= 1e-13 * n0


def jacobi_eig(A, max_sweeps=30, tol=1e-15):
    n = A.shape[0]
    G, V = A.copy(), np.eye(n)
    nbk = -(-n // CB)
    for sweep in range(max_sweeps):
        rotated = False
        for I in range(nbk):
            for J in range(I + 1, nbk):
                cols = [I * CB + c for c in range(CB)] + [J * CB + c for c in range(CB)]
                for step in range(2 * CB - 1):
                    for k in range(CB):
                        p, q = round_robin(step, k)
                        gp, gq = cols[p], cols[q]
                        if gp >= n or gq >= n:
                            continue                                 # padding columns of the last block: zero, never rotate
                        rotated |= rotate(G[:, gp], G[:, gq], V[:, gp], V[:, gq], tol)
        if not rotated:
            break
    lam = np.linalg.norm(G, axis=0)
    order = np.argsort(lam, kind="stable")
    return lam[order], V[:, order], sweep + 1


def test_block_jacobi_converges_to_the_eigen_decomposition():
    rng = np.random.default_rng(7)
    for n in (32, 40, 64):
        Q, _ = np.linalg.qr(rng.normal(size=(n, n)))
        lam_true = np.sort(np.concatenate([rng.uniform(1e-3, 1.0, n - 4), rng.uniform(50, 4000, 4)]))   # stiff 1-D operator spectrum
        A = (Q * lam_true) @ Q.T
        A = 0.5 * (A + A.T)
        lam, V, sweeps = jacobi_eig(A)
        assert sweeps < 30
        assert np.allclose(lam, lam_true, rtol=1e-9, atol=1e-12 * lam_true[-1])
        assert np.linalg.norm(V.T @ V - np.eye(n)) < 1e-12
        assert np.linalg.norm(A @ V - V * lam) < 1e-10 * lam_true[-1]
