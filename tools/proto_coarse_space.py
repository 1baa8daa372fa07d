"""Prototype (CPU, oracle only) for the next step of the trace solve: a two-level preconditioner for
B = D - Fbar^T M^-1 Fbar on an nb x nb mesh of warped blocks.  Compares CG iteration counts to 1e-10 for
  jacobi        D                                   (round-1 start)
  face blocks   exact diagonal blocks B_ff          (round-1 end, hsbp_trace_precond_setup)
  two-level     face blocks + coarse space of q polynomials per face (Legendre modes 0 .. q-1), additive
usage: python tools/proto_coarse_space.py [nb] [N] [p]"""
import sys
import numpy as np
import scipy.sparse as sp
sys.path.insert(0, ".")
from hybridsbp_b200 import synthetic
from hybridsbp_b200.host import connectivityarrays
from oracle import hybrid as orc
from tests.util import warped_metrics

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 6
N = int(sys.argv[2]) if len(sys.argv) > 2 else 17
p = int(sys.argv[3]) if len(sys.argv) > 3 else 4
_, EToV, EToF, FToB = synthetic.block_grid_connectivity(nb, nb)
FToE, FToLF, EToO, EToS = connectivityarrays(EToV, EToF)
ne = nb * nb
bcs = [[FToB[f - 1] for f in EToF[:, e]] for e in range(ne)]
lops = [orc.locoperator(p, N, N, warped_metrics(p, N, N, e % nb, e // nb, nb, nb), bcs[e]) for e in range(ne)]
M, FbarT, D, vstarts, Fl = orc.LocalGlobalOperators(lops, [N] * ne, [N] * ne, FToB, FToE, FToLF, EToO, EToS)
B = orc.assemblelambdamatrix(Fl, vstarts, EToF, FToB, M.F, D, FbarT).toarray()
B = 0.5 * (B + B.T)
n = B.shape[0]
starts = np.asarray(Fl) - 1
faces = [(a, b) for a, b in zip(starts[:-1], starts[1:]) if b > a]
rng = np.random.default_rng(0)
rhs = rng.uniform(-1, 1, n)


def pcg(apply_prec, tol=1e-10, maxit=20000):
    x = np.zeros(n); r = rhs.copy(); z = apply_prec(r); q = z.copy(); rz = r @ z; b2 = rhs @ rhs
    for it in range(1, maxit + 1):
        Bq = B @ q
        al = rz / (q @ Bq)
        x += al * q; r -= al * Bq
        if np.sqrt(r @ r / b2) <= tol:
            return it
        z = apply_prec(r); rz2 = r @ z
        q = z + (rz2 / rz) * q; rz = rz2
    return maxit


blocks = [np.linalg.inv(B[a:b, a:b]) for a, b in faces]


def face_blocks(r):
    z = np.zeros(n)
    for (a, b), Bi in zip(faces, blocks):
        z[a:b] = Bi @ r[a:b]
    return z


print("mesh %d x %d blocks of %d x %d points, p = %d: %d lambda points on %d faces, cond(B) = %.2e" %
      (nb, nb, N + 1, N + 1, p, n, len(faces), np.linalg.cond(B)))
print("jacobi (D)      : %5d iterations" % pcg(lambda r: r / D))
print("face blocks     : %5d iterations" % pcg(face_blocks))
s = np.linspace(-1, 1, N + 1)
for q in (1, 2, 3):
    cols = []
    for a, b in faces:
        for k in range(q):
            v = np.zeros(n); v[a:b] = np.polynomial.legendre.Legendre.basis(k)(s); cols.append(v)
    Z = np.array(cols).T
    Ac = np.linalg.inv(Z.T @ B @ Z)
    print("two-level, q = %d: %5d iterations   (coarse problem %d x %d)" %
          (q, pcg(lambda r: face_blocks(r) + Z @ (Ac @ (Z.T @ r))), Z.shape[1], Z.shape[1]))
