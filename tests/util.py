"""Shared helpers for the parity tests (oracle side = checker only)."""
import numpy as np

from oracle import hybrid as orc


def random_spd_metrics(p, Nr, Ns, rng, scale2=1e-4):
    """Random SPD coefficient tensor field, recipe of local_op_eigenvalues.jl:32-38."""
    m = orc.create_metrics(p, Nr, Ns)
    l1 = rng.random((Nr + 1, Ns + 1))
    l2 = rng.random((Nr + 1, Ns + 1)) * scale2
    q = np.pi * rng.random((Nr + 1, Ns + 1))
    m.crr = l1 * np.cos(q) ** 2 + l2 * np.sin(q) ** 2
    m.css = l1 * np.sin(q) ** 2 + l2 * np.cos(q) ** 2
    m.crs = (l2 - l1) * np.cos(q) * np.sin(q)
    return m


def warped_metrics(p, Nr, Ns, bx=0, by=0, nbx=1, nby=1, amp=None):
    """Block (bx, by) of the synthetic warped multiblock mesh of SURVEY.md section 8(d):
    x = xi + A sin(2 pi xi/L) sin(2 pi eta/L), y = eta - A sin(...) sin(...), on [0,nbx]x[0,nby]."""
    L = float(max(nbx, nby))
    A = L / 40.0 if amp is None else amp
    k = 2 * np.pi / L

    def xf(r, s):
        xi = bx + (r + 1) / 2
        et = by + (s + 1) / 2
        w = A * np.sin(k * xi) * np.sin(k * et)
        wxi = A * k * np.cos(k * xi) * np.sin(k * et)
        wet = A * k * np.sin(k * xi) * np.cos(k * et)
        return xi + w, (1 + wxi) / 2, wet / 2

    def yf(r, s):
        xi = bx + (r + 1) / 2
        et = by + (s + 1) / 2
        w = A * np.sin(k * xi) * np.sin(k * et)
        wxi = A * k * np.cos(k * xi) * np.sin(k * et)
        wet = A * k * np.sin(k * xi) * np.cos(k * et)
        return et - w, -wxi / 2, (1 - wet) / 2

    return orc.create_metrics(p, Nr, Ns, xf, yf)


def flat(a):
    return np.asarray(a).reshape(-1, order="F")


def upload_blocks(hs, ctx, p, lops_metrics, bcs, tauscale=2.0):
    """Make a device Blocks object from oracle metrics (one per block) and bc codes."""
    Nr = [m.crr.shape[0] - 1 for m in lops_metrics]
    Ns = [m.crr.shape[1] - 1 for m in lops_metrics]
    blk = hs.Blocks(ctx, p, Nr, Ns)
    blk.set_metrics(np.concatenate([flat(m.crr) for m in lops_metrics]),
                    np.concatenate([flat(m.css) for m in lops_metrics]),
                    np.concatenate([flat(m.crs) for m in lops_metrics]))
    blk.set_bc(np.asarray(bcs, dtype=np.int64).reshape(-1))
    blk.compute_tau(tauscale)
    return blk


def rel_err_apply(y, yref, Mt, u):
    """normwise error ||y - yref||_inf / || |M| |u| ||_inf  (SURVEY.md section 8d)."""
    scale = np.max(abs(Mt) @ np.abs(u))
    return np.max(np.abs(y - yref)) / scale
