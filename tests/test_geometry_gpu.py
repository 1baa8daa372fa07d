"""Device-side geometry (hsbp_blocks_blend_dev, hsbp_blocks_set_geometry_dev, hsbp_blocks_set_synthetic_warp) against the
oracle's transfinite_blend / create_metrics (global_curved.jl:19-51, 136-209) and against the host-side synthetic mesh."""
import numpy as np
import pytest

from oracle import hybrid as orc
from tests.util import flat

pytestmark = pytest.mark.gpu


def curved_block(e, Nr, Ns):
    """edge curves of a quadrilateral with bulging edges (keeps the corners): callables + analytic derivatives"""
    rng = np.random.default_rng(40 + e)
    vx = np.array([0.0, 1.0, 0.1, 1.2]) + 0.1 * rng.uniform(-1, 1, 4) + 2 * e
    vy = np.array([0.0, 0.1, 1.0, 1.1]) + 0.1 * rng.uniform(-1, 1, 4)
    amp = 0.05 * rng.uniform(-1, 1, (2, 4))

    def curves(v, am):
        pairs = ((0, 2), (1, 3), (0, 1), (2, 3))
        f = [(lambda t, a=v[i], b=v[j], c=am[k]: a * (1 - t) / 2 + b * (1 + t) / 2 + c * (1 - t * t)) for k, (i, j) in enumerate(pairs)]
        df = [(lambda t, a=v[i], b=v[j], c=am[k]: (b - a) / 2 - 2 * c * t + 0 * t) for k, (i, j) in enumerate(pairs)]
        return f, df
    return curves(vx, amp[0]), curves(vy, amp[1])


def test_blend_and_metrics_on_the_device_match_the_oracle(ctx):
    import hybridsbp_b200 as hs
    p = 4
    sizes = [(17, 21), (33, 33), (12, 40)]
    blk = hs.Blocks(ctx, p, [a for a, _ in sizes], [b for _, b in sizes])
    ex, ey, mets = [], [], []
    for e, (Nr, Ns) in enumerate(sizes):
        (fx, dfx), (fy, dfy) = curved_block(e, Nr, Ns)
        r1, s1 = np.linspace(-1, 1, Nr + 1), np.linspace(-1, 1, Ns + 1)
        for f, df, out in ((fx, dfx, ex), (fy, dfy, ey)):
            out.append(np.concatenate([f[0](s1), f[1](s1), f[2](r1), f[3](r1), df[0](s1), df[1](s1), df[2](r1), df[3](r1)]))
        xf = lambda r, s, f=fx, df=dfx: orc.transfinite_blend(f[0], f[1], f[2], f[3], df[0], df[1], df[2], df[3], r, s)
        yf = lambda r, s, f=fy, df=dfy: orc.transfinite_blend(f[0], f[1], f[2], f[3], df[0], df[1], df[2], df[3], r, s)
        mets.append(orc.create_metrics(p, Nr, Ns, xf, yf))
    V, Fn = blk.VNp, blk.FNp
    dx, dxr, dxs = ctx.empty(V), ctx.empty(V), ctx.empty(V)
    dy, dyr, dys = ctx.empty(V), ctx.empty(V), ctx.empty(V)
    blk.blend_dev(ctx.array(np.concatenate(ex)), dx, dxr, dxs)
    blk.blend_dev(ctx.array(np.concatenate(ey)), dy, dyr, dys)
    dJ, dsJ, dnx, dny = ctx.empty(V), ctx.empty(Fn), ctx.empty(Fn), ctx.empty(Fn)
    blk.set_geometry_dev(dxr, dxs, dyr, dys, J=dJ, sJ=dsJ, nx=dnx, ny=dny)
    x, y, J = dx.get(), dy.get(), dJ.get()
    sJ, nx, ny = dsJ.get(), dnx.get(), dny.get()
    for e, m in enumerate(mets):
        sl = blk.vol_slice(e)
        assert np.max(np.abs(x[sl] - flat(m.coord[0]))) <= 1e-13 * np.max(np.abs(m.coord[0]))
        assert np.max(np.abs(y[sl] - flat(m.coord[1]))) <= 1e-13 * max(1.0, np.max(np.abs(m.coord[1])))
        assert np.max(np.abs(J[sl] - flat(m.J))) <= 1e-12 * np.max(np.abs(m.J))
        for lf in range(1, 5):
            fs = blk.face_slice(e, lf)
            assert np.max(np.abs(sJ[fs] - m.sJ[lf - 1])) <= 1e-12 * np.max(m.sJ[lf - 1])
            assert np.max(np.abs(nx[fs] - m.nx[lf - 1])) <= 1e-12 and np.max(np.abs(ny[fs] - m.ny[lf - 1])) <= 1e-12
    # the coefficient fields went straight into the operator: compare M-tilde u with the oracle's assembled matrix
    blk.set_bc(np.tile([1, 0, 2, 7], len(sizes)))
    blk.compute_tau(2.0)
    rng = np.random.default_rng(0)
    u = rng.uniform(-1, 1, V)
    yv = blk.apply_host(u)
    for e, (m, (Nr, Ns)) in enumerate(zip(mets, sizes)):
        lop = orc.locoperator(p, Nr, Ns, m, (1, 0, 2, 7))
        sl = blk.vol_slice(e)
        ref = lop.Mt @ u[sl]
        assert np.max(np.abs(yv[sl] - ref)) <= 1e-11 * np.max(abs(lop.Mt) @ np.abs(u[sl]))
    # edge curves that do not meet at the corners are refused (global_curved.jl:25)
    bad = np.concatenate(ex).copy()
    bad[0] += 0.3
    from hybridsbp_b200._lib import HsbpError
    with pytest.raises(HsbpError) as err:
        blk.blend_dev(ctx.array(bad), dx, dxr, dxs)
    assert "corners" in str(err.value)
    blk.close()


def test_synthetic_warp_on_the_device_matches_the_host_generator(ctx):
    import hybridsbp_b200 as hs
    from hybridsbp_b200 import synthetic
    p, N, nbx, nby, bx0, L = 4, 31, 3, 2, 5, 16.0
    crr, css, crs = synthetic.warped_coefficients(nbx, nby, N, L=L, A=L / 40.0, bx0=bx0)
    bcs = np.tile([1, 1, 2, 2], nbx * nby)
    a = hs.Blocks(ctx, p, [N] * (nbx * nby), [N] * (nbx * nby))
    a.set_metrics(crr, css, crs); a.set_bc(bcs); a.compute_tau(2.0)
    b = hs.Blocks(ctx, p, [N] * (nbx * nby), [N] * (nbx * nby))
    dx, dy = ctx.empty(b.VNp), ctx.empty(b.VNp)
    b.set_synthetic_warp(nbx, bx0, L, L / 40.0, dx, dy)
    b.set_bc(bcs); b.compute_tau(2.0)
    assert np.max(np.abs(a.get_tau() - b.get_tau())) <= 1e-11 * np.max(a.get_tau())
    u = np.random.default_rng(1).uniform(-1, 1, a.VNp)
    ya, yb = a.apply_host(u), b.apply_host(u)
    assert np.max(np.abs(ya - yb)) <= 1e-11 * np.max(np.abs(ya))
    xf, yf = synthetic.warp_maps(bx0 + 1, 1, L, L / 40.0)           # block e = 1 + nbx * 1
    m = orc.create_metrics(p, N, N, xf, yf)
    sl = b.vol_slice(1 + nbx)
    assert np.max(np.abs(dx.get()[sl] - flat(m.coord[0]))) <= 1e-13 * np.max(np.abs(m.coord[0]))
    assert np.max(np.abs(dy.get()[sl] - flat(m.coord[1]))) <= 1e-13 * np.max(np.abs(m.coord[1]))
    a.close(); b.close()
