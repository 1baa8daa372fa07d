"""SEAS BP1 on a MULTIBLOCK mesh: the antiplane earthquake-cycle problem of seas/BP1/BP1.jl on
`meshes/BP1_v1.inp` -- 194 blocks covering both sides of the fault, [-400, 400] x [-400, 0] km -- which the reference
ships without a driver.  Written in the style of the reference's own drivers:

  * side set 1 (x = +-400 km)   Dirichlet, u = +- Vp t / 2                      (BP1.jl:36 for the single block)
  * side set 2 (y = 0, -400 km) Neumann, traction free
  * side set 7 (x = 0, 0-40 km) frictional fault: a JUMP interface whose jump is the slip delta from the ODE state
  * side set 8 (x = 0, below)   steady sliding: a jump interface with delta = Vp t
  the jump enters through the jump branch of locbcarray! (global_curved.jl:614-617) with the `in_jump` sign conventions of
  square_circle.jl:335-350; displacement and lambda come from the trace solve (square_circle.jl:376-388, K4); the shear
  stress on the fault from computetraction on the minus side (global_curved.jl:638-644); the slip rate and the state
  evolution per fault node from rateandstate / newtbndv exactly as in odefun.jl:69-108 (K5).

Everything from the boundary data to the stress change is linear in (delta, t): `FaultOperator` either performs one trace
solve per right-hand-side evaluation (mode "solve") or forms dtau = A delta + t b once with n + 1 trace solves and lets
the device evaluate a stage as one small kernel (mode "condensed", hsbp_fault_rhs).
"""
import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from . import bp1, host
from ._lib import Bp1Params, Bp1Stats, _f64, lib
from .blocks import Blocks, Trace, LOCAL_CHOLESKY, LOCAL_BAND

BC_FAULT, BC_STEADY = 7, 8


def default_mesh_path():
    here = os.path.dirname(os.path.abspath(__file__))
    return os.path.join(os.path.dirname(here), "meshes", "BP1_v1.inp")


@dataclass
class MultiblockSetup:
    p: int
    N: int
    mesh: tuple
    conn: tuple
    mets: list
    params: dict
    fault_faces: np.ndarray      # global face ids (0-based) of the frictional faces, by increasing depth
    steady_faces: np.ndarray
    sign: dict                   # face -> +1 if the plus side of the face is the x > 0 block (mesh jump = physical slip), else -1
    depth: np.ndarray            # depth (km, positive) of the fault nodes, face by face in minus-side orientation
    RSa: np.ndarray
    psi_delta0: np.ndarray


def setup(N=17, SBPp=2, filename=None):
    """mesh, geometry and the rate-and-state fields (BP1.jl:8-23, 96-118 with depth = -y on the fault)"""
    verts, EToV, EToF, FToB, _ = host.read_inp_2d(filename or default_mesh_path())
    conn = host.connectivityarrays(EToV, EToF)
    FToE, FToLF, EToO, EToS = conn
    ne = EToV.shape[1]
    mets = []
    for e in range(ne):
        vx, vy = verts[0, EToV[:, e] - 1], verts[1, EToV[:, e] - 1]
        xf = lambda r, s, v=vx: host.transfinite_blend(v[0], v[1], v[2], v[3], r, s)
        yf = lambda r, s, v=vy: host.transfinite_blend(v[0], v[1], v[2], v[3], r, s)
        mets.append(host.create_metrics(SBPp, N, N, xf, yf))
    base = bp1.setup(N=4, SBPp=2)              # physical constants only (BP1.jl:8-23)
    P = dict(base.params)
    cx = np.array([verts[0, EToV[:, e] - 1].mean() for e in range(ne)])
    faces = lambda code: np.array([f for f in range(len(FToB)) if FToB[f] == code], dtype=np.int64)
    ydepth = lambda f: -mets[FToE[0, f] - 1].facecoord[1][FToLF[0, f] - 1]
    fault = faces(BC_FAULT)
    fault = fault[np.argsort([ydepth(f).mean() for f in fault])]
    steady = faces(BC_STEADY)
    sign = {int(f): (1.0 if cx[FToE[1, f] - 1] > 0 else -1.0) for f in np.concatenate([fault, steady])}
    depth = np.concatenate([ydepth(f) for f in fault])
    RSamin, RSamax, RSb, RSH1, RSH2 = 0.01, 0.025, 0.015, 15.0, 18.0
    RSa = RSamin - (RSamin - RSamax) * np.minimum(1.0, np.maximum(0.0, (RSH1 - depth) / (RSH1 - RSH2)))      # BP1.jl:99-102
    RSV0, RSVinit, RSDc, RSf0, sn, eta = P["V0"], 1e-9, P["Dc"], P["f0"], P["sigma_n"], P["eta"]
    theta = (RSDc / RSV0) * np.exp((RSa / RSb) * np.log((2 * RSV0 / RSVinit) * np.sinh((P["tau_z0"] - eta * RSVinit) / (RSa * sn)))
                                   - RSf0 / RSb)
    psi0 = RSf0 + RSb * np.log(RSV0 * theta / RSDc)                                                           # BP1.jl:108-113
    n = len(depth)
    return MultiblockSetup(SBPp, N, (verts, EToV, EToF, FToB), conn, mets, P, fault, steady, sign, depth, RSa,
                           np.concatenate([psi0, np.zeros(n)]))


def boundary_data(su: MultiblockSetup, tau_faces, Hw, delta, t):
    """block-face data v (what F_k multiplies) and g_delta for slip `delta` on the frictional faces at time t:
    locbcarray! (global_curved.jl:596-623) with in_jump as square_circle.jl:335-350.  tau_faces[e][lf]: penalty of block e.
    Returns (v as {(e, lf): array}, gd in the lambda layout of FTols)."""
    verts, EToV, EToF, FToB = su.mesh
    FToE, FToLF, EToO, EToS = su.conn
    P, N = su.params, su.N
    nl = N + 1
    jump = {}
    for i, f in enumerate(su.fault_faces):
        jump[int(f)] = su.sign[int(f)] * delta[i * nl:(i + 1) * nl]                 # mesh jump (plus - minus), minus-side orientation
    for f in su.steady_faces:
        jump[int(f)] = su.sign[int(f)] * np.full(nl, P["Vp"] * t)
    v = {}
    gd_parts = []
    for e in range(EToV.shape[1]):
        for lf in range(4):
            f = int(EToF[lf, e] - 1)
            code = FToB[f]
            if code == host.BC_DIRICHLET:
                x = su.mets[e].facecoord[0][lf]
                v[(e, lf)] = np.sign(x) * (P["Vp"] * t / 2)
            elif code == host.BC_NEUMANN:
                v[(e, lf)] = np.zeros(nl)
            elif code >= host.BC_JUMP_INTERFACE:
                d = jump[f]
                if EToS[lf, e] == 1:
                    dj = -d
                else:
                    dj = d if EToO[lf, e] else d[::-1]
                vf = dj / 2
                v[(e, lf)] = vf
                contrib = Hw * tau_faces[e][lf] * vf
                gd_parts.append((f, contrib if EToO[lf, e] else contrib[::-1]))
    return v, gd_parts


class FaultOperator:
    """The multiblock BP1 right-hand side on the GPU."""

    def __init__(self, ctx, su: MultiblockSetup, mode="condensed", local_mode=None, tol=1e-13):
        self.ctx, self.su, self.mode, self.tol = ctx, su, mode, tol
        verts, EToV, EToF, FToB = su.mesh
        FToE, FToLF, EToO, EToS = su.conn
        ne, N, p = EToV.shape[1], su.N, su.p
        fl = lambda a: np.asarray(a).reshape(-1, order="F")
        blk = Blocks(ctx, p, [N] * ne, [N] * ne)
        blk.set_metrics(np.concatenate([fl(m.crr) for m in su.mets]), np.concatenate([fl(m.css) for m in su.mets]),
                        np.concatenate([fl(m.crs) for m in su.mets]))
        blk.set_bc(np.array([[FToB[f - 1] for f in EToF[:, e]] for e in range(ne)], dtype=np.int64).reshape(-1))
        blk.compute_tau(2.0)
        blk.local_setup(local_mode or (LOCAL_CHOLESKY if (N + 1) ** 2 <= 2500 else LOCAL_BAND), tol=1e-14, maxit=200000)
        tr = Trace(blk, FToB, FToE, FToLF, EToO, EToS)
        tr.condense()
        tr.precond_setup(1)
        tr.coarse_setup(2)
        self.blk, self.tr = blk, tr
        self.FTols = tr.FTolambdastarts
        tau = blk.get_tau()
        self.tau_faces = [[tau[blk.face_slice(e, lf + 1)] for lf in range(4)] for e in range(ne)]
        self.Hw = host.norm_weights(p, N)
        self.dg, self.dv, self.dgd = ctx.empty(blk.VNp), ctx.empty(blk.FNp), ctx.empty(tr.lNp)
        self.dlam, self.du, self.dtr = ctx.empty(tr.lNp), ctx.empty(blk.VNp), ctx.empty(blk.FNp)
        self.n = len(su.depth)
        self.last_stats, self.trace_stats = None, None
        prm = Bp1Params(**su.params)
        a, pa = _f64(su.RSa)
        A = b = None
        if mode == "condensed":
            # dtau = A delta + t b: the unit responses of the linear chain boundary data -> trace solve -> traction
            b = self.stress_change(np.zeros(self.n), 1.0)
            A = np.empty((self.n, self.n))
            for m in range(self.n):
                e_m = np.zeros(self.n); e_m[m] = 1.0
                A[:, m] = self.stress_change(e_m, 0.0)
        h = C.c_void_p()
        self._A = np.asfortranarray(A) if A is not None else None           # column-major, kept alive for the call
        self._b = np.ascontiguousarray(b) if b is not None else None
        pA = C.c_void_p(self._A.ctypes.data) if A is not None else None
        pb = C.c_void_p(self._b.ctypes.data) if b is not None else None
        ctx._check(lib().hsbp_fault_create(ctx.h, self.n, pA, pb, pa, C.byref(prm), C.byref(h)))
        self.h = h
        self.A, self.b = A, b
        self._ddtau = ctx.empty(self.n)

    def solve_displacement(self, delta, t):
        """lambda and u for slip delta at time t (device-resident; returns the statistics of hsbp_trace_solve)"""
        su, blk, tr = self.su, self.blk, self.tr
        v, gd_parts = boundary_data(su, self.tau_faces, self.Hw, np.asarray(delta, float), float(t))
        vface = np.zeros(blk.FNp)
        for (e, lf), vf in v.items():
            vface[blk.face_slice(e, lf + 1)] = vf
        gd = np.zeros(tr.lNp)
        for f, contrib in gd_parts:
            gd[self.FTols[f] - 1:self.FTols[f + 1] - 1] -= contrib
        self.dv.set(vface)
        self.dgd.set(gd)
        self.dg.zero()
        blk.face_F_add(self.dv, -1.0, self.dg)                          # g_e = - sum_k F_k v_k
        st = tr.solve(self.dg, self.dgd, self.dlam, self.du, tol=self.tol, maxit=2000)
        self.trace_stats = st
        return st

    def stress_change(self, delta, t):
        """dtau on the fault nodes: mu * nx * computetraction(minus side) (global_curved.jl:638-644; odefun.jl:59 for the sign)"""
        su, blk = self.su, self.blk
        FToE, FToLF, EToO, EToS = su.conn
        st = self.solve_displacement(delta, t)
        if st["converged"] != 1:
            raise RuntimeError("trace solve did not converge: %r" % (st,))
        blk.face_traction(self.du, self.dtr)
        trv, lam = self.dtr.get(), self.dlam.get()
        nl = su.N + 1
        out = np.empty(self.n)
        for i, f in enumerate(su.fault_faces):
            e1, lf1 = FToE[0, f] - 1, FToLF[0, f] - 1
            m = su.mets[e1]
            d_mesh = su.sign[int(f)] * np.asarray(delta, float)[i * nl:(i + 1) * nl]
            lamf = lam[self.FTols[f] - 1:self.FTols[f + 1] - 1]
            T = (trv[blk.face_slice(e1, lf1 + 1)] + self.tau_faces[e1][lf1] * (lamf - d_mesh / 2)) / m.sJ[lf1]
            out[i * nl:(i + 1) * nl] = su.params["mu_shear"] * m.nx[lf1] * T
        return out

    def rhs(self, t, y):
        """(dy, rejected) = odefun(y, t) for y = [psi; delta] on the fault nodes"""
        y, py = _f64(y)
        out = np.empty(2 * self.n)
        st = Bp1Stats()
        if self.mode == "condensed":
            self.ctx._check(lib().hsbp_fault_rhs(self.h, float(t), py, C.c_void_p(out.ctypes.data), C.byref(st)))
        else:
            self._ddtau.set(self.stress_change(y[self.n:], t))
            self.ctx._check(lib().hsbp_fault_stage(self.h, self._ddtau.ptr, py, C.c_void_p(out.ctypes.data), C.byref(st)))
        self.last_stats = st.as_dict()
        return out, bool(st.rejected)

    def close(self):
        if self.h is not None:
            if self.ctx.h is not None:
                lib().hsbp_fault_destroy(self.h)
            self.h = None
            self.tr.close(); self.blk.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
