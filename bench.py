#!/usr/bin/env python3
"""Benchmark of the hot path: fp64 matrix-free SBP operator apply (y = M-tilde u) on the synthetic
warped multiblock mesh of BASELINE.json config 4 (1024 blocks x 256x256 points, p = 4) per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one operator apply over all blocks of this rank.  Prints ONE JSON line (rank 0).
Under torchrun every rank holds its own 1024 blocks (a 32-block-wide strip of a 32G x 32 mesh):
weak scaling; the operator apply has no exchange step, so there is no collective in the timed
region besides the barriers.

The reference arm (--impl reference) times the reference's CPU algorithm for this path -- the
sparse product of the assembled M-tilde (global_curved.jl:470-492) -- through the oracle port
(oracle/, C + OpenMP SpMV over the matrices the restated locoperator assembles) on a bounded
sample of the same workload, on all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_DOF = 40.0      # u 8 + crr, css, crs 24 + y 8  


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def bind_to_gpu_numa(local):
    """Pin this process (and the pages it touches from now on) to the NUMA node of its GPU: host buffers of the end-to-end
    path are then allocated next to the GPU's PCIe root instead of all ranks sharing node 0."""
    try:
        import ctypes
        cudart = ctypes.CDLL("libcudart.so")
        buf = ctypes.create_string_buffer(64)
        if cudart.cudaDeviceGetPCIBusId(buf, 64, int(local)) != 0:
            return None
        bus = buf.value.decode().lower()
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def cpu_baseline(p, N, nblk_sample, seconds, threads):
    """Oracle port of the reference's CPU operator apply on `nblk_sample` blocks of the workload."""
    from oracle import hybrid as orc                      # CPU baseline leg only
    from oracle.cbuild import BlockSpMV
    from hybridsbp_b200 import host, synthetic
    t0 = time.time()
    mats = []
    for b in range(nblk_sample):
        xf, yf = synthetic.warp_maps(b, 5, 32.0, 32.0 / 40.0)
        m = orc.create_metrics(p, N, N, xf, yf)
        mats.append(orc.locoperator(p, N, N, m, (1 if b == 0 else 0, 0, 0, 0)).Mt)
    S = BlockSpMV(mats)
    t_asm = time.time() - t0
    nthreads = (os.cpu_count() or S.max_threads) if threads <= 0 else threads      # explicit: torchrun exports OMP_NUM_THREADS=1
    u = np.random.default_rng(778).uniform(-1, 1, S.n)
    y = np.empty(S.n)
    S(u, y, nthreads)
    reps, t0 = 0, time.time()
    while True:
        S(u, y, nthreads)
        reps += 1
        if time.time() - t0 > seconds:
            break
    dt = (time.time() - t0) / reps
    return S, {"value": S.n / dt / 1e9, "unit": "GDOF/s", "cores": int(nthreads), "kind": "port",
               "sample": "%d of 1024 blocks (256x256 points, p=%d): assembled sparse M-tilde (%.1f nnz/row, %.1f s to "
                         "assemble with the oracle), CSR SpMV in C/OpenMP, %d repetitions" %
                         (nblk_sample, p, S.nnz / S.n, t_asm, reps)}


def run_reference(args):
    rank, world, local = dist_env()
    if rank != 0:
        return 0
    p, N = args.p, args.n
    t0 = time.time()
    S, base = cpu_baseline(p, N, args.cpu_blocks, 0.5, 0)
    u = np.random.default_rng(778).uniform(-1, 1, S.n)
    y = np.empty(S.n)
    nthr = base["cores"]
    for _ in range(args.warmup):
        S(u, y, nthr)
    t1 = time.perf_counter()
    for _ in range(args.steps):
        S(u, y, nthr)
    dt = (time.perf_counter() - t1) / args.steps
    val = S.n / dt / 1e9
    base["value"] = val
    line = {"impl": "reference", "metric": "fp64 SBP operator-apply GDOF/s", "value": val, "unit": "GDOF/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, world), "cpu_baseline": base,
            "e2e": {"value": val, "unit": "GDOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference CPU path = sparse M-tilde * u (global_curved.jl:470-492) via the oracle port; "
                    "each step is one SpMV over a bounded sample of the workload's blocks"}
    print(json.dumps(line))
    return 0


def workload_config(args, world):
    return {"workload": "synthetic warped multiblock mesh, %d blocks x %dx%d points per GPU, p=%d (BASELINE config 4%s)"
                        % (args.blocks, args.n + 1, args.n + 1, args.p, "" if world == 1 else "/5"),
            "blocks_per_gpu": args.blocks, "points_per_block": (args.n + 1) ** 2, "sbp_order": args.p,
            "dof_per_gpu": args.blocks * (args.n + 1) ** 2,
            "cache": "inputs (%.2f GB per apply) exceed the 126 MB L2; no flush needed" %
                     (args.blocks * (args.n + 1) ** 2 * 32 / 1e9),
            "parallelism": "blocks partitioned across GPUs, no exchange in operator apply"}


def run_ours(args):
    rank, world, local = dist_env()
    # NCCL writes its version / debug lines to stdout by default; stdout of this program is the one JSON line
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    from hybridsbp_b200 import build as _build
    if rank == 0 or world == 1:
        _build.build()
    import hybridsbp_b200 as hs
    from hybridsbp_b200 import synthetic
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    numa_node = bind_to_gpu_numa(local) if world > 1 else None
    ctx = hs.Context(local)
    p, N = args.p, args.n
    nbx = nby = int(round(args.blocks ** 0.5))
    assert nbx * nby == args.blocks, "--blocks must be a square number"
    Lglob = float(max(nbx * world, nby))
    # this rank's strip of the (nbx*world) x nby mesh
    verts, EToV, EToF, FToB = synthetic.block_grid_connectivity(nbx, nby)
    if world > 1:                         # the strip's left/right outer faces are interfaces unless at the mesh edge
        for by in range(nby):
            if rank > 0:
                FToB[EToF[0, 0 + nbx * by] - 1] = 0
            if rank < world - 1:
                FToB[EToF[1, nbx - 1 + nbx * by] - 1] = 0
    blk = hs.Blocks(ctx, p, [N] * args.blocks, [N] * args.blocks)
    blk.set_synthetic_warp(nbx, rank * nbx, Lglob, Lglob / 40.0)      # metrics generated on the device (hsbp_blocks_set_synthetic_warp)
    blk.set_bc(synthetic.block_bcs(EToF, FToB))
    blk.compute_tau(2.0)
    blk.set_option("sweep_points_per_thread", args.sweep_r)
    blk.set_option("sweep_chunks_per_side", args.sweep_ncs)
    blk.set_option("sweep_deep", args.sweep_deep)
    blk.set_option("sweep_p6_regs", args.sweep_p6_regs)
    blk.set_option("force_generic", 1 if args.generic else 0)
    if args.sweep_no_pdl:
        blk.set_option("sweep_no_pdl", 1)
    try:
        blk.set_option("sweep_fold_faces", 0 if args.no_fold else 1)
    except hs.HsbpError:
        pass                                                # older experiment builds (HSBP_LIB) lack the knob
    rng = np.random.default_rng(778 + rank)
    u_host = rng.uniform(-1, 1, blk.VNp)
    u = ctx.array(u_host)
    y = ctx.empty(blk.VNp)
    dof = blk.VNp

    def barrier():
        ctx.sync()
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        blk.apply(u, y)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ctx.timer_start()
    for _ in range(args.steps):
        blk.apply(u, y)
    ms_total = ctx.timer_stop()
    barrier()
    # dominant-kernel time, live, with events between the stages of the same call
    stage = np.zeros(3)
    nrep = max(3, min(args.steps, 10))
    for _ in range(nrep):
        stage += blk.apply_timed(u, y)
    stage /= nrep
    clocks = sampler.stop() if rank == 0 else None
    barrier()
    ms_step = ms_total / args.steps
    if dist is not None:
        import torch
        t = torch.tensor([ms_step], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step = float(t.item())

    # end to end through the C-ABI with host buffers (pinned): H2D u, apply, D2H y every step
    y_host = np.empty(blk.VNp)
    ctx.host_register(u_host); ctx.host_register(y_host)
    blk.apply_host_pinned(u_host, y_host)
    e2e_steps = max(2, min(args.steps, 5))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        blk.apply_host_pinned(u_host, y_host)
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if dist is not None:
        import torch
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    checksum = float(np.abs(y_host).sum())
    ctx.host_unregister(u_host); ctx.host_unregister(y_host)
    # The timed region of the headline is a few milliseconds at boost clocks.  The same loop held for more than a second runs into the
    # board's power limit (fp64 at ~1 kW): report that sustained rate next to the headline, with its own clock samples.
    sustained = None
    if rank == 0 and args.sustained_steps > 0:
        s2 = ClockSampler(local)
        s2.start()
        ctx.timer_start()
        for _ in range(args.sustained_steps):
            blk.apply(u, y)
        ms2 = ctx.timer_stop() / args.sustained_steps
        sustained = {"steps": args.sustained_steps, "ms_per_step": ms2, "value": dof / (ms2 * 1e-3) / 1e9, "unit": "GDOF/s per GPU",
                     "clocks": s2.stop()}
    variant = blk.apply_variant()
    u.free(); y.free()
    blk.close()                                         # the trace solves below build their own blocks
    del u_host, y_host

    # the same operator apply at the other orders and at a line length of the reference's own drivers (N = 200 per block, BP1:
    # 201 points per line -- odd, served by the pitched variant of k_sweep); rank 0 only, a few launches each
    def apply_case(pp, nn, nblocks):
        tnb = int(round(nblocks ** 0.5))
        b2 = hs.Blocks(ctx, pp, [nn] * nblocks, [nn] * nblocks)
        b2.set_synthetic_warp(tnb, 0, float(tnb), tnb / 40.0)
        _, _, eToF, fToB = synthetic.block_grid_connectivity(tnb, tnb)
        b2.set_bc(synthetic.block_bcs(eToF, fToB))
        b2.compute_tau(2.0)
        uu = ctx.array(np.random.default_rng(5).uniform(-1, 1, b2.VNp))
        yy = ctx.empty(b2.VNp)
        for _ in range(3):
            b2.apply(uu, yy)
        st2 = np.zeros(3)
        for _ in range(5):
            st2 += b2.apply_timed(uu, yy)
        st2 /= 5
        ctx.timer_start()
        for _ in range(10):
            b2.apply(uu, yy)
        ms = ctx.timer_stop() / 10
        out = {"sbp_order": pp, "blocks": nblocks, "points_per_block": (nn + 1) ** 2, "apply_variant": b2.apply_variant(),
               "gdof_per_s": b2.VNp / (ms * 1e-3) / 1e9, "ms_per_apply": ms, "k_sweep_ms": float(st2[0]),
               "other_kernels_ms": float(st2[1]),
               "k_sweep_algorithmic_gbs": BYTES_PER_DOF * b2.VNp / (st2[0] * 1e-3) / 1e9,
               "whole_apply_algorithmic_gbs": BYTES_PER_DOF * b2.VNp / (ms * 1e-3) / 1e9}
        uu.free(); yy.free(); b2.close()
        return out

    other_cases = []
    if rank == 0 and not args.no_other_orders:
        for pp, nn, nbk in ((2, args.n, args.blocks), (6, args.n, args.blocks), (4, 200, args.blocks), (6, 136, args.blocks)):
            if (pp, nn) != (p, N):
                other_cases.append(apply_case(pp, nn, nbk))
    barrier()

    # second half of BASELINE's metric: hybrid trace-CG solve time, through hsbp_trace_solve -- the library's device-resident,
    # two-level preconditioned CG; on N > 1 GPUs the cut-face exchange (ncclSend / ncclRecv) and the reductions (ncclAllReduce)
    # run inside the library on its own stream.  Weak scaling over strips of blocks.
    trace = None
    trace_small = None
    trace_large = None
    if not args.no_trace:
        from hybridsbp_b200 import dist_trace
        if dist is not None and ctx.world == 1:
            ctx.comm_init_torch(dist)
        names = {1: "batched Jacobi-PCG (K2b)", 2: "batched dense Cholesky (K2a)", 3: "batched banded Cholesky (K2c)",
                 4: "batched PCG with fast-diagonalisation preconditioner (K2d)"}

        def timed_trace_solve(nblocks, n_per_block):
            tnb = int(round(nblocks ** 0.5))
            assert tnb * tnb == nblocks, "trace solves need a square number of blocks per GPU"
            t0 = time.perf_counter()
            tm = {}
            pr = dist_trace.StripProblem(ctx, rank, world, tnb, tnb, n_per_block, p, condense=not args.no_condense,
                                         coarse_modes=args.trace_coarse_modes, timings=tm)
            ctx.sync()
            t_setup = time.perf_counter() - t0
            pr.solve(tol=1e-2, maxit=4)                                  # warm-up
            barrier()
            t0 = time.perf_counter()
            ctx.timer_start()
            st_t = pr.solve(tol=args.trace_tol, maxit=20000)
            ms_dev = ctx.timer_stop()
            t_solve = time.perf_counter() - t0
            # the CG loop alone (no right-hand side, no back-substitution): CUDA events inside the library around the iterations
            per_it = st_t["cg_loop_ms"] / max(1, st_t["outer_iterations"])
            if dist is not None:
                import torch
                tt = torch.tensor([t_solve, ms_dev, t_setup, per_it], dtype=torch.float64, device="cuda")
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                t_solve, ms_dev, t_setup, per_it = [float(v) for v in tt]
            if st_t["converged"] != 1 or not st_t["true_rel_residual"] <= 100 * args.trace_tol:
                raise RuntimeError("trace solve did not converge: %r" % (st_t,))
            info = pr.info
            comm_path = pr.tr.comm_path()
            out = {"seconds": t_solve, "device_ms": ms_dev, "setup_seconds": t_setup,
                   "setup_breakdown_seconds": {k: round(v, 3) for k, v in tm.items()},
                   "outer_iterations": st_t["outer_iterations"], "issued_iterations": st_t["issued_iterations"],
                   "cg_loop_ms": st_t["cg_loop_ms"], "ms_per_iteration": per_it,
                   "converged": st_t["converged"], "rel_residual": st_t["rel_residual"], "true_rel_residual": st_t["true_rel_residual"],
                   "failed_local_blocks": st_t["failed_local_blocks"], "tol": args.trace_tol,
                   "coarse_dofs": st_t["coarse_dofs"],
                   "config": "%d blocks x %dx%d points per GPU, p=%d, local solver: %s; %s; first level: %s; second level: %s" %
                             (info["blocks"], n_per_block + 1, n_per_block + 1, p, names[info["local_mode"]],
                              "statically condensed (dense S_e = F^T M^-1 F per block formed during setup)" if info["condensed"] else
                              "matrix-free Schur matvec (one batched local solve per CG iteration)",
                              "exact face blocks B_ff (explicit inverses)" if info["face_blocks"] else "D",
                              "%d Legendre modes per face" % info["coarse_modes"] if info["coarse_modes"] else "none"),
                   "lambda_points_per_gpu": info["lambda_points"], "cut_faces_per_gpu": info["cut_faces"],
                   "volume_points_per_gpu": info["volume_points"],
                   "communication": {0: "none (1 GPU)",
                                     1: "inside libhsbp: ncclSend/ncclRecv of cut-face contributions + 2 ncclAllReduce per iteration",
                                     2: "inside libhsbp, no NCCL call in the iteration loop: cut-face contributions and the partial sums of "
                                        "the CG scalars are written into the partners' device memory over NVLink (cudaIpc-mapped "
                                        "mailboxes, flags, rank-ordered sums) by kernels of the same CUDA graph"}[comm_path]}
            pr.close()
            return out

        # (i) BASELINE config 4 / 5 in full: 1024 blocks x 256x256 points per GPU (507 904 lambda points per GPU)
        if args.trace_full:
            trace = timed_trace_solve(args.blocks, args.n)
        # (ii) small-block variant of SURVEY.md section 8d: 1024 blocks x 18x18 points per GPU, dense Cholesky factors
        trace_small = timed_trace_solve(args.trace_blocks, args.trace_n)
        # (iii) optional: a bounded number of blocks of config 4's size
        if args.trace_large_blocks > 0:
            trace_large = timed_trace_solve(args.trace_large_blocks, args.n)

    if rank == 0:
        peak, peak_src = load_peaks()
        achieved = BYTES_PER_DOF * dof / (stage[0] * 1e-3) / 1e9
        traffic = None                                 # DRAM bytes per k_sweep launch from the committed ncu capture
        try:
            with open(os.path.join(ROOT, "profiles", "k_sweep_traffic.json")) as f:
                tj = json.load(f)
            if variant == 1 and (args.blocks, args.n, args.p) == (1024, 255, 4):
                traffic = tj["dram_bytes_per_launch"]
        except Exception:
            pass
        line = {"metric": "fp64 SBP operator-apply GDOF/s", "value": world * dof / (ms_step * 1e-3) / 1e9,
                "unit": "GDOF/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": traffic,
                             "kernel": ("k_sweep (line-marching kernel: volume operator, all closures and the folded face terms in one pass)" if variant == 1 else
                                        "k_cross_pre + k_vol_apply (two-pass generic volume stage)"),
                             "algorithmic_bytes_per_launch": BYTES_PER_DOF * dof,
                             "kernel_ms": float(stage[0]),
                             "other_kernels_ms": ({"k_edge_prep": float(stage[1])} if (variant == 1 and not args.no_fold) else
                                                  {"k_face_gather": float(stage[1]), "k_face_scatter": float(stage[2])}),
                             "peak_source": peak_src,
                             "other_cases": [dict(c, frac_k_sweep=c["k_sweep_algorithmic_gbs"] / peak,
                                                  frac_whole_apply=c["whole_apply_algorithmic_gbs"] / peak) for c in other_cases],
                             "frac_whole_apply": BYTES_PER_DOF * dof / (ms_step * 1e-3) / 1e9 / peak},
                "e2e": {"value": world * dof / e2e_s / 1e9, "unit": "GDOF/s",
                        "h2d_bytes_per_step": 8 * dof, "d2h_bytes_per_step": 8 * dof,
                        "ms_per_step": e2e_s * 1e3, "steps": e2e_steps, "numa_node_rank0": numa_node,
                        "note": "hsbp_apply_host on pinned host buffers: H2D of u, apply, D2H of y inside each call"},
                "gpu_launches": args.steps * (2 if variant == 1 else 4),      # kernels of the timed region (k_edge_prep + k_sweep per apply)
                "clocks": clocks, "sustained": sustained, "apply_variant": variant, "checksum_abs_y": checksum, "trace_solve": trace,
                "trace_solve_small_blocks": trace_small, "trace_solve_large_blocks": trace_large}
        if world == 1 and not args.no_cpu:
            _, base = cpu_baseline(p, N, args.cpu_blocks, args.cpu_seconds, 1)
            line["cpu_baseline"] = base
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--blocks", type=int, default=1024)
    ap.add_argument("--n", type=int, default=255, help="N per block (points = N+1)")
    ap.add_argument("--p", type=int, default=4)
    ap.add_argument("--cpu-blocks", type=int, default=8)
    ap.add_argument("--cpu-seconds", type=float, default=5.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--sustained-steps", type=int, default=2000,
                    help="extra, separately reported run of this many applies (rank 0) to show the rate under the power limit; 0 = skip")
    ap.add_argument("--no-other-orders", action="store_true", help="skip the p = 2 / 6 and odd-line-length operator-apply entries")
    ap.add_argument("--no-trace", action="store_true", help="skip the trace-CG solve-time measurement")
    ap.add_argument("--trace-blocks", type=int, default=1024, help="blocks per GPU of the trace solve (a square number)")
    ap.add_argument("--trace-n", type=int, default=17, help="N per block of the trace solve")
    ap.add_argument("--trace-tol", type=float, default=1e-10)
    ap.add_argument("--trace-coarse-modes", type=int, default=2,
                    help="trace solves: Legendre modes per face of the additive coarse space (0 = off; 2 makes the CG iteration "
                         "count independent of the number of blocks)")
    ap.add_argument("--no-trace-full", dest="trace_full", action="store_false",
                    help="skip the full config-4 trace solve (1024 blocks x 256x256 points per GPU; about two minutes of setup)")
    ap.add_argument("--no-condense", action="store_true", help="trace solves: matrix-free Schur matvec instead of static condensation")
    ap.add_argument("--trace-large-blocks", type=int, default=0,
                    help="blocks per GPU (a square number) of the trace solve at the operator-apply block size; 0 = skip")
    ap.add_argument("--sweep-r", type=int, default=0, help="points per thread of k_sweep (0 = heuristic)")
    ap.add_argument("--sweep-deep", type=int, default=1, help="1: css / crs windows of k_sweep in shared-memory rings, 0: in registers")
    ap.add_argument("--sweep-p6-regs", type=int, default=168)
    ap.add_argument("--sweep-ncs", type=int, default=0, help="chunks per side of k_sweep (0 = heuristic)")
    ap.add_argument("--generic", action="store_true", help="force the generic two-pass kernels")
    ap.add_argument("--sweep-no-pdl", action="store_true", help="k_sweep without programmatic dependent launch behind k_edge_prep")
    ap.add_argument("--no-fold", action="store_true", help="face terms by separate gather / scatter kernels")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
