"""bench.py -- headline measurement of the shift-and-invert eigensolve path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload cfg2|cfg1|...]

One "step" = one pass of the hot path over one pencil: numeric LU of A - sigma M, Krylov-Schur for
the `nev` direct modes nearest sigma, and the adjoint (left) modes at conj(sigma) on the same
factors (BASELINE config 2: "2D backward-facing step Re=500, Taylor-Hood fine mesh (~1M DOFs),
nev=20 with adjoint modes").  Data: synthetic structured Taylor-Hood pencil of that shape
(lsa_fw_b200/pencils.py), seeded start vectors.

`value`  : device seconds per step with values resident in HBM (CUDA events inside the library).
`e2e`    : seconds per step through the reference-facing API `EigenSolver(A, M, cfg).solve()` with HOST
           CSR buffers (value upload, factor, direct + adjoint eigensolve, eigenvector download); the
           symbolic analysis is reused across steps exactly as across a Reynolds / shift sweep.
`roofline`: the triangular-solve sweep (HBM bound) -- the dominant device time of the step -- plus
           `roofline_lu` for the factorisation against the measured FP64 GEMM rate.
`cpu_baseline` / `--impl reference`: the reference path cannot run here (PETSc/SLEPc are not
           installable), so the CPU arm is the SciPy SuperLU + ARPACK port of Solver/eigen2.py
           (oracle/), timed on a bounded sample of the same pencil family and scaled to the workload.

N > 1: independent replicas (one pencil of a Reynolds sweep per GPU, no data-path collective);
scaling "weak".
"""

from __future__ import annotations

import argparse
import json
import math
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (builder, kwargs, sigma, nev, ncv, description)
    "cfg1": ("cylinder_wake_2d", dict(nx=110, ny=50, re=50.0), 0.05 + 0.74j, 10, 80,
             "2D cylinder-wake surrogate Re=50, Taylor-Hood 110x50 (50 303 DOFs)"),
    "cfg2": ("backward_step_2d", dict(nx=667, ny=167, re=500.0), -0.35 + 0.1j, 20, 80,
             "2D backward-facing-step surrogate Re=500, Taylor-Hood 667x167 (~1.0 M DOFs)"),
    "cfg2_small": ("backward_step_2d", dict(nx=167, ny=42, re=500.0), -0.35 + 0.1j, 20, 80,
                   "2D backward-facing-step surrogate Re=500, Taylor-Hood 167x42 (64 174 DOFs)"),
    "cav3d": ("cavity_3d", dict(n=20, re=100.0), 0.1 + 0.3j, 10, 80,
              "3D lid-driven-cavity surrogate, Taylor-Hood 20^3 x 6 tets (216 024 DOFs = cube.py size)"),
}
TOL = 1e-11
MAX_RESTARTS = 100
METRIC = "shift-invert eigensolve s (direct+adjoint modes, LU included)"


def build_pencil(name: str, rank: int = 0):
    from lsa_fw_b200 import pencils

    fn, kw, sigma, nev, ncv, desc = WORKLOADS[name]
    kw = dict(kw)
    if rank:
        kw["re"] = kw["re"] * (1.0 + 0.02 * rank)  # Reynolds sweep across replicas, same pattern
    return getattr(pencils, fn)(**kw), sigma, nev, ncv, desc


# --------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu: int) -> None:
        self.gpu, self.proc, self.path = gpu, None, None

    def start(self) -> None:
        if not shutil.which("nvidia-smi"):
            return
        fd, self.path = tempfile.mkstemp(suffix=".csv")
        os.close(fd)
        self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-lms", "200"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------- CPU arm
def cpu_sample(workload: str, budget_s: float):
    """Times the SciPy port (oracle) on a bounded member of the workload's pencil family and scales
    it to the workload.  Returns the cpu_baseline object."""
    from lsa_fw_b200 import pencils
    from oracle import eigen_oracle as O

    fn, kw, sigma, nev, ncv, desc = WORKLOADS[workload]
    full_n = None
    if fn == "backward_step_2d":
        table = [((84, 21), 6.0), ((118, 30), 14.0), ((167, 42), 40.0)]
        pick = table[0][0]
        for shape, cost in table:
            if cost <= budget_s:
                pick = shape
        pc = pencils.backward_step_2d(pick[0], pick[1], re=kw["re"])
        full_n = pencils.th_dofs((kw["nx"], kw["ny"]))
        sample_desc = f"same pencil family at {pick[0]}x{pick[1]} cells ({pc.n} DOFs)"
        dim = 2
    elif fn == "cylinder_wake_2d":
        pc = pencils.cylinder_wake_2d(**kw)
        full_n = pc.n
        sample_desc = f"the full workload ({pc.n} DOFs)"
        dim = 2
    else:
        nn = 8 if budget_s < 30 else 10
        pc = pencils.cavity_3d(nn, re=kw["re"])
        full_n = pencils.th_dofs((kw["n"],) * 3)
        sample_desc = f"same pencil family at {nn}^3 cells ({pc.n} DOFs)"
        dim = 3
    t0 = time.perf_counter()
    direct = O.shift_invert_arpack(pc.A, pc.M, sigma, nev, ncv=ncv, tol=TOL, maxiter=MAX_RESTARTS * ncv)
    t_direct = time.perf_counter() - t0
    t0 = time.perf_counter()
    # adjoint modes the way the reference obtains them: a second factorisation of (A^H, M^H)
    # (Sensitivity/__init__.py:246-262)
    AH, MH = pc.A.conj().T.tocsr(), pc.M.conj().T.tocsr()
    adj = O.shift_invert_arpack(AH, MH, np.conj(sigma), nev, ncv=ncv, tol=TOL, maxiter=MAX_RESTARTS * ncv)
    t_adj = time.perf_counter() - t0
    measured = t_direct + t_adj
    ratio = full_n / pc.n
    # scaling law of sparse LU with nested-dissection-like orderings: flops ~ n^1.5 (2-D) / n^2 (3-D),
    # factor entries (solve cost) ~ n log n (2-D) / n^(4/3) (3-D)
    f_fac = ratio ** (1.5 if dim == 2 else 2.0)
    f_sol = ratio * (math.log2(full_n) / math.log2(pc.n)) if dim == 2 else ratio ** (4.0 / 3.0)
    fac_s = direct.seconds["factor"] + adj.seconds["factor"]
    eig_s = direct.seconds["eigs"] + adj.seconds["eigs"]
    scaled = fac_s * f_fac + eig_s * f_sol
    # SuperLU and ARPACK are sequential codes; the dense kernels underneath (OpenBLAS) may use every host thread
    # they are given, and nothing here restricts them
    try:
        from threadpoolctl import threadpool_info

        blas_threads = max([int(i.get("num_threads", 1)) for i in threadpool_info()] or [1])
    except Exception:
        blas_threads = os.cpu_count() or 1
    return {
        "value": scaled, "unit": "s", "cores": blas_threads, "kind": "port",
        "sample": (f"SciPy 1.18 SuperLU (COLAMD) + ARPACK port of Solver/eigen2.py: sequential SuperLU/ARPACK on top of "
                   f"OpenBLAS with {blas_threads} threads (unrestricted), "
                   f"{os.cpu_count()} host cores present; timed on {sample_desc}: direct+adjoint = {measured:.2f} s "
                   f"(factor {fac_s:.2f} s, eigs {eig_s:.2f} s, {direct.n_op_applies + adj.n_op_applies} OP applies); "
                   f"scaled to the workload ({full_n} DOFs) with factor x{f_fac:.1f} (flops ~ n^{1.5 if dim == 2 else 2.0}) "
                   f"and solves x{f_sol:.1f}"),
        "measured_sample_seconds": measured, "sample_dofs": pc.n, "extrapolated": ratio != 1.0,
        "sample_eig0": [float(direct.eigenvalues[0].real), float(direct.eigenvalues[0].imag)],
    }


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step_budget = max(5.0, 150.0 / max(1, args.steps + args.warmup))
    vals = []
    cb = None
    for i in range(args.warmup + args.steps):
        cb = cpu_sample(args.workload, per_step_budget)
        if i >= args.warmup:
            vals.append(cb["value"])
    v = float(np.mean(vals))
    cb["value"] = v
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": v * 1e3, "higher_is_better": False, "scaling": "weak",
        "vs_baseline": None, "dtype": "c128", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload][5], "nev": WORKLOADS[args.workload][3],
                   "ncv": WORKLOADS[args.workload][4], "tol": TOL, "sigma": str(WORKLOADS[args.workload][2])},
        "cpu_baseline": cb,
        "e2e": {"value": v, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------- GPU arm
def measure_fp64_peak(device) -> float:
    import torch

    n = 6144
    a = torch.randn(n, n, dtype=torch.float64, device=device)
    b = torch.randn(n, n, dtype=torch.float64, device=device)
    a @ b
    torch.cuda.synchronize(device)
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        a @ b
        e1.record()
        torch.cuda.synchronize(device)
        best = min(best, e0.elapsed_time(e1))
    del a, b
    torch.cuda.empty_cache()
    return 2 * n**3 / (best * 1e-3) / 1e12


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default=os.environ.get("LSA_BENCH_WORKLOAD", "cfg2"), choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    import lsa_fw_b200 as L
    from lsa_fw_b200 import _lib

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"

    t0 = time.perf_counter()
    pc, sigma, nev, ncv, desc = build_pencil(args.workload, rank)
    t_assemble = time.perf_counter() - t0
    n = pc.n
    fp64_peak = measure_fp64_peak(device)

    # ------------- device-resident arm: C ABI directly, values already in HBM
    h = _lib.Handle(n, local_rank)
    flag = ((pc.A.diagonal() - sigma * pc.M.diagonal()) == 0).astype(np.uint8)
    t0 = time.perf_counter()
    info = h.analyze(pc.A.indptr, pc.A.indices, pc.M.indptr, pc.M.indices, leaf_size=64, order_last=flag)
    t_symbolic = time.perf_counter() - t0
    h.set_values(pc.A.data, pc.M.data)
    v0 = np.random.default_rng(1234 + rank).standard_normal(n).astype(np.complex128)

    def device_step():
        fs = h.factor(1.0, -sigma, _lib.LSA_C128, 1e-13)
        rd = h.eigs(nev=nev, ncv=ncv, tol=TOL, max_restarts=MAX_RESTARTS, which="TARGET_MAGNITUDE",
                    transform=_lib.LSA_ST_SINVERT, sigma=sigma, v0=v0)
        nconv_d = rd.nconv
        lam_d = h.eigenvalues(min(nev, nconv_d))
        ra = h.eigs(nev=nev, ncv=ncv, tol=TOL, max_restarts=MAX_RESTARTS, which="TARGET_MAGNITUDE",
                    transform=_lib.LSA_ST_SINVERT, sigma=sigma, adjoint=True, v0=v0)
        return fs, rd, ra, lam_d

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    for _ in range(args.warmup):
        device_step()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    acc = dict(factor=0.0, eigs=0.0, solve=0.0, spmv=0.0, ortho=0.0, rr=0.0, restart=0.0, applies=0, restarts=0,
               kernels=0, flops=0.0, reorth=0)
    for _ in range(args.steps):
        fs, rd, ra, lam_d = device_step()
        acc["factor"] += fs.seconds
        acc["flops"] += fs.flops
        acc["kernels"] += fs.n_kernels + rd.n_kernels + ra.n_kernels
        for r in (rd, ra):
            acc["eigs"] += r.seconds
            acc["solve"] += r.seconds_solve
            acc["spmv"] += r.seconds_spmv
            acc["ortho"] += r.seconds_ortho
            acc["rr"] += r.seconds_rr
            acc["restart"] += r.seconds_restart
            acc["applies"] += r.n_op_applies
            acc["restarts"] += r.n_restarts
            acc["reorth"] += r.n_reorth
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    dev_s = (acc["factor"] + acc["eigs"]) / args.steps
    # parity gate (outside the timed region): residuals of the direct pairs, then of the adjoint pairs
    rd = h.eigs(nev=nev, ncv=ncv, tol=TOL, max_restarts=MAX_RESTARTS, which="TARGET_MAGNITUDE",
                transform=_lib.LSA_ST_SINVERT, sigma=sigma, v0=v0)
    resid_direct = float(h.residuals(min(nev, rd.nconv)).max()) if rd.nconv else None
    lam_direct = h.eigenvalues(min(nev, rd.nconv))
    ra = h.eigs(nev=nev, ncv=ncv, tol=TOL, max_restarts=MAX_RESTARTS, which="TARGET_MAGNITUDE",
                transform=_lib.LSA_ST_SINVERT, sigma=sigma, adjoint=True, v0=v0)
    resid_adj = float(h.residuals(min(nev, ra.nconv)).max()) if ra.nconv else None
    lam_adj = h.eigenvalues(min(nev, ra.nconv))
    conj_mismatch = float(max(min(abs(np.conj(l) - lam_direct)) / abs(l) for l in lam_adj)) if len(lam_adj) and len(lam_direct) else None
    conj_mismatch5 = float(max(min(abs(np.conj(l) - lam_direct)) / abs(l) for l in lam_adj[:5])) if len(lam_adj) and len(lam_direct) else None
    # accuracy of the triangular solves themselves (both sweeps) at full size
    Csh = (pc.A - sigma * pc.M).tocsr()
    bb = np.random.default_rng(99).standard_normal(n) + 1j * np.random.default_rng(98).standard_normal(n)
    xs = h.solve(bb)
    solve_resid_n = float(np.linalg.norm(Csh @ xs - bb) / np.linalg.norm(bb))
    xs = h.solve(bb, _lib.LSA_OP_H)
    solve_resid_h = float(np.linalg.norm(Csh.conj().T @ xs - bb) / np.linalg.norm(bb))
    del Csh, xs, bb
    counters = h.counters()
    solve_mean = acc["solve"] / max(1, acc["applies"])
    spmv_mean = acc["spmv"] / max(1, acc["applies"])
    h.close()

    # ------------- end-to-end arm: reference-facing API, host buffers
    A_c, M_c = L.iPETScMatrix(pc.A), L.iPETScMatrix(pc.M)
    cfg = L.EigensolverConfig(num_eig=nev, atol=TOL, max_it=MAX_RESTARTS, ncv=ncv)

    # the adjoint solve of the API re-factors unless it goes through the `.H` carriers; use them
    AH_c, MH_c = A_c.H, M_c.H

    def e2e_step2():
        es = L.EigenSolver(A_c, M_c, cfg, check_hermitian=False)
        es.solver.set_st_type(L.iSTType.SINVERT)
        es.solver.set_target(sigma)
        es.solver.set_st_pc_type(L.PreconditionerType.LU)
        es.solver.set_backend_options(device=local_rank, v0=v0)
        pairs = es.solve()
        ea = L.EigenSolver(AH_c, MH_c, cfg, check_hermitian=False)      # Sensitivity/__init__.py:246-262
        ea.solver.set_st_type(L.iSTType.SINVERT)
        ea.solver.set_st_pc_type(L.PreconditionerType.LU)
        ea.solver.set_target(np.conj(sigma))
        ea.solver.set_backend_options(device=local_rank, v0=v0)
        pairs_adj = ea.solve()
        return es, pairs, ea, pairs_adj

    for _ in range(max(1, min(args.warmup, 2))):
        es, pairs, ea, pairs_adj = e2e_step2()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        es, pairs, ea, pairs_adj = e2e_step2()
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.steps
    st = es.solver.stats
    h2d = (pc.A.nnz + pc.M.nnz) * 8
    d2h = (len(pairs) + len(pairs_adj)) * n * 16
    lam0 = complex(pairs[0][0]) if pairs else None
    lam0_adj = complex(pairs_adj[0][0]) if pairs_adj else None

    # ------------- reduce over ranks (max time), gather per-rank sanity
    if world > 1:
        t = torch.tensor([dev_s, e2e_s, wall], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_s, e2e_s, wall = (float(x) for x in t.tolist())

    cb = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = cpu_sample(args.workload, 40.0)

    if rank == 0:
        traffic, traffic_src = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload)
            if tj:
                traffic = tj["dram_read_bytes"] + tj["dram_write_bytes"]
                traffic_src = tj["source"]
        except Exception:
            pass
        bytes_solve = counters.bytes_solve
        achieved = bytes_solve / solve_mean / 1e9 if solve_mean > 0 else 0.0
        lu_tflops = acc["flops"] / acc["factor"] / 1e12 if acc["factor"] > 0 else 0.0
        line = {
            "metric": METRIC, "value": dev_s, "unit": "s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_s * 1e3, "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
            "dtype": "c128", "data": "synthetic",
            "config": {"workload": desc, "n_dofs": n, "nnz_A": int(pc.A.nnz), "nnz_M": int(pc.M.nnz), "nev": nev,
                       "ncv": ncv, "tol": TOL, "sigma": str(sigma), "adjoint_modes": True,
                       "parallelism": "1 process per GPU, independent replicas (Reynolds sweep)" if world > 1 else "single GPU",
                       "l2": "inputs_larger_than_L2" if info.nnz_lu * 16 > 126e6 else "factors fit in L2",
                       "symbolic": "host, reused across steps", "ordering": "graph nested dissection (no coordinates)"},
            "e2e": {"value": e2e_s, "unit": "s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(acc["kernels"]),
            "clocks": clocks,
            "roofline": {"kernel": "supernodal triangular-solve sweep (k_front_stream + k_sweep_cluster + k_up_gather + k_down_off, fwd+bwd)", "bound": "hbm",
                         "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bytes_solve, "mean_launch_seconds": solve_mean,
                         "share_of_step": acc["solve"] / max(1e-30, acc["factor"] + acc["eigs"])},
            "roofline_lu": {"kernel": "k_front_gemm (FP64 DMMA) + panel kernels: whole numeric LU", "bound": "tensor",
                            "achieved": lu_tflops, "peak": fp64_peak, "unit": "TFLOP/s", "frac": lu_tflops / fp64_peak,
                            "peak_source": "cuBLAS DGEMM 6144^3 measured in this run (FP64 DMMA pipe)",
                            "seconds": acc["factor"] / args.steps, "share_of_step": acc["factor"] / max(1e-30, acc["factor"] + acc["eigs"])},
            "roofline_spmv": {"bound": "hbm", "achieved": counters.bytes_spmv_m / spmv_mean / 1e9 if spmv_mean > 0 else 0.0,
                              "peak": hbm_peak, "unit": "GB/s"},
            "phases_s_per_step": {k: acc[k] / args.steps for k in ("factor", "eigs", "solve", "spmv", "ortho", "rr", "restart")},
            "op_applies_per_step": acc["applies"] / args.steps, "restarts_per_step": acc["restarts"] / args.steps,
            "reorthogonalised_columns_per_step": acc["reorth"] / args.steps,
            "symbolic": {"seconds": t_symbolic, "phases": list(info.seconds), "fronts": info.n_fronts, "levels": info.n_levels,
                         "nnz_lu": int(info.nnz_lu), "flops_real": info.flops_real, "max_front": info.max_front,
                         "decoupled": info.n_decoupled},
            "e2e_phases": {k: st.get(k) for k in ("symbolic_seconds", "upload_seconds", "factor_seconds", "eigs_seconds", "fetch_seconds", "total_seconds")},
            "parity": {"resid_direct_max": resid_direct, "resid_adjoint_max": resid_adj, "nconv_direct": len(pairs),
                       "nconv_adjoint": len(pairs_adj), "lambda0": [lam0.real, lam0.imag] if lam0 else None,
                       "lambda0_adjoint": [lam0_adj.real, lam0_adj.imag] if lam0_adj else None,
                       "n_perturbed": int(st.get("n_perturbed", -1)),
                       "adjoint_vs_conj_direct_rel": conj_mismatch, "adjoint_vs_conj_direct_rel_leading5": conj_mismatch5,
                       "conditioning_note": ("channel-type pencils (config 2) are strongly non-normal: the SciPy oracle's own "
                                             "direct and adjoint runs agree only to ~1e-4 on the 10 leading eigenvalues there; "
                                             "eigenvalue parity to 1e-8 is asserted on the well-conditioned pencils in tests/"),
                       "solve_resid_N": solve_resid_n,
                       "solve_resid_H": solve_resid_h, "max_multiplier": fs.max_multiplier},
            "wall_s_timed_region": wall, "assemble_s": t_assemble, "fp64_peak_tflops_measured": fp64_peak,
        }
        if cb is not None:
            line["cpu_baseline"] = cb
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
