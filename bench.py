#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the EBEN cross-validation hot path.

Workload (BASELINE.json configs[1]): CrossValidate, prior="binomial", on the reference's bundled
BASISbinomial / yBinomial (500 x 481, tests/golden/inputs_bundled.npz), nFolds = 5, Epis = "no",
the full 20 x 20 (alpha, lambda) grid  ->  2,000 independent EBEN fits per step.
One "step" = one pass of the hot path over that grid: every fit from initialisation to its
hold-out log-likelihood.

  value   whole-job fits/s with the problem already resident in HBM (pareben_run_fits), timed
          with CUDA events on the library's launch stream, max over ranks.
  e2e     the same metric through the host-buffer C-ABI call a user's CrossValidate makes
          (pareben_cv_grid): H2D of BASIS/y/folds/grid from pinned host memory, per-fold layout
          kernels, the fit kernel, D2H of the error table -- all inside the timed region.
  N > 1   one process per GPU (torchrun); weak scaling: every rank runs its own full grid (the
          fits are independent, there is no data-path collective); rank 0 gathers the tiny
          per-rank error tables over NCCL inside the e2e region only.
  --impl reference   times the reference's own C (oracle/_ref, compiled unmodified from
          /root/reference/EBEN_orig/src) -- or the oracle's C restatement when that library is
          absent -- on all host cores over a bounded, stratified sample of the same grid.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")

import numpy as np  # noqa: E402

METRIC = "EBEN CV grid fits/sec"
UNIT = "fits/s"
WORKLOAD = "config2: CrossValidate binomial, bundled BASISbinomial 500x481 / yBinomial, nFolds=5, Epis=no, 20x20 grid = 2000 fits/step"


def load_workload():
    g = np.load(os.path.join(ROOT, "tests", "golden", "inputs_bundled.npz"))
    X = g["BASISbinomial"].astype(np.float64)
    y = g["yBinomial"].astype(np.float64)
    return X, y, 5


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference implementation on the host cores
def _cpu_task(args):
    from oracle import rlayer as R
    X, y, fid, f, lam, a, kind = args
    e, fit = R.fit_one(X, y, fid, f, lam, a, False, "binomial", R.fit_lib(kind))
    return e


def cpu_sample(X, y, n_folds, grid_alpha, grid_lambda, folds, every: int, procs: int):
    """Time a stratified sample of the grid (every `every`-th grid row, all folds: all 20 lambda
    ranks are visited) on `procs` processes, one grid row per task like %dopar%."""
    from multiprocessing import get_context
    from oracle import rlayer as R
    kind = R.available_kind()
    rows = list(range(0, grid_alpha.size, every))
    jobs = [(X, y, folds, f, grid_lambda[r], grid_alpha[r], kind) for r in rows for f in range(1, n_folds + 1)]
    ctx = get_context("fork")
    with ctx.Pool(procs) as pool:
        pool.map(_cpu_task, jobs[:procs], chunksize=1)          # start-up / page-in, untimed
        t0 = time.perf_counter()
        pool.map(_cpu_task, jobs, chunksize=1)
        dt = time.perf_counter() - t0
    return len(jobs), dt, kind, f"every {every}th grid row x {n_folds} folds = {len(jobs)} of {grid_alpha.size * n_folds} fits"


def oracle_grid(X, y, n_folds):
    from oracle import rlayer as R
    ga, gl = R.build_grid(X, y, n_folds)
    return ga, gl, R.assign_to_folds(X.shape[0], n_folds)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    X, y, n_folds = load_workload()
    ga, gl, folds = oracle_grid(X, y, n_folds)
    procs = os.cpu_count() or 1
    every = args.cpu_every
    for _ in range(max(args.warmup, 0) and 1):
        cpu_sample(X, y, n_folds, ga, gl, folds, every * 4, procs)
    total_fits, total_t, kind, sample = 0, 0.0, "port", ""
    for _ in range(args.steps):
        n, dt, kind, sample = cpu_sample(X, y, n_folds, ga, gl, folds, every, procs)
        total_fits += n; total_t += dt
    value = total_fits / total_t
    line = {"metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "bundled BASISbinomial/yBinomial (reference fixture)",
            "config": {"workload": WORKLOAD, "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "reference" if kind == "reference" else "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def run_gpu(args):
    import torch
    import torch.distributed as dist
    import pareben_b200 as pb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly one JSON line: anything libraries print while the run is in progress (NCCL writes its
    # version banner to stdout when the communicator comes up) is sent to stderr instead
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    pb.load()
    if pb.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device (the hot path has no CPU implementation; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    X, y, n_folds = load_workload()
    n, k = X.shape
    folds = pb.AssignToFolds(X, n_folds)
    grid = pb.BuildGrid(X, y, n_folds, device=local)
    n_grid = grid["alpha"].size
    n_fits = n_grid * n_folds
    fold_vec = np.tile(np.arange(1, n_folds + 1, dtype=np.int32), n_grid)
    a_vec = np.repeat(grid["alpha"], n_folds); l_vec = np.repeat(grid["lambda"], n_folds)

    # FP64 peak for the roofline denominator (MEASURED_PEAKS.json has HBM and bf16 only)
    peak_dfma = pb.measure_fp64_peak(local, 0)
    peak_dmma = pb.measure_fp64_peak(local, 1)

    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")       # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    prob = pb.Problem(X, y, folds, n_folds, False, "binomial", device=local)
    for _ in range(args.warmup):
        prob.run_fits(fold_vec, a_vec, l_vec)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    dev_ms, flops, wall0 = [], [], time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()                      # L2 flush between timed iterations
        torch.cuda.synchronize()
        err, st, ns, it = prob.run_fits(fold_vec, a_vec, l_vec)
        fl, ms, launches = prob.counters()
        dev_ms.append(ms); flops.append(fl)
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    my_ms = float(np.sum(dev_ms))
    if world > 1:
        t = torch.tensor([my_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        max_ms = float(t.item())
    else:
        max_ms = my_ms
    value = world * n_fits * args.steps / (max_ms * 1e-3)

    # ---- e2e: host buffers (pinned) -> pareben_cv_grid -> host table, every step ----
    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.from_numpy(np.empty(0, a.dtype)).dtype, pin_memory=True)
        out = t.numpy()
        out[...] = a
        return out, t
    Xp, _k1 = pinned(np.asfortranarray(X).T.copy())      # column-major bytes of X, held as a C-contiguous (k, n) array
    yp, _k2 = pinned(y); fp, _k3 = pinned(folds.astype(np.int32))
    Xcol = Xp.T                                          # (n, k) Fortran-ordered view onto the pinned block
    ap, _k4 = pinned(grid["alpha"]); lp, _k5 = pinned(grid["lambda"])
    def gather_tables(table):
        # the `.combine = rbind` of the reference: per-rank tables collected on rank 0 over NCCL
        tt = torch.from_numpy(table).cuda()
        bucket = [torch.empty_like(tt) for _ in range(world)] if rank == 0 else None
        dist.gather(tt, bucket, dst=0)
        torch.cuda.synchronize()

    for _ in range(min(args.warmup, 2)):
        w_err, _, _ = pb.cv_grid(Xcol, yp, fp, n_folds, ap, lp, prior="binomial", device=local)
        if world > 1:
            gather_tables(w_err)          # also brings the NCCL communicator up outside the timed region
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e_err, e_st, e_ns = pb.cv_grid(Xcol, yp, fp, n_folds, ap, lp, prior="binomial", device=local)
        if world > 1:
            gather_tables(e_err)
    barrier()
    e2e_t = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_t], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_t = float(t.item())
    e2e_value = world * n_fits * args.steps / e2e_t
    h2d = X.nbytes + y.nbytes + folds.size * 4 + 2 * n_grid * 8 + n_fits * 32
    d2h = n_fits * (8 + 12) + 8

    assert np.array_equal(e_err.ravel(), err), "e2e and resident paths disagree"

    line = None
    if rank == 0:
        avg_ms = float(np.mean(dev_ms)); avg_fl = float(np.mean(flops))
        achieved = avg_fl / (avg_ms * 1e-3) / 1e12
        peak = max(peak_dfma, peak_dmma)
        cpu = None
        if world == 1 and not args.no_cpu:
            from oracle import rlayer as R
            ga, gl, ofolds = oracle_grid(X, y, n_folds)
            procs = os.cpu_count() or 1
            nj, dt, kind, sample = cpu_sample(X, y, n_folds, ga, gl, ofolds, args.cpu_every, procs)
            cpu = {"value": nj / dt, "unit": UNIT, "cores": procs, "kind": "reference" if kind == "reference" else "port",
                   "sample": sample + f", {dt:.1f} s wall"}
        traffic = None
        tp = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": max_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "bundled BASISbinomial/yBinomial (reference fixture, tests/golden/inputs_bundled.npz)",
                "config": {"workload": WORKLOAD, "fits_per_step_per_gpu": n_fits, "l2_flush_between_steps": True,
                           "status_nonzero": int((st != 0).sum()), "max_active_set": int(ns.max())},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "ms_per_step": 1e3 * e2e_t / args.steps},
                "gpu_launches": int(args.steps),
                "roofline": {"bound": "tensor", "pipe": "fp64 (DMMA mma.sync.m8n8k4.f64 for the contraction, DFMA elsewhere)",
                             "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                             "traffic": traffic, "kernel": "eben_fit_kernel<binomial,main>",
                             "algorithmic_flops_per_launch": avg_fl, "launch_ms": avg_ms,
                             "peak_source": f"on-box probe at bench start: DFMA {peak_dfma:.1f}, DMMA {peak_dmma:.1f} TFLOP/s "
                                            "(MEASURED_PEAKS.json has no FP64 entry)"},
                "cpu_baseline": cpu, "wall_s_timed_region": wall}
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    prob.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-every", type=int, default=25, help="CPU arms time every n-th grid row (all folds)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
