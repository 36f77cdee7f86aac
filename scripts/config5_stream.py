"""GPU box: BASELINE config 5's shape through the streaming kernels.
usage: python scripts/config5_stream.py K grid_step n_folds [N] [main|epis grid]
Prints rounds, scan time, achieved FP64 TFLOP/s of the score contraction, fits/s."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pareben_b200 as pb
from pareben_b200.synth import config5

K = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
step = int(sys.argv[2]) if len(sys.argv) > 2 else 10
nf = int(sys.argv[3]) if len(sys.argv) > 3 else 10
N = int(sys.argv[4]) if len(sys.argv) > 4 else 1000
which = sys.argv[5] if len(sys.argv) > 5 else "epis"
t = time.time(); d = config5(N, K); t_gen = time.time() - t
X = d["X"].astype(np.float64); y = d["y"]
folds = pb.AssignToFolds(X, nf)
t = time.time(); grid = pb.BuildGrid(X, y, nf, "yes" if which == "epis" else "no"); t_grid = time.time() - t
rows = np.arange(0, 400, step)
fold = np.tile(np.arange(1, nf + 1), rows.size); a = np.repeat(grid["alpha"][rows], nf); l = np.repeat(grid["lambda"][rows], nf)
pb.set_mode(pb.MODE_STREAMING)
t = time.time()
with pb.Problem(X, y, folds, nf, True, "gaussian") as p:
    t_create = time.time() - t
    t = time.time(); err, st, ns, it = p.run_fits(fold, a, l); dt = time.time() - t
    fl, ms, launches = p.counters()
    scan_ms, scan_fl, scans, rounds = p.stream_counters()
print(f"config5-shape N={N} K={K} Kc={K*(K+1)//2} folds={nf} grid={which} lambda_max*10={grid['lambda'][0]:.4g}: {fold.size} fits; "
      f"synth {t_gen:.1f}s, lambda_max+grid {t_grid:.2f}s, create {t_create:.2f}s, run {dt:.2f}s ({fold.size/dt:.1f} fits/s), "
      f"kernels {ms:.0f} ms in {launches} launches, {rounds} rounds; scan {scan_ms:.0f} ms in {scans} launches = "
      f"{scan_fl/1e12:.1f} TFLOP -> {scan_fl/max(scan_ms,1e-9)/1e9:.2f} TFLOP/s; maxM {ns.max()}, status {np.unique(st)}, outer iters max {it.max()}")
