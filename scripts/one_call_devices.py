"""GPU box: ONE pareben_cv_grid call driving n_devices GPUs (the in-library fan-out an R session uses), bundled Gaussian
1000 x 481, 10 folds, 4,000 fits: wall time through the host-buffer C-ABI for 1 .. n_devices GPUs and bitwise equality of
the tables.  usage: python scripts/one_call_devices.py N"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pareben_b200 as pb
nmax = int(sys.argv[1]) if len(sys.argv) > 1 else pb.device_count()
g = np.load("tests/golden/inputs_bundled.npz")
X, y = g["BASIS"].astype(float), g["y"].astype(float)
folds = pb.AssignToFolds(X, 10)
grid = pb.BuildGrid(X, y, 10)
ref = None
for nd in [n for n in (1, 2, 4, 8) if n <= nmax]:
    pb.cv_grid(X, y, folds, 10, grid["alpha"], grid["lambda"], n_devices=nd)          # warm-up: contexts, pools
    t = time.perf_counter(); err, st, ns = pb.cv_grid(X, y, folds, 10, grid["alpha"], grid["lambda"], n_devices=nd); dt = time.perf_counter() - t
    same = True if ref is None else bool(np.array_equal(err, ref))
    ref = err if ref is None else ref
    print(f"one call, {nd} GPU(s): 4000 Gaussian fits in {dt * 1e3:.1f} ms ({4000 / dt:.0f} fits/s), table identical to 1 GPU: {same}, status!=0: {(st != 0).sum()}")
