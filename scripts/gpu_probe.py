"""Quick on-box probe: FP64 peaks, and timings of a few workloads through the C-ABI."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pareben_b200 as pb

print("devices", pb.device_count())
print("DFMA peak TF/s", pb.measure_fp64_peak(0, 0))
print("DMMA peak TF/s", pb.measure_fp64_peak(0, 1))
g = np.load("tests/golden/inputs_bundled.npz")
X, y = g["BASIS"].astype(float), g["y"]
for (n, k, nf, step) in ((50, 100, 3, 1), (1000, 481, 3, 23), (1000, 481, 10, 4)):
    Xs, ys = X[:n, :k], y[:n]
    folds = pb.AssignToFolds(Xs, nf)
    grid = pb.BuildGrid(Xs, ys, nf)
    rows = np.arange(0, 400, step)
    with pb.Problem(Xs, ys, folds, nf) as p:
        fold = np.tile(np.arange(1, nf + 1), rows.size)
        a = np.repeat(grid["alpha"][rows], nf); l = np.repeat(grid["lambda"][rows], nf)
        for rep in range(2):
            t = time.time(); err, st, ns, it = p.run_fits(fold, a, l); dt = time.time() - t
            fl, ms, _ = p.counters()
            print(f"N={n} K={k} folds={nf} fits={fold.size}: wall {dt*1e3:.1f} ms kernel {ms:.1f} ms  "
                  f"{fold.size/dt:.0f} fits/s  alg {fl/1e9:.2f} GFLOP -> {fl/ms/1e9:.3f} TFLOP/s  maxM {ns.max()} status {np.unique(st)}")
