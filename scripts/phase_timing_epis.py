"""Experiments only: per-phase cycle totals for an Epis Gaussian grid (timing build of the library).
usage: PAREBEN_LIB=pareben_b200/libpareben_timing.so python scripts/phase_timing_epis.py [K]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pareben_b200 as pb
lib = pb.load()
lib.pareben_phase_cycles.restype = ctypes.c_int
lib.pareben_phase_cycles.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
names = ["vfill", "contract", "quad", "gram", "sweep", "irls", "delta_ml", "actions", "resid", "refresh"]
g = np.load("tests/golden/inputs_bundled.npz")
K = int(sys.argv[1]) if len(sys.argv) > 1 else 100
X, y, nf = g["BASIS"][:, :K].astype(float), g["y"], 10
folds = pb.AssignToFolds(X, nf); grid = pb.BuildGrid(X, y, nf, "no")
rows = np.arange(0, 400, 16)
fold = np.tile(np.arange(1, nf + 1), rows.size); a = np.repeat(grid["alpha"][rows], nf); l = np.repeat(grid["lambda"][rows], nf)
with pb.Problem(X, y, folds, nf, True, "gaussian") as p:
    buf = (ctypes.c_ulonglong * 32)()
    lib.pareben_phase_cycles(buf, 1)
    err, st, ns, it = p.run_fits(fold, a, l)
    fl, ms, _ = p.counters()
    n = lib.pareben_phase_cycles(buf, 1)
    cyc = np.array(buf[:n], dtype=float); calls = np.array(buf[n:2 * n], dtype=float)
    tot = ms * 1e-3 * 1.965e9 * min(296, fold.size)
    print(f"Epis K={K} Kc={K*(K+1)//2}: kernel {ms:.0f} ms, {fold.size} fits, maxM {ns.max()}, alg {fl/ms/1e9:.2f} TFLOP/s")
    for nm, c, k in zip(names, cyc, calls):
        print(f"  {nm:10s} {100*c/tot:5.1f} % of block time   calls {k:.0f}  cycles/call {c/max(k,1):.0f}")
