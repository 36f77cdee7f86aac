"""GPU box: run the golden-vector workloads through the C-ABI and write per-fit relative
differences vs the reference (tests/golden) to gpurun_out/parity_<name>.npz + a text summary."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pareben_b200 as pb

G = "tests/golden/"
inp = np.load(G + "inputs_bundled.npz")
cases = [
    ("config1_gaussian", inp["BASIS"][:50, :100], inp["y"][:50], 3, False, "gaussian", "config1_gaussian.npz"),
    ("bundled_gaussian_rows", inp["BASIS"], inp["y"], 3, False, "gaussian", "gauss_bundled_sample.npz"),
    ("gaussian_epis_slice", inp["BASIS"][:120, :25], inp["y"][:120], 3, True, "gaussian", "gauss_epis_slice.npz"),
    ("config2_binomial", inp["BASISbinomial"], inp["yBinomial"], 5, False, "binomial", "config2_binomial.npz"),
    ("binomial_epis_slice", inp["BASISbinomial"][::4, :20], inp["yBinomial"][::4], 3, True, "binomial", "binom_epis_slice.npz"),
]
os.makedirs("gpurun_out", exist_ok=True)
lines = []
for name, X, y, nf, epis, prior, gf in cases:
    g = np.load(G + gf)
    rows = g["rows"]
    err, st, ns = pb.cv_grid(X.astype(float), y.astype(float), g["fold_id"], nf, g["grid_alpha"][rows], g["grid_lambda"][rows],
                             epis=epis, prior=prior)
    rel = np.abs(err - g["fold_err"]) / np.maximum(np.abs(g["fold_err"]), 1e-300)
    np.savez_compressed(f"gpurun_out/parity_{name}.npz", err=err, status=st, n_selected=ns, rel=rel)
    q = np.quantile(rel, [0.5, 0.9, 0.99, 1.0])
    lines.append(f"| {name} | {err.size} | {int((ns != g['n_selected']).sum())} | {int((st != 0).sum())} | "
                 f"{q[0]:.1e} | {q[1]:.1e} | {q[2]:.1e} | {q[3]:.1e} | {int((rel > 1e-8).sum())} |")
# streaming organisation: the K = 2,000 slice of config 5 (2,001,000 candidates) against the reference's own C
from pareben_b200.synth import config5
g = np.load(G + "stream_k2000.npz")
d = config5(int(g["n"]), int(g["k"]), int(g["seed"]))
gi, fi = np.nonzero(np.isfinite(g["fold_err"]))
pb.set_mode(pb.MODE_STREAMING)
with pb.Problem(d["X"].astype(float), d["y"], g["fold_id"], int(g["n_folds"]), epis=True) as prob:
    err, st, ns, it = prob.run_fits(fi + 1, g["alpha"][gi], g["lam"][gi])
pb.set_mode(pb.MODE_AUTO)
want = g["fold_err"][gi, fi]
rel = np.abs(err - want) / np.abs(want)
q = np.quantile(rel, [0.5, 0.9, 0.99, 1.0])
lines.append(f"| stream_k2000_epis (streaming kernels, 2,001,000 candidates, active sets 1..{int(ns.max())}) | {err.size} | "
             f"{int((ns != g['n_selected'][gi, fi]).sum())} | {int((st != 0).sum())} | {q[0]:.1e} | {q[1]:.1e} | {q[2]:.1e} | {q[3]:.1e} | {int((rel > 1e-8).sum())} |")
hdr = ("| workload | fits | support-size mismatches | status != 0 | median rel | p90 | p99 | max | fits > 1e-8 |\n"
       "|---|---|---|---|---|---|---|---|---|")
open("gpurun_out/parity.md", "w").write(hdr + "\n" + "\n".join(lines) + "\n")
print(hdr); print("\n".join(lines))
