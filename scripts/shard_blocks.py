"""Experiment: one 1/8 shard of a grid (what a GPU holds in the 8-GPU strong-scaling run) with two resident blocks per SM
(default) against one (PAREBEN_BLOCKS_PER_SM=1).  usage: python scripts/shard_blocks.py [binomial|gaussian] [n_shards]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pareben_b200 as pb

g = np.load("tests/golden/inputs_bundled.npz")
prior = sys.argv[1] if len(sys.argv) > 1 else "binomial"
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
if prior == "binomial":
    X, y, nf = g["BASISbinomial"].astype(float), g["yBinomial"].astype(float), 5
else:
    X, y, nf = g["BASIS"].astype(float), g["y"], 10
folds = pb.AssignToFolds(X, nf)
grid = pb.BuildGrid(X, y, nf)
mine = pb.shard_plan(grid["lambda"], nf, 0, world)
f = (mine % nf + 1).astype(np.int32); a = grid["alpha"][mine // nf]; l = grid["lambda"][mine // nf]
for bps in ("2", "1", "auto"):          # auto = the library's own choice (one block per SM for small binomial launches)
    if bps == "auto": os.environ.pop("PAREBEN_BLOCKS_PER_SM", None)
    else: os.environ["PAREBEN_BLOCKS_PER_SM"] = bps
    with pb.Problem(X, y, folds, nf, False, prior) as p:
        for rep in range(3):
            err, st, ns, it = p.run_fits(f, a, l)
            fl, ms, _ = p.counters()
        print(f"{prior}: shard 0 of {world} = {mine.size} fits, blocks per SM <= {bps}: kernel {ms:.1f} ms")
    pb.release_cache() if hasattr(pb, "release_cache") else None
