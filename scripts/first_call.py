import sys, time, numpy as np
sys.path.insert(0, '.')
import pareben_b200 as pb
pb.load()
g = np.load("tests/golden/inputs_bundled.npz")
X, y = g["BASISbinomial"].astype(float), g["yBinomial"].astype(float)
folds = pb.AssignToFolds(X, 5)
t = time.time(); grid = pb.BuildGrid(X, y, 5); t_grid = time.time() - t
for rep in range(3):
    t = time.time(); err, st, ns = pb.cv_grid(X, y, folds, 5, grid["alpha"], grid["lambda"], prior="binomial"); dt = time.time() - t
    print(f"call {rep}: cv_grid wall {dt*1e3:.0f} ms (BuildGrid first: {t_grid*1e3:.0f} ms)")
