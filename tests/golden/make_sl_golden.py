"""Generates tests/golden/sl_filter_zeo.npz: the single-locus prefilter pinned to an artifact the reference holds.

Run HERE (the build container), where /root/reference exists:   python tests/golden/make_sl_golden.py
The authors' own output of SL_filter.R's main-effect step on the yeast data is in the checkout:
  paper_materials/Real Data Analysis/Full_Test/filter_matrix_main0.05_Zeo.zip   (3844 strains x 11,396 kept markers)
produced from  genotype_full.txt (Timing Tests/genotype_full.zip, 4390 x 28,220)  and  pheno_Zeo (3844 values, strain
order of the filtered matrix) with the threshold 0.05 in  `which(abs(t(pheno_stand) %*% geno_stand / n) > tau)`
(SL_filter.R:21).  This script re-derives the kept set three ways and requires them to agree:
  (1) the numpy restatement oracle/rlayer.py:sl_filter on the same inputs;
  (2) the columns of the reference-held matrix, matched to genotype columns by content and order;
and stores what the GPU test needs: the 3844 x 28,220 genotype rows (bit-packed), the phenotype, the kept indices.
"""
import os
import sys
import zipfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import rlayer as R  # noqa: E402
from pareben_b200 import io as pio  # noqa: E402

T = "/root/reference/paper_materials/Timing Tests/"
F = "/root/reference/paper_materials/Real Data Analysis/Full_Test/"


def main():
    ids, _names, G = pio.read_genotype_text(T + "genotype_full.zip")
    z = zipfile.ZipFile(F + "filter_matrix_main0.05_Zeo.zip")
    rows = z.read("filter_matrix_main0.05_Zeo.05").decode().strip().split("\n")
    fid = [r.split("\t", 1)[0] for r in rows]
    held = np.array([np.fromstring(r.split("\t", 1)[1], dtype=np.int8, sep="\t") for r in rows])
    y = np.array([float(v) for v in open(F + "pheno_Zeo").read().split()])
    assert held.shape == (3844, 11396) and y.size == 3844
    pos = {s: i for i, s in enumerate(ids)}
    sel = np.array([pos[s] for s in fid])
    X = G[sel]
    main, _pairs, stat, _ = R.sl_filter(X.astype(np.float64), y, 0.05, 0.0, False)
    kept = main - 1
    assert kept.size == held.shape[1], (kept.size, held.shape)
    assert np.array_equal(X[:, kept], held), "restatement's kept columns differ from the reference-held matrix"
    np.savez_compressed(HERE + "/sl_filter_zeo.npz", bits=np.packbits(X == 1, axis=1), n=X.shape[0], k=X.shape[1], y=y,
                        tau_main=0.05, kept=kept.astype(np.int32), stat=stat)
    print("kept", kept.size, "of", X.shape[1], "== the reference-held matrix; file bytes", os.path.getsize(HERE + "/sl_filter_zeo.npz"))


if __name__ == "__main__":
    main()
