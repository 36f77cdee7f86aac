"""Readers for the on-disk inputs either side of the hot path, no R needed.

  * tab/space-separated genotype text with a header row and a sample-id column, the format the reference's scripts
    feed to read.table / read.delim (paper_materials/Real Data Analysis/Full_Test/dataprep.R:3-4,
    paper_materials/Timing Tests/test_time_Gaus.R:13-19) -> int8 matrix (genotype codes) ready for the device;
  * two-column phenotype text (sample id, value), pheno_left.txt;
  * .rda / .RDS (R's XDR serialisation; the package's bundled BASIS.rda, y.rda, ...), via pareben_b200.rdata;
  * .npy / .npz for convenience.
`load_problem` applies the row selection of the published workflow (rows of the genotype table whose sample id occurs
in the phenotype table, in phenotype order: SL_filter.R:14 `genotype[pheno[,1],]`)."""
from __future__ import annotations

import io as _io
import os
import zipfile

import numpy as np

from . import rdata


def _open_text(path: str, member: str | None = None):
    if path.endswith(".zip"):
        z = zipfile.ZipFile(path)
        name = member or z.namelist()[0]
        return z.open(name)
    return open(path, "rb")


def read_genotype_text(path: str, member: str | None = None, dtype=np.int8):
    """-> (sample_ids [n], marker_names [k], matrix [n, k]).  First row = header (first field names the id column),
    first field of every other row = sample id, the rest numeric codes.  Values are parsed as numbers and stored as
    `dtype` (int8 for genotype codes; pass np.float64 for real-valued designs); a non-integer value with an integer
    dtype raises."""
    with _open_text(path, member) as f:
        header = f.readline().decode().rstrip("\r\n")
        sep = "\t" if "\t" in header else None
        names = header.split(sep)[1:]
        ids, rows = [], []
        for raw in f:
            line = raw.decode().rstrip("\r\n")
            if not line:
                continue
            first, _, rest = line.partition("\t") if sep else line.partition(" ")
            v = np.fromstring(rest, dtype=np.float64, sep="\t" if sep else " ")
            if v.size != len(names):
                raise ValueError(f"{path}: row '{first}' has {v.size} values, header has {len(names)}")
            if np.issubdtype(dtype, np.integer):
                w = v.astype(dtype)
                if not np.array_equal(w, v):
                    raise ValueError(f"{path}: row '{first}' holds values that are not {np.dtype(dtype).name} codes")
                v = w
            ids.append(first.strip('"'))
            rows.append(v)
    return np.array(ids), np.array([n.strip('"') for n in names]), np.vstack(rows) if rows else np.zeros((0, len(names)), dtype)


def read_phenotype_text(path: str):
    """-> (sample_ids [n], values [n]) from a headerless two-column table; a single-column file gives ids = None."""
    ids, vals = [], []
    with _open_text(path) as f:
        for raw in f:
            parts = raw.decode().split()
            if not parts:
                continue
            if len(parts) == 1:
                vals.append(float(parts[0]))
            else:
                ids.append(parts[0].strip('"')); vals.append(float(parts[1]))
    return (np.array(ids) if ids else None), np.array(vals)


def read_matrix(path: str, name: str | None = None):
    """A numeric matrix or vector from .rda / .RDS / .npy / .npz / delimited text.  For containers with several objects
    `name` picks one (default: the only one, or the one named like the file)."""
    ext = os.path.splitext(path)[1].lower()
    if ext == ".rda" or ext == ".rdata":
        objs = rdata.read_rda(path)
        key = name or (next(iter(objs)) if len(objs) == 1 else os.path.splitext(os.path.basename(path))[0])
        return np.asarray(objs[key], dtype=np.float64)
    if ext == ".rds":
        return np.asarray(rdata.read_rds(path), dtype=np.float64)
    if ext == ".npy":
        return np.load(path)
    if ext == ".npz":
        z = np.load(path)
        return z[name or z.files[0]]
    ids, _names, M = read_genotype_text(path, dtype=np.float64)
    return M


def load_problem(basis_path: str, target_path: str, basis_name: str | None = None, target_name: str | None = None):
    """(BASIS [n, k], Target [n]) from a pair of files.  Genotype text + two-column phenotype text are joined on the
    sample id (rows in phenotype order, SL_filter.R:14); everything else is taken row for row."""
    text_like = lambda p: os.path.splitext(p)[1].lower() in (".txt", ".tsv", ".zip", ".tab", "")  # noqa: E731
    if text_like(basis_path) and text_like(target_path):
        pid, y = read_phenotype_text(target_path)
        ids, _names, X = read_genotype_text(basis_path)
        if pid is not None:
            pos = {s: i for i, s in enumerate(ids)}
            missing = [s for s in pid if s not in pos]
            if missing:
                raise ValueError(f"{len(missing)} phenotype ids are absent from the genotype table (first: {missing[0]})")
            X = X[[pos[s] for s in pid]]
        return X, y
    X = read_matrix(basis_path, basis_name)
    y = read_matrix(target_path, target_name) if not text_like(target_path) else read_phenotype_text(target_path)[1]
    return X, np.asarray(y, dtype=np.float64).ravel()
