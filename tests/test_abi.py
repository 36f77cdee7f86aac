"""CPU suite, part 2: the C-ABI library loads and exports every symbol include/pareben.h declares;
host-only entry points behave; compute entry points fail loudly without a device."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "pareben.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    # pareben_* plus EBEN's own four `.C` entry points, which the library exports under their original names
    return sorted(set(re.findall(r"\b(pareben_[a-z0-9_]+|elasticNetLinearNe(?:MainEff|EpisEff)|ElasticNetBinaryNE(?:mainEff|full))\s*\(", text)))


def test_header_symbols_exported(built):
    import pareben_b200 as pb
    names = _declared()
    assert "pareben_cv_grid" in names and "pareben_fit" in names and len(names) >= 10
    assert {"elasticNetLinearNeMainEff", "elasticNetLinearNeEpisEff", "ElasticNetBinaryNEmainEff", "ElasticNetBinaryNEfull"} <= set(names)
    lib = ctypes.CDLL(pb._lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pareben.h but not exported"
    assert sorted(pb._lib.SIGNATURES) == names, "ctypes binding table out of sync with the header"
    assert pb.load().pareben_version() == 2


def test_cited_reference_lines_in_header():
    text = open(os.path.join(ROOT, "include", "pareben.h")).read()
    for needle in ("elasticNetLinearNeMainEff.c:55-57", "ElasticNetBinaryNEmainEff.c:236-238", "R/CrossValidate.R:66-70"):
        assert needle in text


def test_shard_plan_partitions(built):
    import pareben_b200 as pb
    lam = np.repeat(np.exp(np.linspace(1, -6, 20)), 20)
    for n_folds in (3, 10):
        for w in (1, 2, 4, 8):
            parts = [pb.shard_plan(lam, n_folds, r, w) for r in range(w)]
            allidx = np.concatenate(parts)
            assert np.array_equal(np.sort(allidx), np.arange(400 * n_folds))
            sizes = [len(p) for p in parts]
            assert max(sizes) - min(sizes) <= 1
            # cost mix: every shard sees (almost) the same number of fits from every lambda rank
            for p in parts:
                per_lambda = np.bincount((p // n_folds) // 20, minlength=20)
                assert per_lambda.max() - per_lambda.min() <= 2 if w <= 4 else True


def test_compute_fails_loudly_without_device(built):
    import pareben_b200 as pb
    if pb.device_count() > 0:
        pytest.skip("a device is present")
    with pytest.raises(pb.ParebenError):
        pb.Problem(np.zeros((10, 3)), np.arange(10.0))
    with pytest.raises(pb.ParebenError):
        pb.cv_grid(np.zeros((10, 3)), np.arange(10.0), np.tile([1, 2], 5), 2, np.ones(2), np.ones(2))
    with pytest.raises(pb.ParebenError):
        pb.CrossValidate(np.zeros((10, 3)), np.arange(10.0), 2)
