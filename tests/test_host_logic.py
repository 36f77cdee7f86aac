"""CPU suite, part 3: the host-side mirror of the R layer (pareben_b200/cross_validate.py, rcompat.py)
against the oracle's independent restatement, plus the world_size-2 gloo merge path."""
import math
import os
import subprocess
import sys
import textwrap

import numpy as np

from conftest import ROOT, golden
from oracle import rlayer as R


def test_assign_to_folds_matches_oracle():
    import pareben_b200 as pb
    g = golden("rrng.npz")
    assert np.array_equal(pb.AssignToFolds(np.zeros((50, 2)), 3), g["folds_50_3"])
    assert np.array_equal(pb.AssignToFolds(np.zeros((500, 2)), 5), g["folds_500_5"])
    assert np.array_equal(pb.AssignToFolds(np.zeros((1000, 2)), 10), g["folds_1000_10"])
    assert np.array_equal(pb.AssignToFolds(np.zeros((50, 2)), 3, sample_kind="Rounding"), g["folds_50_3_rounding"])
    mine = np.arange(50) % 3 + 1
    assert np.array_equal(pb.AssignToFolds(np.zeros((50, 2)), 3, mine), mine)       # user foldId of length N passes through


def test_rcompat_numerics():
    from pareben_b200 import rcompat
    x = np.array([2246.40804244, 2004.15893851, 1558.65730397])
    assert rcompat.mean(x) == R.r_mean(x) and rcompat.sd(x) == R.r_sd(x)
    assert np.array_equal(rcompat.seq(1.0, 0.05, -0.05), R.r_seq(1.0, 0.05, -0.05))
    a = rcompat.seq(1.0, 0.05, -0.05)
    assert a.size == 20 and a[-1] == 0.05 and a[-2] == 1 + 18 * (-0.05)


def test_summary_and_local_search_against_oracle(monkeypatch):
    """Feed the host layer the reference's config-1 error table through a stubbed device call and
    compare every derived object with the oracle's restatement."""
    import pareben_b200 as pb
    from pareben_b200 import cross_validate as cv
    g = golden("config1_gaussian.npz")
    X = golden("inputs_bundled.npz")["BASIS"][:50, :100].astype(float)
    y = golden("inputs_bundled.npz")["y"][:50]
    monkeypatch.setattr(cv, "BuildGrid", lambda *a, **k: {"alpha": g["grid_alpha"], "lambda": g["grid_lambda"]})
    monkeypatch.setattr(cv, "_grid_errors", lambda *a, **k: (g["fold_err"].copy(), np.zeros((400, 3), np.int32), g["n_selected"]))
    out = pb.CrossValidate(X, y, 3)
    assert out["alpha.optimal"] == float(g["alpha_optimal"]) and out["lambda.optimal"] == float(g["lambda_optimal"])
    assert np.array_equal(out["Results.Summary"]["MSE"], g["summary_mse"])
    assert np.array_equal(out["Results.Summary"]["SE"], g["summary_se"])
    assert np.array_equal(out["Results.Detail"]["foldId"][:4], [1, 2, 3, 1])
    assert np.array_equal(out["Results.Detail"]["MSE"], g["fold_err"].ravel())
    loc = pb.CrossValidate(X, y, 3, search="local")
    assert np.array_equal(loc["CrossValidation"], g["local_cv"]) and np.array_equal(loc["fullCV"], g["local_full"])
    assert loc["alpha.optimal"] == float(g["local_alpha"]) and loc["lambda.optimal"] == float(g["local_lambda"])


def test_two_rank_gloo_merge(tmp_path):
    """world_size 2 on CPU (gloo): each rank fills only its shard (as pareben_cv_grid does) and the
    host layer's all-reduce must rebuild the full table on both ranks."""
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys
        sys.path.insert(0, {ROOT!r})
        import numpy as np, torch.distributed as dist
        import pareben_b200 as pb
        from pareben_b200 import cross_validate as cv, _lib
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        lam = np.repeat(np.exp(np.linspace(1, -6, 20)), 20); alpha = np.tile(np.linspace(1, .05, 20), 20)
        truth = np.sin(np.arange(1200.0)).reshape(400, 3) + 3
        def fake(X, y, f, nf, a, l, epis, prior, device, shard=0, n_shards=1):
            mine = pb.shard_plan(l, nf, shard, n_shards)
            err = np.zeros(1200); st = np.zeros(1200, np.int32); ns = np.zeros(1200, np.int32)
            err[mine] = truth.ravel()[mine]; ns[mine] = mine % 7; st[mine] = (mine % 11 == 0)
            return err.reshape(400, 3), st.reshape(400, 3), ns.reshape(400, 3)
        _lib.cv_grid = fake
        err, st, ns = cv._grid_errors(None, None, None, 3, alpha, lam, False, "gaussian", 0)
        assert np.array_equal(err, truth), "merged table differs"
        assert np.array_equal(ns.ravel(), np.arange(1200) % 7) and np.array_equal(st.ravel(), (np.arange(1200) % 11 == 0).astype(np.int32))
        dist.destroy_process_group()
        print("rank", rank, "ok")
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                       capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2
