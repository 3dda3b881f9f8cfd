"""The C++ host mirror of shared.predictions (include/mrs_predictions.hpp) over the C ABI.

CPU: the header and its driver compile and link against libmrs_b200.so, and without a GPU the program fails loudly
(no CPU fallback).  GPU: the driver's output is compared with the oracle on the same files -- averages bit-exact,
neighbour lists identical, predictions / MAE within 1e-6 relative."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "movie-recommender-system_b200")
REL = 1e-6


@pytest.fixture(scope="module")
def binary(tmp_path_factory):
    import __graft_entry__ as G
    if not os.path.exists(os.path.join(PKG, "libmrs_b200.so")):
        G.build()
    out = str(tmp_path_factory.mktemp("cpp") / "mirror_check")
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "mirror_check.cpp"), "-L", PKG, "-lmrs_b200", f"-Wl,-rpath,{PKG}", "-o", out],
                   check=True)
    return out


def _write(path, data, sep):
    u, i, r = data
    with open(path, "w") as f:
        f.write(f"user{sep}item{sep}rating\n")  # a header row: dropped because column 0 is not an Int (P:41-47)
        for a, b, c in zip(u.tolist(), i.tolist(), r.tolist()):
            f.write(f"{a}{sep}{b}{sep}{c}\n")


def test_mirror_compiles_links_and_fails_loudly_without_gpu(binary, tmp_path, small):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the parity test")
    _write(tmp_path / "tr.csv", small["train"], ",")
    _write(tmp_path / "te.csv", small["test"], ",")
    p = subprocess.run([binary, str(tmp_path / "tr.csv"), str(tmp_path / "te.csv"), ",", "10", "1"], capture_output=True, text=True)
    assert p.returncode == 1 and "mirror_check:" in p.stderr and p.stdout.strip() == ""


@pytest.mark.gpu
@pytest.mark.parametrize("name,sep,k", [("small", ",", 10), ("ml100k", "\t", 300)])
def test_mirror_matches_oracle(binary, tmp_path, small, ml100k, name, sep, k):
    from oracle import oracle as O
    d = small if name == "small" else ml100k
    tr, te = d["train"], d["test"]
    _write(tmp_path / "tr.data", tr, sep)
    _write(tmp_path / "te.data", te, sep)
    user = int(tr[0][0])
    p = subprocess.run([binary, str(tmp_path / "tr.data"), str(tmp_path / "te.data"), sep, str(k), str(user)],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr
    got = json.loads(p.stdout)
    o = O.Oracle(*tr)
    close = lambda a, b: abs(a - b) <= REL * max(1.0, abs(b))  # noqa: E731

    assert got["n_train"] == tr[0].size and got["n_test"] == te[0].size
    assert got["global_avg"] == o.global_avg
    assert got["n_users"] == np.unique(tr[0]).size
    u0 = int(np.min(tr[0]))
    assert got["user_avg_first"] == [u0, o.user_avg(u0)]
    for key, kind in [("mae_global", O.GLOBAL), ("mae_user", O.USER), ("mae_item", O.ITEM), ("mae_itemdev", O.ITEMDEV),
                      ("mae_baseline", O.BASELINE), ("mae_baseline_spark", O.BASELINE)]:
        assert close(got[key], o.mae(te, kind)), key
    assert close(got["pred_baseline_1_1"], o.predict(1, 1, O.BASELINE))
    assert close(got["pred_unknown_user"], o.predict(1 << 20, 1, O.BASELINE))
    for key, sk in [("mae_uniform", O.SIM_UNIFORM), ("mae_cosine", O.SIM_COSINE), ("mae_jaccard", O.SIM_JACCARD)]:
        assert close(got[key], o.mae(te, O.PERSONALIZED, sk, 0)), key
    assert got["sim_cosine_1_2"] == o.cosine(1, 2) and got["sim_jaccard_1_2"] == o.jaccard(1, 2)
    assert close(got["mae_knn"], o.mae(te, O.PERSONALIZED, O.SIM_COSINE, k))
    assert close(got["pred_knn_1_1"], o.predict(1, 1, O.PERSONALIZED, O.SIM_COSINE, k))
    assert close(got["wsd_knn_1_1"], o.wsd(1, 1, O.SIM_COSINE, k))
    ids, sims = o.neighbors(user, k)
    assert [a for a, _ in got["neighbors"]] == ids.tolist() and [b for _, b in got["neighbors"]] == sims.tolist()
    items, scores = o.recommend(user, 5, O.PERSONALIZED, O.SIM_COSINE, k)
    assert [a for a, _ in got["recommendations"]] == items.tolist()
    assert all(close(b, s) for (_, b), s in zip(got["recommendations"], scores.tolist()))
    assert got["scale"] == [2.0, 2.0, 1.0] and got["std"] == O.std([1, 2, 3, 4])
