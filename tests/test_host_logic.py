"""Host-side logic that needs no GPU: reference-style utilities of the predictions mirror, the synthetic
generators and the text format they write."""
import numpy as np
import pytest

import mrs_b200  # noqa: F401
from mrs_b200 import predictions as P
from mrs_b200 import synth
import oracle


def test_scale_mean_std_match_oracle():
    for x, y in ((4, 3.5), (3, 3.5), (3.5, 3.5), (0.5, 0.75)):
        assert P.scale(x, y) == oracle.scale(x, y)
    xs = [1.0, 2.5, 4.0, 4.0]
    assert P.mean(xs) == oracle.mean(xs) and P.std(xs) == pytest.approx(oracle.std(xs), abs=1e-15)
    assert P.mean([]) == 0.0 and P.std([]) == 0.0


def test_untagged_functions_are_rejected():
    with pytest.raises(P.UnsupportedOperationError):
        P.MAE(lambda u, i: 3.0, P.RatingSet.__new__(P.RatingSet))


def test_ml100k_shape(ml100k):
    u, i, r = ml100k["all"]
    assert u.size == 100_000 and ml100k["train"][0].size == 80_000 and ml100k["test"][0].size == 20_000
    assert u.min() == 1 and u.max() == 943 and i.min() >= 1 and i.max() <= 1682
    assert np.unique(u.astype(np.int64) * 4096 + i).size == u.size
    assert set(np.unique(r)) <= {1.0, 2.0, 3.0, 4.0, 5.0}
    d2 = synth.ml100k()
    assert all(np.array_equal(a, b) for a, b in zip(d2["all"], ml100k["all"]))  # seeded => reproducible


def test_writer_round_trip(tmp_path, small):
    u, i, r = small["train"]
    p = tmp_path / "r.csv"
    synth.write_ratings(str(p), u, i, r, sep=",", header="userId,movieId,rating,timestamp")
    rows = [l.split(",") for l in p.read_text().splitlines()[1:]]
    assert len(rows) == u.size
    assert [int(x[0]) for x in rows] == u.tolist() and [float(x[2]) for x in rows] == r.tolist()
