"""GPU parity of the baseline family against the CPU oracle, through the C ABI (engine binding) and through
the predictions mirror.  Tolerances (BASELINE.json north_star): averages bit-exact, deviations / predictions /
MAE within 1e-6 relative."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import mrs_b200  # noqa: F401,E402
from mrs_b200 import engine as E  # noqa: E402
from mrs_b200 import predictions as P  # noqa: E402
from mrs_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402
import oracle  # noqa: E402

REL = 1e-6


def close(a, b, rel=REL, floor=1e-12):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return bool(np.all(np.abs(a - b) <= rel * np.maximum(np.abs(b), floor) + floor))


@pytest.fixture(scope="module")
def eng():
    e = E.Engine(0)
    yield e
    e.close()


def fit_both(eng, d):
    tr, te = d["train"], d["test"]
    R, T = eng.ratings(*tr), eng.ratings(*te)
    return R, T, E.Model(eng, R), O.Oracle(*tr)


def check_model(m, o, te):
    assert m.global_avg == o.global_avg                                    # bit-exact (A.1)
    ua, uc = m.vector(E.USER_AVG)
    ia, ic = m.vector(E.ITEM_AVG)
    idv, _ = m.vector(E.ITEM_AVG_DEV)
    for u in range(ua.size):
        assert ua[u] == o.user_avg(u) and uc[u] == o.user_count(u)         # bit-exact
    for i in range(ia.size):
        assert ia[i] == o.item_avg(i) and ic[i] == o.item_count(i)         # bit-exact
    assert close(idv, [o.item_avg_dev(i) for i in range(idv.size)])
    for kind, okind in ((E.PRED_GLOBAL, O.GLOBAL), (E.PRED_USER, O.USER), (E.PRED_ITEM, O.ITEM),
                        (E.PRED_ITEMDEV, O.ITEMDEV), (E.PRED_BASELINE, O.BASELINE)):
        assert close(m.predict(te[0], te[1], kind), o.predict_batch(te[0], te[1], kind=okind))


def test_ml100k_vectors_predictions_mae(eng, ml100k):
    R, T, m, o = fit_both(eng, ml100k)
    assert R.value_kind == 0 and R.n == 80_000
    check_model(m, o, ml100k["test"])
    for kind, okind in ((E.PRED_GLOBAL, O.GLOBAL), (E.PRED_USER, O.USER), (E.PRED_ITEM, O.ITEM), (E.PRED_BASELINE, O.BASELINE)):
        assert m.mae(T, kind) == pytest.approx(o.mae(ml100k["test"], kind=okind), rel=REL)
    # single lookups with fallbacks
    assert m.lookup(E.USER_AVG, 1) == (o.user_avg(1), True)
    assert m.lookup(E.USER_AVG, 5000) == (o.global_avg, False)
    assert m.lookup(E.ITEM_AVG, 99999) == (o.global_avg, False)
    assert m.lookup(E.ITEM_AVG_DEV, 99999) == (0.0, False)
    p = m.predict([5000, 1, 5000, -3], [1, 99999, 99999, 1], E.PRED_BASELINE)
    assert p[0] == o.global_avg and p[1] == o.user_avg(1) and p[2] == o.global_avg and p[3] == o.global_avg
    # refit on the same buffers is bit-reproducible
    before = m.vector(E.ITEM_AVG_DEV)[0].copy()
    m.refit(); eng.sync()
    assert np.array_equal(before, m.vector(E.ITEM_AVG_DEV)[0])


def test_half_star_small_and_fp64_values(eng, small):
    R, T, m, o = fit_both(eng, small)
    assert R.value_kind == 0
    check_model(m, o, small["test"])
    # ratings that are not multiples of 0.5 take the fp64 value path
    tr, te = small["train"], small["test"]
    rng = np.random.default_rng(3)
    tr2 = (tr[0], tr[1], np.round(tr[2] + rng.uniform(-0.4, 0.4, tr[2].size), 3))
    te2 = (te[0], te[1], np.round(te[2] + rng.uniform(-0.4, 0.4, te[2].size), 3))
    R2, T2 = eng.ratings(*tr2), eng.ratings(*te2)
    assert R2.value_kind == 1
    m2, o2 = E.Model(eng, R2), O.Oracle(*tr2)
    assert m2.global_avg == pytest.approx(o2.global_avg, rel=1e-13)
    assert close(m2.vector(E.USER_AVG)[0], [o2.user_avg(u) for u in range(R2.n_users_dim)], rel=1e-12)
    assert close(m2.vector(E.ITEM_AVG_DEV)[0], [o2.item_avg_dev(i) for i in range(R2.n_items_dim)])
    assert m2.mae(T2, E.PRED_BASELINE) == pytest.approx(o2.mae(te2, kind=O.BASELINE), rel=REL)


def test_edge_cases(eng, small):
    tr = small["train"]
    R = eng.ratings(*tr)
    m = E.Model(eng, R)
    empty = eng.ratings(np.array([], np.int32), np.array([], np.int32), np.array([], np.float64))
    assert math.isnan(m.mae(empty, E.PRED_BASELINE))                      # 0.0/0 (P:85)
    m0 = E.Model(eng, empty)
    assert m0.global_avg == 0.0                                           # mean of an empty Seq (P:18)
    assert m0.predict([1], [1], E.PRED_BASELINE)[0] == 0.0
    one = eng.ratings([7], [9], [4.5])
    m1 = E.Model(eng, one)
    assert m1.global_avg == 4.5 and m1.lookup(E.ITEM_AVG_DEV, 9)[0] == 0.0  # r == avg -> 0/1 (A.2)
    assert m1.predict([7], [9], E.PRED_BASELINE)[0] == 4.5
    with pytest.raises(E.MrsError) as ei:
        eng.ratings([1, 1, 2], [5, 5, 5], [3.0, 4.0, 5.0])
    assert ei.value.status == -4
    with pytest.raises(E.MrsError):
        eng.ratings([1, -2], [5, 5], [3.0, 4.0])
    with pytest.raises(E.MrsError):
        m.mae(R, E.PRED_PERSONALIZED)                                    # no similarity handle, no fallback


def test_text_loader_matches_coo(eng, small, tmp_path):
    u, i, r = small["train"]
    p = tmp_path / "u.base"
    synth.write_ratings(str(p), u, i, r, sep="\t")
    Rf = eng.ratings_from_file(str(p), "\t")
    q = tmp_path / "r.csv"
    with open(q, "w", newline="") as f:
        f.write("userId,movieId,rating,timestamp\r\n")                   # header row is dropped (P:40), CRLF tolerated
        for a, b, c in zip(u.tolist(), i.tolist(), r.tolist()):
            f.write(f" {a} ,{b}, {c},123\r\n")
        f.write("\r\n")
    Rc = eng.ratings_from_file(str(q), ",")
    Rm = eng.ratings(u, i, r)
    assert Rf.n == Rc.n == Rm.n == u.size
    a, b, c = E.Model(eng, Rf), E.Model(eng, Rc), E.Model(eng, Rm)
    for k in (E.USER_AVG, E.ITEM_AVG, E.ITEM_AVG_DEV):
        assert np.array_equal(a.vector(k)[0], c.vector(k)[0]) and np.array_equal(b.vector(k)[0], c.vector(k)[0])
    bad = tmp_path / "bad.csv"
    bad.write_text("1,2\n")
    with pytest.raises(E.MrsError) as ei:
        eng.ratings_from_file(str(bad), ",")
    assert ei.value.status == -5
    with pytest.raises(E.MrsError):
        eng.ratings_from_file(str(tmp_path / "missing"), ",")


def test_predictions_mirror_reads_like_the_reference(eng, ml100k):
    P.set_default_engine(eng)
    train = P.RatingSet.from_arrays(*ml100k["train"])
    test = P.RatingSet.from_arrays(*ml100k["test"])
    o = O.Oracle(*ml100k["train"])
    te = ml100k["test"]
    # predict/Baseline.scala:46-67,93-103
    assert P.MAE(P.computeAvgRating(train), test) == pytest.approx(o.mae(te, kind=O.GLOBAL), rel=REL)
    assert P.MAE(P.computeUserAvg(train), test) == pytest.approx(o.mae(te, kind=O.USER), rel=REL)
    assert P.MAE(P.computeItemAvg(train), test) == pytest.approx(o.mae(te, kind=O.ITEM), rel=REL)
    assert P.MAE(P.computePrediction(train), test) == pytest.approx(o.mae(te, kind=O.BASELINE), rel=REL)
    assert P.average(train) == o.global_avg
    assert P.computeUserAvg(train)(1, 1) == o.user_avg(1)
    assert P.computeItemAvg(train)(1, 1) == o.item_avg(1)
    assert P.computeItemAvgDev(train)(1, 1) == pytest.approx(o.item_avg_dev(1), rel=REL)
    assert P.computePrediction(train)(1, 1) == pytest.approx(o.predict(1, 1), rel=REL)
    # distributed/DistributedBaseline.scala:46,70-75
    assert P.MeanAbsoluteErrorSpark(P.baselinePredictorSpark(train), test) == pytest.approx(
        oracle.spark_baseline_mae(ml100k["train"], te, nthreads=4)[0], rel=REL)
    assert P.getGlobalAvg(train) == o.global_avg
    assert P.getUsersAvg(train).getOrElse(1, -1.0) == o.user_avg(1)
    assert P.getItemsAvg(train)[1] == o.item_avg(1)
    assert 5000 not in P.usersAvg(train)


def test_ml25m_shape_full_size(eng):
    """Full BASELINE.json size: oracle's Spark twin for the MAE, exact integer arithmetic for the averages and
    linearity properties for the deviation sums."""
    d = synth.cached("ml25m")
    tr, te = d["train"], d["test"]
    R, T = eng.ratings(*tr), eng.ratings(*te)
    assert R.value_kind == 0 and R.n == 20_000_076 and T.n == 5_000_019
    m = E.Model(eng, R)
    mae = m.mae(T, E.PRED_BASELINE)
    ref_mae, ref_g = oracle.spark_baseline_mae(tr, te, nthreads=oracle.max_threads())
    assert m.global_avg == ref_g
    assert mae == pytest.approx(ref_mae, rel=REL)
    # averages: exact in any order for half-star data
    ua, uc = m.vector(E.USER_AVG)
    cnt = np.bincount(tr[0], minlength=ua.size)
    s = np.bincount(tr[0], weights=tr[2], minlength=ua.size)
    known = cnt > 0
    assert np.array_equal(uc, cnt) and np.array_equal(ua[known], s[known] / cnt[known])
    ia, ic = m.vector(E.ITEM_AVG)
    cnt_i = np.bincount(tr[1], minlength=ia.size)
    s_i = np.bincount(tr[1], weights=tr[2], minlength=ia.size)
    ki = cnt_i > 0
    assert np.array_equal(ic, cnt_i) and np.array_equal(ia[ki], s_i[ki] / cnt_i[ki])
    # deviations: numpy restatement of P:167 + a checksum of sums (order-free to 1e-9)
    a = ua[tr[0]]
    sc = np.where(tr[2] > a, 5 - a, np.where(tr[2] < a, a - 1, 1.0))
    dev = (tr[2] - a) / sc
    ds = np.bincount(tr[1], weights=dev, minlength=ia.size)
    idv, _ = m.vector(E.ITEM_AVG_DEV)
    assert close(idv[ki], ds[ki] / cnt_i[ki])
    assert float((idv * cnt_i).sum()) == pytest.approx(float(dev.sum()), rel=1e-9, abs=1e-6)
    p = m.predict(te[0][:200000], te[1][:200000], E.PRED_BASELINE)
    assert p.min() >= 0.5 - 1e-9 and p.max() <= 5.0 + 1e-9


def test_device_text_parser_follows_the_reference_rules(eng):
    """`load` (P:35-49): split on the separator, trim every column, keep the row iff column 0 is an Int; column 1 must be
    an Int and column 2 a Double.  Parsed on the device; odd number forms go through the host parser with the same result."""
    def parse(text, sep):
        R = eng.ratings_from_text(text, sep)
        m = E.Model(eng, R)
        n = R.n
        ua, uc = m.vector(E.USER_AVG)
        ia, ic = m.vector(E.ITEM_AVG)
        out = (n, {u: ua[u] for u in np.flatnonzero(uc)}, {i: ia[i] for i in np.flatnonzero(ic)})
        m.close(); R.close()
        return out

    plain = b"userId::movieId::rating::ts\n 1 :: 10 :: 4.5 ::99\n2::10::3\n\n+3::11::0.5::x::y\r\nabc::1::1\n2::11:: 5.0\n7::12::.5\n7::13::2."
    n, ua, ia = parse(plain, "::")
    assert n == 6                                                       # header, blank line and the `abc` row are dropped
    assert ua == {1: 4.5, 2: 4.0, 3: 0.5, 7: 1.25} and ia == {10: 3.75, 11: 2.75, 12: 0.5, 13: 2.0}
    odd = plain + b"\n9::14::1e0\n9::15::0x1p1\n"                      # exponent / hex forms: host parser, same rows plus two
    n2, ua2, ia2 = parse(odd, "::")
    assert n2 == 8 and ua2[9] == 1.5 and ia2[15] == 2.0 and {k: v for k, v in ua2.items() if k != 9} == ua
    assert parse(b"1\t2\t3.25\n", "\t")[0] == 1 and eng.ratings_from_text(b"", ",").n == 0
    assert parse(b"1,2,0.1\n1,3,0.7\n", ",")[1] == {1: (0.1 + 0.7) / 2}                                    # correctly rounded
    assert parse(b"1,2,0.1\n1,3,0.30000000000000004\n", ",")[1] == {1: (0.1 + 0.30000000000000004) / 2}   # 17 digits: host parser
    for bad in (b"1,2\n", b"1,x,3\n", b"5,6,7\n1,2,\n", b"1,2,abc\n"):
        with pytest.raises(E.MrsError) as ei:
            eng.ratings_from_text(bad, ",")
        assert ei.value.status == -5
    with pytest.raises(E.MrsError):
        eng.ratings_from_text(b"-1,2,3\n", ",")                        # negative id: rejected by the build, as from arrays


def test_cost_cut_user_tiles(eng, small, monkeypatch):
    """MRS_TILES=1 (opt-in): the item pass' user tiles are cut so that every CTA gets the same number of ratings instead of the
    same number of users; the fit must not change by a bit."""
    R, T, m, o = fit_both(eng, small)
    want = m.vector(E.ITEM_AVG_DEV)[0].copy()
    want_mae = m.mae(T, E.PRED_BASELINE)
    monkeypatch.setenv("MRS_TILES", "1")
    tr = small["train"]
    R2 = eng.ratings(*tr)          # a new rating set: the tiled layout is built on its first fit, with the cut tiles
    m2 = E.Model(eng, R2)
    assert np.array_equal(m2.vector(E.ITEM_AVG_DEV)[0], want)
    assert m2.mae(T, E.PRED_BASELINE) == want_mae
    check_model(m2, o, small["test"])
    for h in (m2, R2):
        h.close()
