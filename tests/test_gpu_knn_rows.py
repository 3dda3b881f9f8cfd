"""GPU parity of the row-block kNN path (knn_rows.cu): neighbour lists without a similarity matrix, for user counts
above the dense path's limit and for a rank that owns a range of the rows (BASELINE config 5).

Forced onto the ml-100k shape it must reproduce the oracle exactly like the dense path does (neighbour ids, order and
similarity bits; predictions and MAE within 1e-6); on 20,000 users (above the 16,384 limit) mrs_fit_similarity takes it
by itself and is spot-checked against the oracle per user."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import mrs_b200  # noqa: F401,E402
from mrs_b200 import engine as E  # noqa: E402
from mrs_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

REL = 1e-6
ALL = (0, 2**31 - 1)


@pytest.fixture(scope="module")
def eng():
    e = E.Engine(0)
    yield e
    e.close()


@pytest.fixture(scope="module")
def fitted(eng, ml100k):
    tr, te = ml100k["train"], ml100k["test"]
    R, T = eng.ratings(*tr), eng.ratings(*te)
    m = E.Model(eng, R)
    return R, T, m, O.Oracle(*tr), tr, te


@pytest.mark.parametrize("kind,okind", [(E.SIM_COSINE, O.SIM_COSINE), (E.SIM_JACCARD, O.SIM_JACCARD)])
@pytest.mark.parametrize("k", [10, 300, 942])
def test_lists_bit_exact_on_ml100k_shape(fitted, kind, okind, k):
    R, T, m, o, tr, te = fitted
    s = m.similarity(kind, k, rows=ALL)
    d = m.similarity(kind, k)  # dense path
    for u in list(range(1, 944, 11)) + [943]:
        ids, sims = s.neighbors(u, k)
        oi, os_ = o.neighbors(u, k, okind)
        assert ids.tolist() == oi.tolist()                               # index sets AND order
        assert sims.tolist() == os_.tolist()                             # similarity bits
        di, ds = d.neighbors(u, k)
        assert ids.tolist() == di.tolist() and sims.tolist() == ds.tolist()
    s.close()
    d.close()


def test_lists_predictions_mae_and_prefix_property(fitted):
    R, T, m, o, tr, te = fitted
    s = m.similarity(E.SIM_COSINE, 300, rows=ALL)
    n = 2500
    for k in (10, 30, 300):
        s.set_k(k)                                                       # any k below the fitted one is a prefix (A.6)
        assert m.mae(T, E.PRED_PERSONALIZED, s) == pytest.approx(o.mae(te, kind=O.PERSONALIZED, simkind=O.SIM_COSINE, k=k), rel=REL)
        p = m.predict(te[0][:n], te[1][:n], E.PRED_PERSONALIZED, s)
        ref = o.predict_batch(te[0][:n], te[1][:n], kind=O.PERSONALIZED, simkind=O.SIM_COSINE, k=k)
        assert np.allclose(p, ref, rtol=REL, atol=0)
    with pytest.raises(E.MrsError):
        s.set_k(301)                                                     # more than was kept: refused, not approximated
    with pytest.raises(E.MrsError):
        s.set_k(0)
    s.set_k(30)
    us, is_ = te[0][:400], te[1][:400]
    w = m.predict(us, is_, E.PRED_WSD, s)
    ref = np.array([o.wsd(int(u), int(i), k=30) for u, i in zip(us, is_)])
    assert np.allclose(w, ref, rtol=REL, atol=1e-15)
    p = m.predict([5000, 1, 5000], [1, 99999, 99999], E.PRED_PERSONALIZED, s)
    assert p[0] == o.global_avg and p[1] == o.user_avg(1) and p[2] == o.global_avg
    # s_k(u, v) (P:638-641)
    ids, sims = s.neighbors(1, 30)
    assert s(1, int(ids[0])) == sims[0] == o.similarity(1, int(ids[0]), k=30)
    assert s(1, 1) == 0.0
    far = int(o.neighbors(1, 942)[0][-1])
    assert s(1, far) == 0.0
    items, scores = m.recommend(400, 5, E.PRED_PERSONALIZED, s)
    oi, os_ = o.recommend(400, 5, k=30)
    assert items.tolist() == oi.tolist() and np.allclose(scores, os_, rtol=REL, atol=0)
    s.close()


def test_row_ranges_partition_the_work(fitted, eng):
    """Two handles over disjoint user ranges == one handle over all users (what two ranks of a sharded run hold)."""
    R, T, m, o, tr, te = fitted
    k, cut = 50, 480
    lo = m.similarity(E.SIM_COSINE, k, rows=(0, cut))
    hi = m.similarity(E.SIM_COSINE, k, rows=(cut, 10**6))
    for u in (1, 7, cut - 1):
        assert lo.neighbors(u, k)[0].tolist() == o.neighbors(u, k)[0].tolist()
        with pytest.raises(E.MrsError):
            hi.neighbors(u, k)
    for u in (cut, 700, 943):
        assert hi.neighbors(u, k)[0].tolist() == o.neighbors(u, k)[0].tolist()
        with pytest.raises(E.MrsError):
            lo.neighbors(u, k)
    assert np.isnan(m.predict([943], [1], E.PRED_PERSONALIZED, lo)[0])   # another rank's row: loud, not a guess
    sel = te[0] < cut
    parts = []
    for s, mask in ((lo, sel), (hi, ~sel)):
        Ts = eng.ratings(te[0][mask], te[1][mask], te[2][mask])
        parts.append((m.mae(Ts, E.PRED_PERSONALIZED, s) * int(mask.sum()), int(mask.sum())))
        Ts.close()
    total = sum(a for a, _ in parts) / sum(b for _, b in parts)
    assert total == pytest.approx(o.mae(te, kind=O.PERSONALIZED, simkind=O.SIM_COSINE, k=k), rel=REL)
    lo.close()
    hi.close()


def test_above_the_dense_limit(eng):
    d = synth.cached("mid")
    tr, te = d["train"], d["test"]
    assert np.unique(tr[0]).size > 16384
    R = eng.ratings(*tr)
    m = E.Model(eng, R)
    o = O.Oracle(*tr)
    with pytest.raises(E.MrsError):
        m.similarity(E.SIM_COSINE, 0)                                    # the full matrix is not available at this size
    k = 50
    s = m.similarity(E.SIM_COSINE, k)                                    # takes the row-block path by itself
    rng = np.random.default_rng(3)
    for u in [1, 2, 16384, 16385, 20000] + rng.integers(1, 20001, 40).tolist():
        ids, sims = s.neighbors(int(u), k)
        oi, os_ = o.neighbors(int(u), k)
        assert ids.tolist() == oi.tolist() and sims.tolist() == os_.tolist()
        assert all(sims[j] > sims[j + 1] or (sims[j] == sims[j + 1] and ids[j] < ids[j + 1]) for j in range(len(ids) - 1))
    n = 600
    sub = (te[0][:n], te[1][:n], te[2][:n])
    T = eng.ratings(*sub)
    assert m.mae(T, E.PRED_PERSONALIZED, s) == pytest.approx(o.mae(sub, kind=O.PERSONALIZED, simkind=O.SIM_COSINE, k=k), rel=REL)
    p = m.predict(sub[0], sub[1], E.PRED_PERSONALIZED, s)
    assert np.allclose(p, o.predict_batch(sub[0], sub[1], kind=O.PERSONALIZED, simkind=O.SIM_COSINE, k=k), rtol=REL, atol=0)
    j = m.similarity(E.SIM_JACCARD, k)
    for u in (3, 19999):
        ids, sims = j.neighbors(u, k)
        oi, os_ = o.neighbors(u, k, O.SIM_JACCARD)
        assert ids.tolist() == oi.tolist() and sims.tolist() == os_.tolist()
    for h in (s, j, T, m, R):
        h.close()
