"""C oracle vs the statement-by-statement Python restatement (tests/ref_literal.py) on small sets,
plus the dataset-independent invariants of SURVEY A.9."""
import math

import numpy as np
import pytest

import oracle
from oracle import oracle as O
import ref_literal as L


def as_list(t):
    return list(zip(t[0].tolist(), t[1].tolist(), t[2].tolist()))


@pytest.fixture(scope="module")
def fitted(small):
    tr, te = small["train"], small["test"]
    return O.Oracle(*tr), as_list(tr), as_list(te), tr, te


def test_baseline_family_matches_literal(fitted):
    o, tr, te, _, tet = fitted
    assert o.global_avg == L.average(tr)
    ua, ia, idev = L.users_avg(tr), L.items_avg(tr), L.items_avg_dev(tr)
    for u, v in ua.items():
        assert o.user_avg(u) == v
    for i, v in ia.items():
        assert o.item_avg(i) == v
    for i, v in idev.items():
        assert o.item_avg_dev(i) == v
    nd = L.normalize_deviation(tr)
    assert np.array_equal(o.deviations(), np.array([nd[(u, i)] for (u, i, _) in tr]))
    pred = L.compute_prediction(tr)
    for (u, i, _) in te[:200]:
        assert o.predict(u, i) == pred(u, i)
    assert o.mae(tet, kind=O.BASELINE) == L.mae(pred, te)
    assert o.mae(tet, kind=O.GLOBAL) == L.mae(lambda u, i: L.average(tr), te)
    assert o.mae(tet, kind=O.USER) == L.mae(lambda u, i: ua.get(u, L.average(tr)), te)
    assert o.mae(tet, kind=O.ITEM) == L.mae(lambda u, i: ia.get(i, L.average(tr)), te)


def test_fallbacks(fitted):
    o, tr, _, _, _ = fitted
    g = L.average(tr)
    big = 10_000
    assert o.user_avg(big) == g and o.item_avg(big) == g and o.item_avg_dev(big) == 0.0
    assert o.predict(big, 1) == g                       # unknown user -> global average (P:222-224)
    u0 = tr[0][0]
    assert o.predict(u0, big) == o.user_avg(u0)          # unknown item -> user average exactly (A.3)
    assert o.predict(big, big, kind=O.PERSONALIZED, simkind=O.SIM_COSINE, k=5) == g
    assert o.predict(u0, big, kind=O.PERSONALIZED, simkind=O.SIM_COSINE, k=5) == o.user_avg(u0)


def test_cosine_and_neighbours_match_literal(fitted):
    o, tr, te, _, _ = fitted
    cos = L.adjusted_cosine(tr)
    users = sorted({r[0] for r in tr})
    for u in users[:12]:
        for v in users:
            assert o.cosine(u, v) == cos(u, v)
    pre = L.preprocessed_rating(tr)
    for (u, i, _) in tr[:100]:
        assert o.pair_values(u, i)[1] == pre[(u, i)]
    for k in (1, 3, 10, len(users) - 1, len(users) + 5):
        nn = L.get_neighbors(tr, k, cos)
        for u in users[:8]:
            ids, sims = o.neighbors(u, k)
            ref = nn(u)
            assert ids.tolist() == [x[0] for x in ref]
            assert sims.tolist() == [x[1] for x in ref]


@pytest.mark.parametrize("k", [0, 3, 10, 1000])
def test_personalized_predictions_match_literal(fitted, k):
    o, tr, te, _, tet = fitted
    cos = L.adjusted_cosine(tr)
    sim = cos if k == 0 else L.get_similarity(tr, k, cos)
    pred = L.predictor(tr, L.weighted_sum_deviation(tr, sim))
    for (u, i, _) in te[:120]:
        assert o.predict(u, i, kind=O.PERSONALIZED, simkind=O.SIM_COSINE, k=k) == pytest.approx(pred(u, i), rel=0, abs=1e-14)
    assert o.mae(tet, kind=O.PERSONALIZED, simkind=O.SIM_COSINE, k=k) == pytest.approx(L.mae(pred, te), abs=1e-13)


def test_jaccard_matches_literal(fitted):
    o, tr, te, _, tet = fitted
    jac = L.jaccard(tr)
    users = sorted({r[0] for r in tr})
    for u in users[:10]:
        for v in users[:30]:
            assert o.jaccard(u, v) == jac(u, v)
    pred = L.predictor(tr, L.weighted_sum_deviation(tr, jac))
    assert o.mae(tet, kind=O.PERSONALIZED, simkind=O.SIM_JACCARD) == pytest.approx(L.mae(pred, te), abs=1e-13)


def test_recommendations_match_literal(fitted):
    o, tr, _, _, _ = fitted
    cos = L.adjusted_cosine(tr)
    pred = L.predictor(tr, L.weighted_sum_deviation(tr, L.get_similarity(tr, 7, cos)))
    rec = L.recommendations(tr, pred)
    for user in (tr[0][0], tr[5][0]):
        items, scores = o.recommend(user, 5, k=7)
        ref = rec(user, 5)
        assert items.tolist() == [x[0] for x in ref]
        assert np.allclose(scores, [x[1] for x in ref], rtol=0, atol=1e-14)


def test_invariants_a9(ml100k):
    tr, te = ml100k["train"], ml100k["test"]
    o = O.Oracle(*tr)
    n = 3000
    sub = (te[0][:n], te[1][:n], te[2][:n])
    base = o.predict_batch(sub[0], sub[1], kind=O.BASELINE)
    uni = o.predict_batch(sub[0], sub[1], kind=O.PERSONALIZED, simkind=O.SIM_UNIFORM)
    assert np.allclose(base, uni, rtol=1e-12, atol=0)                     # (1) uniform == baseline
    cos = o.predict_batch(sub[0], sub[1], kind=O.PERSONALIZED, simkind=O.SIM_COSINE)
    knn_all = o.predict_batch(sub[0], sub[1], kind=O.PERSONALIZED, simkind=O.SIM_COSINE, k=943)
    assert np.allclose(cos, knn_all, rtol=1e-12, atol=0)                  # (2) k >= U-1 == plain cosine (test pairs are not in train)
    for u in (1, 2, 500, 943):
        assert o.similarity(u, u, k=10) == 0.0                            # (3)
        assert abs(o.cosine(u, u) - 1.0) < 1e-12                          # (4)
        for v in (3, 77, 400):
            assert abs(o.cosine(u, v)) <= 1 + 1e-12 and o.cosine(u, v) == o.cosine(v, u)   # (5)
    for kind in (O.BASELINE,):
        p = o.predict_batch(te[0], te[1], kind=kind)
        assert p.min() >= 1.0 - 1e-12 and p.max() <= 5.0 + 1e-12           # (6)
    # (7) Spark twin == collections version
    mae_spark, g = oracle.spark_baseline_mae(tr, te, nthreads=1)
    assert g == o.global_avg
    assert mae_spark == pytest.approx(o.mae(te, kind=O.BASELINE), rel=1e-12)
    mae_spark4, _ = oracle.spark_baseline_mae(tr, te, nthreads=4)
    assert mae_spark4 == pytest.approx(mae_spark, rel=1e-12)
    o.close()


def test_mean_std_and_empty():
    assert oracle.mean([]) == 0.0 and oracle.std([]) == 0.0
    assert oracle.mean([1.0, 2.0, 4.0]) == 7.0 / 3
    assert oracle.std([1.0, 3.0]) == 1.0  # population std (P:23)
    o = O.Oracle(np.array([1], np.int32), np.array([1], np.int32), np.array([3.0]))
    e = (np.array([], np.int32), np.array([], np.int32), np.array([], np.float64))
    assert math.isnan(o.mae(e))  # 0.0/0 in applyAndMean (P:85)


def _tie_key(i):
    h = (i + (~(i << 9))) & 0xFFFFFFFF
    h ^= h >> 14
    h = (h + (h << 4)) & 0xFFFFFFFF
    h ^= h >> 10
    key = 0
    for k in range(7):
        key = (key << 5) | ((h >> (5 * k)) & 31)
    return key


def test_tie_order_only_moves_exact_ties(small):
    """(sim desc, tie_rank asc); tie_rank = user id (default) or the Scala 2.11 HashSet iteration order."""
    tr = small["train"]
    o = O.Oracle(*tr)
    n = int(np.unique(tr[0]).size) - 1
    moved = 0
    for u in np.unique(tr[0])[:12]:
        ids0, s0 = o.neighbors(int(u), n)
        o.set_tie_order(1)
        ids1, s1 = o.neighbors(int(u), n)
        o.set_tie_order(0)
        assert s0.tolist() == s1.tolist()                       # the similarities are the same sorted sequence
        assert sorted(ids0.tolist()) == sorted(ids1.tolist())
        j = 0
        while j < n:                                            # group by exactly equal similarity
            k = j
            while k < n and s0[k] == s0[j]:
                k += 1
            g0, g1 = ids0[j:k].tolist(), ids1[j:k].tolist()
            assert sorted(g0) == sorted(g1)                     # only members of one tie group change places
            assert g0 == sorted(g0) and g1 == sorted(g1, key=_tie_key)
            moved += g0 != g1
            j = k
    assert moved > 0                                            # the data do have ties (zero-similarity blocks)
