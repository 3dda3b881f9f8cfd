"""GPU parity of the personalized / kNN path against the CPU oracle: similarities bit-exact, neighbour index
sets (and order) exact, predictions and MAE within 1e-6 relative; plus the invariants of SURVEY A.9."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import mrs_b200  # noqa: F401,E402
from mrs_b200 import engine as E  # noqa: E402
from mrs_b200 import predictions as P  # noqa: E402
from oracle import oracle as O  # noqa: E402

REL = 1e-6
KS = [10, 30, 50, 100, 200, 300, 400, 800, 943]  # predict/kNN.scala:73


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


def test_deviation_and_preprocessed_maps_bit_exact(fitted):
    R, T, m, o, tr, te = fitted
    s = m.similarity(E.SIM_COSINE, 0)
    u, i, dev = s.entry_values(0)
    _, _, pre = s.entry_values(1)
    rng = np.random.default_rng(1)
    for j in rng.integers(0, u.size, 400):
        d, p = o.pair_values(int(u[j]), int(i[j]))
        assert dev[j] == d and pre[j] == p


def test_cosine_similarities_bit_exact(fitted):
    R, T, m, o, tr, te = fitted
    s = m.similarity(E.SIM_COSINE, 0)
    rng = np.random.default_rng(2)
    for u, v in rng.integers(1, 944, size=(600, 2)):
        assert s(int(u), int(v)) == o.cosine(int(u), int(v))
    for u in (1, 2, 300, 943):
        assert abs(s(u, u) - 1.0) < 1e-12                                # A.9 (4)
        assert s(u, 7) == s(7, u)                                         # A.9 (5)
    assert s(1, 5000) == 0.0 and s(5000, 1) == 0.0


@pytest.mark.parametrize("k", [10, 300, 942, 2000])
def test_neighbour_lists_exact(fitted, k):
    R, T, m, o, tr, te = fitted
    s = m.similarity(E.SIM_COSINE, k)
    for u in list(range(1, 944, 7)) + [943]:
        ids, sims = s.neighbors(u, k)
        oi, os_ = o.neighbors(u, k)
        assert ids.tolist() == oi.tolist()                               # index sets AND order
        assert sims.tolist() == os_.tolist()


def test_knn_similarity_semantics(fitted):
    R, T, m, o, tr, te = fitted
    s = m.similarity(E.SIM_COSINE, 10)
    assert s(1, 1) == 0.0                                                 # knn-100k.json:8 analogue (A.5)
    ids, sims = s.neighbors(1, 10)
    assert s(1, int(ids[0])) == sims[0] == o.similarity(1, int(ids[0]), k=10)
    far = int(o.neighbors(1, 942)[0][-1])
    assert s(1, far) == 0.0 and o.similarity(1, far, k=10) == 0.0
    s.set_k(0)
    assert s(1, 1) == o.cosine(1, 1)


def test_knn_mae_sweep_and_predictions(fitted):
    R, T, m, o, tr, te = fitted
    s = m.similarity(E.SIM_COSINE, 300)
    n = 2500
    for k in KS:
        s.set_k(k)
        got = m.mae(T, E.PRED_PERSONALIZED, s)
        assert got == pytest.approx(o.mae(te, kind=O.PERSONALIZED, simkind=O.SIM_COSINE, k=k), rel=REL)
        p = m.predict(te[0][:n], te[1][:n], E.PRED_PERSONALIZED, s)
        ref = o.predict_batch(te[0][:n], te[1][:n], kind=O.PERSONALIZED, simkind=O.SIM_COSINE, k=k)
        assert np.allclose(p, ref, rtol=REL, atol=0)
    # k >= U-1 == plain cosine on test pairs (A.9 (2)); knn-100k.json:47-49 == personalized-100k.json:14
    s.set_k(943)
    a = m.predict(te[0], te[1], E.PRED_PERSONALIZED, s)
    s.set_k(0)
    b = m.predict(te[0], te[1], E.PRED_PERSONALIZED, s)
    assert np.allclose(a, b, rtol=1e-12, atol=0)
    assert m.mae(T, E.PRED_PERSONALIZED, s) == pytest.approx(o.mae(te, kind=O.PERSONALIZED, simkind=O.SIM_COSINE), rel=REL)


def test_uniform_equals_baseline_and_jaccard(fitted):
    R, T, m, o, tr, te = fitted
    uni = m.similarity(E.SIM_UNIFORM, 0)
    a = m.predict(te[0], te[1], E.PRED_PERSONALIZED, uni)
    b = m.predict(te[0], te[1], E.PRED_BASELINE)
    assert np.allclose(a, b, rtol=1e-12, atol=0)                          # A.9 (1); personalized-100k.json:8-9
    assert m.mae(T, E.PRED_PERSONALIZED, uni) == pytest.approx(m.mae(T, E.PRED_BASELINE), rel=1e-12)
    jac = m.similarity(E.SIM_JACCARD, 0)
    rng = np.random.default_rng(5)
    for u, v in rng.integers(1, 944, size=(200, 2)):
        assert jac(int(u), int(v)) == o.jaccard(int(u), int(v))
    assert m.mae(T, E.PRED_PERSONALIZED, jac) == pytest.approx(o.mae(te, kind=O.PERSONALIZED, simkind=O.SIM_JACCARD), rel=REL)


def test_wsd_and_fallbacks(fitted):
    R, T, m, o, tr, te = fitted
    s = m.similarity(E.SIM_COSINE, 30)
    us, is_ = te[0][:500], te[1][:500]
    w = m.predict(us, is_, E.PRED_WSD, s)
    ref = np.array([o.wsd(int(u), int(i), k=30) for u, i in zip(us, is_)])
    assert np.allclose(w, ref, rtol=REL, atol=1e-15)
    p = m.predict([5000, 1, 5000], [1, 99999, 99999], E.PRED_PERSONALIZED, s)
    assert p[0] == o.global_avg and p[1] == o.user_avg(1) and p[2] == o.global_avg


def test_recommendations(fitted, eng, ml100k):
    R, T, m, o, tr, te = fitted
    s = m.similarity(E.SIM_COSINE, 300)
    for user in (1, 944 - 1, 400):
        items, scores = m.recommend(user, 5, E.PRED_PERSONALIZED, s)
        oi, os_ = o.recommend(user, 5, k=300)
        assert items.tolist() == oi.tolist()
        assert np.allclose(scores, os_, rtol=REL, atol=0)
    items, _ = m.recommend(1, 3, E.PRED_BASELINE)
    assert items.tolist() == o.recommend(1, 3, kind=O.BASELINE, simkind=O.SIM_UNIFORM, k=0)[0].tolist()


def test_mirror_knn_reads_like_the_reference(eng, ml100k):
    P.set_default_engine(eng)
    train = P.RatingSet.from_arrays(*ml100k["train"])
    test = P.RatingSet.from_arrays(*ml100k["test"])
    o = O.Oracle(*ml100k["train"])
    te = ml100k["test"]
    # predict/kNN.scala:43-44
    mae = P.MAE(P.predictor(train, P.weightedSumDeviation(train, P.getSimilarity(train, 300, P.adjustedCosineSimilarityFunction(train)))), test)
    assert mae == pytest.approx(o.mae(te, kind=O.PERSONALIZED, simkind=O.SIM_COSINE, k=300), rel=REL)
    # predict/kNN.scala:66-70
    sim10 = P.getSimilarity(train, 10, P.adjustedCosineSimilarityFunction(train))
    assert sim10(1, 1) == 0.0
    nn = P.getNeighbors(train, 10, P.adjustedCosineSimilarityFunction(train))(1)
    assert [x[0] for x in nn] == o.neighbors(1, 10)[0].tolist()
    assert sim10(1, nn[0][0]) == nn[0][1]
    p11 = P.predictor(train, P.weightedSumDeviation(train, sim10))(1, 1)
    assert p11 == pytest.approx(o.predict(1, 1, kind=O.PERSONALIZED, simkind=O.SIM_COSINE, k=10), rel=REL)
    # predict/Personalized.scala:61-67
    ones = P.predictor(train, P.weightedSumDeviation(train, P.similarityOne))
    assert P.MAE(ones, test) == pytest.approx(o.mae(te, kind=O.BASELINE), rel=REL)
    cos = P.adjustedCosineSimilarityFunction(train)
    assert cos(1, 2) == o.cosine(1, 2)
    # recommend/Recommender.scala:82-88
    rec = P.recommendations(train, P.predictor(train, P.weightedSumDeviation(train, P.getSimilarity(train, 300, cos))))(1, 3)
    assert [x[0] for x in rec] == o.recommend(1, 3, k=300)[0].tolist()
    with pytest.raises(P.UnsupportedOperationError):
        P.weightedSumDeviation(train, lambda u, v: 0.5)


def test_answer_documents_have_the_reference_schema(eng, ml100k, tmp_path):
    """the JSON documents of the five mains, key for key (values checked against the oracle)."""
    import json
    from mrs_b200 import answers
    P.set_default_engine(eng)
    train = P.RatingSet.from_arrays(*ml100k["train"])
    test = P.RatingSet.from_arrays(*ml100k["test"])
    o = O.Oracle(*ml100k["train"])
    te = ml100k["test"]
    b = answers.baseline(train, test, num_measurements=1)
    assert list(b) == ["Meta", "B.1", "B.2", "B.3"]
    assert list(b["B.1"]) == ["1.GlobalAvg", "2.User1Avg", "3.Item1Avg", "4.Item1AvgDev", "5.PredUser1Item1"]
    assert b["B.1"]["1.GlobalAvg"] == o.global_avg and b["B.1"]["2.User1Avg"] == o.user_avg(1) and b["B.1"]["3.Item1Avg"] == o.item_avg(1)
    assert b["B.2"]["4.BaselineMAE"] == pytest.approx(o.mae(te, kind=O.BASELINE), rel=REL)
    assert set(b["B.3"]["4.Baseline"]) == {"average (ms)", "stddev (ms)"}
    d = answers.distributed(train, test, num_measurements=1)
    assert list(d["D.1"]) == ["1.GlobalAvg", "2.User1Avg", "3.Item1Avg", "4.Item1AvgDev", "5.PredUser1Item1", "6.Mae"]
    assert d["D.1"]["6.Mae"] == pytest.approx(b["B.2"]["4.BaselineMAE"], rel=1e-12)          # A.9 (7)
    p = answers.personalized(train, test)
    assert p["P.1"]["2.OnesMAE"] == pytest.approx(b["B.2"]["4.BaselineMAE"], rel=1e-9)        # A.9 (1)
    assert p["P.2"]["1.AdjustedCosineUser1User2"] == o.cosine(2, 1)
    assert p["P.3"]["1.JaccardUser1User2"] == o.jaccard(1, 2)
    k = answers.knn(train, test, num_measurements=1, sweep=[10, 943])
    assert k["N.1"]["1.k10u1v1"] == 0.0
    assert k["N.2"]["1.kNN-Mae"][1][1] == pytest.approx(p["P.2"]["3.AdjustedCosineMAE"], rel=1e-9)   # k=943 == cosine, A.9 (2)
    personal = tmp_path / "personal.csv"
    personal.write_bytes(b"id,title,rating\r\n1,Toy Story (1995),5\r\n2,GoldenEye (1995),\r\n3,Four Rooms (1995),2\r\n")
    r = answers.recommender(ml100k["all"], str(personal))
    assert list(r) == ["Meta", "R.1", "R.2"] and len(r["R.2"]) == 3 and all(len(x) == 3 for x in r["R.2"])
    assert all(x[0] not in (1, 3) for x in r["R.2"])                                          # rated items are never recommended
    json.dumps([b, d, p, k, r])


@pytest.mark.parametrize("n_users", [300, 1500, 3000])
def test_sort_kernels_at_other_sizes(eng, n_users):
    """The neighbour sort has a register/shuffle form for up to 2,048 users (512-, 1,024-, 2,048-element networks) and a
    shared-memory form above: every one must give the oracle's lists."""
    from mrs_b200 import synth
    d = synth.small(seed=n_users, n_users=n_users, n_items=400, n_ratings=n_users * 25)
    tr = d["train"]
    R = eng.ratings(*tr)
    m = E.Model(eng, R)
    o = O.Oracle(*tr)
    s = m.similarity(E.SIM_COSINE, 0)
    users = np.unique(tr[0])
    for u in users[:: max(1, users.size // 25)]:
        k = int(users.size)
        ids, sims = s.neighbors(int(u), k)
        oi, os_ = o.neighbors(int(u), k)
        assert ids.tolist() == oi.tolist() and sims.tolist() == os_.tolist()
    for h in (s, m, R):
        h.close()


def test_tie_order_matches_the_oracle_in_both_modes(small):
    """mrs_model_set_tie_order: neighbours with exactly equal similarity in user-id order (default) or in the Scala 2.11
    HashSet iteration order that the reference's stable sort keeps; dense path and row-block path."""
    tr = small["train"]
    eng = E.Engine(0)
    R = eng.ratings(*tr)
    m = E.Model(eng, R)
    o = O.Oracle(*tr)
    n = int(np.unique(tr[0]).size) - 1
    users = [int(u) for u in np.unique(tr[0])[:10]]
    differ = 0
    for mode in (1, 0):
        m.set_tie_order(mode)
        o.set_tie_order(mode)
        s = m.similarity(E.SIM_COSINE, n)
        sr = m.similarity(E.SIM_COSINE, n, rows=(0, 2**31 - 1))
        for u in users:
            oi, os_ = o.neighbors(u, n)
            for h in (s, sr):
                ids, sims = h.neighbors(u, n)
                assert ids.tolist() == oi.tolist() and sims.tolist() == os_.tolist()
            if mode == 1:
                o.set_tie_order(0)
                differ += o.neighbors(u, n)[0].tolist() != oi.tolist()
                o.set_tie_order(1)
        sr.close(); s.close()
    assert differ > 0
    o.set_tie_order(0)
    m.close(); R.close(); eng.close()


@pytest.mark.parametrize("kind", ["cosine", "jaccard"])
def test_similarity_kernel_with_several_item_phases(eng, kind):
    """The similarity kernel keeps 16 row users x all items in shared memory; above 1,746 items the items are cut into
    phases and the partial sums wait in the matrix between two phases.  5,000 items = three phases: same bits as the oracle,
    both triangles (only one is computed, the other one is its copy)."""
    from mrs_b200 import synth
    d = synth.small(seed=77, n_users=150, n_items=5000, n_ratings=9000)
    tr = d["train"]
    R = eng.ratings(*tr)
    m = E.Model(eng, R)
    o = O.Oracle(*tr)
    simkind = E.SIM_COSINE if kind == "cosine" else E.SIM_JACCARD
    s = m.similarity(simkind, 0)
    users = [int(u) for u in np.unique(tr[0])]
    rng = np.random.default_rng(5)
    for _ in range(400):
        u, v = (int(x) for x in rng.choice(users, 2))
        want = o.cosine(u, v) if kind == "cosine" else o.jaccard(u, v)
        assert s(u, v) == want and s(v, u) == s(u, v)
    if kind == "cosine":
        for u in users[::15]:
            ids, sims = s.neighbors(u, len(users))
            oi, os_ = o.neighbors(u, len(users))
            assert ids.tolist() == oi.tolist() and sims.tolist() == os_.tolist()
    for h in (s, m, R):
        h.close()


def test_users_only_refit_is_the_fit_of_the_knn_closure(eng, ml100k):
    """mrs_fit_users_async: user averages and global average bit-identical to the full fit, item deviations untouched, and the
    kNN MAE after a users-only refit equal to the one after a full refit (P:557-586 never reads the item deviations)."""
    tr, te = ml100k["train"], ml100k["test"]
    R, T = eng.ratings(*tr), eng.ratings(*te)
    m = E.Model(eng, R)
    ua, idev, g = m.vector(E.USER_AVG)[0].copy(), m.vector(E.ITEM_AVG_DEV)[0].copy(), m.global_avg
    s = m.similarity(E.SIM_COSINE, 30)
    want = m.mae(T, E.PRED_PERSONALIZED, s)
    for _ in range(3):          # K1's sums are re-armed by the users-only kernel: repeated refits stay exact
        m.refit_users()
    s.refit(30)
    eng.sync()
    assert np.array_equal(m.vector(E.USER_AVG)[0], ua) and m.global_avg == g
    assert np.array_equal(m.vector(E.ITEM_AVG_DEV)[0], idev)
    assert m.mae(T, E.PRED_PERSONALIZED, s) == want
    m.refit(); eng.sync()        # and a full fit after it finds its accumulators armed
    assert np.array_equal(m.vector(E.ITEM_AVG_DEV)[0], idev) and np.array_equal(m.vector(E.USER_AVG)[0], ua)
    assert m.mae(T, E.PRED_BASELINE) == pytest.approx(O.Oracle(*tr).mae(te, kind=O.BASELINE), rel=1e-6)
    for h in (s, m, T, R):
        h.close()
