"""GPU tests of the engine features around the kernels: CUDA-graph replay, the item-average switch, the sharded
wrapper on one rank, per-kernel profiling, layout diagnostics and the device block cache."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import mrs_b200  # noqa: F401,E402
from mrs_b200 import engine as E  # noqa: E402
from mrs_b200 import sharded  # noqa: E402
from oracle import oracle as O  # noqa: E402


@pytest.fixture(scope="module")
def eng():
    import torch
    stream = torch.cuda.Stream()
    e = E.Engine(0, stream=stream.cuda_stream)
    e._keep = stream
    yield e
    e.close()


def test_graph_replay_matches_direct_launch(eng, ml100k):
    import torch
    R, T = eng.ratings(*ml100k["train"]), eng.ratings(*ml100k["test"])
    m = E.Model(eng, R)
    direct = m.mae(T, E.PRED_BASELINE)
    out = torch.zeros(2, dtype=torch.float64, device="cuda")

    def pass_():
        m.refit()
        m.mae_async(T, out.data_ptr())

    pass_(); eng.sync()
    g = eng.capture(pass_)
    for _ in range(3):
        out.zero_()
        g.launch()
        eng.sync()
        r = out.cpu().numpy()
        assert r[0] / r[1] == direct and r[1] == 20_000       # bit-reproducible across replays
    g.close()


def test_item_average_switch(eng, ml100k):
    R, T = eng.ratings(*ml100k["train"]), eng.ratings(*ml100k["test"])
    m = E.Model(eng, R)
    ref_dev = m.vector(E.ITEM_AVG_DEV)[0].copy()
    ref_avg = m.vector(E.ITEM_AVG)[0].copy()
    m.set_item_averages(False)
    m.refit(); eng.sync()
    assert np.array_equal(m.vector(E.ITEM_AVG_DEV)[0], ref_dev)   # the deviation pass is unchanged
    with pytest.raises(E.MrsError):
        m.vector(E.ITEM_AVG)
    with pytest.raises(E.MrsError):
        m.mae(T, E.PRED_ITEM)
    m.set_item_averages(True)
    m.refit(); eng.sync()
    assert np.array_equal(m.vector(E.ITEM_AVG)[0], ref_avg, equal_nan=True)


def test_sharded_wrapper_single_rank_and_graphs(eng, ml100k):
    R, T = eng.ratings(*ml100k["train"]), eng.ratings(*ml100k["test"])
    o = O.Oracle(*ml100k["train"])
    ref = o.mae(ml100k["test"], kind=O.BASELINE)
    sb = sharded.ShardedBaseline(eng, R, T)
    sb.fit()
    assert sb.mae() == pytest.approx(ref, rel=1e-6)
    sb.capture()
    sb.step(); eng.sync()
    r = sb.out2.cpu().numpy()
    assert r[0] / r[1] == pytest.approx(ref, rel=1e-6)
    # exchange buffer prefix = [dev sums | counts | global sum, count]
    x = sb.xbuf.cpu().numpy()
    n_items = R.n_items_dim
    assert x.size == sharded.mandatory_size(n_items)
    assert x[2 * n_items + 1] == 80_000 and x[2 * n_items] == ml100k["train"][2].sum()
    assert np.array_equal(x[n_items:2 * n_items], np.bincount(ml100k["train"][1], minlength=n_items))


def test_profile_labels_layout_info_and_block_cache(eng, ml100k):
    R, T = eng.ratings(*ml100k["train"]), eng.ratings(*ml100k["test"])
    m = E.Model(eng, R)
    m.mae(T, E.PRED_BASELINE)
    eng.profile_begin()
    m.refit()
    labels = [k for k, _ in eng.profile_end()]
    assert labels == ["user_sum", "item_tiled", "item_tiled_finalize"]
    info = R.layout_info()
    assert info["user_tiles"] == 1 and info["units"] > 0 and info["tiled_slots"] >= 80_000
    assert T.layout_info()["item_tiles"] == 1
    b = R.bytes()
    assert b["item_major"] == 4 * 80_000 and T.bytes()["sorted_coo"] == 8 * 20_000
    # released blocks are reused: rebuilding the same set must not grow device memory
    import torch
    free0 = torch.cuda.mem_get_info()[0]
    for _ in range(3):
        R2 = eng.ratings(*ml100k["train"]); m2 = E.Model(eng, R2); m2.close(); R2.close()
    assert torch.cuda.mem_get_info()[0] >= free0 - (64 << 20)
    before = E.launch_count()
    m.refit(); eng.sync()
    assert E.launch_count() - before == 3


def test_peer_exchange_single_rank_is_identity_and_graph_capturable(eng):
    """world = 1 exercises the whole kernel path (publish, flag, wait, rank-ordered reduce, device-side epoch)."""
    import torch
    x = E.PeerExchange(eng, 1001, 0, 1, lambda b: [b])
    with torch.cuda.stream(eng._keep):
        buf = torch.arange(1001, dtype=torch.float64, device="cuda") * 0.25
        ref = buf.clone()
        for _ in range(3):                      # alternating parity buffers
            x.allreduce_async(buf.data_ptr(), 1001)
        eng.sync()
        assert torch.equal(buf, ref) and not x.timed_out()
        small = torch.tensor([3.5, 7.0], dtype=torch.float64, device="cuda")
        g = eng.capture(lambda: (x.allreduce_async(buf.data_ptr(), 1001), x.allreduce_async(small.data_ptr(), 2)))
        for _ in range(4):
            g.launch()
        eng.sync()
        assert torch.equal(buf, ref) and small.tolist() == [3.5, 7.0] and not x.timed_out()
        g.close()
        # indexed form: only the listed slots take part (odd and even counts, one-shot sizes)
        for idx_list in ([5, 17, 900], [0, 2, 4, 1000]):
            idx = torch.tensor(idx_list, dtype=torch.int32, device="cuda")
            x.allreduce_indexed_async(buf.data_ptr(), idx.data_ptr(), idx.numel())
            eng.sync()
            assert torch.equal(buf, ref) and not x.timed_out()
    x.close()


def test_staged_upload_equals_direct_build(eng, ml100k):
    """mrs_upload_begin + mrs_ratings_from_upload (copies on the copy stream, sort started when the ids are in) build the
    same rating sets as mrs_ratings_from_coo; an upload that is never consumed can be dropped; bad input still fails loudly."""
    import torch
    tr, te = ml100k["train"], ml100k["test"]
    pin = [torch.from_numpy(np.ascontiguousarray(x)).pin_memory() for x in (*tr, *te)]
    hu, hi, hr, tu, ti, tv = [p.numpy() for p in pin]
    up_r, up_t = eng.upload(hu, hi, hr), eng.upload(tu, ti, tv)       # both in flight at once
    R, T = up_r.ratings(), up_t.ratings()
    R0, T0 = eng.ratings(*tr), eng.ratings(*te)
    assert (R.n, R.n_users_dim, R.n_items_dim, R.value_kind) == (R0.n, R0.n_users_dim, R0.n_items_dim, R0.value_kind)
    m, m0 = E.Model(eng, R), E.Model(eng, R0)
    assert m.mae(T) == m0.mae(T0) == pytest.approx(O.Oracle(*tr).mae(te, kind=O.BASELINE), rel=1e-6)
    assert np.array_equal(m.vector(E.USER_AVG)[0], m0.vector(E.USER_AVG)[0])
    eng.upload(hu, hi, hr).close()                                    # dropped without building
    bad = hu.copy(); bad[5] = -3
    with pytest.raises(E.MrsError):
        eng.upload(bad, hi, hr).ratings()
    fr = tr[2] + 0.25                                                 # not half-star codes: the fp64 value path
    Rf = eng.upload(tr[0], tr[1], fr).ratings()
    assert Rf.value_kind == 1
    for h in (m, m0, R, T, R0, T0, Rf):
        h.close()


def test_compact_code_upload_matches_fp64_upload(eng, ml100k):
    """mrs_ratings_from_coo_codes: (int32, int32, uint8 code = 2 x rating), 9 bytes per rating over PCIe instead of 16."""
    tr, te = ml100k["train"], ml100k["test"]
    R, T = eng.ratings(*tr), eng.ratings(*te)
    Rc = eng.ratings_from_codes(tr[0], tr[1], (tr[2] * 2).astype(np.uint8))
    Tc = eng.upload_codes(te[0], te[1], (te[2] * 2).astype(np.uint8)).ratings()
    m, mc = E.Model(eng, R), E.Model(eng, Rc)
    assert mc.mae(Tc, E.PRED_BASELINE) == m.mae(T, E.PRED_BASELINE)
    for kind in (E.USER_AVG, E.ITEM_AVG, E.ITEM_AVG_DEV):
        assert mc.vector(kind)[0].tolist() == m.vector(kind)[0].tolist()
    with pytest.raises(E.MrsError):
        eng.ratings_from_codes(tr[0][:10], tr[1][:10], np.full(10, 255, dtype=np.uint8))   # 255 is not a legal code
    for h in (mc, m, Tc, Rc, T, R):
        h.close()


def test_fit_mae_closure_equals_fit_then_mae(eng):
    """mrs_fit_mae_async: the test pass finishes the fit in its own prologue (no K2b); same MAE bits, same model arrays."""
    import torch
    from mrs_b200 import synth
    d = synth.ml25m(seed=8, n_users=30000, n_items=9000, n_ratings=1_200_000, max_item_id=30000)   # several item tiles and user tiles
    R, T = eng.ratings(*d["train"]), eng.ratings(*d["test"])
    m = E.Model(eng, R)
    m.set_item_averages(False)
    with torch.cuda.stream(eng._keep):
        out = torch.zeros(2, dtype=torch.float64, device="cuda")
        m.refit()
        m.mae_async(T, out.data_ptr())
        eng.sync()
        ref = out.cpu().numpy().copy()
        idev, ua, g = m.vector(E.ITEM_AVG_DEV)[0], m.vector(E.USER_AVG)[0], m.global_avg
        for _ in range(3):                              # repeated: the accumulators and the user sums are re-armed correctly
            out.zero_()
            m.fit_mae_async(T, out.data_ptr())
            eng.sync()
            assert out.cpu().numpy().tolist() == ref.tolist()
        assert m.vector(E.ITEM_AVG_DEV)[0].tolist() == idev.tolist() and m.vector(E.USER_AVG)[0].tolist() == ua.tolist() and m.global_avg == g
        m.refit(); m.mae_async(T, out.data_ptr()); eng.sync()    # and the four-kernel form still works afterwards
        assert out.cpu().numpy().tolist() == ref.tolist()
    assert ref[0] / ref[1] == pytest.approx(O.Oracle(*d["train"]).mae(d["test"], kind=O.BASELINE), rel=1e-6)
    for h in (m, T, R):
        h.close()
