"""Host logic of the sharded (N>1) baseline pass on CPU: user partition, exchange-buffer protocol and the MAE combine,
run as two real processes over `torch.distributed` with the gloo backend.  The rank-local partial sums that
`mrs_fit_local` writes on a GPU are produced here by a numpy statement of the same quantities (test helper only)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import mrs_b200  # noqa: F401
from mrs_b200 import sharded, synth


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _local_exchange(tr, n_users, n_items):
    """numpy statement of the buffer mrs_fit_local leaves on a rank (P:113, P:167, P:180-185 restricted to the shard)."""
    u, i, r = tr
    cnt_u = np.bincount(u, minlength=n_users).astype(np.float64)
    sum_u = np.bincount(u, weights=r, minlength=n_users)
    avg = np.where(cnt_u > 0, sum_u / np.where(cnt_u > 0, cnt_u, 1), -1.0)
    a = avg[u]
    sc = np.where(r > a, 5 - a, np.where(r < a, a - 1, 1.0))
    dev = (r - a) / sc
    buf = np.zeros(sharded.exchange_size(n_items))
    d, rs, c, _, _ = sharded.split_exchange(buf, n_items)
    d += np.bincount(i, weights=dev, minlength=n_items)
    rs += np.bincount(i, weights=r, minlength=n_items)
    c += np.bincount(i, minlength=n_items)
    buf[2 * n_items] = r.sum()
    buf[2 * n_items + 1] = r.size
    return buf, avg


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        d = synth.cached("ml100k")
        tr, te = d["train"], d["test"]
        n_users, n_items = int(tr[0].max()) + 1, int(max(tr[1].max(), te[1].max())) + 1
        bounds = sharded.partition_users(np.bincount(tr[0], minlength=n_users), world)
        mtr, mte = sharded.shard_of(tr[0], bounds, rank), sharded.shard_of(te[0], bounds, rank)
        ltr = tuple(x[mtr] for x in tr)
        lte = tuple(x[mte] for x in te)
        buf, avg = _local_exchange(ltr, n_users, n_items)
        t = torch.from_numpy(buf)
        sharded.all_reduce_sum(t)                                   # the one collective of the fit
        idev, iavg, gavg = sharded.finish_from_exchange(t.numpy(), n_items)
        # rank-local prediction + |err| for its own test pairs, then the 16-byte combine
        ua = avg[lte[0]]
        dv = idev[lte[1]]
        s = ua + dv
        sc = np.where(s > ua, 5 - ua, np.where(s < ua, ua - 1, 1.0))
        pred = np.where(ua < 0, gavg, ua + dv * sc)
        out2 = torch.tensor([np.abs(lte[2] - pred).sum(), float(lte[2].size)], dtype=torch.float64)
        sharded.all_reduce_sum(out2)
        q.put((rank, bounds, int(mtr.sum()), int(mte.sum()), float(out2[0] / out2[1]), float(gavg), idev[:50].tolist()))
    finally:
        dist.destroy_process_group()


def test_partition_is_contiguous_complete_and_balanced():
    rng = np.random.default_rng(0)
    counts = rng.integers(0, 200, size=1000)
    for world in (1, 2, 3, 8):
        b = sharded.partition_users(counts, world)
        assert b[0] == 0 and b[-1] == counts.size and all(x <= y for x, y in zip(b, b[1:]))
        loads = [counts[b[r]:b[r + 1]].sum() for r in range(world)]
        assert sum(loads) == counts.sum()
        assert max(loads) - min(loads) <= 2 * counts.max()
    assert sharded.partition_users(np.zeros(10, dtype=int), 4)[-1] == 10


def test_exchange_layout_helpers():
    n = 7
    buf = np.arange(sharded.exchange_size(n), dtype=np.float64)
    d, r, c, gs, gc = sharded.split_exchange(buf, n)
    assert d[0] == 0 and c[0] == 7 and gs == 14 and gc == 15 and r[0] == 16 and sharded.mandatory_size(n) == 16
    buf2 = np.zeros(sharded.exchange_size(2)); buf2[:] = [1.0, 0.0, 2.0, 0.0, 8.0, 2.0, 8.0, 0.0]
    idev, iavg, g = sharded.finish_from_exchange(buf2, 2)
    assert idev.tolist() == [0.5, 0.0] and iavg[0] == 4.0 and np.isnan(iavg[1]) and g == 4.0


@pytest.mark.timeout(180)
def test_two_rank_gloo_matches_single_process_oracle(ml100k):
    from oracle import oracle as O
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=150) for _ in range(world))
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    o = O.Oracle(*ml100k["train"])
    ref = o.mae(ml100k["test"], kind=O.BASELINE)
    assert res[0][2] + res[1][2] == 80_000 and res[0][3] + res[1][3] == 20_000      # shards are a partition
    for r in res:
        assert r[4] == pytest.approx(ref, rel=1e-9)                                  # same MAE on every rank
        assert r[5] == o.global_avg
        assert np.allclose(r[6], [o.item_avg_dev(i) for i in range(50)], rtol=1e-9, atol=1e-12)


def _knn_worker(rank, world, port, q):
    """Row-sharded kNN (BASELINE config 5): the train set is replicated, a rank evaluates the test pairs of its user range
    (the rank-local kNN pass is the oracle here -- on a GPU it is ShardedKnn.fit/mae_local) and the ranks add {sum, n}."""
    from oracle import oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        d = synth.cached("ml100k")
        tr, te = d["train"], tuple(x[:3000] for x in d["test"])
        bounds = sharded.partition_rows(np.bincount(tr[0], minlength=int(tr[0].max()) + 1), world)
        mte = sharded.shard_of(te[0], bounds, rank)
        lte = tuple(x[mte] for x in te)
        o = O.Oracle(*tr)
        out2 = torch.tensor([o.mae(lte, kind=O.PERSONALIZED, simkind=O.SIM_COSINE, k=30) * lte[0].size, float(lte[0].size)],
                            dtype=torch.float64)
        sharded.all_reduce_sum(out2)
        q.put((rank, bounds, int(mte.sum()), float(out2[0] / out2[1])))
    finally:
        dist.destroy_process_group()


def test_row_partition_balances_length_plus_constant():
    counts = np.array([0, 100, 1, 1, 1, 1, 0, 50, 50, 2])
    b = sharded.partition_rows(counts, 2)
    assert b[0] == 0 and b[-1] == counts.size and 0 < b[1] < counts.size
    w = counts + (counts > 0) * int(counts[counts > 0].mean())
    assert abs(w[:b[1]].sum() - w[b[1]:].sum()) <= 2 * w.max()


@pytest.mark.timeout(180)
def test_two_rank_gloo_row_sharded_knn(ml100k):
    from oracle import oracle as O
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_knn_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=150) for _ in range(world))
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    te = tuple(x[:3000] for x in ml100k["test"])
    ref = O.Oracle(*ml100k["train"]).mae(te, kind=O.PERSONALIZED, simkind=O.SIM_COSINE, k=30)
    assert res[0][1] == res[1][1] and res[0][2] + res[1][2] == 3000
    for r in res:
        assert r[3] == pytest.approx(ref, rel=1e-9)
