"""GPU parity of the USER-SHARDED baseline pass (BASELINE config 4; distributed/DistributedBaseline.scala:30-47, the
reduceByKey/collect of P:267-268 and the sum/count of P:247 being the cross-rank exchange).

Every rank's data are distinct, so a pass whose exchange does nothing cannot reproduce the oracle:
  * one process, the ranks one after the other on cuda:0, the exchange done by the test on the host (sum of the
    exchange buffers in rank order): checks partition_users / shard_of with global table sizes, mrs_fit_local,
    mrs_fit_finish and the rank-local MAE against the single-GPU fit and the oracle;
  * two processes on two GPUs (skipped on a one-GPU box), through tools/sharded_check.py: the library's own NVLink
    peer-memory exchange kernel (and NCCL) on a strong-scaling and a weak-scaling workload.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import mrs_b200  # noqa: F401,E402
from mrs_b200 import engine as E, sharded, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REL = 1e-6


def close(a, b, rel=REL, floor=1e-12):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return bool(np.all(np.abs(a - b) <= rel * np.maximum(np.abs(b), floor)))


@pytest.fixture(scope="module")
def eng():
    import torch
    stream = torch.cuda.Stream()
    e = E.Engine(0, stream=stream.cuda_stream)
    e._keep = stream
    yield e
    e.close()


def _sequential_ranks(eng, shards, nu, ni):
    """fit_local on every shard, host-side sum of the exchange buffers in rank order, fit_finish, rank-local MAE."""
    import torch
    with torch.cuda.stream(eng._keep):                     # torch's copies and sums on the engine's stream
        models, sets = [], []
        for tr, te in shards:
            R, T = eng.ratings(*tr, nu, ni), eng.ratings(*te, nu, ni)
            sb = sharded.ShardedBaseline(eng, R, T)          # no process group: the exchange is done below
            sb.fit_local()
            models.append(sb); sets.append((R, T))
        eng.sync()
        total = torch.zeros_like(models[0].xbuf)
        local = []
        for sb in models:                                      # rank order
            local.append(sb.xbuf.clone())
            total += sb.xbuf
        err = cnt = 0.0
        for sb in models:
            sb.xbuf.copy_(total)                               # what the all-reduce leaves on every rank
            sb.fit_finish()
            sb.mae_local()
            eng.sync()
            r = sb.out2.cpu().numpy()
            err += r[0]; cnt += r[1]
    return models, sets, local, err / cnt


@pytest.mark.parametrize("world", [2, 5])
def test_user_partitioned_set_matches_single_fit_and_oracle(eng, world):
    d = synth.ml25m(seed=3, n_users=20000, n_items=4000, n_ratings=800_000, max_item_id=15000)
    tr, te = d["train"], d["test"]
    nu = int(max(tr[0].max(), te[0].max())) + 1
    ni = int(max(tr[1].max(), te[1].max())) + 1
    bounds = sharded.partition_users(np.bincount(tr[0], minlength=nu), world)
    shards = []
    for r in range(world):
        a, b = sharded.shard_of(tr[0], bounds, r), sharded.shard_of(te[0], bounds, r)
        shards.append((tuple(x[a] for x in tr), tuple(x[b] for x in te)))
    assert sum(s[0][0].size for s in shards) == tr[0].size and sum(s[1][0].size for s in shards) == te[0].size
    models, sets, local, mae = _sequential_ranks(eng, shards, nu, ni)
    o = O.Oracle(*tr)
    assert mae == pytest.approx(o.mae(te, kind=O.BASELINE), rel=REL)
    R1 = eng.ratings(*tr, nu, ni)
    m1 = E.Model(eng, R1)
    idev1, cnt1 = m1.vector(E.ITEM_AVG_DEV)
    ua1, uc1 = m1.vector(E.USER_AVG)
    items = np.flatnonzero(cnt1 > 0)
    assert close(idev1[items], [o.item_avg_dev(int(i)) for i in items])
    for r, sb in enumerate(models):
        idev, cnt = sb.model.vector(E.ITEM_AVG_DEV)
        assert cnt.tolist() == cnt1.tolist()               # global counts after the exchange
        assert close(idev, idev1)
        assert sb.model.global_avg == m1.global_avg         # exact: integer code sums
        ua, uc = sb.model.vector(E.USER_AVG)
        mine = np.arange(bounds[r], bounds[r + 1])
        known = mine[uc1[mine] > 0]
        assert ua[known].tolist() == ua1[known].tolist()    # a rank's own users: bit-identical averages
        other = np.setdiff1d(np.flatnonzero(uc1 > 0), mine)
        assert (uc[other] == 0).all()                       # users of other ranks are unknown here
    # the check bites: without the exchange a rank's item deviations are wrong for most items
    n_items = ni
    lone = sharded.finish_from_exchange(np.concatenate([local[0].cpu().numpy(), np.zeros(n_items)]), n_items)[0]
    assert (np.abs(lone[items] - idev1[items]) > 1e-6 * np.maximum(np.abs(idev1[items]), 1e-12)).mean() > 0.5
    for sb, (R, T) in zip(models, sets):
        sb.close(); T.close(); R.close()
    m1.close(); R1.close()


def test_weak_shards_match_oracle_on_the_union(eng):
    d = synth.ml25m(seed=4, n_users=12000, n_items=3000, n_ratings=500_000, max_item_id=9000)
    world = 3
    parts = [synth.weak_shard(d, r) for r in range(world)]
    nu, ni = world * parts[0]["user_stride"] + 1, parts[0]["max_item_id"] + 1
    models, sets, local, mae = _sequential_ranks(eng, [(p["train"], p["test"]) for p in parts], nu, ni)
    u = synth.weak_union(d, world)
    o = O.Oracle(*u["train"])
    assert mae == pytest.approx(o.mae(u["test"], kind=O.BASELINE), rel=REL)
    items = np.unique(u["train"][1])
    ref = [o.item_avg_dev(int(i)) for i in items]
    for sb in models:
        assert close(sb.model.vector(E.ITEM_AVG_DEV)[0][items], ref)
        assert sb.model.global_avg == o.global_avg
    for sb, (R, T) in zip(models, sets):
        sb.close(); T.close(); R.close()


def _n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("flag", [[], ["--nccl"], ["--fused"], ["--fused", "--closure"]])
def test_two_processes_two_gpus(flag):
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", {"": "29615", "--nccl": "29617", "--fused": "29619", "--fused--closure": "29621"}["".join(flag)], os.path.join(ROOT, "tools", "sharded_check.py")] + flag
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-3000:]
    out = json.loads([ln for ln in p.stdout.splitlines() if ln.startswith("{")][-1])
    for tag in ("strong", "weak"):
        r = out[tag]
        assert r["mae_rel_err"] <= REL and r["item_dev_worst_rel"] <= REL and r["global_avg_equal"], (tag, r)
        assert r["mae"] == r["mae_eager"]                      # graph replay == eager launches, bit for bit
        assert r["items_wrong_without_exchange"] > r["items"] // 2   # the exchange is doing the work
        assert not r["timed_out"]


@pytest.mark.parametrize("n_dev", [1, 2])
def test_one_process_drives_the_devices(n_dev):
    """mrs_multi_*: ONE process (what a JVM is) owns the devices; peers are plain pointers after cudaDeviceEnablePeerAccess."""
    if _n_gpus() < n_dev:
        pytest.skip(f"needs {n_dev} GPUs")
    d = synth.ml25m(seed=6, n_users=15000, n_items=3000, n_ratings=600_000, max_item_id=9000)
    tr, te = d["train"], d["test"]
    me = E.MultiEngine(list(range(n_dev)))
    me.load(tr, te)
    o = O.Oracle(*tr)
    ref = o.mae(te, kind=O.BASELINE)
    assert me.baseline_mae() == pytest.approx(ref, rel=REL)
    assert me.baseline_mae() == pytest.approx(ref, rel=REL)            # a second pass on the same buffers
    items = np.unique(tr[1])
    oid = [o.item_avg_dev(int(i)) for i in items]
    for slot in range(n_dev):
        m = me.model(slot)
        assert close(m.vector(E.ITEM_AVG_DEV)[0][items], oid)
        assert m.global_avg == o.global_avg
    u = int(tr[0][0])
    assert me.model(me.owner(u)).lookup(E.USER_AVG, u)[0] == o.user_avg(u)
    me.close()
