#!/usr/bin/env python
"""Multi-rank parity check of the user-sharded baseline pass (run under torch.distributed.run, one rank per GPU).

Two workloads, both with DISTINCT data per rank so that a dead exchange cannot pass:
  strong: ONE set, users partitioned over the ranks (sharded.partition_users / shard_of, global table sizes);
  weak:   one shard per rank (synth.weak_shard), the union is the reference set.
For each: fit_local -> exchange (own peer-memory kernel, or NCCL with --nccl) -> fit_finish -> MAE -> 16-byte exchange;
rank 0 compares the MAE with the CPU oracle and the per-item average deviations with the oracle's, and prints one JSON
line.  Used by tests/test_gpu_sharded.py (needs >= 2 GPUs) and by hand:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/sharded_check.py
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch
    import torch.distributed as dist
    import mrs_b200  # noqa: F401
    from mrs_b200 import engine as E, sharded, synth

    ap = argparse.ArgumentParser()
    ap.add_argument("--nccl", action="store_true")
    ap.add_argument("--fused", action="store_true", help="exchanges fused into the pass' kernels (mrs_fit_local_push ...)")
    ap.add_argument("--closure", action="store_true", help="with --fused: the whole step is mrs_fit_mae_push_async (three kernels)")
    ap.add_argument("--users", type=int, default=30000)
    ap.add_argument("--items", type=int, default=6000)
    ap.add_argument("--ratings", type=int, default=1_500_000)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    eng = E.Engine(local, stream=stream.cuda_stream)
    d = synth.ml25m(seed=5, n_users=args.users, n_items=args.items, n_ratings=args.ratings, max_item_id=4 * args.items)
    out = {"world": world, "exchange": "nccl" if args.nccl else (("closure" if args.closure else "fused") if args.fused else "peer")}

    def run(tr, te, nu, ni, ref_tr, ref_te, tag):
        with torch.cuda.stream(stream):
            R, T = eng.ratings(*tr, nu, ni), eng.ratings(*te, nu, ni)
            sb = sharded.ShardedBaseline(eng, R, T, peer_exchange=not args.nccl, fused=args.fused, closure=args.closure)
            assert sb.fused == (args.fused and not args.nccl) and sb.closure == (sb.fused and args.closure)
            sb.step()
            mae_eager = sb.result()
            sb.capture()
            for _ in range(3):
                sb.step()
            mae = sb.result()
            idev = sb.model.vector(E.ITEM_AVG_DEV)[0]
            gavg = sb.model.global_avg
            # what this rank would get WITHOUT the exchange (local sums only): must differ, or the check is vacuous
            was_fused, sb.fused = sb.fused, False
            sb.fit_local(); sb.fit_finish()
            sb.fused = was_fused
            idev_local = sb.model.vector(E.ITEM_AVG_DEV)[0]
            if sb.closure:                                 # ... and the closure still works on the model after that detour
                sb.step()
                assert sb.result() == mae, (sb.result(), mae)
            torch.cuda.synchronize(dev)
        if rank == 0:
            from oracle import oracle as O
            o = O.Oracle(*ref_tr)
            ref = o.mae(ref_te, kind=O.BASELINE)
            items = np.unique(ref_tr[1])
            oid = np.array([o.item_avg_dev(int(i)) for i in items])
            err = np.abs(idev[items] - oid) / np.maximum(np.abs(oid), 1e-12)
            err_local = np.abs(idev_local[items] - oid) / np.maximum(np.abs(oid), 1e-12)
            out[tag] = {"mae": mae, "mae_eager": mae_eager, "oracle_mae": ref, "mae_rel_err": abs(mae - ref) / abs(ref),
                        "global_avg_equal": gavg == o.global_avg, "item_dev_worst_rel": float(err.max()),
                        "items_wrong_without_exchange": int((err_local > 1e-6).sum()), "items": int(items.size),
                        "timed_out": None if sb.peer is None else sb.peer.timed_out()}
        torch.cuda.synchronize(dev)
        dist.barrier()
        for g in ("_g_all", "_g1", "_g2"):
            if getattr(sb, g, None) is not None:
                getattr(sb, g).close()
        sb.close(close_peer=True)
        T.close(); R.close()

    # strong: one set, users partitioned
    tr, te = d["train"], d["test"]
    nu = int(max(tr[0].max(), te[0].max())) + 1
    ni = int(max(tr[1].max(), te[1].max())) + 1
    bounds = sharded.partition_users(np.bincount(tr[0], minlength=nu), world)
    a, b = sharded.shard_of(tr[0], bounds, rank), sharded.shard_of(te[0], bounds, rank)
    run(tuple(x[a] for x in tr), tuple(x[b] for x in te), nu, ni, tr, te, "strong")
    # weak: a distinct shard per rank
    s = synth.weak_shard(d, rank)
    u = synth.weak_union(d, world)
    run(s["train"], s["test"], world * s["user_stride"] + 1, s["max_item_id"] + 1, u["train"], u["test"], "weak")
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
