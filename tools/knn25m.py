#!/usr/bin/env python
"""kNN (k = 300) at ml-25m shape on ONE GPU over a range of the rows (BASELINE config 5, the share of one rank):
    python tools/knn25m.py [--parts 8] [--part 0] [--k 300] [--check 3] [--scale 1.0]
Prints per-kernel time sums (CUDA events) and, for --check users, compares the neighbour lists with the CPU oracle."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import mrs_b200  # noqa: F401,E402
from mrs_b200 import engine as E, sharded, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--parts", type=int, default=8)
ap.add_argument("--part", type=int, default=0)
ap.add_argument("--k", type=int, default=300)
ap.add_argument("--check", type=int, default=3)
ap.add_argument("--kind", default="cosine")
args = ap.parse_args()

d = synth.cached("ml25m")
tr, te = d["train"], d["test"]
n_users_dim = int(max(tr[0].max(), te[0].max())) + 1
bounds = sharded.partition_rows(np.bincount(tr[0], minlength=n_users_dim), args.parts)
lo, hi = int(bounds[args.part]), int(bounds[args.part + 1])
eng = E.Engine(0)
R = eng.ratings(*tr)
m = E.Model(eng, R)
sel = (te[0] >= lo) & (te[0] < hi)
T = eng.ratings(te[0][sel], te[1][sel], te[2][sel])
kind = E.SIM_COSINE if args.kind == "cosine" else E.SIM_JACCARD
out = {"users": [lo, hi], "test_pairs": int(sel.sum()), "k": args.k}
for rep in range(2):
    eng.profile_begin()
    t0 = time.perf_counter()
    if rep == 0:
        s = m.similarity(kind, args.k, rows=(lo, hi))
    else:
        s.refit()
        eng.sync()
    t1 = time.perf_counter()
    mae = m.mae(T, E.PRED_PERSONALIZED, s)
    t2 = time.perf_counter()
    prof = {}
    for name, ms in eng.profile_end():
        prof[name] = prof.get(name, 0.0) + ms
    out[f"rep{rep}"] = {"fit_similarity_s": t1 - t0, "mae_s": t2 - t1, "mae": mae, "kernel_ms": {a: round(b, 3) for a, b in prof.items()}}
    print(json.dumps(out[f"rep{rep}"]), flush=True)
if args.check:
    from oracle import oracle as O
    t0 = time.perf_counter()
    o = O.Oracle(*tr)
    okind = O.SIM_COSINE if args.kind == "cosine" else O.SIM_JACCARD
    rng = np.random.default_rng(1)
    users = np.unique(tr[0][(tr[0] >= lo) & (tr[0] < hi)])
    ok = True
    for u in rng.choice(users, args.check, replace=False):
        ids, sims = s.neighbors(int(u), args.k)
        oi, os_ = o.neighbors(int(u), args.k, okind)
        same = ids.tolist() == oi.tolist() and sims.tolist() == os_.tolist()
        ok &= same
        print("user", int(u), "neighbour list identical to the oracle:", same, flush=True)
    out["oracle_check"] = {"users": args.check, "identical": bool(ok), "seconds": time.perf_counter() - t0}
print(json.dumps(out))
