#!/usr/bin/env python
"""End-to-end step (compact host form -> layouts -> fit -> MAE -> D2H) timed in both upload orders, interleaved:
train set first (the sort of the train ids overlaps the rest of the copies) against test set first (its layouts are
built while the train ids travel)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import mrs_b200
from mrs_b200 import engine as E, synth

d = synth.cached("ml25m")
eng = E.Engine(0)
def pin(a):
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory(); return t.numpy(), t
hu, ku = pin(d["train"][0]); hi, ki = pin(d["train"][1]); hc, kc = pin((d["train"][2] * 2).astype(np.uint8))
tu, kt = pin(d["test"][0]); ti, kj = pin(d["test"][1]); tc, kd = pin((d["test"][2] * 2).astype(np.uint8))
out = torch.zeros(2, dtype=torch.float64, device="cuda")

def step(test_first):
    if test_first:
        ut = eng.upload_codes(tu, ti, tc); ur = eng.upload_codes(hu, hi, hc)
        T = ut.ratings(); R = ur.ratings(); m = E.Model(eng, R, sync=False)
    else:
        ur = eng.upload_codes(hu, hi, hc); ut = eng.upload_codes(tu, ti, tc)
        R = ur.ratings(); m = E.Model(eng, R, sync=False); T = ut.ratings()
    m.mae_async(T, out.data_ptr()); r = out.cpu().numpy()
    m.close(); T.close(); R.close()
    return r[0] / r[1]

for tf in (False, True):
    step(tf); step(tf)
for rnd in range(3):
    for tf in (False, True):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5):
            mae = step(tf)
        torch.cuda.synchronize()
        print(f"round {rnd} {'test first ' if tf else 'train first'}: {1e3 * (time.perf_counter() - t0) / 5:.3f} ms/step  mae={mae:.9f}")
