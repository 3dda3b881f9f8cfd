#!/usr/bin/env python
"""Three end-to-end steps (compact host form -> layouts -> fit -> MAE), nothing else: the short command for an ncu launch list."""
import os, sys
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
for rep in range(3):
    ur = eng.upload_codes(hu, hi, hc); ut = eng.upload_codes(tu, ti, tc)
    R = ur.ratings(); m = E.Model(eng, R, sync=False); T = ut.ratings()
    m.mae_async(T, out.data_ptr()); r = out.cpu().numpy()
    print("mae", r[0] / r[1])
    m.close(); T.close(); R.close()
