#!/usr/bin/env python
"""Wall-clock phases of the end-to-end path (host COO -> layouts -> fit -> MAE) on the ml-25m shape."""
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
host = [pin(x) for x in (*d["train"], *d["test"])]
hu, hi, hr, tu, ti, tv = [h[0] for h in host]
out = torch.zeros(2, dtype=torch.float64, device="cuda")
for rep in range(3):
    t = [time.perf_counter()]
    R = eng.ratings(hu, hi, hr); eng.sync(); t.append(time.perf_counter())
    T = eng.ratings(tu, ti, tv); eng.sync(); t.append(time.perf_counter())
    m = E.Model(eng, R, sync=False); eng.sync(); t.append(time.perf_counter())
    m.mae_async(T, out.data_ptr()); eng.sync(); t.append(time.perf_counter())
    r = out.cpu().numpy(); t.append(time.perf_counter())
    print("rep", rep, "train build %.2f ms | test build %.2f | first fit (tiled layout + kernels) %.2f | first mae (mae layout + kernel) %.2f | d2h %.2f | total %.2f ms  mae=%.6f" % (
        *(1e3 * (t[k + 1] - t[k]) for k in range(5)), 1e3 * (t[-1] - t[0]), r[0] / r[1]))
    t0 = time.perf_counter(); m.close(); T.close(); R.close(); print("  destroy %.2f ms" % (1e3 * (time.perf_counter() - t0)))
for rep in range(3):   # staged form: both uploads start at once (mrs_upload_begin), builds overlap the copies
    t = [time.perf_counter()]
    ur = eng.upload(hu, hi, hr); ut = eng.upload(tu, ti, tv); t.append(time.perf_counter())
    R = ur.ratings(); t.append(time.perf_counter())
    m = E.Model(eng, R, sync=False); t.append(time.perf_counter())
    T = ut.ratings(); t.append(time.perf_counter())
    m.mae_async(T, out.data_ptr()); r = out.cpu().numpy(); t.append(time.perf_counter())
    print("staged rep", rep, "upload calls %.2f ms | train build %.2f | fit enqueue %.2f | test build %.2f | mae + d2h %.2f | total %.2f ms  mae=%.6f" % (
        *(1e3 * (t[k + 1] - t[k]) for k in range(5)), 1e3 * (t[-1] - t[0]), r[0] / r[1]))
    m.close(); T.close(); R.close()

hc, _k1 = pin((d["train"][2] * 2).astype(np.uint8)); tc, _k2 = pin((d["test"][2] * 2).astype(np.uint8))
for rep in range(3):   # compact form (int32, int32, uint8), phases separated by syncs
    t = [time.perf_counter()]
    R = eng.ratings_from_codes(hu, hi, hc); eng.sync(); t.append(time.perf_counter())
    T = eng.ratings_from_codes(tu, ti, tc); eng.sync(); t.append(time.perf_counter())
    m = E.Model(eng, R, sync=False); eng.sync(); t.append(time.perf_counter())
    m.mae_async(T, out.data_ptr()); eng.sync(); t.append(time.perf_counter())
    print("codes rep", rep, "train build %.2f ms | test build %.2f | first fit (tiled layout + kernels) %.2f | first mae (mae layout + kernel) %.2f | total %.2f ms" % (
        *(1e3 * (t[k + 1] - t[k]) for k in range(4)), 1e3 * (t[-1] - t[0])))
    m.close(); T.close(); R.close()
for rep in range(3):   # compact + staged
    t = [time.perf_counter()]
    ur = eng.upload_codes(hu, hi, hc); ut = eng.upload_codes(tu, ti, tc); t.append(time.perf_counter())
    R = ur.ratings(); t.append(time.perf_counter())
    m = E.Model(eng, R, sync=False); t.append(time.perf_counter())
    T = ut.ratings(); t.append(time.perf_counter())
    m.mae_async(T, out.data_ptr()); r = out.cpu().numpy(); t.append(time.perf_counter())
    print("codes staged rep", rep, "upload calls %.2f ms | train build %.2f | fit enqueue %.2f | test build %.2f | mae + d2h %.2f | total %.2f ms  mae=%.6f" % (
        *(1e3 * (t[k + 1] - t[k]) for k in range(5)), 1e3 * (t[-1] - t[0]), r[0] / r[1]))
    m.close(); T.close(); R.close()
