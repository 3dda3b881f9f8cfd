#!/usr/bin/env python
"""The kNN k=300 closure at ml-100k shape, a few times (stream launches): the short command for ncu captures; prints the
per-kernel durations (CUDA events between the launches) and the closure time of a graph replay."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import mrs_b200
from mrs_b200 import engine as E, synth
d = synth.cached("ml100k")
stream = torch.cuda.Stream()
eng = E.Engine(0, stream=stream.cuda_stream)
R, T = eng.ratings(*d["train"]), eng.ratings(*d["test"])
m = E.Model(eng, R)
s = m.similarity(E.SIM_COSINE, 300)
out2 = torch.zeros(2, dtype=torch.float64, device="cuda")
def closure():
    (m.refit() if os.environ.get("FULL_FIT") else m.refit_users()); s.refit(300); m.mae_async(T, out2.data_ptr(), E.PRED_PERSONALIZED, s)
for _ in range(int(os.environ.get("REPS", "3"))):
    closure()
torch.cuda.synchronize()
if not os.environ.get("NO_TIMING"):
    per = {}
    for _ in range(5):
        eng.profile_begin(); closure()
        for name, t in eng.profile_end():
            per.setdefault(name, []).append(t)
    print({k: round(1e3 * sum(v) / len(v), 1) for k, v in per.items()}, "us")
    g = eng.capture(closure); g.launch(); torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
    for a, b in ev:
        a.record(stream); g.launch(); b.record(stream)
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    print(f"closure (graph replay): median {1e3 * ms[len(ms) // 2]:.1f} us, min {1e3 * ms[0]:.1f} us")
r = out2.cpu().numpy(); print("mae", r[0] / r[1])
if os.environ.get("MRS_TIMELINE"):
    st = np.zeros(1024, dtype=np.uint64)
    closure(); torch.cuda.synchronize()
    E._check(E.lib().mrs_debug_cta_stamps(eng._h, st.ctypes.data))
    closure(); torch.cuda.synchronize()
    E._check(E.lib().mrs_debug_cta_stamps(eng._h, st.ctypes.data))
    s4 = st.reshape(4, 256).astype(np.int64)
    used = s4[3] > 0
    t0 = s4[0][used].min()
    rel = (s4[:, used] - t0) / 1e3
    print("similarity per-CTA stamps (us after the first CTA starts): start / phase-0 streaming / phase-0 end / done")
    for b in range(rel.shape[1]):
        if b < 24 or b % 8 == 0 or b >= rel.shape[1] - 4:
            print(f"  cta {b:3d}: {rel[0, b]:6.1f} {rel[1, b]:6.1f} {rel[2, b]:6.1f} {rel[3, b]:6.1f}")
    print("  done: min/mean/max", rel[3].min(), rel[3].mean(), rel[3].max())
