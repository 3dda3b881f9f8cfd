#!/usr/bin/env python
"""Where the item pass spends its time, per CTA (clock64 stamps recorded by the kernel when MRS_PASS_DEBUG=1)."""
import os
import sys
os.environ.setdefault("MRS_PASS_DEBUG", "1")   # 1 = item pass, 2 = test pass
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import mrs_b200  # noqa: F401,E402
from mrs_b200 import engine as E, synth  # noqa: E402

eng = E.Engine(0)
d = synth.cached("ml25m")
R = eng.ratings(*d["train"])
T = eng.ratings(*d["test"])
m = E.Model(eng, R)
m.set_item_averages(False)
import torch  # noqa: E402
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    if not os.environ.get("MRS_WARM"):
        flush.zero_()
        torch.cuda.synchronize()
    m.refit()
    m.mae(T, E.PRED_BASELINE)
eng.sync()
n = 148
out = np.zeros(16 * n, dtype=np.int64)
E._check(E.lib().mrs_debug_pass_stamps(out.ctypes.data, n))
s = out.reshape(n, 16)
cnt = s[:, 15]
rel = (s[:, :13] - s[:, :1]) / 1.965e3   # microseconds at 1965 MHz
g0, g1 = s[:, 13], s[:, 14]
print("globaltimer: first CTA start -> last CTA start %.2f us; first start -> last end %.2f us; per-CTA span mean %.2f max %.2f" % (
    (g0.max() - g0.min()) / 1e3, (g1.max() - g0.min()) / 1e3, ((g1 - g0) / 1e3).mean(), ((g1 - g0) / 1e3).max()))
print("layout", R.layout_info())
names = ["start", "dep wait passed", "averages written"]
for b in list(range(0, n, 12)) + [n - 1]:
    k = int(cnt[b])
    print(f"cta {b:3d} stamps {k}:", " ".join(f"{x:6.2f}" for x in rel[b, :k]))
ends = np.array([rel[b, int(cnt[b]) - 1] for b in range(n)])
print("end of CTA (us after its start): min %.2f mean %.2f max %.2f" % (ends.min(), ends.mean(), ends.max()))
print("dep wait passed: mean %.2f max %.2f; averages written: mean %.2f" % (rel[:, 1].mean(), rel[:, 1].max(), rel[:, 2].mean()))
