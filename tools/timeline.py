#!/usr/bin/env python
"""Timeline of one baseline pass replayed as a CUDA graph (cold L2): first block start / last block end of every kernel,
from %globaltimer stamps the kernels record when the engine is created with MRS_TIMELINE=1."""
import os
import sys
os.environ["MRS_TIMELINE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import mrs_b200  # noqa: F401,E402
from mrs_b200 import engine as E, synth  # noqa: E402

stream = torch.cuda.Stream()
eng = E.Engine(0, stream=stream.cuda_stream)
d = synth.cached("ml25m")
names = ["user_sum", "item_pass", "item_finalize", "test_pass"]
with torch.cuda.stream(stream):
    R, T = eng.ratings(*d["train"]), eng.ratings(*d["test"])
    m = E.Model(eng, R)
    m.set_item_averages(False)
    out2 = torch.zeros(2, dtype=torch.float64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    flush2 = torch.zeros(256 << 20, dtype=torch.uint8, device="cuda")

    def pass_():
        if os.environ.get("MRS_NO_FOLD"):
            m.refit()
            m.mae_async(T, out2.data_ptr())
        else:
            m.fit_mae_async(T, out2.data_ptr())
    pass_(); torch.cuda.synchronize()
    g = eng.capture(pass_)
    buf = np.zeros(32, dtype=np.uint64)
    E._check(E.lib().mrs_debug_timeline(eng._h, buf.ctypes.data))
    rows = []
    for it in range(6):
        flush.zero_(); flush2.sum(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); g.launch(); b.record(stream); torch.cuda.synchronize()
        E._check(E.lib().mrs_debug_timeline(eng._h, buf.ctypes.data))
        ks = [k for k in range(4) if int(buf[2 * k + 1]) > 0]
        t0 = min(int(buf[2 * k]) for k in ks)
        line = f"step {a.elapsed_time(b) * 1e3:6.1f} us |"
        for k in ks:
            line += f" {names[k]} {(int(buf[2 * k]) - t0) / 1e3:5.1f}-{(int(buf[2 * k + 1]) - t0) / 1e3:5.1f} |"
        print(line)
        if os.environ.get("MRS_WARP_STAMPS") and it == 5:
            ws = np.zeros(16384, dtype=np.uint64)
            E._check(E.lib().mrs_debug_warp_stamps(eng._h, ws.ctypes.data))
            end = (ws[:8192].astype(np.int64) - t0) / 1e3
            rows = (ws[8192:] >> np.uint64(32)).astype(np.int64); slices = (ws[8192:] & np.uint64(0xffffffff)).astype(np.int64)
            for cta in (0, 3, 7, 8, 14, 15, 22, 60, 100):
                sl = slice(cta * 32, cta * 32 + 32)
                print(f"cta {cta}: rows/warp {rows[sl].min()}-{rows[sl].max()} slices/warp {slices[sl].min()}-{slices[sl].max()} warp end min/mean/max "
                      f"{end[sl].min():.1f}/{end[sl].mean():.1f}/{end[sl].max():.1f}")
                print("   ", " ".join(f"{x:.0f}" for x in end[sl]))
            used = rows > 0
            x = np.stack([rows[used], slices[used], np.ones(used.sum())], axis=1).astype(np.float64)
            coef, *_ = np.linalg.lstsq(x, end[used] - 17.0, rcond=None)
            print("least squares: busy us =", coef, "(per row, per slice, const)")
        if os.environ.get("MRS_CTA_STAMPS") and it == 5:
            st = np.zeros(1024, dtype=np.uint64)
            E._check(E.lib().mrs_debug_cta_stamps(eng._h, st.ctypes.data))
            st = st.astype(np.int64).reshape(4, 256)
            for nm, a, b in (("item", 0, 1), ("test", 2, 3)):
                used = st[b] > 0
                s0, s1 = (st[a][used] - t0) / 1e3, (st[b][used] - t0) / 1e3
                print(f"{nm} pass per-CTA: stream start min/mean/max {s0.min():.1f}/{s0.mean():.1f}/{s0.max():.1f} | end min/mean/max "
                      f"{s1.min():.1f}/{s1.mean():.1f}/{s1.max():.1f} | busy mean {(s1 - s0).mean():.1f} max {(s1 - s0).max():.1f}")
                print("  start:", " ".join(f"{x:.0f}" for x in s0))
                print("  end:  ", " ".join(f"{x:.0f}" for x in s1))
