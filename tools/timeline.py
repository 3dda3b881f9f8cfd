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
