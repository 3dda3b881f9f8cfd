#!/usr/bin/env python
"""Timeline of one SHARDED baseline pass (weak-scaling shards, fused exchange unless --no-fused) replayed as a CUDA graph on
every rank; rank 0 prints first-block-start / last-block-end of each kernel (MRS_TIMELINE=1, %globaltimer).
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29681 tools/timeline_sharded.py"""
import os
import sys
os.environ["MRS_TIMELINE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import mrs_b200  # noqa: F401,E402
from mrs_b200 import engine as E, sharded, synth  # noqa: E402

fused = "--no-fused" not in sys.argv
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
stream = torch.cuda.Stream(device=dev)
eng = E.Engine(local, stream=stream.cuda_stream)
d = synth.cached("ml25m")
s = synth.weak_shard(d, rank)
nu, ni = world * s["user_stride"] + 1, s["max_item_id"] + 1
names = {0: "user_sum", 1: "item_pass", 2: "finalize/push", 4: "finish_pull", 3: "test_pass"}
with torch.cuda.stream(stream):
    R, T = eng.ratings(*s["train"], nu, ni), eng.ratings(*s["test"], nu, ni)
    sb = sharded.ShardedBaseline(eng, R, T, peer_exchange=True, fused=fused)
    sb.step(); torch.cuda.synchronize(dev)
    sb.capture()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    flush2 = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
    buf = np.zeros(32, dtype=np.uint64)
    E._check(E.lib().mrs_debug_timeline(eng._h, buf.ctypes.data))
    for it in range(6):
        flush.zero_(); flush2.sum(); torch.cuda.synchronize(dev); dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); sb.step(); b.record(stream); torch.cuda.synchronize(dev)
        E._check(E.lib().mrs_debug_timeline(eng._h, buf.ctypes.data))
        if rank == 0:
            ks = [k for k in (0, 1, 2, 4, 3) if int(buf[2 * k + 1]) > 0]
            t0 = min(int(buf[2 * k]) for k in ks)
            print(f"step {a.elapsed_time(b) * 1e3:6.1f} us |" + "".join(
                f" {names[k]} {(int(buf[2 * k]) - t0) / 1e3:5.1f}-{(int(buf[2 * k + 1]) - t0) / 1e3:5.1f} |" for k in ks), flush=True)
    mae = sb.result()
    if rank == 0:
        print("mae", mae, "fused", sb.fused)
dist.barrier()
dist.destroy_process_group()
