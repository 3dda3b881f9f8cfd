#!/usr/bin/env python
"""bench.py -- the reference's headline benchmark on B200: the ml-25m-shape baseline "MAE pass"
(distributed/DistributedBaseline.scala:45-47: MeanAbsoluteErrorSpark(baselinePredictorSpark(train), test))
in ratings/s, plus the kNN k=300 ml-100k-shape fit+predict+MAE time (predict/kNN.scala:42-45) as an extra key.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run, one rank per GPU)
  python bench.py --impl reference ...                      (the reference's CPU algorithm on the host cores)

One JSON line on stdout (rank 0).  Data are synthetic MovieLens-shaped sets (mrs_b200/synth.py, seed 449).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FLUSH_MODE = ["write+read"]
METRIC = "ml25m_baseline_mae_pass_ratings_per_s"
UNIT = "ratings/s"
# distributed-25m-4.json:16-21, 66,147.47 ms for ~25,000,095 ratings on Spark local[4]
# (hardware unstated, text parsing inside the timer)
PUBLISHED_RATINGS_PER_S = 25_000_095 / 66.14747


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the GPU is under load (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,utilization.gpu,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for r in self.rows if len(r) >= 9]
        busy = [r for r in rows if r[4].isdigit() and int(r[4]) >= 50] or rows
        for r in busy:
            try:
                sm.append(float(r[1])); smax.append(float(r[2])); power.append(float(r[3]))
            except ValueError:
                continue
            for j, nme in enumerate(names):
                if r[5 + j].lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(busy), "reasons": sorted(reasons)}


def load_workload(rank):
    import mrs_b200  # noqa: F401
    from mrs_b200 import synth
    t0 = time.time()
    d = synth.cached("ml25m")
    log(f"[rank {rank}] ml25m-shaped synthetic set ready in {time.time() - t0:.1f}s "
        f"(train {d['train'][0].size}, test {d['test'][0].size})")
    return d


def host_threads():
    """Host threads this process may use (torch.distributed.run exports OMP_NUM_THREADS=1: the CPU arms pass their
    thread count explicitly to the oracle's `num_threads` clauses instead of trusting the OpenMP default)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def rank_workload(d, rank, world):
    """(train, test, n_users_dim, n_items_dim) of one rank of the headline (weak-scaling) workload: at N > 1 every rank
    holds one ml-25m-shaped shard of DISTINCT users with its own item popularity (synth.weak_shard), laid out on the
    global id space so that the per-item exchange buffers of all ranks line up."""
    if world == 1:
        return d["train"], d["test"], 0, 0
    from mrs_b200 import synth
    s = synth.weak_shard(d, rank)
    return s["train"], s["test"], world * s["user_stride"] + 1, s["max_item_id"] + 1


def whole_workload(d, world):
    """The union of all ranks' shards: what a CPU reference has to process for the same total work."""
    if world == 1:
        return d["train"], d["test"]
    from mrs_b200 import synth
    u = synth.weak_union(d, world)
    return u["train"], u["test"]


# ----------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference's CPU algorithm (oracle port of the Spark twin, partitions = host threads) on the same TOTAL workload
    as our arm at this N: the union of the N ranks' shards, all host threads, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle
    world = max(int(os.environ.get("WORLD_SIZE", "1")), args.gpus, 1)
    d = load_workload(0)
    tr, te = whole_workload(d, world)
    n = tr[0].size + te[0].size
    threads = host_threads()
    for _ in range(max(args.warmup, 0)):
        oracle.spark_baseline_mae(tr, te, nthreads=threads)
    times = []
    mae = None
    for _ in range(max(args.steps, 1)):
        t0 = time.perf_counter()
        mae, _g = oracle.spark_baseline_mae(tr, te, nthreads=threads)
        times.append(time.perf_counter() - t0)
    total = sum(times)
    value = n * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": args.warmup, "ms_per_step": 1000.0 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": value / PUBLISHED_RATINGS_PER_S, "dtype": "f64", "data": "synthetic ml-25m shape (seed 449)",
        "config": workload_config(d, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"whole workload per step: the union of the {world} rank shard(s) (fit on {tr[0].size:,} train + MAE on "
                                   f"{te[0].size:,} test ratings), arrays already parsed"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "mae": mae,
        "note": "the Scala/Spark reference cannot run here (no JVM); this is oracle/mrs_oracle.c orc_baseline_mae_spark, "
                "the C restatement of P:246-391 with partitions = threads",
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(d, world):
    return {
        "workload": "distributed.DistributedBaseline on synthetic ml-25m shape: baselinePredictorSpark(train) fit + "
                    "MeanAbsoluteErrorSpark on test (BASELINE.json configs[3])",
        "train_ratings_per_gpu": int(d["train"][0].size), "test_ratings_per_gpu": int(d["test"][0].size),
        "users_per_gpu": int(d["n_users"]), "items": int(d["n_items"]),
        "layout": "train: user-major codes padded to 16 B vectors (1 B/rating + 4 B/vector), user-tiled item-major sliced-ELL "
                  "(4 B/rating: valid|code|16-bit local user); test: item-tiled, one packed 8-byte word per rating (int32 user | 16-bit local item | code)",
        "l2": "flushed between timed iterations, outside the event pair: 256 MiB write" +
              (" + 256 MiB read of a second buffer (cold and clean L2: no dirty-line write-backs inside the timed region)" if FLUSH_MODE[0] != "write" else ""),
        "parallelism": (f"user-sharded x{world}: every rank holds one ml-25m-shaped shard of distinct users with its own item "
                        "popularity (synth.weak_shard); one all-reduce of the per-item exchange buffer (own NVLink peer-memory kernel)")
                       if world > 1 else "single GPU",
    }


# ----------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import mrs_b200  # noqa: F401
    from mrs_b200 import engine as E

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        log(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ.pop("NCCL_DEBUG")   # keeps NCCL's version banner off stdout (one JSON line only)
        dist.init_process_group("nccl", device_id=dev)
    d = load_workload(rank)
    tr, te, nu_dim, ni_dim = rank_workload(d, rank, world)   # N > 1: a distinct shard per rank, on the global id space
    n_step = int(tr[0].size + te[0].size)

    stream = torch.cuda.Stream(device=dev)
    eng = E.Engine(local_rank, stream=stream.cuda_stream)
    R = eng.ratings(*tr, nu_dim, ni_dim)
    T = eng.ratings(*te, nu_dim, ni_dim)
    model = E.Model(eng, R) if world == 1 else None
    # the timed closure is baselinePredictorSpark + MeanAbsoluteErrorSpark: it never forms per-item rating averages
    # (P:362-391), so that optional part of the fit is switched off (it costs about 8 % of the item pass)
    if model is not None:
        model.set_item_averages(False)
        model.refit()
    bytes_r, bytes_t = R.bytes(), T.bytes()
    class _Flush:
        """L2 flush between timed iterations, outside the event pairs: a 256 MiB write (twice the 126 MB L2) followed, unless
        --flush write, by a 256 MiB read of a second buffer.  The write alone leaves the L2 full of DIRTY lines, whose
        write-backs then compete with the timed kernels' reads for HBM; the read evicts them, so the timed region starts
        from a cold AND clean L2 (none of the workload's data is resident either way)."""

        def __init__(self):
            self.a = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
            self.b = torch.zeros(256 << 20, dtype=torch.uint8, device=dev) if args.flush != "write" else None
            self.sink = torch.zeros(1, dtype=torch.int64, device=dev)

        def zero_(self):
            self.a.zero_()
            if self.b is not None:
                self.sink += self.b.view(torch.int64).sum()
    flush = _Flush()

    class _Align:
        """N>1: a 16-byte all-reduce through the library's own peer-memory kernel right before a timed step, OUTSIDE the event
        pair: the ranks leave it within a flag round trip of each other (about 2 us), so a step's time is the step and not the
        skew with which the ranks finished flushing their L2 (the exchange inside the step waits for the slowest rank).
        MRS_BENCH_ALIGN=0 switches it off."""
        def __init__(self):
            self.peer = None
            self.buf = torch.zeros(2, dtype=torch.float64, device=dev)
            if world > 1 and not args.nccl and os.environ.get("MRS_BENCH_ALIGN", "1") != "0":
                def gather(b):
                    out = [None] * world
                    dist.all_gather_object(out, b)
                    return out
                self.peer = E.PeerExchange(eng, 2, rank, world, gather)

        def __call__(self):
            if self.peer is not None:
                self.peer.allreduce_async(self.buf.data_ptr(), 2)

        def close(self):
            if self.peer is not None:
                self.peer.close()
                self.peer = None
    align = _Align()
    out2 = torch.zeros(2, dtype=torch.float64, device=dev)

    from mrs_b200 import sharded
    sb = sharded.ShardedBaseline(eng, R, T, peer_exchange=not args.nccl, fused=not args.no_fused, closure=args.closure) if world > 1 else None

    def enqueue():
        if sb is not None and sb.closure:  # three kernels: both exchanges happen inside the test pass (mrs_fit_mae_push_async)
            sb.closure_step()
        elif sb is not None:     # local pass -> all-reduce of the exchange buffer -> finish -> MAE -> 16-byte all-reduce
            sb.fit()
            sb.mae_async()
        elif args.no_fold:
            model.refit()
            model.mae_async(T, out2.data_ptr())
        else:                    # the whole timed closure in one call: the test pass finishes the fit itself (3 kernels)
            model.fit_mae_async(T, out2.data_ptr())

    graph = None

    def step():
        if graph is not None:
            graph.launch()   # one cudaGraphLaunch replays the whole pass
        else:
            enqueue()

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    with torch.cuda.stream(stream):
        enqueue()                      # first use allocates layouts and buffers
        torch.cuda.synchronize(dev)
        if world == 1 and not args.no_graph:
            graph = eng.capture(enqueue)
        elif not args.no_graph:
            sb.capture()               # N>1: kernels between the two NCCL all-reduces replay as two graphs
            class _SG:
                def launch(self):
                    sb.step()
            graph = _SG()
        for _ in range(max(args.warmup, 3)):
            flush.zero_()
            step()
        barrier()
        l0 = E.launch_count(); enqueue(); kernels_per_step = E.launch_count() - l0
        torch.cuda.synchronize(dev)
        launches0 = E.launch_count()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        t_wall0 = time.perf_counter()
        for a, b in evs:
            flush.zero_()          # L2 flush, outside the timed pair
            align()                # N>1: the ranks start the step together (outside the timed pair as well)
            a.record(stream)
            step()
            b.record(stream)
        barrier()
        t_wall = time.perf_counter() - t_wall0
        launches = E.launch_count() - launches0
        step_ms = [a.elapsed_time(b) for a, b in evs]
        total_ms = sum(step_ms)
        tot = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        total_ms = float(tot.item())
        res = (sb.out2 if sb is not None else out2).cpu().numpy()
        mae = float(res[0] / res[1])

        # ---- per-kernel durations (CUDA events on the launching stream) for the roofline
        per_kernel = {}
        reps = 10
        for _ in range(reps):
            flush.zero_()
            eng.profile_begin()
            if sb is None:
                model.refit()
                model.mae_async(T, out2.data_ptr())
            elif sb.closure:
                sb.closure_step()
            else:            # same number of exchanges on every rank
                sb.fit_local(); sb.exchange()
                if sb.peer is not None and not sb.fused and rank == 0 and _ == reps - 1:
                    log(f"[rank 0] big exchange stamps (ns after start: published, barrier1, reduced, barrier2, done): {sb.peer.stamps()}")
                sb.fit_finish(); sb.mae_local(); sb.mae_exchange()
                if sb.peer is not None and not sb.fused and rank == 0 and _ == reps - 1:
                    log(f"[rank 0] small exchange stamps: {sb.peer.stamps()}")
            for name, ms in eng.profile_end():
                per_kernel.setdefault(name, []).append(ms)
        per_kernel = {k: sum(v) / len(v) for k, v in per_kernel.items()}

        # ---- sustained repetition so that nvidia-smi (100 ms sampling) sees the same step under load
        # (a fixed count derived from the max-reduced step time: every rank must issue the same number of exchanges)
        n_sustain = int(min(20000, max(50, 1.5e3 / max(total_ms / args.steps, 1e-3))))
        for k in range(n_sustain):
            step()
            if k % 50 == 49:
                torch.cuda.synchronize(dev)
        torch.cuda.synchronize(dev)
    clocks = sampler.stop() if sampler else None

    value = world * n_step * args.steps / (total_ms / 1000.0)
    peak, peak_src = measured_peak()
    bytes_r, bytes_t = R.bytes(), T.bytes()  # the tiled layouts exist once the first pass has run
    alg = {"user_sum": bytes_r["user_major"], "item_tiled": bytes_r["item_major"], "predict_mae_tiled": bytes_t["sorted_coo"]}
    alg_total = sum(alg.values())
    dom = max((k for k in per_kernel if k in alg), key=lambda k: per_kernel[k])
    dom_gbs = alg[dom] / (per_kernel[dom] * 1e-3) / 1e9
    step_ms_mean = total_ms / args.steps
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(dom)
    except Exception:
        pass
    roofline = {
        "bound": "hbm", "kernel": dom, "achieved": dom_gbs, "peak": peak, "unit": "GB/s", "frac": dom_gbs / peak,
        "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg[dom],
        "kernel_ms": per_kernel[dom], "per_kernel_ms": per_kernel,
        "step": {"algorithmic_bytes": alg_total, "achieved": alg_total / (step_ms_mean * 1e-3) / 1e9,
                 "frac": alg_total / (step_ms_mean * 1e-3) / 1e9 / peak},
        # SURVEY 8(d) "also report test-only": the test pass alone (predict + |error| + reduction over this rank's test ratings)
        "test_only": ({"ratings_per_s": float(te[0].size) / (per_kernel["predict_mae_tiled"] * 1e-3), "kernel": "predict_mae_tiled",
                       "kernel_ms": per_kernel["predict_mae_tiled"], "algorithmic_bytes_per_launch": alg["predict_mae_tiled"],
                       "achieved": alg["predict_mae_tiled"] / (per_kernel["predict_mae_tiled"] * 1e-3) / 1e9,
                       "frac": alg["predict_mae_tiled"] / (per_kernel["predict_mae_tiled"] * 1e-3) / 1e9 / peak}
                      if per_kernel.get("predict_mae_tiled") else None),
    }

    # ---- end to end through the public API with HOST buffers (pinned), H2D + layout build + fit + MAE + D2H per step
    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t.numpy(), t
    host = [pinned(x) for x in (*tr, *te)]
    hu, hi, hr, tu, ti, tv = [h[0] for h in host]
    # the compact host form: the rating as a 1-byte half-star code (2 x rating), what a JNI shim writes while it fills its
    # pinned buffers from the JVM's Rating objects -- 9 bytes per rating over PCIe instead of 16
    hc, _hc_keep = pinned((tr[2] * 2).astype(np.uint8))
    tc, _tc_keep = pinned((te[2] * 2).astype(np.uint8))
    e2e_steps = max(1, min(args.steps, 5))
    keep = []

    def e2e_step(compact=True):
        # both sets start travelling at once on the engine's copy stream, the (small) test set first: its layouts are built
        # while the train ids are still in flight, the train set is sorted while its ratings arrive
        test_first = not os.environ.get("MRS_E2E_TRAIN_FIRST")
        if test_first:
            up_t = eng.upload_codes(tu, ti, tc) if compact else eng.upload(tu, ti, tv)
        up_r = eng.upload_codes(hu, hi, hc) if compact else eng.upload(hu, hi, hr)
        if not test_first:
            up_t = eng.upload_codes(tu, ti, tc) if compact else eng.upload(tu, ti, tv)
        if test_first and world == 1:
            T2 = up_t.ratings(nu_dim, ni_dim)
        R2 = up_r.ratings(nu_dim, ni_dim)
        if world == 1:
            m2 = E.Model(eng, R2, sync=False)
            if not test_first:
                T2 = up_t.ratings(nu_dim, ni_dim)
            m2.mae_async(T2, out2.data_ptr())
            r = out2.cpu().numpy()       # D2H read of the result
            handles = (m2, T2, R2)
        else:
            T2 = up_t.ratings(nu_dim, ni_dim)
            # the exchange object (IPC-mapped peer buffers) is long-lived and shared with the pass above; the rating sets
            # are new every step, so the whole exchange buffer travels (no set-up collective for the slots in use)
            s2 = sharded.ShardedBaseline(eng, R2, T2, peer=sb.peer, indexed=False)
            s2.fit()
            s2.mae_async()
            r = s2.out2.cpu().numpy()
            handles = (s2.model, T2, R2)
        for h in handles:                # a caller drops the handles; their device blocks go back to the engine's cache
            h.close()
        return float(r[0] / r[1])

    def e2e_run(compact):
        with torch.cuda.stream(stream):
            e2e_step(compact)
            e2e_step(compact)
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                mae_ = e2e_step(compact)
            barrier()
            secs = time.perf_counter() - t0
        t_ = torch.tensor([secs], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
        return float(t_.item()), mae_

    e2e_f64_s, e2e_f64_mae = e2e_run(False)
    e2e_s, e2e_mae = e2e_run(True)
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    e2e_value = world * n_step * e2e_steps / e2e_s
    h2d = sum(int(x.nbytes) for x in (hu, hi, hc, tu, ti, tc))
    h2d_f64 = sum(int(x.nbytes) for x in (hu, hi, hr, tu, ti, tv))
    for hs in keep:
        for h in hs:
            h.close()

    knn25m = None
    if not args.no_knn25m:
        with torch.cuda.stream(stream):
            knn25m = bench_knn25m(eng, stream, torch, dist, d, rank, world, dev, peer=not args.nccl)

    strong = None
    if world > 1 and not args.no_strong:
        with torch.cuda.stream(stream):
            strong = bench_strong(eng, stream, torch, dist, d, rank, world, dev, flush, args, peer=not args.nccl, align=align)

    line = None
    if rank == 0:
        import oracle
        if world == 1:
            # ---- CPU baseline on this box's host cores: the oracle's Spark-twin port with 1, 4 and all host threads
            # (the local[1] / local[4] / local[N] stand-ins of BASELINE.md section 3), 3 measurements each, mean and
            # population sigma like the reference's own statistics (P:18-25)
            by_threads = {}
            cpu_mae = None
            for j in sorted({1, min(4, host_threads()), host_threads()}):
                ts = []
                for _ in range(3):
                    t0 = time.perf_counter()
                    cpu_mae, _g = oracle.spark_baseline_mae(tr, te, nthreads=j)
                    ts.append(time.perf_counter() - t0)
                mean_s = sum(ts) / len(ts)
                by_threads[str(j)] = {"ratings_per_s": n_step / mean_s, "mean_ms": 1000.0 * mean_s,
                                      "stddev_ms": 1000.0 * (sum((t - mean_s) ** 2 for t in ts) / len(ts)) ** 0.5, "runs": len(ts)}
            top = str(host_threads())
            cpu = {"value": by_threads[top]["ratings_per_s"], "unit": UNIT, "cores": host_threads(), "kind": "port",
                   "sample": "whole workload (fit on 20,000,076 train + MAE on 5,000,019 test), arrays already parsed; 3 runs per thread count",
                   "by_threads": by_threads, "mae": cpu_mae, "host_cores_available": os.cpu_count()}
        else:
            # N > 1: no CPU baseline line (rank 0 at N=1 only), but the parity check needs the oracle's MAE of the UNION of
            # all ranks' shards -- distinct data per rank, so an exchange that did nothing cannot match it
            wtr, wte = whole_workload(d, world)
            t0 = time.perf_counter()
            cpu_mae, _g = oracle.spark_baseline_mae(wtr, wte, nthreads=host_threads())
            cpu = {"value": None, "unit": UNIT, "cores": host_threads(), "kind": "port",
                   "sample": f"parity only: oracle MAE of the union of the {world} shards ({wtr[0].size + wte[0].size:,} ratings) "
                             f"in {time.perf_counter() - t0:.2f} s; the timed CPU baseline is reported at N=1 and by --impl reference",
                   "mae": cpu_mae}
            del wtr, wte
        knn = bench_knn(eng, stream, torch) if world == 1 else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": step_ms_mean, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": value / PUBLISHED_RATINGS_PER_S, "dtype": "f64", "data": "synthetic ml-25m shape (seed 449)",
            "config": workload_config(d, world),
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 16, "steps": e2e_steps,
                    "ms_per_step": 1000.0 * float(e2e_t.item()) / e2e_steps, "mae": e2e_mae,
                    "note": "pinned host COO in the compact form (int32 user, int32 item, uint8 half-star code: 9 B per rating) -> H2D on a copy "
                            "stream (mrs_upload_begin_codes; test set and train codes travel while the train ids are sorted) -> CSR/CSC + "
                            "kernel layouts build -> fit -> MAE -> D2H -> handles released, per step (2 untimed warm-up steps fill the "
                            "engine's device block cache)",
                    "f64_form": {"value": world * n_step * e2e_steps / e2e_f64_s, "ms_per_step": 1000.0 * e2e_f64_s / e2e_steps,
                                 "h2d_bytes_per_step": h2d_f64, "mae": e2e_f64_mae,
                                 "note": "the same with fp64 ratings on the host (int32,int32,f64: 16 B per rating, mrs_upload_begin)"}},
            "gpu_launches": int(launches) if graph is None else int(args.steps * kernels_per_step),
            "launch_mode": ("cuda graph replay (1 cudaGraphLaunch per step)" if world == 1 else
                            ("1 cuda graph per step (exchanges fused into the kernels)" if not args.nccl and not args.no_fused else
                             "1 cuda graph per step incl. the 2 peer-memory exchange kernels" if not args.nccl else
                             "2 cuda graphs + 2 NCCL all-reduces per step")) if graph is not None else "stream launches", "clocks": clocks, "mae": mae, "mae_matches_cpu_port": abs(mae - cpu_mae) <= 1e-6 * abs(cpu_mae),
            "wall_ms_per_step_incl_flush": 1000.0 * t_wall / args.steps,
            "exchange": None if sb is None else ("nccl" if sb.peer is None else
                                                 {"kind": ("both exchanges inside the test pass kernel (mrs_fit_mae_push_async): its CTAs push the per-item partial sums into every rank's receive buffer over NVLink, wait for all flags and sum locally; 3 kernels per step"
                                                           if sb.closure else "fused into the pass' own kernels: partial sums pushed into every rank's receive buffer over NVLink")
                                                          if sb.fused else "own NVLink peer-memory all-reduce kernel", "timed_out": sb.peer.timed_out()}),
            "step_ms_min_max": [min(step_ms), max(step_ms)],
            "rank_alignment": ("16-byte peer-memory all-reduce before every timed step, outside the event pair (the ranks start a step "
                               "within a flag round trip of each other)") if align.peer is not None else None,
            "strong_scaling": strong,
            "knn": knn,
            "knn25m": knn25m,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.cuda.synchronize(dev)
        dist.barrier()
        align.close()
        dist.destroy_process_group()
    sys.stdout.flush()
    return 0


def bench_strong(eng, stream, torch, dist, d, rank, world, dev, flush, args, peer, align=lambda: None):
    """BASELINE config 4 as worded: ONE ml-25m-shaped set, its users partitioned over the ranks
    (distributed/DistributedBaseline.scala:30-47 with --master local[N]; reduceByKey/collect of P:267-268 = the exchange).
    A rank holds the train rows and the test pairs of its user range (sharded.partition_users / shard_of), the tables are
    laid out on the global id space.  Strong scaling: the job is fixed, N ranks split it.  Parity: the MAE against the
    oracle and the per-item average deviations against a single-GPU fit of the whole set on rank 0."""
    import numpy as np
    from mrs_b200 import engine as E, sharded
    tr, te = d["train"], d["test"]
    nu_dim = int(max(tr[0].max(), te[0].max())) + 1
    ni_dim = int(max(tr[1].max(), te[1].max())) + 1
    bounds = sharded.partition_users(np.bincount(tr[0], minlength=nu_dim), world)
    mtr, mte = sharded.shard_of(tr[0], bounds, rank), sharded.shard_of(te[0], bounds, rank)
    R = eng.ratings(tr[0][mtr], tr[1][mtr], tr[2][mtr], nu_dim, ni_dim)
    T = eng.ratings(te[0][mte], te[1][mte], te[2][mte], nu_dim, ni_dim)
    sb = sharded.ShardedBaseline(eng, R, T, peer_exchange=peer, fused=not args.no_fused, closure=args.closure)
    sb.step()
    torch.cuda.synchronize(dev)
    if not args.no_graph:
        sb.capture()
    for _ in range(max(args.warmup, 3)):
        flush.zero_()
        sb.step()
    torch.cuda.synchronize(dev)
    dist.barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in evs:
        flush.zero_()
        align()
        a.record(stream)
        sb.step()
        b.record(stream)
    torch.cuda.synchronize(dev)
    dist.barrier()
    tot = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], dtype=torch.float64, device=dev)
    dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    ms = float(tot.item()) / args.steps
    mae = sb.result()
    n_total = int(tr[0].size + te[0].size)
    out = None
    idev = sb.model.vector(E.ITEM_AVG_DEV)[0]
    if rank == 0:
        import oracle
        ref_mae, _g = oracle.spark_baseline_mae(tr, te, nthreads=host_threads())
        R1 = eng.ratings(*tr, nu_dim, ni_dim)
        m1 = E.Model(eng, R1)
        idev1 = m1.vector(E.ITEM_AVG_DEV)[0]
        tol = 1e-6 * np.maximum(np.abs(idev1), 1e-12)
        worst = float(np.max(np.abs(idev - idev1) / np.maximum(np.abs(idev1), 1e-12)))
        out = {"metric": METRIC, "value": n_total / (ms / 1000.0), "unit": UNIT, "n_gpus": world, "ms_per_step": ms,
               "scaling": "strong", "steps": args.steps,
               "workload": "ONE synthetic ml-25m-shaped set (20,000,076 train / 5,000,019 test), users partitioned over the ranks "
                           "(sharded.partition_users: contiguous id ranges balanced by rating count)",
               "user_bounds": [int(b) for b in bounds], "train_ratings_rank0": int(mtr.sum()), "test_ratings_rank0": int(mte.sum()),
               "mae": mae, "oracle_mae": ref_mae, "mae_matches_cpu_port": bool(abs(mae - ref_mae) <= 1e-6 * abs(ref_mae)),
               "item_avg_dev_matches_single_gpu_fit": bool(np.all(np.abs(idev - idev1) <= tol)), "item_avg_dev_worst_rel": worst,
               "exchange": "nccl" if sb.peer is None else (("both exchanges inside the test pass kernel (mrs_fit_mae_push_async, 3 kernels per step, NVLink push)" if sb.closure else "fused into the pass' kernels (NVLink push)") if sb.fused else "own NVLink peer-memory kernel")}
        m1.close(); R1.close()
    torch.cuda.synchronize(dev)
    dist.barrier()                 # nobody may still be reading our symmetric buffers when they are unmapped
    for g in ("_g_all", "_g1", "_g2"):
        if getattr(sb, g, None) is not None:
            getattr(sb, g).close()
    sb.close(close_peer=True)
    for h in (T, R):
        h.close()
    return out


def bench_knn(eng, stream, torch):
    """kNN k=300 on ml-100k shape: the closure timed by predict/kNN.scala:42-45 (similarities + top-k + predict + MAE)."""
    from mrs_b200 import engine as E, synth
    d = synth.cached("ml100k")
    tr, te = d["train"], d["test"]
    R, T = eng.ratings(*tr), eng.ratings(*te)
    m = E.Model(eng, R)
    s = m.similarity(E.SIM_COSINE, 300)
    out2 = torch.zeros(2, dtype=torch.float64, device=f"cuda:{eng.device}")
    reps = 30

    def closure():
        # the fit of this closure is the users-only one: predictor(train, wsd(train, getSimilarity(train, k, cosine))) takes the
        # user averages and the global average from it (P:557-586), never the per-item average deviation of the baseline
        m.refit_users()
        s.refit(300)
        m.mae_async(T, out2.data_ptr(), E.PRED_PERSONALIZED, s)

    for _ in range(5):
        closure()
    torch.cuda.synchronize()
    # like the baseline pass, the closure is captured once and replayed: 8 kernels of 7-100 us each, the launch gaps matter
    graph, run, mode = None, closure, "stream launches"
    try:
        graph = eng.capture(closure)
        run, mode = graph.launch, "cuda graph replay (1 cudaGraphLaunch per closure)"
        run()
        torch.cuda.synchronize()
    except Exception as ex:  # capture refused: time the plain launches
        log(f"kNN closure not captured ({ex}); timing stream launches")
        graph, run = None, closure
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:
        a.record(stream)
        run()
        b.record(stream)
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in evs]
    if graph is not None:
        graph.close()
    r = out2.cpu().numpy()
    per_kernel = {}
    for _ in range(5):
        eng.profile_begin()
        closure()
        for name, t in eng.profile_end():
            per_kernel.setdefault(name, []).append(t)
    per_kernel = {k: sum(v) / len(v) for k, v in per_kernel.items()}
    # CPU port of the same closure (single thread)
    from oracle import oracle as O
    t0 = time.perf_counter()
    o = O.Oracle(*tr)
    cpu_mae = o.mae(te, kind=O.PERSONALIZED, simkind=O.SIM_COSINE, k=300)
    cpu_ms = 1000.0 * (time.perf_counter() - t0)
    mae = float(r[0] / r[1])
    for h in (s, m, T, R):
        h.close()
    # roofline of the similarity SpGEMM: an L2-resident working set (about 1 MB of inputs, S = 943^2 x 8 B = 7.1 MB); what it
    # is short of is shared-memory bandwidth, not fp64 rate (both are reported)
    import numpy as np
    cnt_i = np.bincount(tr[1]).astype(np.float64)
    cnt_u = np.bincount(tr[0])
    # Products the kernel executes: row users in blocks of 16 positions of the length-sorted order, block b meets the users at
    # positions >= 16 b (one triangle, the other one is copied); a step = one entry of a column user against the 16 row users.
    lens = np.sort(cnt_u[cnt_u > 0])[::-1].astype(np.int64)
    steps = float((lens * (np.arange(lens.size) // 16 + 1)).sum())
    executed = 16.0 * steps
    useful = float((cnt_i * cnt_i).sum())                    # products over item intersections (both triangles)
    fp64 = eng.fp64_fma_per_s()
    sim_s = per_kernel.get("similarity", 0.0) * 1e-3
    # Bound: shared-memory wavefronts.  A warp step (two column users x 16 rows = 32 products) reads two 16-byte records
    # (2 wavefronts: one address per half warp) and two 128-byte tile rows (2 wavefronts); an SM serves one wavefront per clock.
    wavefronts = 4.0 * steps / 2.0
    sm_clock = 1.965e9
    lsu_peak = 148 * sm_clock
    knn_roof = {"bound": "shared-memory wavefronts (LSU): 4 per warp step of 32 products, 1 per clock per SM", "kernel": "similarity",
                "achieved": wavefronts / sim_s if sim_s else None, "peak": lsu_peak, "unit": "shared-memory wavefronts/s",
                "frac": (wavefronts / sim_s / lsu_peak) if sim_s else None, "executed_products": executed, "useful_products": useful,
                "kernel_ms": per_kernel.get("similarity"),
                "fp64": {"achieved_macs_per_s": executed / sim_s if sim_s else None, "peak": fp64,
                         "peak_source": "measured in this run (mrs_debug_fp64_fma_per_s: 8 DFMA chains per thread, all SMs)"},
                "peak_source": "148 SMs x 1.965 GHz x 1 wavefront per clock (ncu: l1tex__data_pipe_lsu_wavefronts, profiles/r02_ncu_summary_knn.txt)",
                "traffic": None}
    # SURVEY 8(d) work counts of the closure: similarity products, keys ranked, prediction gathers, compulsory bytes
    n_known = int((cnt_u > 0).sum())
    ti_ok = te[1][te[1] < cnt_i.size]
    knn_work = {"similarity_products_sparse": useful, "similarity_products_dense": float(n_known) ** 2 * float(cnt_i.size),
                "similarity_products_executed": executed, "topk_keys": float(n_known) * float(n_known - 1),
                "predict_gathers": float(cnt_i[ti_ok].sum()),
                "compulsory_bytes": 12.0 * float(tr[0].size) + 12.0 * float(te[0].size) + 12.0 * float(n_known) * 300.0}
    return {"roofline": knn_roof, "work": knn_work, "metric": "knn_k300_ml100k_fit_predict_mae_ms", "value": statistics.median(ms), "unit": "ms", "min_ms": min(ms),
            "mean_ms": sum(ms) / len(ms), "reps": reps, "launch_mode": mode, "mae": mae, "per_kernel_ms": per_kernel,
            "closure": "mrs_fit_users_async (user + global averages: all that predictor/weightedSumDeviation read from the fit, P:489-586) -> "
                       "mrs_fit_similarity_async(cosine, k=300) -> mrs_mae_async(PERSONALIZED)",
            "l2": "not flushed: the whole working set (about 25 MB) is L2-resident by design",
            "cpu_port_ms": cpu_ms, "cpu_port_mae": cpu_mae, "published_reference_ms": 26198.54,
            "mae_matches_cpu_port": abs(mae - cpu_mae) <= 1e-6 * abs(cpu_mae)}


def bench_knn25m(eng, stream, torch, dist, d, rank, world, dev, peer, k=300, reps=2):
    """BASELINE config 5: kNN k=300 at ml-25m shape, similarity rows sharded over the ranks (train set replicated), fused
    predict + MAE over each rank's test pairs, one 16-byte exchange.  Strong scaling: the job is fixed, N ranks split it."""
    from mrs_b200 import sharded
    tr, te = d["train"], d["test"]
    sk = sharded.ShardedKnn(eng, tr, te, k=k, rank=rank, world=world, peer_exchange=peer)
    sk.step()                                   # first use builds the layouts and fills the block cache
    torch.cuda.synchronize(dev)
    times, prof = [], {}
    for _ in range(reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng.profile_begin()
        a.record(stream)
        sk.step()
        b.record(stream)
        torch.cuda.synchronize(dev)
        prof = {}
        for name, ms in eng.profile_end():
            prof[name] = prof.get(name, 0.0) + ms
        t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t.item()))
    mae = sk.mae()
    out = None
    if rank == 0:
        # neighbour lists of a few of this rank's users against the CPU oracle, at full size
        from oracle import oracle as O
        import numpy as np
        t0 = time.perf_counter()
        o = O.Oracle(*tr)
        users = np.unique(tr[0][(tr[0] >= sk.user_lo) & (tr[0] < sk.user_hi)])
        picks = np.random.default_rng(5).choice(users, 3, replace=False)
        same = True
        for u in picks:
            ids, sims = sk.sim.neighbors(int(u), k)
            oi, os_ = o.neighbors(int(u), k)
            same &= ids.tolist() == oi.tolist() and sims.tolist() == os_.tolist()
        ms = statistics.median(times)
        cnt_i = np.bincount(tr[1]).astype(np.float64)
        products = float((cnt_i * cnt_i).sum())                  # sum_i cnt_i^2 pair products, all ranks together
        fp64 = eng.fp64_fma_per_s()
        rows_s = prof.get("knn_rows", 0.0) * 1e-3
        roof = {"bound": "fp64 FMA rate as the reference; the kernel is latency bound (L2 gathers of column slices + a shared-memory "
                         "read-modify-write per product)", "kernel": "knn_rows", "achieved": products / world / rows_s if rows_s else None,
                "peak": fp64, "unit": "fp64 multiply-adds/s per GPU", "frac": (products / world / rows_s / fp64) if rows_s else None,
                "pair_products": products, "kernel_ms_rank0": prof.get("knn_rows"),
                "peak_source": "measured in this run (mrs_debug_fp64_fma_per_s)", "traffic": None}
        out = {"roofline": roof, "metric": "knn_k300_ml25m_fit_predict_mae_s", "value": ms / 1000.0, "unit": "s", "n_gpus": world, "reps": reps,
               "scaling": "strong", "mae": mae, "times_ms": times, "rows_rank0": [sk.user_lo, sk.user_hi],
               "test_pairs_rank0": sk.n_test_local, "pairs_per_s": float(te[0].size) / (ms / 1000.0),
               "rank0_kernel_ms": {a: round(b, 3) for a, b in prof.items()},
               "exchange": "16 bytes {sum |err|, n}: " + ("own NVLink peer-memory kernel" if sk.peer is not None else
                                                          ("nccl" if world > 1 else "none (1 rank)")),
               "oracle_check": {"users": [int(u) for u in picks], "neighbour_lists_identical": bool(same),
                                "seconds": time.perf_counter() - t0},
               "note": "train set replicated on every rank (240 MB), no all-gather needed; similarity work = sum_i cnt_i^2 pair "
                       "products, exact fp64 in the oracle's order"}
    sk.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nccl", action="store_true", help="N>1: use NCCL all-reduces instead of the library's peer-memory exchange kernel")
    ap.add_argument("--no-knn25m", action="store_true", help="skip the kNN k=300 leg at ml-25m shape (BASELINE config 5)")
    ap.add_argument("--flush", default="write+read", choices=["write", "write+read"],
                    help="L2 flush between timed iterations: 256 MiB write, or write followed by a 256 MiB read (cold and clean L2)")
    ap.add_argument("--no-fold", action="store_true", help="N=1: mrs_fit_async + mrs_mae_async (4 kernels) instead of mrs_fit_mae_async (3 kernels)")
    ap.add_argument("--closure", action="store_true", help="N>1: the whole step as mrs_fit_mae_push_async (3 kernels, both exchanges inside the test pass) instead of the "
                    "fused exchange with its own delivering / finishing kernels (5 kernels); measured slower at 2 ranks (96.1 vs 93.2 us), kept as an option")
    ap.add_argument("--no-fused", action="store_true", help="N>1: separate exchange kernels between the pass' kernels instead of the fused push exchange")
    ap.add_argument("--no-strong", action="store_true", help="N>1: skip the strong-scaling leg (ONE ml-25m set user-sharded over the ranks)")
    ap.add_argument("--no-graph", action="store_true", help="launch the kernels of a step one by one instead of replaying a CUDA graph")
    args = ap.parse_args()
    args.steps = max(args.steps, 1)
    FLUSH_MODE[0] = args.flush
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
