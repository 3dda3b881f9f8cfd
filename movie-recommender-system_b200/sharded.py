"""User-sharded baseline pass: one process per GPU, one collective.

The reference distributes the baseline with Spark: `reduceByKey` + `collect` per keyed average (P:267-268) and
`sum`/`count` for the global mean (P:247).  Here users are partitioned into contiguous id ranges balanced by rating
count; a rank holds the train rows and the test pairs of its users, so user sums/averages are local.  The only exchange
is ONE all-reduce (sum) of the per-item exchange buffer written by ``mrs_fit_local``:

    [ sum of deviations per item | count per item | sum of all ratings | count | sum of ratings per item ]   (3*I+2 fp64;
    the first 2*I+2 values are all the baseline predictor needs)

followed by ``mrs_fit_finish`` on every rank, and a 16-byte all-reduce of {sum |err|, n} for the MAE.
The collective itself is `torch.distributed` (NCCL on GPUs; gloo in the CPU tests of this host logic).
"""
import numpy as np


def partition_users(user_counts, world):
    """Cut points b[0]=0 <= ... <= b[world]=len(user_counts): rank r owns user ids [b[r], b[r+1]); contiguous ranges
    whose rating counts are as equal as a contiguous split allows."""
    counts = np.asarray(user_counts, dtype=np.int64)
    total = int(counts.sum())
    csum = np.cumsum(counts)
    bounds = [0]
    for r in range(1, world):
        target = total * r // world
        cut = int(np.searchsorted(csum, target, side="left")) + 1 if total else 0
        bounds.append(min(max(cut, bounds[-1]), counts.size))
    bounds.append(int(counts.size))
    return bounds


def partition_rows(user_counts, world):
    """Cut points for the similarity rows of a sharded kNN run: the cost of a row is about (its length) for the
    similarity pass plus a constant for the selection pass, so ranges are balanced on count + mean count."""
    counts = np.asarray(user_counts, dtype=np.int64)
    known = counts > 0
    mean = int(counts[known].mean()) if known.any() else 0
    return partition_users(counts + known * mean, world)


def shard_of(users, bounds, rank):
    """Boolean mask of the entries owned by `rank` (its users are [bounds[rank], bounds[rank+1]))."""
    u = np.asarray(users)
    return (u >= bounds[rank]) & (u < bounds[rank + 1])


def exchange_size(n_items_dim):
    return 3 * int(n_items_dim) + 2


def mandatory_size(n_items_dim):
    """Length of the prefix that must be exchanged when per-item rating averages are switched off."""
    return 2 * int(n_items_dim) + 2


def split_exchange(buf, n_items_dim):
    """Views of an exchange buffer: (dev_sum[I], rating_sum[I], count[I], global_sum, global_count)."""
    n = int(n_items_dim)
    return buf[:n], buf[2 * n + 2:3 * n + 2], buf[n:2 * n], buf[2 * n], buf[2 * n + 1]


def finish_from_exchange(buf, n_items_dim):
    """What ``mrs_fit_finish`` computes from the (all-reduced) buffer -- numpy statement used by the CPU tests of the
    exchange protocol: item average deviation (0.0 for unknown items, P:197), item average, global average."""
    dev, rate, cnt, gs, gc = split_exchange(np.asarray(buf, dtype=np.float64), n_items_dim)
    known = cnt > 0
    idev = np.where(known, dev / np.where(known, cnt, 1.0), 0.0)
    iavg = np.where(known, rate / np.where(known, cnt, 1.0), np.nan)
    gavg = gs / gc if gc > 0 else 0.0
    return idev, iavg, gavg


def all_reduce_sum(tensor, group=None):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
    return tensor


class _CudaView:
    """Zero-copy `__cuda_array_interface__` view of a device buffer owned by the engine (n fp64 values)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


class ShardedBaseline:
    """Fit + MAE of the baseline predictor over this rank's shard, with the one exchange of the module docstring.

    `engine` must have been created on the CUDA stream the collectives run on (``torch.cuda.current_stream()``)."""

    def __init__(self, engine, train, test, group=None, item_averages=False, peer_exchange=False, peer=None, indexed=True, fused=False,
                 closure=False):
        """``peer_exchange``: create the library's peer-memory exchange object for the two all-reduces (else NCCL through
        torch.distributed).  ``peer``: an existing PeerExchange to reuse instead (its buffers are long-lived IPC mappings;
        a pass that is rebuilt every step, like the end-to-end arm of bench.py, must not create one per step);
        ``indexed=False`` then exchanges the whole buffer, so no set-up collective is needed for the new rating sets.
        ``fused`` (with ``peer_exchange``): the exchanges happen INSIDE the pass' own kernels -- the fit's last kernel
        delivers its partial sums into every rank's receive buffer over NVLink, the finishing kernel adds the deliveries
        from its own memory, the test pass' last block exchanges {sum |err|, n} itself (``mrs_fit_local_push`` /
        ``mrs_fit_finish_pull`` / ``mrs_mae_push_async``): two launches and the remote-load round trip fewer per step.
        ``closure`` (with ``fused``): the whole step is ``mrs_fit_mae_push_async`` -- three kernels; the test pass' CTAs deliver
        the per-item partial sums themselves (an equal share each), wait for all ranks and build their tile's deviations
        from the deliveries: no separate delivering or finishing kernel."""
        import torch
        from . import engine as E
        self.E, self.torch, self.group = E, torch, group
        self.engine, self.train, self.test = engine, train, test
        self.model = E.Model(engine, train, sync=False)
        self.model.set_item_averages(item_averages)
        ptr, n = self.model.exchange_buffer()
        self.device = torch.device("cuda", engine.device)
        self.xbuf = torch.as_tensor(_CudaView(ptr, n), device=self.device)
        if not item_averages:            # only [dev sums | counts | global sum, count] has to travel
            self.xbuf = self.xbuf[:mandatory_size(train.n_items_dim)]
        self.out2 = torch.zeros(2, dtype=torch.float64, device=self.device)
        # peer_exchange: the two all-reduces run as the library's own NVLink peer-memory kernel instead of NCCL
        self.peer = peer
        self.peer_small = None           # fused mode: the 16-byte exchange has its own handle (its own epoch counter)
        self.fused = False
        self.closure = False
        self.known = None
        self.xidx = None
        if peer is not None and indexed:
            import torch.distributed as dist
            self.xidx = self._slots_in_use(dist, item_averages)
        if peer_exchange and peer is None:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
                world, rank = dist.get_world_size(group), dist.get_rank(group)

                def gather(b):
                    out = [None] * world
                    dist.all_gather_object(out, b, group=group)
                    return out
                self.peer = E.PeerExchange(engine, max(int(self.xbuf.numel()), 2), rank, world, gather)
                self.xidx = self._slots_in_use(dist, item_averages)
                if fused and not item_averages:
                    self.peer_small = E.PeerExchange(engine, 2, rank, world, gather)
                    self.fused = True
                    self.closure = bool(closure)

    def _slots_in_use(self, dist, item_averages):
        """Positions of the exchange buffer that are non-zero on SOME rank: the slots of the items that occur in some rank's
        train shard (+ the two global sums).  Known once the shards are loaded; only these travel in each exchange."""
        torch = self.torch
        n_items = int(self.train.n_items_dim)
        self.fit_local()                                   # leaves this rank's per-item counts in xbuf[I:2I]
        torch.cuda.synchronize(self.device)
        used = (self.xbuf[n_items:2 * n_items] > 0).to(torch.int32)
        dist.all_reduce(used, op=dist.ReduceOp.MAX, group=self.group)      # set-up time only
        known = torch.nonzero(used, as_tuple=False).flatten().to(torch.int32)
        self.known = known.contiguous()
        parts = [known, known + n_items, torch.tensor([2 * n_items, 2 * n_items + 1], dtype=torch.int32, device=self.device)]
        if item_averages:
            parts.append(known + (2 * n_items + 2))
        idx = torch.cat(parts)
        if idx.numel() % 2:                                # the two-shot form works on pairs: repeat a slot (harmless)
            idx = torch.cat([idx, idx[-1:]])
        return idx.contiguous()

    # ---- the five pieces of a step (all asynchronous on the engine's stream)
    def fit_local(self):
        if self.fused:                                 # the last kernel of the local pass delivers the partial sums itself
            self.E._check(self.E.lib().mrs_fit_local_push(self.engine._h, self.train._h, self.E.C.byref(self.model._h), self.peer._h,
                                                          self.E.C.c_void_p(self.known.data_ptr()), int(self.known.numel())))
            return
        self.E._check(self.E.lib().mrs_fit_local(self.engine._h, self.train._h, self.E.C.byref(self.model._h)))

    def exchange(self):                                # THE collective of the fit (P:267-268, P:247)
        if self.fused:
            return                                     # (inside fit_local / fit_finish)
        if self.peer is not None and self.xidx is not None:
            self.peer.allreduce_indexed_async(self.xbuf.data_ptr(), self.xidx.data_ptr(), self.xidx.numel())
        elif self.peer is not None:
            self.peer.allreduce_async(self.xbuf.data_ptr(), self.xbuf.numel())
        else:
            all_reduce_sum(self.xbuf, self.group)

    def fit_finish(self):
        if self.fused:                                 # waits for every rank's delivery, adds them from local memory
            self.E._check(self.E.lib().mrs_fit_finish_pull(self.model._h, self.peer._h))
            return
        self.E._check(self.E.lib().mrs_fit_finish(self.model._h))

    def mae_local(self):
        if self.fused:                                 # the last block of the test pass exchanges {sum |err|, n} itself
            self.E._check(self.E.lib().mrs_mae_push_async(self.model._h, self.test._h, self.peer_small._h,
                                                          self.E.C.c_void_p(self.out2.data_ptr())))
            return
        self.model.mae_async(self.test, self.out2.data_ptr(), self.E.PRED_BASELINE)

    def mae_exchange(self):                            # 16 bytes: {sum |err|, n}
        if self.fused:
            return
        if self.peer is not None:
            self.peer.allreduce_async(self.out2.data_ptr(), 2)
        else:
            all_reduce_sum(self.out2, self.group)

    def capture(self):
        """Capture the kernel runs between the collectives as two CUDA graphs; `step()` then issues
        graph, all-reduce, graph, all-reduce instead of ~10 separate launches."""
        self.step()                                    # first use allocates layouts
        self.torch.cuda.synchronize(self.device)
        if self.peer is not None:                      # our exchange kernels are plain launches: the whole step is ONE graph
            if self.closure:
                self._g_all = self.engine.capture(self.closure_step)
                return
            self._g_all = self.engine.capture(lambda: (self.fit_local(), self.exchange(), self.fit_finish(), self.mae_local(),
                                                       self.mae_exchange()))
            return
        self._g1 = self.engine.capture(self.fit_local)
        self._g2 = self.engine.capture(lambda: (self.fit_finish(), self.mae_local()))

    def closure_step(self):
        """fit + both exchanges + MAE of this rank's shard as ONE call (three kernels)."""
        E = self.E
        E._check(E.lib().mrs_fit_mae_push_async(self.engine._h, self.train._h, E.C.byref(self.model._h), self.test._h, self.peer._h,
                                                self.peer_small._h, E.C.c_void_p(self.known.data_ptr()), int(self.known.numel()),
                                                E.C.c_void_p(self.out2.data_ptr())))

    def step(self):
        if getattr(self, "_g_all", None) is not None:
            self._g_all.launch()
        elif self.closure:
            self.closure_step()
        elif getattr(self, "_g1", None) is not None:
            self._g1.launch(); self.exchange(); self._g2.launch(); self.mae_exchange()
        else:
            self.fit_local(); self.exchange(); self.fit_finish(); self.mae_local(); self.mae_exchange()

    def fit(self):
        """Enqueue: local pass -> all-reduce of the exchange buffer -> finish (no host sync)."""
        self.fit_local(); self.exchange(); self.fit_finish()

    def mae_async(self):
        self.mae_local(); self.mae_exchange()

    def close(self, close_peer=False):
        """Release the model (and, if asked, the exchange object -- only the pass that created it should)."""
        if self.model is not None:
            self.model.close()
            self.model = None
        if close_peer and self.peer is not None:
            self.peer.close()
            self.peer = None
        if close_peer and self.peer_small is not None:
            self.peer_small.close()
            self.peer_small = None

    def check(self):
        """Raise if a peer-memory exchange of this pass timed out (host sync).  Call it wherever a result is read."""
        if self.peer is not None:
            self.peer.check()
        if self.peer_small is not None:
            self.peer_small.check()

    def result(self):
        """{sum |err|, n} of the last step as an MAE (host sync); raises if an exchange timed out."""
        r = self.out2.cpu().numpy()
        self.check()
        return float(r[0] / r[1])

    def mae(self):
        self.mae_async()
        return self.result()


class ShardedKnn:
    """kNN predictor over a box of GPUs (BASELINE config 5): similarity ROWS are sharded, the ratings are not.

    Every rank holds the whole train set (240 MB at ml-25m shape against 180 GB of HBM: re-deriving the normalised
    matrix locally costs 0.2 ms, less than all-gathering it) and fits the baseline model on it -- the kernels are
    deterministic, so all ranks hold bit-identical averages and r~ without any exchange.  A rank computes the neighbour
    lists of its user range only, evaluates the test pairs of those users, and the ranks add {sum |err|, n}: 16 bytes,
    the only collective of the path (our peer-memory kernel when ``peer_exchange``, else ``torch.distributed``).

    ``train`` / ``test`` are host COO triples (users, items, ratings); ``rank`` / ``world`` default to the process group."""

    def __init__(self, engine, train, test, k=300, kind=None, rank=None, world=None, group=None, peer_exchange=False):
        import torch
        from . import engine as E
        self.E, self.torch, self.group = E, torch, group
        dist = torch.distributed
        live = dist.is_available() and dist.is_initialized()
        self.world = int(world) if world is not None else (dist.get_world_size(group) if live else 1)
        self.rank = int(rank) if rank is not None else (dist.get_rank(group) if live else 0)
        self.engine, self.k = engine, int(k)
        self.kind = E.SIM_COSINE if kind is None else int(kind)
        tu = np.asarray(train[0])
        n_users_dim = int(max(tu.max(initial=0), np.asarray(test[0]).max(initial=0))) + 1
        self.bounds = partition_rows(np.bincount(tu, minlength=n_users_dim), self.world)
        self.user_lo, self.user_hi = int(self.bounds[self.rank]), int(self.bounds[self.rank + 1])
        self.train = engine.ratings(*train)
        self.model = E.Model(engine, self.train)
        mask = shard_of(test[0], self.bounds, self.rank)
        self.test = engine.ratings(np.asarray(test[0])[mask], np.asarray(test[1])[mask], np.asarray(test[2])[mask])
        self.n_test_local = int(mask.sum())
        self.sim = None
        self.device = torch.device("cuda", engine.device) if torch.cuda.is_available() else None
        self.out2 = torch.zeros(2, dtype=torch.float64, device=self.device)
        self.peer = None
        if peer_exchange and live and self.world > 1:
            def gather(b):
                out = [None] * self.world
                dist.all_gather_object(out, b, group=group)
                return out
            self.peer = E.PeerExchange(engine, 2, self.rank, self.world, gather)

    def fit(self):
        """Enqueue deviations, r~, the similarity rows of this rank's users and their selection (no host sync)."""
        self.model.refit()
        if self.sim is None:
            self.sim = self.E.Sim(self.model, self.kind, self.k, sync=False, rows=(self.user_lo, self.user_hi))
        else:
            self.sim.refit()

    def mae_local(self):
        self.model.mae_async(self.test, self.out2.data_ptr(), self.E.PRED_PERSONALIZED, self.sim)

    def mae_exchange(self):
        if self.peer is not None:
            self.peer.allreduce_async(self.out2.data_ptr(), 2)
        else:
            all_reduce_sum(self.out2, self.group)

    def step(self):
        self.fit(); self.mae_local(); self.mae_exchange()

    def mae(self):
        r = self.out2.cpu().numpy()
        if self.peer is not None:
            self.peer.check()
        return float(r[0] / r[1])

    def close(self):
        for h in (self.sim, self.test, self.model, self.train):
            if h is not None:
                h.close()
