"""ctypes binding of libmrs_b200.so (the C ABI in include/mrs_b200.h).

There is deliberately no fallback: if the shared library is missing or no CUDA device is present,
everything here raises.  PyTorch is not imported; pass ``stream=torch.cuda.current_stream().cuda_stream``
to run on torch's stream when torch owns the timing events or the NCCL collectives.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MRS_LIB") or os.path.join(_HERE, "libmrs_b200.so")   # MRS_LIB: another build of the same library (A/B timing)

# mrs_vec_kind / mrs_pred_kind / mrs_sim_kind
GLOBAL_AVG, USER_AVG, ITEM_AVG, ITEM_AVG_DEV = range(4)
PRED_GLOBAL, PRED_USER, PRED_ITEM, PRED_ITEMDEV, PRED_BASELINE, PRED_PERSONALIZED, PRED_WSD = range(7)
SIM_UNIFORM, SIM_COSINE, SIM_JACCARD = range(3)

ERR_NAMES = {-1: "MRS_ERR_INVALID", -2: "MRS_ERR_CUDA", -3: "MRS_ERR_NOMEM", -4: "MRS_ERR_DUPLICATE",
             -5: "MRS_ERR_IO", -6: "MRS_ERR_UNSUPPORTED"}

EXPORTS = [
    "mrs_last_error", "mrs_version", "mrs_launch_count", "mrs_engine_create", "mrs_engine_destroy", "mrs_engine_sync", "mrs_debug_timeline", "mrs_debug_cta_stamps", "mrs_debug_warp_stamps", "mrs_debug_fp64_fma_per_s",
    "mrs_graph_begin", "mrs_graph_end", "mrs_graph_launch", "mrs_graph_destroy", "mrs_profile_begin", "mrs_profile_end", "mrs_upload_begin", "mrs_upload_begin_codes", "mrs_ratings_from_coo_codes", "mrs_ratings_from_upload", "mrs_upload_destroy", "mrs_ratings_from_coo", "mrs_ratings_from_file", "mrs_ratings_from_text", "mrs_ratings_info", "mrs_ratings_bytes", "mrs_ratings_layout_info", "mrs_ratings_destroy",
    "mrs_fit", "mrs_fit_local", "mrs_fit_async", "mrs_fit_users_async", "mrs_model_set_item_averages", "mrs_model_exchange_buffer", "mrs_fit_finish", "mrs_model_destroy", "mrs_exchange_create", "mrs_exchange_connect", "mrs_exchange_connect_local", "mrs_multi_create", "mrs_multi_load", "mrs_multi_baseline_mae", "mrs_multi_model", "mrs_multi_owner", "mrs_multi_destroy", "mrs_exchange_allreduce_async", "mrs_exchange_allreduce_indexed_async", "mrs_fit_local_push", "mrs_fit_finish_pull", "mrs_mae_push_async", "mrs_exchange_status", "mrs_exchange_set_timeout_ms", "mrs_exchange_stamps", "mrs_exchange_destroy",
    "mrs_model_scalar",
    "mrs_model_lookup", "mrs_model_vector", "mrs_fit_similarity", "mrs_fit_similarity_async", "mrs_fit_similarity_rows_async", "mrs_model_set_tie_order", "mrs_sim_set_k",
    "mrs_similarity", "mrs_neighbors", "mrs_sim_entry_values", "mrs_sim_destroy", "mrs_predict", "mrs_mae",
    "mrs_mae_async", "mrs_fit_mae_async", "mrs_fit_mae_push_async", "mrs_recommend",
]


class MrsError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"{ERR_NAMES.get(status, status)}: {message}")
        self.status = status


def build_library(force=False):
    """Compile csrc/*.cu for sm_100a with nvcc (works without a GPU) into libmrs_b200.so, in-tree."""
    src = os.path.join(_HERE, "csrc")
    newest = max(os.path.getmtime(os.path.join(src, f)) for f in os.listdir(src) if f.endswith((".cu", ".cuh")))
    newest = max(newest, os.path.getmtime(os.path.join(_HERE, "..", "include", "mrs_b200.h")))
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < newest:
        subprocess.run(["make", "-C", src, "-j8"], check=True, stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def lib():
    """Load libmrs_b200.so; raises if it has not been built (no silent fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MrsError(-6, f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    P = C.POINTER
    sig = {
        "mrs_last_error": (C.c_char_p, []),
        "mrs_version": (C.c_char_p, []),
        "mrs_launch_count": (i64, []),
        "mrs_engine_create": (i32, [i32, vp, P(vp)]),
        "mrs_engine_destroy": (None, [vp]),
        "mrs_engine_sync": (i32, [vp]),
        "mrs_debug_timeline": (i32, [vp, vp]),
        "mrs_debug_cta_stamps": (i32, [vp, vp]),
        "mrs_debug_warp_stamps": (i32, [vp, vp]),
        "mrs_debug_fp64_fma_per_s": (i32, [vp, P(dbl)]),
        "mrs_graph_begin": (i32, [vp]),
        "mrs_graph_end": (i32, [vp, P(vp)]),
        "mrs_graph_launch": (i32, [vp]),
        "mrs_graph_destroy": (None, [vp]),
        "mrs_profile_begin": (i32, [vp]),
        "mrs_profile_end": (i32, [vp, C.c_char_p, i64, P(C.c_float), i32, P(i32)]),
        "mrs_ratings_from_coo": (i32, [vp, vp, vp, vp, i64, i32, i32, P(vp)]),
        "mrs_upload_begin": (i32, [vp, vp, vp, vp, i64, P(vp)]),
        "mrs_upload_begin_codes": (i32, [vp, vp, vp, vp, i64, P(vp)]),
        "mrs_ratings_from_coo_codes": (i32, [vp, vp, vp, vp, i64, i32, i32, P(vp)]),
        "mrs_ratings_from_upload": (i32, [vp, i32, i32, P(vp)]),
        "mrs_upload_destroy": (None, [vp]),
        "mrs_ratings_from_file": (i32, [vp, C.c_char_p, C.c_char_p, P(vp)]),
        "mrs_ratings_from_text": (i32, [vp, C.c_char_p, i64, C.c_char_p, P(vp)]),
        "mrs_ratings_info": (i32, [vp, P(i64), P(i32), P(i32), P(i32)]),
        "mrs_ratings_bytes": (i32, [vp, P(i64)]),
        "mrs_ratings_layout_info": (i32, [vp, P(i64)]),
        "mrs_ratings_destroy": (None, [vp]),
        "mrs_fit": (i32, [vp, vp, P(vp)]),
        "mrs_fit_local": (i32, [vp, vp, P(vp)]),
        "mrs_fit_async": (i32, [vp, vp, P(vp)]),
        "mrs_fit_users_async": (i32, [vp, vp, P(vp)]),
        "mrs_model_set_item_averages": (i32, [vp, i32]),
        "mrs_model_exchange_buffer": (i32, [vp, P(vp), P(i64)]),
        "mrs_fit_finish": (i32, [vp]),
        "mrs_model_destroy": (None, [vp]),
        "mrs_exchange_create": (i32, [vp, i64, i32, i32, vp, P(vp)]),
        "mrs_exchange_connect": (i32, [vp, vp]),
        "mrs_exchange_connect_local": (i32, [vp, i32]),
        "mrs_multi_create": (i32, [vp, i32, P(vp)]),
        "mrs_multi_load": (i32, [vp, vp, vp, vp, i64, vp, vp, vp, i64]),
        "mrs_multi_baseline_mae": (i32, [vp, P(dbl)]),
        "mrs_multi_model": (i32, [vp, i32, P(vp)]),
        "mrs_multi_owner": (i32, [vp, i32, P(i32)]),
        "mrs_multi_destroy": (None, [vp]),
        "mrs_exchange_allreduce_async": (i32, [vp, vp, i64]),
        "mrs_exchange_allreduce_indexed_async": (i32, [vp, vp, vp, i64]),
        "mrs_fit_local_push": (i32, [vp, vp, P(vp), vp, vp, i32]),
        "mrs_fit_finish_pull": (i32, [vp, vp]),
        "mrs_mae_push_async": (i32, [vp, vp, vp, vp]),
        "mrs_exchange_status": (i32, [vp, P(i32)]),
        "mrs_exchange_set_timeout_ms": (i32, [vp, i64]),
        "mrs_exchange_stamps": (i32, [vp, P(C.c_uint64)]),
        "mrs_exchange_destroy": (None, [vp]),
        "mrs_model_scalar": (i32, [vp, i32, P(dbl)]),
        "mrs_model_lookup": (i32, [vp, i32, i32, P(dbl), P(i32)]),
        "mrs_model_vector": (i32, [vp, i32, vp, vp, i64, P(i64)]),
        "mrs_fit_similarity": (i32, [vp, i32, i32, P(vp)]),
        "mrs_fit_similarity_async": (i32, [vp, i32, i32, P(vp)]),
        "mrs_fit_similarity_rows_async": (i32, [vp, i32, i32, i32, i32, P(vp)]),
        "mrs_model_set_tie_order": (i32, [vp, i32]),
        "mrs_sim_set_k": (i32, [vp, i32]),
        "mrs_similarity": (i32, [vp, i32, i32, P(dbl)]),
        "mrs_neighbors": (i32, [vp, i32, i32, vp, vp, i32, P(i32)]),
        "mrs_sim_entry_values": (i32, [vp, i32, vp, vp, vp, i64, P(i64)]),
        "mrs_sim_destroy": (None, [vp]),
        "mrs_predict": (i32, [vp, vp, i32, vp, vp, i64, vp]),
        "mrs_mae": (i32, [vp, vp, i32, vp, P(dbl)]),
        "mrs_mae_async": (i32, [vp, vp, i32, vp, vp]),
        "mrs_fit_mae_async": (i32, [vp, vp, P(vp), vp, vp]),
        "mrs_fit_mae_push_async": (i32, [vp, vp, P(vp), vp, vp, vp, vp, i32, vp]),
        "mrs_recommend": (i32, [vp, vp, i32, i32, i32, vp, vp, P(i32)]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    _lib = L
    return L


def _check(status):
    if status != 0:
        raise MrsError(status, lib().mrs_last_error().decode("utf-8", "replace"))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def launch_count():
    return int(lib().mrs_launch_count())


class Engine:
    """One CUDA device + one stream."""

    def __init__(self, device=0, stream=None):
        self._h = C.c_void_p()
        _check(lib().mrs_engine_create(int(device), C.c_void_p(stream) if stream else None, C.byref(self._h)))
        self.device = int(device)

    def sync(self):
        _check(lib().mrs_engine_sync(self._h))

    def fp64_fma_per_s(self):
        """Measured fp64 FMA rate of this device (FMA/s): the denominator of the kNN similarity rooflines."""
        out = C.c_double()
        _check(lib().mrs_debug_fp64_fma_per_s(self._h, C.byref(out)))
        return out.value

    def capture(self, fn):
        """Run ``fn()`` (asynchronous engine calls only) under CUDA-graph capture and return a replayable Graph."""
        _check(lib().mrs_graph_begin(self._h))
        try:
            fn()
        finally:
            h = C.c_void_p()
            status = lib().mrs_graph_end(self._h, C.byref(h))
        _check(status)
        return Graph(h)

    def profile_begin(self):
        _check(lib().mrs_profile_begin(self._h))

    def profile_end(self):
        """[(kernel label, milliseconds)] for every launch since profile_begin (CUDA events on the engine's stream)."""
        cap = 4096
        names = C.create_string_buffer(cap * 24)
        ms = (C.c_float * cap)()
        n = C.c_int32()
        _check(lib().mrs_profile_end(self._h, names, len(names), ms, cap, C.byref(n)))
        labels = names.value.decode().split("\n")[:n.value]
        return [(labels[j], float(ms[j])) for j in range(min(n.value, cap))]

    def close(self):
        if getattr(self, "_h", None):
            lib().mrs_engine_destroy(self._h)
            self._h = None

    def ratings(self, users, items, ratings, n_users_dim=0, n_items_dim=0):
        return Ratings(self, users, items, ratings, n_users_dim, n_items_dim)

    def upload_codes(self, users, items, codes):
        """Staged upload of the compact form: ``codes`` = 2 x rating as uint8 (9 bytes per rating over PCIe instead of 16)."""
        return Upload(self, users, items, codes, codes=True)

    def ratings_from_codes(self, users, items, codes, n_users_dim=0, n_items_dim=0):
        return self.upload_codes(users, items, codes).ratings(n_users_dim, n_items_dim)

    def upload(self, users, items, ratings):
        """Start the host -> device copies of a rating set on the engine's copy stream and return at once; build the set
        with ``Upload.ratings()``.  Lets the copies of a second set (test) run while the first (train) is being built."""
        return Upload(self, users, items, ratings)

    def ratings_from_text(self, text, sep):
        return Ratings.from_text(self, text, sep)

    def ratings_from_file(self, path, sep):
        return Ratings.from_file(self, path, sep)


class MultiEngine:
    """Several GPUs driven from this one process (``mrs_multi_*``): users sharded over the devices, one exchange."""

    def __init__(self, device_ids):
        ids = np.ascontiguousarray(device_ids, dtype=np.int32)
        self._h = C.c_void_p()
        _check(lib().mrs_multi_create(_ptr(ids), ids.size, C.byref(self._h)))
        self.n = int(ids.size)

    def load(self, train, test):
        a = [np.ascontiguousarray(x, dtype=t) for x, t in zip((*train, *test), (np.int32, np.int32, np.float64) * 2)]
        _check(lib().mrs_multi_load(self._h, _ptr(a[0]), _ptr(a[1]), _ptr(a[2]), a[0].size, _ptr(a[3]), _ptr(a[4]), _ptr(a[5]), a[3].size))

    def baseline_mae(self):
        out = C.c_double()
        _check(lib().mrs_multi_baseline_mae(self._h, C.byref(out)))
        return out.value

    def model(self, slot):
        """The model of one device slot as a (borrowed) Model for the query methods."""
        h = C.c_void_p()
        _check(lib().mrs_multi_model(self._h, int(slot), C.byref(h)))
        m = Model.__new__(Model)
        m.engine, m.train, m._h = None, None, h
        m.close = lambda: None          # owned by the multi handle
        return m

    def owner(self, user):
        s = C.c_int32()
        _check(lib().mrs_multi_owner(self._h, int(user), C.byref(s)))
        return s.value

    def close(self):
        if getattr(self, "_h", None):
            lib().mrs_multi_destroy(self._h)
            self._h = None


class PeerExchange:
    """Sum fp64 device buffers across the ranks of one box with the library's own NVLink peer-memory kernel.

    ``all_gather(bytes) -> list[bytes]`` is any host-side transport of the 64-byte IPC handles (e.g.
    ``torch.distributed.all_gather_object``)."""

    def __init__(self, engine, n_doubles, rank, world, all_gather):
        self._h = C.c_void_p()
        handle = C.create_string_buffer(64)
        _check(lib().mrs_exchange_create(engine._h, int(n_doubles), int(rank), int(world), handle, C.byref(self._h)))
        handles = all_gather(handle.raw)
        assert len(handles) == world and all(len(h) == 64 for h in handles)
        blob = C.create_string_buffer(b"".join(handles), 64 * world)
        _check(lib().mrs_exchange_connect(self._h, blob))

    def allreduce_async(self, device_ptr, n_doubles):
        _check(lib().mrs_exchange_allreduce_async(self._h, C.c_void_p(int(device_ptr)), int(n_doubles)))

    def allreduce_indexed_async(self, device_ptr, device_idx_ptr, n_idx):
        """Sum only the positions ``idx[0..n_idx)`` (int32 on the device, identical on every rank) of the buffer."""
        _check(lib().mrs_exchange_allreduce_indexed_async(self._h, C.c_void_p(int(device_ptr)), C.c_void_p(int(device_idx_ptr)), int(n_idx)))

    def stamps(self):
        """Nanosecond stamps of the last exchange, relative to its start: published, barrier 1, reduced, barrier 2, done."""
        o = (C.c_uint64 * 8)()
        _check(lib().mrs_exchange_stamps(self._h, o))
        return [int(o[k]) - int(o[0]) if o[k] >= o[0] and o[k] else None for k in range(1, 6)]  # older = left by a previous call

    def timed_out(self):
        """True if a peer never arrived in some exchange since creation (synchronises the stream).  The kernel has then
        overwritten the exchanged buffer with NaN, and the handle refuses further exchanges."""
        t = C.c_int32()
        _check(lib().mrs_exchange_status(self._h, C.byref(t)))
        return bool(t.value)

    def check(self):
        """Raise if an exchange timed out (call where the host synchronises anyway: after reading a result)."""
        if self.timed_out():
            raise MrsError(-2, "peer-memory exchange timed out: a rank did not arrive; results are NaN and the handle is dead")

    def set_timeout_ms(self, ms):
        _check(lib().mrs_exchange_set_timeout_ms(self._h, int(ms)))

    def close(self):
        if getattr(self, "_h", None):
            lib().mrs_exchange_destroy(self._h)
            self._h = None


class Graph:
    """A captured sequence of kernels (cudaGraphExec); ``launch()`` replays it on the engine's stream."""

    def __init__(self, handle):
        self._h = handle

    def launch(self):
        _check(lib().mrs_graph_launch(self._h))

    def close(self):
        if getattr(self, "_h", None):
            lib().mrs_graph_destroy(self._h)
            self._h = None


class Upload:
    """A rating set on its way to the device (``mrs_upload_begin``); the host arrays are kept alive until it is consumed."""

    def __init__(self, engine, users, items, ratings, codes=False):
        self.engine = engine
        u = np.ascontiguousarray(users, dtype=np.int32)
        i = np.ascontiguousarray(items, dtype=np.int32)
        r = np.ascontiguousarray(ratings, dtype=np.uint8 if codes else np.float64)
        if not (u.shape == i.shape == r.shape and u.ndim == 1):
            raise ValueError("users, items, ratings must be 1-D arrays of equal length")
        self._keep = (u, i, r)
        self._h = C.c_void_p()
        begin = lib().mrs_upload_begin_codes if codes else lib().mrs_upload_begin
        _check(begin(engine._h, _ptr(u), _ptr(i), _ptr(r), u.size, C.byref(self._h)))

    def ratings(self, n_users_dim=0, n_items_dim=0):
        h, self._h = self._h, None
        out = C.c_void_p()
        try:
            _check(lib().mrs_ratings_from_upload(h, int(n_users_dim), int(n_items_dim), C.byref(out)))   # consumes the upload
        finally:
            self._keep = None
        return Ratings(self.engine, None, None, None, _handle=out)

    def close(self):
        if getattr(self, "_h", None):
            lib().mrs_upload_destroy(self._h)
            self._h = None
            self._keep = None


class Ratings:
    """Device-resident rating set (user-major CSR + item-major CSC + sorted COO)."""

    def __init__(self, engine, users, items, ratings, n_users_dim=0, n_items_dim=0, _handle=None):
        self.engine = engine
        self._h = C.c_void_p()
        if _handle is not None:
            self._h = _handle
        else:
            u = np.ascontiguousarray(users, dtype=np.int32)
            i = np.ascontiguousarray(items, dtype=np.int32)
            r = np.ascontiguousarray(ratings, dtype=np.float64)
            if not (u.shape == i.shape == r.shape and u.ndim == 1):
                raise ValueError("users, items, ratings must be 1-D arrays of equal length")
            _check(lib().mrs_ratings_from_coo(engine._h, _ptr(u), _ptr(i), _ptr(r), u.size, int(n_users_dim), int(n_items_dim),
                                              C.byref(self._h)))
        n, nu, ni, vk = C.c_int64(), C.c_int32(), C.c_int32(), C.c_int32()
        _check(lib().mrs_ratings_info(self._h, C.byref(n), C.byref(nu), C.byref(ni), C.byref(vk)))
        self.n, self.n_users_dim, self.n_items_dim, self.value_kind = n.value, nu.value, ni.value, vk.value

    @classmethod
    def from_file(cls, engine, path, sep):
        h = C.c_void_p()
        _check(lib().mrs_ratings_from_file(engine._h, os.fsencode(path), sep.encode(), C.byref(h)))
        return cls(engine, None, None, None, _handle=h)

    @classmethod
    def from_text(cls, engine, text, sep):
        """Parse ``text`` (bytes) on the device with the rules of the reference's ``load`` (P:35-49)."""
        data = bytes(text)
        h = C.c_void_p()
        _check(lib().mrs_ratings_from_text(engine._h, data, len(data), sep.encode(), C.byref(h)))
        return cls(engine, None, None, None, _handle=h)

    def bytes(self):
        b = (C.c_int64 * 3)()
        _check(lib().mrs_ratings_bytes(self._h, b))
        return {"user_major": b[0], "item_major": b[1], "sorted_coo": b[2]}

    def layout_info(self):
        o = (C.c_int64 * 8)()
        _check(lib().mrs_ratings_layout_info(self._h, o))
        keys = ("user_tiles", "units", "slices", "tiled_slots", "item_tiles", "mae_chunks", "mae_slots", "code_vectors")
        return dict(zip(keys, [int(x) for x in o]))

    def __len__(self):
        return self.n

    def close(self):
        if getattr(self, "_h", None):
            lib().mrs_ratings_destroy(self._h)
            self._h = None


class Model:
    """Fitted baseline model of a train set (global / user / item averages, item average deviations)."""

    def __init__(self, engine, train, sync=True):
        self.engine, self.train = engine, train
        self._h = C.c_void_p()
        if sync:
            _check(lib().mrs_fit(engine._h, train._h, C.byref(self._h)))
        else:
            self.refit()

    def refit(self, between=None):
        """Enqueue the fit again on the same buffers (no host sync). ``between(ptr, n_doubles)`` is called after the
        local pass with the device exchange buffer -- a sharded run all-reduces it there."""
        if between is None:
            _check(lib().mrs_fit_async(self.engine._h, self.train._h, C.byref(self._h)))
            return
        _check(lib().mrs_fit_local(self.engine._h, self.train._h, C.byref(self._h)))
        between(*self.exchange_buffer())
        _check(lib().mrs_fit_finish(self._h))

    def refit_users(self):
        """Enqueue a users-only refit (user averages + global average): what the personalized / kNN predictors take from
        the fit (P:557-586); the per-item average deviations keep the values of the last full fit."""
        _check(lib().mrs_fit_users_async(self.engine._h, self.train._h, C.byref(self._h)))

    def set_item_averages(self, enabled):
        """Per-item rating averages are not needed by the baseline predictor; switching them off saves work in the fit."""
        _check(lib().mrs_model_set_item_averages(self._h, 1 if enabled else 0))

    def set_tie_order(self, mode):
        """0: neighbours with exactly equal similarity in ascending user id; 1: in the Scala 2.11 HashSet order the reference's
        stable sort keeps.  Applies to similarities fitted afterwards."""
        _check(lib().mrs_model_set_tie_order(self._h, int(mode)))

    def exchange_buffer(self):
        p, n = C.c_void_p(), C.c_int64()
        _check(lib().mrs_model_exchange_buffer(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    @property
    def global_avg(self):
        out = C.c_double()
        _check(lib().mrs_model_scalar(self._h, GLOBAL_AVG, C.byref(out)))
        return out.value

    def lookup(self, kind, ident):
        out, known = C.c_double(), C.c_int32()
        _check(lib().mrs_model_lookup(self._h, kind, int(ident), C.byref(out), C.byref(known)))
        return out.value, bool(known.value)

    def vector(self, kind):
        n = C.c_int64()
        _check(lib().mrs_model_vector(self._h, kind, None, None, 0, C.byref(n)))
        vals = np.empty(n.value, dtype=np.float64)
        cnts = np.empty(n.value, dtype=np.int32)
        _check(lib().mrs_model_vector(self._h, kind, _ptr(vals), _ptr(cnts), n.value, C.byref(n)))
        return vals, cnts

    def predict(self, users, items, kind=PRED_BASELINE, sim=None):
        u = np.ascontiguousarray(users, dtype=np.int32)
        i = np.ascontiguousarray(items, dtype=np.int32)
        out = np.empty(u.size, dtype=np.float64)
        _check(lib().mrs_predict(self._h, sim._h if sim is not None else None, int(kind), _ptr(u), _ptr(i), u.size, _ptr(out)))
        return out

    def mae(self, test, kind=PRED_BASELINE, sim=None):
        out = C.c_double()
        _check(lib().mrs_mae(self._h, sim._h if sim is not None else None, int(kind), test._h, C.byref(out)))
        return out.value

    def mae_async(self, test, device_out_ptr, kind=PRED_BASELINE, sim=None):
        """Enqueue predict+|err| reduction; {sum, count} (2 fp64) land at ``device_out_ptr``; no host sync."""
        _check(lib().mrs_mae_async(self._h, sim._h if sim is not None else None, int(kind), test._h, C.c_void_p(device_out_ptr)))

    def fit_mae_async(self, test, device_out_ptr):
        """The closure MeanAbsoluteErrorSpark(baselinePredictorSpark(train), test) in one call: refit + fused MAE, the test pass
        finishing the fit itself (three kernels); {sum |err|, count} land at ``device_out_ptr``; no host sync."""
        _check(lib().mrs_fit_mae_async(self.engine._h, self.train._h, C.byref(self._h), test._h, C.c_void_p(device_out_ptr)))

    def similarity(self, kind=SIM_COSINE, k=0, sync=True, rows=None):
        return Sim(self, kind, k, sync=sync, rows=rows)

    def recommend(self, user, n, kind=PRED_PERSONALIZED, sim=None):
        items = np.empty(max(n, 1), dtype=np.int32)
        scores = np.empty(max(n, 1), dtype=np.float64)
        w = C.c_int32()
        _check(lib().mrs_recommend(self._h, sim._h if sim is not None else None, int(kind), int(user), int(n), _ptr(items),
                                   _ptr(scores), C.byref(w)))
        return items[:w.value].copy(), scores[:w.value].copy()

    def close(self):
        if getattr(self, "_h", None):
            lib().mrs_model_destroy(self._h)
            self._h = None


class Sim:
    """User-user similarity of a fitted model (uniform | cosine | jaccard), optionally restricted to k neighbours."""

    def __init__(self, model, kind=SIM_COSINE, k=0, sync=True, rows=None):
        """``rows=(user_lo, user_hi)`` selects the row-block path: neighbour lists (first k) only for users with an id in
        that range -- the rows a rank of a sharded run owns."""
        self.model, self.kind, self.k = model, int(kind), int(k)
        self.rows = None if rows is None else (int(rows[0]), int(rows[1]))
        self._h = C.c_void_p()
        if sync and self.rows is None:
            _check(lib().mrs_fit_similarity(model._h, self.kind, self.k, C.byref(self._h)))
        else:
            self.refit()
            if sync:
                model.engine.sync()

    def refit(self, k=None):
        if k is not None:
            self.k = int(k)
        if self.rows is not None:
            _check(lib().mrs_fit_similarity_rows_async(self.model._h, self.kind, self.k, self.rows[0], self.rows[1], C.byref(self._h)))
        else:
            _check(lib().mrs_fit_similarity_async(self.model._h, self.kind, self.k, C.byref(self._h)))

    def set_k(self, k):
        self.k = int(k)
        _check(lib().mrs_sim_set_k(self._h, self.k))

    def __call__(self, u, v):
        out = C.c_double()
        _check(lib().mrs_similarity(self._h, int(u), int(v), C.byref(out)))
        return out.value

    def neighbors(self, u, k):
        cap = max(int(k), 1)
        ids = np.empty(cap, dtype=np.int32)
        sims = np.empty(cap, dtype=np.float64)
        w = C.c_int32()
        _check(lib().mrs_neighbors(self._h, int(u), int(k), _ptr(ids), _ptr(sims), cap, C.byref(w)))
        return ids[:w.value].copy(), sims[:w.value].copy()

    def entry_values(self, which):
        n = C.c_int64()
        _check(lib().mrs_sim_entry_values(self._h, int(which), None, None, None, 0, C.byref(n)))
        u = np.empty(n.value, dtype=np.int32)
        i = np.empty(n.value, dtype=np.int32)
        v = np.empty(n.value, dtype=np.float64)
        _check(lib().mrs_sim_entry_values(self._h, int(which), _ptr(u), _ptr(i), _ptr(v), n.value, C.byref(n)))
        return u, i, v

    def close(self):
        if getattr(self, "_h", None):
            lib().mrs_sim_destroy(self._h)
            self._h = None
