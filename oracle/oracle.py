"""ctypes binding of oracle/libmrs_oracle.so (the C restatement of predictions.scala).

TEST INFRASTRUCTURE ONLY.  Parity status: "partially pinned" -- see mrs_oracle.c.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

GLOBAL, USER, ITEM, ITEMDEV, BASELINE, PERSONALIZED = range(6)
SIM_UNIFORM, SIM_COSINE, SIM_JACCARD = range(3)

_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def build_oracle(force=False):
    """Compile the oracle with its Makefile (gcc, -ffp-contract=off). Building is not using."""
    so = os.path.join(_HERE, "libmrs_oracle.so")
    src = os.path.join(_HERE, "mrs_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
    return so


def _lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    so = build_oracle()
    L = C.CDLL(so)
    vp = C.c_void_p
    sig = {
        "orc_scale": (C.c_double, [C.c_double, C.c_double]),
        "orc_combine": (C.c_double, [C.c_double, C.c_double]),
        "orc_mean": (C.c_double, [_f64p, C.c_int64]),
        "orc_std": (C.c_double, [_f64p, C.c_int64]),
        "orc_fit": (vp, [_i32p, _i32p, _f64p, C.c_int64]),
        "orc_free": (None, [vp]),
        "orc_global_avg": (C.c_double, [vp]),
        "orc_umax": (C.c_int32, [vp]),
        "orc_imax": (C.c_int32, [vp]),
        "orc_user_count": (C.c_int32, [vp, C.c_int32]),
        "orc_item_count": (C.c_int32, [vp, C.c_int32]),
        "orc_user_avg": (C.c_double, [vp, C.c_int32]),
        "orc_item_avg": (C.c_double, [vp, C.c_int32]),
        "orc_item_avg_dev": (C.c_double, [vp, C.c_int32]),
        "orc_user_norm": (C.c_double, [vp, C.c_int32]),
        "orc_deviations": (None, [vp, _f64p]),
        "orc_pair_values": (C.c_int32, [vp, C.c_int32, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
        "orc_cosine": (C.c_double, [vp, C.c_int32, C.c_int32]),
        "orc_jaccard": (C.c_double, [vp, C.c_int32, C.c_int32]),
        "orc_neighbors": (C.c_int32, [vp, C.c_int, C.c_int32, C.c_int32, _i32p, _f64p, C.c_int32]),
        "orc_similarity": (C.c_double, [vp, C.c_int, C.c_int32, C.c_int32, C.c_int32]),
        "orc_wsd": (C.c_double, [vp, C.c_int, C.c_int32, C.c_int32, C.c_int32]),
        "orc_predict": (C.c_double, [vp, C.c_int, C.c_int, C.c_int32, C.c_int32, C.c_int32]),
        "orc_predict_batch": (None, [vp, C.c_int, C.c_int, C.c_int32, _i32p, _i32p, C.c_int64, _f64p]),
        "orc_mae": (C.c_double, [vp, C.c_int, C.c_int, C.c_int32, _i32p, _i32p, _f64p, C.c_int64]),
        "orc_recommend": (C.c_int32, [vp, C.c_int, C.c_int, C.c_int32, C.c_int32, C.c_int32, _i32p, _f64p]),
        "orc_baseline_mae_spark": (C.c_double, [_i32p, _i32p, _f64p, C.c_int64, _i32p, _i32p, _f64p, C.c_int64,
                                                C.c_int32, C.POINTER(C.c_double)]),
        "orc_max_threads": (C.c_int32, []),
        "orc_set_tie_order": (None, [C.c_void_p, C.c_int32]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _LIB = L
    return L


def _a32(x):
    return np.ascontiguousarray(x, dtype=np.int32)


def _a64(x):
    return np.ascontiguousarray(x, dtype=np.float64)


def scale(x, y):
    return _lib().orc_scale(float(x), float(y))


def combine(avg, dev):
    """avg + dev * scale(avg + dev, avg)  (P:229, P:383, P:578)"""
    return _lib().orc_combine(float(avg), float(dev))


def mean(xs):
    a = _a64(xs)
    return _lib().orc_mean(a, a.size)


def std(xs):
    a = _a64(xs)
    return _lib().orc_std(a, a.size)


def max_threads():
    return int(_lib().orc_max_threads())


def spark_baseline_mae(train, test, nthreads=1):
    """MeanAbsoluteErrorSpark(baselinePredictorSpark(train), test), partitions = threads."""
    u, i, r = _a32(train[0]), _a32(train[1]), _a64(train[2])
    tu, ti, tr = _a32(test[0]), _a32(test[1]), _a64(test[2])
    g = C.c_double(0.0)
    mae = _lib().orc_baseline_mae_spark(u, i, r, u.size, tu, ti, tr, tu.size, int(nthreads), C.byref(g))
    return mae, g.value


class Oracle:
    """A fitted train set. Method names follow predictions.scala."""

    def __init__(self, users, items, ratings):
        self._L = _lib()
        u, i, r = _a32(users), _a32(items), _a64(ratings)
        assert u.shape == i.shape == r.shape
        self.n = int(u.size)
        self._h = self._L.orc_fit(u, i, r, self.n)

    def close(self):
        if self._h:
            self._L.orc_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- baseline family
    @property
    def global_avg(self):
        return self._L.orc_global_avg(self._h)

    @property
    def umax(self):
        return self._L.orc_umax(self._h)

    @property
    def imax(self):
        return self._L.orc_imax(self._h)

    def user_avg(self, u):
        return self._L.orc_user_avg(self._h, int(u))

    def item_avg(self, i):
        return self._L.orc_item_avg(self._h, int(i))

    def item_avg_dev(self, i):
        return self._L.orc_item_avg_dev(self._h, int(i))

    def user_count(self, u):
        return self._L.orc_user_count(self._h, int(u))

    def item_count(self, i):
        return self._L.orc_item_count(self._h, int(i))

    def user_norm(self, u):
        return self._L.orc_user_norm(self._h, int(u))

    def user_avg_vector(self):
        return np.array([self.user_avg(u) for u in range(self.umax + 1)])

    def item_avg_vector(self):
        return np.array([self.item_avg(i) for i in range(self.imax + 1)])

    def item_avg_dev_vector(self):
        return np.array([self.item_avg_dev(i) for i in range(self.imax + 1)])

    def deviations(self):
        out = np.empty(self.n, dtype=np.float64)
        self._L.orc_deviations(self._h, out)
        return out

    def pair_values(self, u, i):
        d, p = C.c_double(), C.c_double()
        ok = self._L.orc_pair_values(self._h, int(u), int(i), C.byref(d), C.byref(p))
        return (d.value, p.value) if ok else None

    # --- similarities / neighbours
    def cosine(self, u, v):
        return self._L.orc_cosine(self._h, int(u), int(v))

    def jaccard(self, u, v):
        return self._L.orc_jaccard(self._h, int(u), int(v))

    def set_tie_order(self, mode):
        """0: equal similarities in ascending user id; 1: in the Scala 2.11 HashSet iteration order."""
        self._L.orc_set_tie_order(self._h, int(mode))

    def neighbors(self, u, k, simkind=SIM_COSINE):
        cap = max(int(k), 1)
        ids = np.empty(cap, dtype=np.int32)
        sims = np.empty(cap, dtype=np.float64)
        w = self._L.orc_neighbors(self._h, simkind, int(k), int(u), ids, sims, cap)
        return ids[:w].copy(), sims[:w].copy()

    def similarity(self, u, v, simkind=SIM_COSINE, k=0):
        return self._L.orc_similarity(self._h, simkind, int(k), int(u), int(v))

    def wsd(self, u, i, simkind=SIM_COSINE, k=0):
        return self._L.orc_wsd(self._h, simkind, int(k), int(u), int(i))

    # --- predictors
    def predict(self, u, i, kind=BASELINE, simkind=SIM_UNIFORM, k=0):
        return self._L.orc_predict(self._h, kind, simkind, int(k), int(u), int(i))

    def predict_batch(self, us, is_, kind=BASELINE, simkind=SIM_UNIFORM, k=0):
        us, is_ = _a32(us), _a32(is_)
        out = np.empty(us.size, dtype=np.float64)
        self._L.orc_predict_batch(self._h, kind, simkind, int(k), us, is_, us.size, out)
        return out

    def mae(self, test, kind=BASELINE, simkind=SIM_UNIFORM, k=0):
        tu, ti, tr = _a32(test[0]), _a32(test[1]), _a64(test[2])
        return self._L.orc_mae(self._h, kind, simkind, int(k), tu, ti, tr, tu.size)

    def recommend(self, user, n, kind=PERSONALIZED, simkind=SIM_COSINE, k=300):
        items = np.empty(max(n, 1), dtype=np.int32)
        scores = np.empty(max(n, 1), dtype=np.float64)
        w = self._L.orc_recommend(self._h, kind, simkind, int(k), int(user), int(n), items, scores)
        return items[:w].copy(), scores[:w].copy()
