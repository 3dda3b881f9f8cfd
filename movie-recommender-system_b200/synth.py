"""Seeded MovieLens-shaped synthetic rating sets.

There is no network and the MovieLens files are not in the reference tree, so every test and
benchmark runs on these.  Pure numpy, deterministic for a given (numpy version, seed).

  ml100k(): U=943 (ids 1..943), I=1682 (ids 1..1682), 100,000 unique (u,i), r in {1..5},
            80,000 / 20,000 random split (u2.base / u2.test analogue), full set = u.data.
  ml25m():  U=162,541 (ids 1..U), I=59,047 distinct ids drawn from 1..209,171, N=25,000,095 unique
            pairs, r in {0.5,...,5.0}, 80/20 split (r2.train / r2.test analogue).
Writers emit the text formats the reference's ``load`` parses (P:35-49): ``u\\ti\\tr\\tts`` and ``u,i,r,ts``.
"""
import os

import numpy as np


def _generate(n_users, item_ids, n_ratings, half_star, seed, min_per_user=20, sigma=1.0, zipf_a=1.0,
              zipf_c=30.0, max_frac=0.45, rare_frac=0.08, mean_rating=3.53):
    rng = np.random.default_rng(seed)
    n_items = len(item_ids)
    # per-user activity: min + log-normal tail, scaled to the target total
    raw = rng.lognormal(mean=0.0, sigma=sigma, size=n_users)
    extra = n_ratings - min_per_user * n_users
    cnt = min_per_user + np.floor(raw / raw.sum() * extra).astype(np.int64)
    cnt = np.minimum(cnt, int(max_frac * n_items))
    # item popularity: shifted Zipf over a random permutation of the item slots, plus a set of
    # rare items (so that a few test items / users are unseen in train and exercise the fallbacks)
    rank = rng.permutation(n_items)
    pop = 1.0 / np.power(rank + zipf_c, zipf_a)
    pop[rng.random(n_items) < rare_frac] *= 0.04
    cdf = np.cumsum(pop)
    cdf /= cdf[-1]

    def draw(per_user):
        uu = np.repeat(np.arange(n_users, dtype=np.int64), per_user)
        ii = np.searchsorted(cdf, rng.random(uu.size, dtype=np.float32).astype(np.float64), side="right")
        np.minimum(ii, n_items - 1, out=ii)
        uu *= n_items
        uu += ii
        return uu

    def uniq(a):
        a.sort()
        keep = np.empty(a.size, dtype=bool)
        keep[0] = True
        np.not_equal(a[1:], a[:-1], out=keep[1:])
        return a[keep]

    keys = uniq(draw(cnt + (cnt // 6) + 2))
    # top up until there are enough unique pairs, then drop random extras to hit n_ratings exactly
    while keys.size < n_ratings:
        need = n_ratings - keys.size
        per = np.bincount(rng.integers(0, n_users, size=int(need * 1.5) + 64), minlength=n_users)
        keys = uniq(np.concatenate([keys, draw(per)]))
    if keys.size > n_ratings:
        drop = rng.choice(keys.size, size=keys.size - n_ratings, replace=False)
        keep = np.ones(keys.size, dtype=bool)
        keep[drop] = False
        keys = keys[keep]
    u = keys // n_items
    it = keys - u * n_items
    # ratings: global mean + user bias + item quality + rank-2 taste + noise
    bias = rng.normal(0.0, 0.45, n_users).astype(np.float32)
    qual = rng.normal(0.0, 0.5, n_items).astype(np.float32)
    fu = rng.normal(0.0, 0.6, (n_users, 2)).astype(np.float32)
    gi = rng.normal(0.0, 0.6, (n_items, 2)).astype(np.float32)
    x = bias[u] + qual[it]
    x += fu[u, 0] * gi[it, 0]
    x += fu[u, 1] * gi[it, 1]
    x += rng.standard_normal(u.size, dtype=np.float32) * np.float32(0.85)
    x += np.float32(mean_rating) - x.mean(dtype=np.float64).astype(np.float32)
    if half_star:
        r = np.clip(np.round(x * 2.0) / 2.0, 0.5, 5.0)
    else:
        r = np.clip(np.round(x), 1.0, 5.0)
    users = (u + 1).astype(np.int32)  # ids are 1-based like MovieLens
    items = np.asarray(item_ids, dtype=np.int32)[it]
    # shuffle into "file order" (u.data is not sorted)
    perm = rng.permutation(users.size)
    return users[perm], items[perm], r[perm].astype(np.float64), rng


def _split(users, items, ratings, n_test, rng):
    n = users.size
    is_test = np.zeros(n, dtype=bool)
    is_test[rng.choice(n, size=n_test, replace=False)] = True
    te, tr = np.flatnonzero(is_test), np.flatnonzero(~is_test)
    return (users[tr], items[tr], ratings[tr]), (users[te], items[te], ratings[te])


def ml100k(seed=449):
    """dict(train=(u,i,r), test=(u,i,r), all=(u,i,r)) with ml-100k u2.base/u2.test/u.data shape."""
    u, i, r, rng = _generate(943, np.arange(1, 1683), 100_000, half_star=False, seed=seed,
                             min_per_user=20, sigma=1.0, zipf_a=1.0, zipf_c=30.0, max_frac=0.45)
    train, test = _split(u, i, r, 20_000, rng)
    return {"train": train, "test": test, "all": (u, i, r), "n_users": 943, "n_items": 1682}


def ml25m(seed=449, n_users=162_541, n_items=59_047, n_ratings=25_000_095, max_item_id=209_171):
    """ml-25m r2.train/r2.test shape; item ids are sparse in 1..max_item_id. Arguments allow scaled copies."""
    rng0 = np.random.default_rng(seed + 1)
    item_ids = np.sort(rng0.choice(np.arange(1, max_item_id + 1), size=n_items, replace=False))
    u, i, r, rng = _generate(n_users, item_ids, n_ratings, half_star=True, seed=seed,
                             min_per_user=20, sigma=1.4, zipf_a=1.0, zipf_c=40.0, max_frac=0.55)
    n_test = n_ratings // 5
    train, test = _split(u, i, r, n_test, rng)
    return {"train": train, "test": test, "all": (u, i, r), "n_users": n_users, "n_items": n_items}


def small(seed=7, n_users=60, n_items=90, n_ratings=1200, half_star=True):
    """A tiny set for unit tests (every code path of the oracle finishes instantly)."""
    u, i, r, rng = _generate(n_users, np.arange(1, n_items + 1), n_ratings, half_star=half_star, seed=seed,
                             min_per_user=5, sigma=0.8, zipf_a=0.8, zipf_c=3.0, max_frac=0.8, rare_frac=0.15)
    train, test = _split(u, i, r, n_ratings // 5, rng)
    return {"train": train, "test": test, "all": (u, i, r), "n_users": n_users, "n_items": n_items}


def mid(seed=11, n_users=20_000, n_items=3_000, n_ratings=600_000):
    """More users than the dense-similarity path takes (16,384): exercises the row-block kNN path at a size the oracle
    still answers per user in milliseconds."""
    u, i, r, rng = _generate(n_users, np.arange(1, n_items + 1), n_ratings, half_star=True, seed=seed,
                             min_per_user=8, sigma=1.0, zipf_a=1.0, zipf_c=20.0, max_frac=0.5)
    train, test = _split(u, i, r, n_ratings // 5, rng)
    return {"train": train, "test": test, "all": (u, i, r), "n_users": n_users, "n_items": n_items}


def weak_shard(d, rank, stride=None):
    """Rank ``rank``'s shard of a weak-scaling workload built from one generated set ``d``: the same shape on every
    rank but DISTINCT data -- users are moved to the id range of the rank (``u + rank * stride``, ``stride`` = number
    of users of ``d``) and the items are rotated through the set's item ids by a rank-specific offset, so which items
    are popular differs from rank to rank.  The union over the ranks is a valid rating set (unique pairs) whose
    per-item sums differ from every single shard's: a run whose cross-rank exchange does nothing cannot reproduce
    the union's item deviations or MAE.  Rank 0's shard is ``d`` itself."""
    stride = int(d["n_users"]) if stride is None else int(stride)
    ids = np.unique(d["all"][1])
    off = (int(rank) * 7919) % ids.size

    def move(t):
        u, i, r = t
        if rank == 0:
            return u, i, r
        idx = np.searchsorted(ids, i)
        idx += off
        idx %= ids.size
        return (u + np.int32(rank * stride)).astype(np.int32), ids[idx].astype(np.int32), r
    return {"train": move(d["train"]), "test": move(d["test"]), "n_users": d["n_users"], "n_items": d["n_items"],
            "user_stride": stride, "max_item_id": int(ids[-1])}


def weak_union(d, world):
    """Concatenation of all ranks' weak_shard()s: the set a CPU reference must process for the same total work."""
    parts = [weak_shard(d, r) for r in range(world)]
    cat = lambda k, j: np.concatenate([p[k][j] for p in parts])
    return {"train": tuple(cat("train", j) for j in range(3)), "test": tuple(cat("test", j) for j in range(3))}


def cached(name, **kw):
    """Generate (or load from a /tmp cache) one of the named sets; ml25m takes ~20 s to generate."""
    fn = {"ml100k": ml100k, "ml25m": ml25m, "small": small, "mid": mid}[name]
    tag = name + "".join(f"_{k}{v}" for k, v in sorted(kw.items()))
    path = os.path.join(os.environ.get("MRS_SYNTH_CACHE", "/tmp/mrs_b200_synth"), tag + f"_np{np.__version__}.npz")
    if os.path.exists(path):
        try:
            z = np.load(path)
            return {"train": (z["tu"], z["ti"], z["tr"]), "test": (z["eu"], z["ei"], z["er"]),
                    "all": (z["au"], z["ai"], z["ar"]), "n_users": int(z["nu"]), "n_items": int(z["ni"])}
        except Exception:
            pass
    d = fn(**kw)
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        tmp = path + f".{os.getpid()}.tmp.npz"
        np.savez(tmp, tu=d["train"][0], ti=d["train"][1], tr=d["train"][2], eu=d["test"][0], ei=d["test"][1],
                 er=d["test"][2], au=d["all"][0], ai=d["all"][1], ar=d["all"][2], nu=d["n_users"], ni=d["n_items"])
        os.replace(tmp, path)
    except OSError:
        pass
    return d


def write_ratings(path, users, items, ratings, sep="\t", header=None):
    """Write ``u<sep>i<sep>r<sep>timestamp`` lines (the 4th column is ignored by the loader, P:41)."""
    with open(path, "w") as f:
        if header:
            f.write(header + "\n")
        ts = 880000000
        for a, b, c in zip(users.tolist(), items.tolist(), ratings.tolist()):
            rs = str(int(c)) if float(c).is_integer() and sep == "\t" else repr(float(c))
            f.write(f"{a}{sep}{b}{sep}{rs}{sep}{ts}\n")
            ts += 1
