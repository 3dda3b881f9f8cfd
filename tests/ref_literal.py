"""A second, independent restatement of shared/predictions.scala in plain Python dicts, written to
follow the Scala text statement by statement (P:<line> = predictions.scala line).  It exists only
to cross-check the C oracle on small inputs; where Scala iterates a HashMap/HashSet (order not
reproducible without a JVM) the canonical orders of oracle/mrs_oracle.c are used.
"""
import math
from collections import OrderedDict


def scale(x, y):  # P:57-61
    if x > y:
        return 5 - y
    elif x < y:
        return y - 1
    else:
        return 1


def mean(s):  # P:18
    if len(s) > 0:
        acc = s[0]
        for v in s[1:]:
            acc = acc + v
        return acc / len(s)
    return 0.0


def average(ratings):  # P:94
    return mean([r[2] for r in ratings])


def group_by(ratings, key):  # Scala groupBy keeps encounter order inside each group
    g = OrderedDict()
    for r in ratings:
        g.setdefault(r[key], []).append(r)
    return g


def users_avg(ratings):  # P:113
    return {u: average(rs) for u, rs in group_by(ratings, 0).items()}


def items_avg(ratings):  # P:134
    return {i: average(rs) for i, rs in group_by(ratings, 1).items()}


def normalize_deviation(ratings):  # P:155-169
    ua = users_avg(ratings)
    g = average(ratings)
    out = OrderedDict()
    for (u, i, r) in ratings:
        a = ua.get(u, g)
        out[(u, i)] = (r - a) / scale(r, a)
    return out


def items_avg_dev(ratings):  # P:176-186
    acc = OrderedDict()
    for (u, i), d in normalize_deviation(ratings).items():
        cur = acc.get(i, (0.0, 0))
        acc[i] = (d + cur[0], 1 + cur[1])
    return {i: s / c for i, (s, c) in acc.items()}


def compute_prediction(ratings):  # P:205-237
    ua = users_avg(ratings)
    dev = items_avg_dev(ratings)
    g = average(ratings)

    def predict(user, item):
        a = ua.get(user, -1.0)
        if a < 0.0:
            return g
        d = dev.get(item, 0.0)
        return a + d * scale(a + d, a)
    return predict


def mae(predict, data):  # P:69-86
    acc, cnt = 0.0, 0
    for (u, i, r) in data:
        acc, cnt = abs(r - predict(u, i)) + acc, cnt + 1
    return acc / cnt if cnt else float("nan")


def preprocessed_rating(ratings):  # P:470-481 (per-user sum in ascending item order)
    nd = normalize_deviation(ratings)
    by_user = {}
    for (u, i), d in nd.items():
        by_user.setdefault(u, []).append((i, d))
    weights = {}
    for u, lst in by_user.items():
        s = 0.0
        for _, d in sorted(lst):
            s = s + d * d
        weights[u] = math.sqrt(s)
    return {k: (d / weights[k[0]] if weights.get(k[0], 0.0) != 0 else 0.0) for k, d in nd.items()}


def adjusted_cosine(ratings):  # P:407-433 (intersection in ascending item order)
    pre = preprocessed_rating(ratings)
    rated = group_by(ratings, 0)

    def sim(u, v):
        ui = {r[1] for r in rated.get(u, [])}
        vi = {r[1] for r in rated.get(v, [])}
        acc = 0.0
        for i in sorted(ui & vi):
            acc = acc + pre.get((u, i), 0.0) * pre.get((v, i), 0.0)
        return acc
    return sim


def jaccard(ratings):  # P:440-464
    rated = group_by(ratings, 0)

    def sim(u, v):
        ur, vr = rated.get(u, []), rated.get(v, [])
        inter = len({r[1] for r in ur} & {r[1] for r in vr})
        den = len(ur) + len(vr) - inter
        return inter / den if den else float("nan")
    return sim


def similarity_one(u, v):  # P:400
    return 1.0


def weighted_sum_deviation(ratings, sim):  # P:489-549
    rated_i = group_by(ratings, 1)
    g = average(ratings)
    ua = users_avg(ratings)

    def wsd(u, i):
        num, den = 0.0, 0.0
        for (xu, _, xr) in rated_i.get(i, []):
            a = ua.get(xu, g)
            d = (xr - a) / scale(xr, a)
            s = sim(u, xu)
            num, den = num + d * s, den + abs(s)
        return num / den if den > 0 else 0.0
    return wsd


def predictor(ratings, wsd):  # P:557-586
    g = average(ratings)
    ua = users_avg(ratings)

    def predict(u, i):
        a = ua.get(u, -1.0)
        if a < 0.0:
            return g
        w = wsd(u, i)
        return a + w * scale(a + w, a)
    return predict


def get_neighbors(ratings, k, sim):  # P:596-617 (candidate order: ascending user id)
    all_users = sorted({r[0] for r in ratings})

    def nn(u):
        others = [x for x in all_users if x != u]
        scored = [(x, sim(u, x)) for x in others]
        scored.sort(key=lambda t: -t[1])  # Python's sort is stable, like sortWith on a Seq
        return scored[:k]
    return nn


def get_similarity(ratings, k, sim):  # P:626-649
    nn = get_neighbors(ratings, k, sim)
    cache = {}

    def s(u, v):
        if u not in cache:
            cache[u] = nn(u)
        return sum((x[1] if x[0] == v else 0.0) for x in cache[u])
    return s


def recommendations(ratings, predict):  # P:651-674
    def rec(user, n):
        not_rated = {r[1] for r in ratings} - {r[1] for r in ratings if r[0] == user}
        scored = [(x, predict(user, x)) for x in not_rated]
        scored.sort(key=lambda t: (-t[1], t[0]))
        return scored[:n]
    return rec
