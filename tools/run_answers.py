#!/usr/bin/env python
"""Writes an answer JSON with the reference's schema (see mrs_b200/answers.py).

  python tools/run_answers.py baseline --train data/ml-100k/u2.base --test data/ml-100k/u2.test --separator "\t" --json out.json
  python tools/run_answers.py knn|personalized|distributed ... ;  python tools/run_answers.py recommender --data u.data --personal personal.csv
Without --train/--test the seeded synthetic ml-100k-shaped set is used.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mrs_b200  # noqa: E402,F401
from mrs_b200 import answers, predictions as P, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("app", choices=["baseline", "distributed", "personalized", "knn", "recommender"])
ap.add_argument("--train"); ap.add_argument("--test"); ap.add_argument("--data"); ap.add_argument("--personal")
ap.add_argument("--separator", default="\t"); ap.add_argument("--num_measurements", type=int, default=3); ap.add_argument("--json")
a = ap.parse_args()
sep = a.separator.encode().decode("unicode_escape")
if a.app == "recommender":
    d = synth.cached("ml100k")["all"] if not a.data else None
    if d is None:
        import numpy as np
        raw = np.loadtxt(a.data, delimiter=sep if sep != "\t" else None)
        d = (raw[:, 0].astype("int32"), raw[:, 1].astype("int32"), raw[:, 2])
    out = answers.recommender(d, a.personal or "/root/reference/data/personal.csv", a.data or "synthetic ml-100k shape")
else:
    if a.train:
        train, test = P.load(None, a.train, sep), P.load(None, a.test, sep)
    else:
        d = synth.cached("ml100k")
        train, test = P.RatingSet.from_arrays(*d["train"]), P.RatingSet.from_arrays(*d["test"])
    out = getattr(answers, a.app)(train, test, a.num_measurements, train_path=a.train or "synthetic", test_path=a.test or "synthetic")
text = json.dumps(out, indent=4)
print(text)
if a.json:
    with open(a.json, "w") as f:
        f.write(text)
