#!/usr/bin/env python
"""Runs the ml-25m-shape baseline MAE pass a few times (nothing else): the short command profiled with ncu.
    python tools/prof_pass.py [--passes 4] [--knn]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import mrs_b200  # noqa: F401,E402
from mrs_b200 import engine as E, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--passes", type=int, default=4)
ap.add_argument("--knn", action="store_true")
args = ap.parse_args()
eng = E.Engine(0)
if args.knn:
    d = synth.cached("ml100k")
    R, T = eng.ratings(*d["train"]), eng.ratings(*d["test"])
    m = E.Model(eng, R)
    s = m.similarity(E.SIM_COSINE, 300)
    for _ in range(args.passes):
        m.refit(); s.refit(300)
        print("knn mae", m.mae(T, E.PRED_PERSONALIZED, s))
else:
    d = synth.cached("ml25m")
    R, T = eng.ratings(*d["train"]), eng.ratings(*d["test"])
    m = E.Model(eng, R)
    if os.environ.get("MRS_NO_ITEM_AVG"):
        m.set_item_averages(False); m.refit()
    m.mae(T, E.PRED_BASELINE)
    print("train layout", R.layout_info(), "test layout", T.layout_info())
    for _ in range(args.passes):
        eng.profile_begin()
        m.refit()
        mae = m.mae(T, E.PRED_BASELINE)
        print("mae", mae, [(k, round(v * 1000, 1)) for k, v in eng.profile_end()])
