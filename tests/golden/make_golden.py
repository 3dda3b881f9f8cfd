"""Collects the answer JSONs the reference commits (real MovieLens runs) into one fixture.

Run in the build container only (reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
"""
import json
import os

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_answers.json")


def main():
    out = {}
    for name in ("baseline-100k.json", "distributed-25m-4.json", "personalized-100k.json", "knn-100k.json"):
        with open(os.path.join(REF, name)) as f:
            d = json.load(f)
        # drop timing blocks and paths: only result values are golden
        out[name] = {k: v for k, v in d.items() if k not in ("Meta", "B.3", "N.3", "D.2")}
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
