#!/bin/bash
# same-box A/B: step time of the baseline pass (tools/timeline.py, median of its steps) for the in-tree library and for
# every variant in _ab/ (tools/build_variant.sh), interleaved, twice
for r in 1 2; do
  for lib in "" _ab/*.so; do
    if [ -z "$lib" ]; then tag=tree; else tag=$(basename "$lib" .so); fi
    med=$(MRS_LIB=${lib:+$PWD/$lib} python tools/timeline.py 2>&1 | grep '^step' | tail -5 | awk '{print $2}' | sort -n | sed -n 3p)
    last=$(MRS_LIB=${lib:+$PWD/$lib} python tools/timeline.py 2>&1 | grep '^step' | tail -1)
    echo "$tag median $med | $last"
  done
done
