#!/bin/bash
# item-pass time (CUDA events, microseconds) for a few settings of the layout / partition knobs (environment variables
# read when the layout is built); usage: tools/k2_sweep.sh "VAR=a,b,c" ...   e.g.  tools/k2_sweep.sh "MRS_W_POP_ROW=14,28,42"
export MRS_NO_ITEM_AVG=1
for spec in "$@"; do
  var="${spec%%=*}"; vals="${spec#*=}"
  for v in ${vals//,/ }; do
    out=$(env "$var=$v" python tools/prof_pass.py --passes 4 2>&1 | tail -1)
    echo "$var=$v  $out"
  done
done
