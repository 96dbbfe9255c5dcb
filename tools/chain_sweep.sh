#!/bin/bash
# usage: tools/chain_sweep.sh <tag> <variant>...   each variant is a quoted env assignment list, e.g. "Y3_CHAIN=0"
tag=$1; shift
i=0
for v in "$@"; do
  i=$((i+1))
  env Y3_PROF_LIB=1 $v timeout 300 python tools/chain_check.py 64 416 4 > gpurun_out/${tag}_$i.json 2> gpurun_out/${tag}_$i.err
  echo "[$v] exit $?"
  python - <<PY
import json
try:
    r = json.load(open("gpurun_out/${tag}_$i.json"))
    print("   eager %.3f ms  graph %.3f ms  hash %s/%s bad %d/%d   A/B runs %.3f  per-layer %.3f" % (r["eager_ms"], r["graph_ms"], r["f32"]["hash"], r["u8_padded"]["hash"], r["f32"]["mismatching_runs"], r["u8_padded"]["mismatching_runs"], r.get("ab_runs_ms", 0), r.get("ab_per_layer_ms", 0)))
except Exception as e:
    print("   no result", e)
PY
done
