#!/bin/bash
# usage: tools/sweep_env.sh VAR v1 v2 ...   -- bench.py (device leg) once per value of an environment knob
var=$1; shift
export ASRK_BENCH_CACHE=${ASRK_BENCH_CACHE:-/tmp/asrk_bench_cache}
for v in "$@"; do
  env $var=$v timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | tail -1 > /tmp/_sweep.json
  python - "$var" "$v" <<'PY'
import json, sys
d = json.load(open("/tmp/_sweep.json"))
print(sys.argv[1], sys.argv[2], "value %.0f" % d["value"], "ms %.4f" % d["ms_per_step"],
      {k: round(v * 1e3, 1) for k, v in d["kernel_ms"].items()})
PY
done
