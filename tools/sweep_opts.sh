#!/bin/bash
# usage: tools/sweep_opts.sh <tag> <bench args...> -- <PEB_OPTS value> [<PEB_OPTS value> ...]
# runs bench.py once per option set and prints value / ms per step / per-iteration launch times (development sweeps)
tag=$1; shift
args=()
while [ "$1" != "--" ]; do args+=("$1"); shift; done
shift
for o in "$@"; do
  f=gpurun_out/${tag}_$(echo "$o" | tr '=,' '__').json
  PEB_OPTS=$o python bench.py "${args[@]}" > "$f" 2> gpurun_out/${tag}.err
  python - "$f" "$o" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], round(d["value"]), round(d["ms_per_step"], 2), round(d["e2e"]["value"]), d["roofline"]["launch_ms_by_iteration"])
PY
done
