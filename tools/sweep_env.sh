#!/bin/bash
# tuning aid: usage sweep_env.sh VAR v1 v2 ... — time the headline bench with VAR=value
var=$1; shift
for v in "$@"; do
  env $var=$v timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null > /tmp/b.json
  python - "$var=$v" <<'PY'
import json, sys
d = json.load(open('/tmp/b.json'))
print(sys.argv[1], "ms/step %.3f" % d["ms_per_step"], "QPS %.0f" % d["value"], "frac %.3f" % d["roofline"]["frac"])
PY
done
