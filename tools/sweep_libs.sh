#!/bin/bash
# tuning aid: time the headline bench for library builds with different register budgets
for lib in "$@"; do
  DAB_LIB_PATH=$lib timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null > /tmp/b.json
  python - "$lib" <<'PY'
import json, sys
d = json.load(open('/tmp/b.json'))
print(sys.argv[1], "ms/step %.3f" % d["ms_per_step"], "QPS %.0f" % d["value"], "frac %.3f" % d["roofline"]["frac"])
PY
done
