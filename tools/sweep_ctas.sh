#!/bin/bash
# tuning aid: time the headline bench for several resident-CTA counts of the v2 search kernel
for c in "$@"; do
  DAB_V2_CTAS_PER_SM=$c timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null > /tmp/b.json
  python - "$c" <<'PY'
import json, sys
d = json.load(open('/tmp/b.json'))
print("ctas/SM", sys.argv[1], "ms/step %.3f" % d["ms_per_step"], "QPS %.0f" % d["value"], "frac %.3f" % d["roofline"]["frac"])
PY
done
