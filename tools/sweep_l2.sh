#!/bin/bash
# tuning aid: A/B of the L2 eviction-priority hints and the persisting access-policy window
for lib in build/lib_hints.so build/lib_nohints.so; do
  for win in 0 1; do
    if [ $win = 1 ]; then export DAB_NO_L2_WINDOW=1; else unset DAB_NO_L2_WINDOW; fi
    DAB_LIB_PATH=$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null > /tmp/b.json
    python - "$lib" "nowindow=$win" <<'PY'
import json, sys
d = json.load(open('/tmp/b.json'))
print(sys.argv[1], sys.argv[2], "ms/step %.3f" % d["ms_per_step"], "QPS %.0f" % d["value"], "recall", d["config"].get("recall_at_10"))
PY
  done
done
