#!/bin/bash
# tuning aid: run the smoke search (checked against the oracle) on each library build, bounded
for lib in "$@"; do
  DAB_LIB_PATH=$lib timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > /tmp/qc.log 2>&1
  echo "$lib rc=$? $(tail -1 /tmp/qc.log | cut -c1-150)"
done
