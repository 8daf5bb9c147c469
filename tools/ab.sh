#!/bin/bash
# A/B of library variants on the bench workloads: tools/ab.sh <variant.so ...>   (prints value / ms / stages per workload)
for lib in "$@"; do
  for w in ${AB_WORKLOADS:-cfg2 cfg1 cfg5 cfg4}; do
    SCB_LIBRARY=$lib python bench.py --workload $w --steps ${AB_STEPS:-20} --no-cpu-baseline > /tmp/ab.json 2> /tmp/ab.err || { echo "$lib $w FAILED"; tail -3 /tmp/ab.err; continue; }
    python - "$lib" "$w" <<'PY'
import json,sys
d=json.load(open("/tmp/ab.json"))
st={k:round(v*1e3,1) for k,v in d["stages_ms"].items()}
print(sys.argv[1].split("/")[-1], sys.argv[2], "Mpix/s %.0f"%d["value"], "ms %.4f"%d["ms_per_step"], "e2e_ms %.4f"%d["e2e"]["ms_per_step"], st)
PY
  done
done
