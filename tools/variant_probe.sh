#!/bin/bash
# A/B of prebuilt library variants on one box.   tools/variant_probe.sh <tag> <lib.so>...   (first one is restored at the end)
# Workloads: WORKLOADS="seaside1024 synthetic256" by default.
tag=${1:-var}; shift
out=gpurun_out; mkdir -p $out
first=$1
for lib in "$@"; do
  cp $lib tangerine_b200/libtangerine_b200.so
  for w in ${WORKLOADS:-seaside1024 synthetic256}; do
    python bench.py --workload $w --steps 5 --warmup 4 --no-cpu-baseline --no-extra-workloads > $out/${tag}_tmp.json 2> $out/${tag}_tmp.err
    python - $out/${tag}_tmp.json $lib $w <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print("%-40s %-13s ms/step %.3f eval %.3f frac %.3f parity %s" % (sys.argv[2], sys.argv[3], d["ms_per_step"], d["stage_ms_rank0"]["evaluate_ms"], d["roofline"]["frac"], (d.get("parity") or {}).get("equal")))
PY
  done
done
cp $first tangerine_b200/libtangerine_b200.so
