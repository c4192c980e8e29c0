#!/bin/bash
# parity tests, then one bench line per workload (short form).   tools/gpu_workloads.sh <tag> <workload>...
tag=${1:-run}; shift
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for w in "$@"; do
  python bench.py --workload $w --steps 4 --warmup 3 --no-cpu-baseline > $out/${tag}_$w.json 2> $out/${tag}_$w.err
  python - $out/${tag}_$w.json $w <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print("%-14s ms/step %.3f  e2e %.3f  bricks %d/%d  %s" % (sys.argv[2], d["ms_per_step"], d["e2e"]["ms_per_step"], d["bricks"]["evaluated"], d["bricks"]["total"], {k: round(v, 3) for k, v in d["stage_ms_rank0"].items()}))
PY
done
