#!/bin/bash
# bench line + full ncu capture of MeshBricksKernel only, reduced to text.   tools/gpu_mb_profile.sh <tag>
tag=${1:-run}
out=gpurun_out
mkdir -p $out
python bench.py --steps 5 --warmup 4 --no-cpu-baseline --no-extra-workloads > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
python - $out/${tag}_bench_n1.json <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print("ms/step %.3f  e2e %.3f  stages %s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], {k: round(v, 3) for k, v in d["stage_ms_rank0"].items()}))
PY
TG_PIPELINE_CHUNKS=1 ncu --set full --clock-control none --import-source on -k regex:"MeshBricksKernel" \
    --launch-skip 2 -c 1 -f -o $out/${tag}_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra-workloads > $out/${tag}_ncu_full.log 2>&1
python tools/ncu_raw.py $out/${tag}_full.ncu-rep > $out/${tag}_raw.txt 2>&1
python tools/ncu_lines.py $out/${tag}_full.ncu-rep MeshBricksKernel tangerine_b200/libtangerine_b200.so 40 > $out/${tag}_MeshBricksKernel_lines.txt 2>&1
head -24 $out/${tag}_raw.txt
