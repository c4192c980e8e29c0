#!/bin/bash
# Short GPU-box visit: parity tests, one bench line, the ncu launch list.   tools/gpu_quick.sh <tag>
tag=${1:-run}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1
tail -3 $out/${tag}_pytest.log
python bench.py --steps 5 --warmup 4 --no-cpu-baseline --no-extra-workloads > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
python - $out/${tag}_bench_n1.json <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print("ms/step %.3f  e2e %.3f  stages %s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], {k: round(v, 3) for k, v in d["stage_ms_rank0"].items()}))
PY
tail -3 $out/${tag}_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-workloads > $out/${tag}_ncu_b.log 2>&1
python tools/launch_list.py $out/${tag}_launches.csv 3 > $out/${tag}_launch_list.txt 2>&1
cat $out/${tag}_launch_list.txt
