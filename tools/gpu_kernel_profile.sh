#!/bin/bash
# Full ncu capture of one kernel (regex) of the bench step, reduced to text.   tools/gpu_kernel_profile.sh <tag> <kernel regex> <launch-skip> [count]
tag=${1:-run}; k=${2:-MeshBricksKernel}; skip=${3:-2}; cnt=${4:-1}
out=gpurun_out
mkdir -p $out
TG_PIPELINE_CHUNKS=1 ncu --set full --clock-control none --import-source on -k regex:"$k" --launch-skip $skip -c $cnt -f -o $out/${tag}_full \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra-workloads > $out/${tag}_ncu_full.log 2>&1
python tools/ncu_raw.py $out/${tag}_full.ncu-rep > $out/${tag}_raw.txt 2>&1
python tools/ncu_lines.py $out/${tag}_full.ncu-rep "$k" tangerine_b200/libtangerine_b200.so 45 > $out/${tag}_lines.txt 2>&1
head -60 $out/${tag}_raw.txt
head -70 $out/${tag}_lines.txt
if [ $(stat -c %s $out/${tag}_full.ncu-rep) -gt 30000000 ]; then rm -f $out/${tag}_full.ncu-rep; fi
