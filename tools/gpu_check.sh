#!/bin/bash
# One GPU-box visit: parity tests, the bench line, the ncu launch list of the same command and a full capture of the
# top kernels, reduced to text on the box (the .ncu-rep stays only if it is small).   tools/gpu_check.sh <tag> [full]
tag=${1:-run}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1
tail -5 $out/${tag}_pytest.log
python bench.py --steps 5 --warmup 4 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
cat $out/${tag}_bench_n1.json
tail -3 $out/${tag}_bench_n1.err
if [ "$2" != "quick" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-workloads > $out/${tag}_ncu_b.log 2>&1
python tools/launch_list.py $out/${tag}_launches.csv 3 > $out/${tag}_launch_list.txt 2>&1
cat $out/${tag}_launch_list.txt
TG_PIPELINE_CHUNKS=1 ncu --set full --clock-control none --import-source on -k regex:"MeshBricksKernel|FinalizeMeshKernel|AttributesKernel|ColorsKernel" \
    --launch-skip 8 -c 4 -f -o $out/${tag}_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra-workloads > $out/${tag}_ncu_full.log 2>&1
python tools/ncu_raw.py $out/${tag}_full.ncu-rep > $out/${tag}_raw.txt 2>&1
for k in MeshBricksKernel AttributesKernel ColorsKernel; do
  python tools/ncu_lines.py $out/${tag}_full.ncu-rep $k tangerine_b200/libtangerine_b200.so 40 > $out/${tag}_${k}_lines.txt 2>&1
done
ls -la $out/${tag}_full.ncu-rep
if [ $(stat -c %s $out/${tag}_full.ncu-rep) -gt 30000000 ]; then rm -f $out/${tag}_full.ncu-rep; fi
fi
du -sh $out
