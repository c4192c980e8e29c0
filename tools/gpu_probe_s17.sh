#!/bin/bash
# diagnostic visit: cull kernels under ncu, e2e pipeline timeline, PCIe D2H ceiling
out=gpurun_out
mkdir -p $out
TG_PIPELINE_CHUNKS=1 ncu --set full --clock-control none --import-source on -k regex:"CullLevelKernel|CullRegionInitKernel" \
    --launch-skip 18 -c 6 -f -o $out/s17_cull python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/s17_ncu_cull.log 2>&1
python tools/ncu_raw.py $out/s17_cull.ncu-rep > $out/s17_cull_raw.txt 2>&1
python tools/e2e_probe.py 1 4 8 > $out/s17_e2e_probe.txt 2>&1
TG_TRACE_HOST=1 python tools/trace_probe.py > $out/s17_trace.txt 2>&1
python - > $out/s17_pcie.txt 2>&1 <<'PY'
import torch, time
n = 300 * 1024 * 1024
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8).pin_memory()
for size in (n, n // 4, n // 16):
    for it in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        h[:size].copy_(d[:size], non_blocking=True); torch.cuda.synchronize()
        t1 = time.perf_counter()
    print("D2H %d MiB: %.3f ms  %.1f GB/s" % (size >> 20, (t1 - t0) * 1e3, size / (t1 - t0) * 1e-9))
    for it in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        d[:size].copy_(h[:size], non_blocking=True); torch.cuda.synchronize()
        t1 = time.perf_counter()
    print("H2D %d MiB: %.3f ms  %.1f GB/s" % (size >> 20, (t1 - t0) * 1e3, size / (t1 - t0) * 1e-9))
PY
cat $out/s17_pcie.txt $out/s17_e2e_probe.txt
tail -30 $out/s17_trace.txt
grep -A8 "====" $out/s17_cull_raw.txt | grep -E "====|duration|grid_size|thread_inst_executed_per|inst_executed.sum"
