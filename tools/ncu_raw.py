#!/usr/bin/env python
"""Key raw metrics of every kernel instance in an ncu report:  python tools/ncu_raw.py <report.ncu-rep> [regex]"""
import csv
import io
import re
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__warps_eligible.avg.per_cycle_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__sass_thread_inst_executed_op_fadd_pred_on.sum', 'smsp__sass_thread_inst_executed_op_fmul_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_ffma_pred_on.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__cycles_active.avg', 'sm__cycles_elapsed.max']


def main():
    rep = sys.argv[1]
    regex = sys.argv[2] if len(sys.argv) > 2 else "."
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    stall = [h for h in hdr if h.startswith("smsp__average_warp") and "issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
    stall2 = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith(".ratio")]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if not re.search(regex, name):
            continue
        print("==== " + name)
        for w in WANT:
            if w in hdr:
                print("  %-72s %s %s" % (w, r[hdr.index(w)], units[hdr.index(w)]))
        st = []
        for h in set(stall + stall2):
            try:
                st.append((float(r[hdr.index(h)].replace(",", "")), h))
            except ValueError:
                pass
        for v, h in sorted(st, reverse=True)[:10]:
            print("  stall %-66s %.3f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("smsp__average_warp_latency_issue_stalled_", ""), v))


if __name__ == "__main__":
    main()
