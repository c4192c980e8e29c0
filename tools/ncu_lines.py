#!/usr/bin/env python
"""Per-source-line hot spots of one kernel from an ncu report (no GUI needed).

    python tools/ncu_lines.py <report.ncu-rep> <kernel regex> <lib.so> [top N]

Joins `ncu --page source --csv` (SASS rows: executed instructions, stall samples) with
`nvdisasm --print-line-info` of the same kernel (needs -lineinfo at compile time) by instruction order.
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def sass_lines(lib, kernel_regex):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if "sm_100" in f][0]
    text = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], check=True, capture_output=True, text=True).stdout
    out = []
    active = False
    cur = ("?", 0)
    for line in text.splitlines():
        m = re.match(r"^\.text\.(\S+):", line)
        if m:
            active = re.search(kernel_regex, m.group(1)) is not None
            continue
        if not active:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            out.append((int(m.group(1), 16), m.group(2).strip(), cur))
    return out


def main():
    rep, regex, lib = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + regex], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    # first kernel instance only
    start = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    hdr = rows[start]
    body = []
    for r in rows[start + 1:]:
        if not r or r[0] in ("Kernel Name", "Address"):
            break
        body.append(r)
    sass = sass_lines(lib, regex)
    if len(sass) != len(body):
        print("warning: %d SASS rows in the report vs %d in the library (different build?)" % (len(body), len(sass)))
    ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
    ct = hdr.index("Thread Instructions Executed")
    by_line = defaultdict(lambda: [0, 0, 0])
    by_op = defaultdict(lambda: [0, 0])
    tot_i = tot_s = 0
    for k, r in enumerate(body):
        n, s, t = int(r[ci]), int(r[cs]), int(r[ct])
        loc = sass[k][2] if k < len(sass) else ("?", 0)
        by_line[loc][0] += n
        by_line[loc][1] += s
        by_line[loc][2] += t
        op = r[1].split()[0] if not r[1].lstrip().startswith("@") else r[1].split()[1]
        op = op.split(".")[0]
        by_op[op][0] += n
        by_op[op][1] += s
        tot_i += n
        tot_s += s
    print("total warp instructions %d, stall samples %d, SASS rows %d" % (tot_i, tot_s, len(body)))
    print("\n== by source line (warp instructions) ==")
    for loc, (n, s, t) in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%-28s %5d  inst %6.2f%%  samples %6.2f%%  thr/inst %4.1f" % (loc[0], loc[1], 100.0 * n / tot_i, 100.0 * s / max(tot_s, 1), t / max(n, 1)))
    print("\n== by source line (stall samples) ==")
    for loc, (n, s, t) in sorted(by_line.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%-28s %5d  inst %6.2f%%  samples %6.2f%%" % (loc[0], loc[1], 100.0 * n / tot_i, 100.0 * s / max(tot_s, 1)))
    print("\n== by opcode ==")
    for op, (n, s) in sorted(by_op.items(), key=lambda kv: -kv[1][0])[:30]:
        print("%-12s inst %6.2f%%  samples %6.2f%%" % (op, 100.0 * n / tot_i, 100.0 * s / max(tot_s, 1)))


if __name__ == "__main__":
    main()
