#!/usr/bin/env python
"""bench.py -- the meshing hot path on N B200s, beside the reference's CPU path on the box's host cores.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload seaside1024] [--impl ours|reference]

A "step" is one whole mesh export of the workload: culling, brick evaluation + cell classification, vertex
numbering, quad emission, refinement + normals + colours.  With N > 1 the export runs INSIDE the library on a
multi-device context (tg_context_create_multi): one process, one host thread + stream per GPU, z-slabs cut from a
host-side work estimate (no planning exports, plan_iters = 0), the per-slab vertex counts all-gathered with NCCL on the
devices, one stitched host mesh.  Under torch.distributed.run rank 0 drives that call; the other ranks join the NCCL
process group (so the launch contract's rendezvous and barrier hold) and wait.  The job is the same fixed grid at
every N ("strong" scaling).

JSON keys (one line on stdout, rank 0):
  value          Mvoxel/s = grid cells / device time per step (max over the devices), model tables resident in HBM,
                 results left in HBM
  e2e            Mvoxel/s through the C ABI the reference would bind (tg_model_upload + tg_export_mesh with host result
                 buffers): per step the model tables go host -> device and the mesh comes back
  whole_export   wall time from the CSG tree to the host mesh (octree build + flattening + upload + export): what a
                 caller of MeshExport waits for
  parity         the benched mesh compared layer by layer with the reference's own mesh of the same grid
                 (tests/golden/slices_<workload>.json), outside the timed region
  roofline       dominant kernel = MeshBricksKernel (evaluation + classification), FP32-pipe bound; `hbm` holds the
                 bandwidth-bound mesh kernels (vertex numbering / scatter / quad emission)
  workloads      (N = 1) the other BASELINE.json configurations, a short run each
  cpu_baseline   the reference's own thunks (oracle/_ref/tangerine_ref, built from /root/reference by
                 oracle/Makefile) on all host threads over a stratified sample of z-slices of the same grid
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(1, os.path.join(ROOT, "tests"))
MODELS = os.path.join(ROOT, "tests", "golden", "models")
REF_TOOL = os.path.join(ROOT, "oracle", "_ref", "tangerine_ref")

# name -> (model file, bounds step as float32, refine iterations, description)
WORKLOADS = {
    "seaside1024": ("seaside_town", 10.0 / 1022.0, 0, "seaside_town.lua 1024^3 (BASELINE.json configs[2], the north-star target grid)"),
    "seaside512": ("seaside_town", 10.0 / 510.0, 0, "seaside_town.lua 512^3"),
    "gear512": ("gear", 8.0 / 510.0, 5, "gear.lua 512x512x34 with 5 refinement iterations (BASELINE.json configs[1])"),
    "colorcube512": ("color-cube", 9.6 / 510.0, 0, "color-cube.lua 513^3 with per-vertex colour (BASELINE.json configs[4])"),
    "basic66": ("basic_thing", 1.0 / 16.0, 0, "basic_thing.lua 66^3 (BASELINE.json configs[0])"),
    # BASELINE.json configs[3]: synthetic random CSG scene, 10k primitives (tg_make_synthetic, seed 1234), dense sweep
    "synthetic256": ("synthetic:10000", 10.0 / 254.0, 0, "synthetic random CSG scene, 10,000 primitives, 256^3"),
    "synthetic512": ("synthetic:10000", 10.0 / 510.0, 0, "synthetic random CSG scene, 10,000 primitives, 512^3"),
    "synthetic1024": ("synthetic:10000", 10.0 / 1022.0, 0, "synthetic random CSG scene, 10,000 primitives, 1024^3"),
    "synthetic2048": ("synthetic:10000", 10.0 / 2046.0, 0, "synthetic random CSG scene, 10,000 primitives, 2048^3"),
}


def model_file(name):
    """.tgm file of a workload's CSG tree.  The synthetic scene (tg_make_synthetic(10000, 1234)) is committed as a file
    too, so that the reference arm needs nothing from this repository's library."""
    return os.path.join(MODELS, name.replace(":", "") + ".tgm")


def load_workload_tree(T, name):
    """Returns (tree, path of a .tgm file of it for the reference tool)."""
    path = model_file(name)
    if name.startswith("synthetic:") and not os.path.exists(path):
        tree = T.Tree.synthetic(int(name.split(":")[1]), 1234)
        tree.save(path)
        return tree, path
    return T.Tree.load(path), path


def export_grid_shape(lo, hi, step):
    """MeshExportThread's grid (export.cpp:324-337) in float32, without the library: (origin, cells per axis)."""
    lo, hi, step = np.asarray(lo, np.float32), np.asarray(hi, np.float32), np.float32(step)
    origin = (lo - step * np.float32(2.0)).astype(np.float32)
    shape = [int(np.ceil(np.float32((hi[i] - origin[i]) / step))) for i in range(3)]
    return origin, shape


def workload_config(workload, shape):
    """The `config` object both arms print (same keys, same values)."""
    name, step, refine, desc = WORKLOADS[workload]
    return {"workload": desc, "model": name.replace(":", "") + ".tgm", "grid": list(shape), "refine_iterations": refine,
            "l2": "inputs larger than L2 (a step writes and re-reads the active-cell bitmap and the result arrays: hundreds of MB at 1024^3); the CUDA arm also "
                  "flushes L2 before every timed step (256 MiB fill per device, outside the per-step event pair)"}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md).

    NVML in-process (nvidia_ml_py) on a background thread: a query costs tens of microseconds.  Spawning
    `nvidia-smi -lms` for the same purpose stalled the CUDA driver for about a millisecond per poll and doubled the
    measured time of sub-millisecond steps (gear.lua 512x512x34: 1.80 ms/step with it, 0.78 without); it is only the
    fallback when NVML cannot be loaded."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device, interval=0.02):
        self.device = device
        self.interval = interval
        self.samples = []          # (sm_mhz, sm_max_mhz, power_w, reasons bitmask)
        self.proc = None
        self.thread = None
        self.stop_flag = threading.Event()
        self.nvml = None
        self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            index = int(visible.split(",")[device]) if visible and visible.split(",")[device].isdigit() else device
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _poll_nvml(self):
        n = self.nvml
        while not self.stop_flag.is_set():
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                smax = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
                # the power read goes out to the board's controller and can hold the driver for milliseconds: once in ten
                power = n.nvmlDeviceGetPowerUsage(self.handle) * 1e-3 if len(self.samples) % 10 == 0 else (self.samples[-1][2] if self.samples else 0.0)
                reasons = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.samples.append((float(sm), float(smax), float(power), int(reasons)))
            except Exception:
                pass
            self.stop_flag.wait(self.interval)

    def start(self):
        self.stop_flag.clear()
        if self.nvml:
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.lines = []
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        if self.nvml:
            self.stop_flag.set()
            if self.thread:
                self.thread.join(timeout=2)
            if not self.samples:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
            n = self.nvml
            bits = {"hw_slowdown": getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8), "hw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                    "sw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20), "sw_power_cap": getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            reasons = sorted(name for name in names if any(s[3] & bits[name] for s in self.samples))
            return {"sm_mhz": float(np.median([s[0] for s in self.samples])), "sm_max_mhz": float(max(s[1] for s in self.samples)),
                    "power_w_max": float(max(s[2] for s in self.samples)), "samples": len(self.samples), "reasons": reasons, "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi"}


def run_reference_sample(tgm_path, lo, hi, step, stride, threads):
    """The reference's FirstLoopInnerThunk / SecondLoopThunk on std::threads over every `stride`-th z-slice."""
    args = [REF_TOOL, "bench", tgm_path] + ["%.9g" % v for v in list(lo) + list(hi)] + ["%.9g" % step, str(threads), str(stride)]
    out = subprocess.run(args, check=True, capture_output=True, text=True).stdout
    return json.loads(out.strip().splitlines()[-1])


def reference_arm(args, workload):
    """--impl reference: the reference CPU implementation (oracle/_ref/tangerine_ref = its own sources), all host threads,
    a bounded sample of the workload per step.  Nothing of this repository's library is loaded here: bounds come from
    `tangerine_ref info`, the grid from the reference's own formula."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    name, step, refine, desc = WORKLOADS[workload]
    if not os.path.exists(REF_TOOL):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/tangerine_ref was not built (run __graft_entry__.build() where /root/reference exists)"}))
        return 0
    tgm = model_file(name)
    if not os.path.exists(tgm):
        print(json.dumps({"impl": "reference", "unavailable": "model file %s is missing" % os.path.basename(tgm)}))
        return 0
    info = json.loads(subprocess.run([REF_TOOL, "info", tgm], check=True, capture_output=True, text=True).stdout.strip().splitlines()[-1])
    lo, hi = info["bounds_min"], info["bounds_max"]
    step32 = float(np.float32(step))
    threads = os.cpu_count() or 1
    _, shape = export_grid_shape(lo, hi, step32)
    stride = args.ref_stride or max(1, shape[2] // 8)
    for _ in range(args.warmup):
        run_reference_sample(tgm, lo, hi, step32, max(stride * 4, 1), threads)
    cells = 0.0
    seconds = 0.0
    last = None
    for _ in range(args.steps):
        last = run_reference_sample(tgm, lo, hi, step32, stride, threads)
        cells += last["cells_timed"]
        seconds += last["loop1_s"]
    value = cells / seconds * 1e-6
    sample = "every %d-th z-slice of the %dx%dx%d grid (%d slices, %.3g Mcells) per step through the reference's FirstLoopInnerThunk on %d std::threads; octree build %.2f s, loop 2 and the serial attribute pass are not in the time" % (
        stride, shape[0], shape[1], shape[2], last["slices_timed"], last["cells_timed"] * 1e-6, threads, last["octree_build_s"])
    line = {
        "impl": "reference", "metric": "mesh export throughput", "value": value, "unit": "Mvoxel/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": seconds / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(workload, shape),
        "cpu_baseline": {"value": value, "unit": "Mvoxel/s", "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "Mvoxel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except OSError:
        return {}


def check_parity(workload, mesh):
    """Layer-by-layer comparison of a host mesh with the reference's mesh of the same grid (committed fixture)."""
    path = os.path.join(ROOT, "tests", "golden", "slices_%s.json" % workload)
    if not os.path.exists(path):
        return {"fixture": None, "equal": None, "note": "no reference fixture for this workload (tests/golden/make_slices.py)"}
    from golden_util import layer_report
    with open(path) as f:
        fixture = json.load(f)
    layers, v, t, bad = layer_report(mesh.positions, mesh.normals, mesh.colors, mesh.triangles, fixture)
    return {"fixture": "tests/golden/" + os.path.basename(path), "against": "the reference's own thunks over the whole grid (oracle/_ref/tangerine_ref slices)",
            "layers": layers, "vertices_compared": v, "triangles_compared": t, "equal": not bad, "first_mismatches": [list(b) for b in bad[:4]],
            "compared": "positions, normals, colours, triangle indices per cell layer, bit for bit up to the sign of zero / NaN payloads"}


def measure_workload(T, ctx, workload, steps, warmup, flags_extra=0, with_parity=True, fp32_peak=None, hbm_peak=None, devices=1, with_fast=True):
    """One workload on `ctx` (single- or multi-device): device-resident steps, end-to-end steps, parity, roofline."""
    name, step, refine, desc = WORKLOADS[workload]
    tree, tgm = load_workload_tree(T, name)
    lo, hi = tree.bounds()
    grid = T.export_grid(lo, hi, np.float32(step))
    sx, sy, sz = grid.shape
    cells_total = sx * sy * sz
    flags = T.MESH_NORMALS | T.MESH_COLORS | flags_extra

    # ---- what a caller of MeshExport waits for: tree -> host mesh, cold (first export of this model on this context) ----
    ctx.synchronize()
    t0 = time.perf_counter()
    model = T.Model(ctx, tree)
    model_seconds = time.perf_counter() - t0
    first = model.export_mesh(grid, flags=flags, refine=refine)
    whole_cold_ms = (time.perf_counter() - t0) * 1e3
    first_export_ms = whole_cold_ms - model_seconds * 1e3
    parity = None
    if with_parity:
        if refine > 0:
            # the reference's mesh export never refines (export.cpp:320-381): its mesh is compared with the unrefined export of
            # the same grid; the refined positions are checked against the C oracle in tests/test_gpu_bench_parity.py
            plain = model.export_mesh(grid, flags=flags, refine=0)
            parity = check_parity(workload, plain)
            parity["note"] = "compared at refine 0 (the reference's MeshExportThread ignores RefineIterations); refined vertices: tests/test_gpu_bench_parity.py::test_gear512_refine5_against_oracle"
            plain.close()
        else:
            parity = check_parity(workload, first)
    first.close()
    stats = model.stats()

    # ---- device-resident throughput (`value`) ----
    def device_step():
        mesh = model.export_mesh(grid, flags=flags | T.MESH_DEVICE_ONLY, refine=refine)
        t = dict(mesh.timings)
        t["vertices"], t["triangles"] = mesh.vertex_count, mesh.triangle_count
        t["ranks"] = mesh.rank_info()
        mesh.close()
        return t

    for _ in range(warmup):
        device_step()
    ctx.synchronize()
    records = []
    ms_total = 0.0
    for _ in range(steps):
        ctx.flush_l2()                 # between timed iterations, outside the timed interval
        ctx.timer_begin()              # CUDA events on every device's own stream, the ones the kernels are launched on
        records.append(device_step())
        ms_total += ctx.timer_end()    # slowest device
    ms_per_step = ms_total / steps
    value = cells_total / (ms_per_step * 1e-3) * 1e-6

    # ---- end to end through the C ABI with host buffers (`e2e`) ----
    def e2e_step():
        model.upload()                                  # host -> device: octree table, regions, both instruction streams
        mesh = model.export_mesh(grid, flags=flags, refine=refine)      # device -> host: one pinned result mesh
        d2h = mesh.vertex_count * (12 + 12 + (3 if mesh.colors is not None else 0)) + mesh.triangle_count * 12
        mesh.close()
        return d2h

    for _ in range(max(1, warmup // 2)):
        e2e_step()
    ctx.synchronize()
    t0 = time.perf_counter()
    d2h = [e2e_step() for _ in range(steps)][-1]
    ctx.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / steps
    e2e_value = cells_total / (e2e_ms * 1e-3) * 1e-6

    # ---- the whole export once more, warm: octree build + flatten + upload + export to the host ----
    t0 = time.perf_counter()
    model2 = T.Model(ctx, tree)
    m2 = model2.export_mesh(grid, flags=flags, refine=refine)
    whole_warm_ms = (time.perf_counter() - t0) * 1e3
    m2.close()
    model2.close()
    # ... and once more after that model was destroyed: what the second MeshExport of an application costs (the context is
    # kept, the model's tables and the result arrays come out of its caches)
    t0 = time.perf_counter()
    model3 = T.Model(ctx, tree)
    model3_ms = (time.perf_counter() - t0) * 1e3
    m3 = model3.export_mesh(grid, flags=flags, refine=refine)
    whole_repeat_ms = (time.perf_counter() - t0) * 1e3
    m3.close()
    model3.close()

    # ---- the opt-in fast arithmetic (TG_MESH_FAST: FMA contraction, approximate sqrt / div), same steps ----
    fast_records = []
    fast_ms = 0.0
    if with_fast:
        for _ in range(max(1, warmup // 2)):
            model.export_mesh(grid, flags=flags | T.MESH_DEVICE_ONLY | T.MESH_FAST, refine=refine).close()
        ctx.synchronize()
        for _ in range(steps):
            ctx.flush_l2()
            ctx.timer_begin()
            mesh = model.export_mesh(grid, flags=flags | T.MESH_DEVICE_ONLY | T.MESH_FAST, refine=refine)
            fast_ms += ctx.timer_end()
            t = dict(mesh.timings)
            t["vertices"], t["triangles"], t["ranks"] = mesh.vertex_count, mesh.triangle_count, mesh.rank_info()
            fast_records.append(t)
            mesh.close()

    # ---- multi-GPU only: the same steps with the opt-in feedback (TG_MESH_REBALANCE moves the cuts by the measured times) ----
    tuned = None
    if devices > 1:
        rounds = 6
        for _ in range(rounds):
            model.export_mesh(grid, flags=flags | T.MESH_DEVICE_ONLY | T.MESH_REBALANCE, refine=refine).close()
        for _ in range(2):   # the last round moved the cuts once more: let the capacities of the final slabs settle, untimed
            model.export_mesh(grid, flags=flags | T.MESH_DEVICE_ONLY, refine=refine).close()
        ctx.synchronize()
        tuned_ms = 0.0
        tuned_ranks = None
        for _ in range(steps):
            ctx.flush_l2()
            ctx.timer_begin()
            mesh = model.export_mesh(grid, flags=flags | T.MESH_DEVICE_ONLY, refine=refine)   # cuts frozen where the feedback left them
            tuned_ms += ctx.timer_end()
            tuned_ranks = mesh.rank_info()
            mesh.close()
        tuned = {"flag": "TG_MESH_REBALANCE during %d untimed exports, then the cuts stay" % rounds, "plan_iters": rounds, "ms_per_step": tuned_ms / steps,
                 "value": cells_total / (tuned_ms / steps * 1e-3) * 1e-6, "slabs": [[b, e] for b, e, _ in tuned_ranks],
                 "per_rank_total_ms": [round(t["total_device_ms"], 4) for _, _, t in tuned_ranks]}

    last = records[-1]

    def mean(key):
        return float(np.mean([r[key] for r in records]))

    vertices, triangles = int(last["vertices"]), int(last["triangles"])
    samples, flops = float(last["samples_evaluated"]), float(last["algorithmic_flops"])
    vertex_evals = vertices * (5 * refine + 4 + (1 if stats["has_paint"] else 0))
    out = {
        "workload": workload, "config": workload_config(workload, grid.shape), "cells": cells_total,
        "ms_per_step": ms_per_step, "value": value,
        "e2e": {"value": e2e_value, "unit": "Mvoxel/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": float(stats["device_bytes"]) * devices, "d2h_bytes_per_step": float(d2h),
                "timed": "host wall clock around tg_model_upload + tg_export_mesh with pinned host results (one call, one host mesh, whatever the device count)"},
        "whole_export": {"cold_ms": whole_cold_ms, "warm_ms": whole_warm_ms, "repeat_ms": whole_repeat_ms, "repeat_model_build_ms": model3_ms, "model_build_ms": model_seconds * 1e3, "host_octree_build_ms": stats["build_seconds"] * 1e3,
                         "first_export_ms": first_export_ms,
                         "spans": "tg_model_create (host octree build + flatten + upload) + tg_export_mesh to host memory; cold = first use of the context (allocations), warm = again with a second model beside the first, repeat = again after that model was destroyed"},
        "mesh": {"vertices": vertices, "triangles": triangles},
        "bricks": {"total": float(last["bricks_total"]), "evaluated": float(last["bricks_evaluated"])},
        "evals_per_s": (samples + vertex_evals) / (ms_per_step * 1e-3),
        "reference_equivalent_evals_per_s": (8.0 * cells_total + 6.0 * vertices) / (ms_per_step * 1e-3),
        "stage_ms": {k: mean(k) for k in ("cull_ms", "evaluate_ms", "compact_ms", "faces_ms", "attributes_ms", "total_device_ms")},
        "octree_nodes": stats["octree_nodes"], "gpu_launches": int(sum(r["kernel_launches"] for r in records)),
        "parity": parity,
    }
    if last["ranks"]:
        keys = ("cull_ms", "evaluate_ms", "compact_ms", "faces_ms", "attributes_ms", "total_device_ms")
        out["per_rank_ms"] = {k: [round(float(np.mean([r["ranks"][i][2][k] for r in records])), 4) for i in range(len(last["ranks"]))] for k in keys}
        out["slabs"] = [[b, e] for b, e, _ in last["ranks"]]
    # roofline of the dominant kernel (evaluation) and of the bandwidth-bound mesh kernels
    eval_ms = mean("evaluate_ms")
    if last["ranks"]:   # multi-device: the ranks evaluate concurrently; achieved = all FLOPs / slowest rank's kernel, peak = N x one device
        eval_ms = max(out["per_rank_ms"]["evaluate_ms"])
    achieved = flops / (eval_ms * 1e-3) * 1e-12 if eval_ms > 0 else 0.0
    peak = (fp32_peak or 0.0) * devices
    mesh_bytes = cells_total / 8.0 + 36.0 * vertices + 12.0 * vertices + 12.0 * triangles + 12.0 * vertices + 3.0 * vertices
    mesh_ms = mean("compact_ms") + mean("faces_ms")
    out["roofline"] = {
        "kernel": "MeshBricksKernel", "bound": "fp32", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
        "kernel_ms": eval_ms, "share_of_step": eval_ms / mean("total_device_ms") if mean("total_device_ms") else None,
        "hbm": {"kernels": "vertex / quad numbering over the bitmap + FinalizeMeshKernel", "bound": "hbm", "achieved": mesh_bytes / (mesh_ms * 1e-3) * 1e-9 if mesh_ms > 0 else None,
                "peak": (hbm_peak or 0.0) * devices, "unit": "GB/s", "frac": (mesh_bytes / (mesh_ms * 1e-3) * 1e-9 / (hbm_peak * devices)) if mesh_ms > 0 and hbm_peak else None, "kernel_ms": mesh_ms},
    }
    if tuned:
        out["tuned"] = tuned
    if fast_records:
        f_eval = float(np.mean([r["evaluate_ms"] for r in fast_records]))
        if fast_records[-1]["ranks"]:
            f_eval = max(float(np.mean([r["ranks"][i][2]["evaluate_ms"] for r in fast_records])) for i in range(len(fast_records[-1]["ranks"])))
        f_flops = float(fast_records[-1]["algorithmic_flops"])
        f_achieved = f_flops / (f_eval * 1e-3) * 1e-12 if f_eval > 0 else 0.0
        out["fast"] = {"flag": "TG_MESH_FAST (opt-in; not bit-identical: FMA contraction, approximate sqrt / div; within the north-star tolerances, tests/test_gpu_fast.py)",
                       "ms_per_step": fast_ms / steps, "value": cells_total / (fast_ms / steps * 1e-3) * 1e-6,
                       "mesh": {"vertices": int(fast_records[-1]["vertices"]), "triangles": int(fast_records[-1]["triangles"])},
                       "roofline": {"kernel": "MeshBricksKernel (fast build)", "bound": "fp32", "achieved": f_achieved, "peak": peak, "unit": "TFLOP/s",
                                    "frac": f_achieved / peak if peak else None, "kernel_ms": f_eval}}
    model.close()
    return out, (tgm, lo, hi, float(np.float32(step)), grid.shape)


def join_ranks(backend, device=None):
    """The launch contract's rendezvous (one rank per GPU, launched by torch.distributed.run): every rank joins one
    process group of `backend` and proves it with an all-reduce; then the ranks > 0 only have to wait for rank 0, which
    they do in a gloo group -- on the host, not with a collective kernel spinning on their GPU.  Returns
    (rank, world, waiting group or None)."""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1:
        return rank, world, None
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"   # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
    if device is not None:
        dist.init_process_group(backend, device_id=device)
    else:
        dist.init_process_group(backend)
    t = torch.ones(1, device=device if device is not None else "cpu")
    dist.all_reduce(t)
    if device is not None:
        torch.cuda.synchronize()
    assert int(t.item()) == world, "rendezvous: %d of %d ranks answered" % (int(t.item()), world)
    waiters = dist.new_group(backend="gloo")
    return rank, world, waiters


def leave_ranks(waiters):
    """Ranks > 0: wait for rank 0 to finish; rank 0: release them.  Then tear the group down."""
    import torch.distributed as dist
    if waiters is not None:
        dist.barrier(group=waiters)
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="seaside1024", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-workloads", action="store_true", help="N = 1: skip the short runs of the other BASELINE configurations")
    ap.add_argument("--ref-stride", type=int, default=0, help="z-slice stride of the CPU sample (0 = auto)")
    ap.add_argument("--no-cull", action="store_true", help="evaluate every brick like the reference does")
    args = ap.parse_args()

    if args.impl == "reference":
        return reference_arm(args, args.workload)

    import torch
    import tangerine_b200 as T

    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: tangerine_b200 has no CPU path")
    torch.cuda.set_device(local)
    rank, world, waiters = join_ranks("nccl", torch.device("cuda", local))
    if rank != 0:
        # The export is ONE call in ONE process driving every GPU of the box (tg_context_create_multi): rank 0 makes it.
        leave_ranks(waiters)
        return 0

    devices = list(range(int(os.environ.get("TG_BENCH_DEVICES", world))))   # (tools: a single process driving several GPUs without torchrun)
    ctx = T.Context(devices=devices) if len(devices) > 1 else T.Context(local)
    peaks = load_peaks()
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    fp32_peak = ctx.fp32_peak_tflops()
    flags_extra = T.MESH_NO_CULL if args.no_cull else 0

    sampler = ClockSampler(local)
    if not os.environ.get("TG_BENCH_NO_SMI"):
        sampler.start()
    head, ref_args = measure_workload(T, ctx, args.workload, args.steps, args.warmup, flags_extra, True, fp32_peak, hbm_peak, len(devices))
    clocks = sampler.stop()

    # DRAM traffic of the dominant kernel: dram__bytes_read.sum + dram__bytes_write.sum of one launch from the committed
    # `ncu --set full` capture of this same workload (profiles/ncu_meshbricks.json names the capture); null otherwise
    traffic, ncu_note = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_meshbricks.json")) as f:
            cap = json.load(f)
        if cap.get("workload") == args.workload and world == 1:
            traffic = float(cap["dram_bytes_read"]) + float(cap["dram_bytes_write"])
            ncu_note = {k: cap[k] for k in ("source", "issue_active_pct", "warp_instructions", "thread_instructions_per_sample") if k in cap}
    except (OSError, ValueError, KeyError):
        pass
    roofline = head["roofline"]
    roofline.update({
        "traffic": traffic, "traffic_unit": "bytes per launch (ncu)", "ncu": ncu_note,
        "peak_source": "FP32 FMA-chain kernel measured in this run x %d device(s) (MEASURED_PEAKS.json has no CUDA-core figure); theoretical 148 SM x 128 lanes x 2 x 1.965 GHz = 74.5 per device" % world,
        "flops_convention": "SURVEY.md 8(d): FMA = 2, sqrt/div/abs/compare = 1, summed over the samples actually evaluated (culled bricks earn nothing)"})
    roofline["hbm"]["peak_source"] = "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"

    # ---- the other BASELINE.json configurations, a short run each (N = 1) ----
    workloads = None
    if world == 1 and not args.no_extra_workloads:
        workloads = {}
        for w in ("basic66", "gear512", "colorcube512", "synthetic256", "synthetic512", "synthetic1024", "synthetic2048"):
            if w == args.workload:
                continue
            try:
                r, _ = measure_workload(T, ctx, w, 3, 2, flags_extra, True, fp32_peak, hbm_peak, 1)
                workloads[w] = {"config": r["config"], "ms_per_step": r["ms_per_step"], "value": r["value"], "e2e_ms_per_step": r["e2e"]["ms_per_step"], "e2e": r["e2e"]["value"],
                                "whole_export_ms": r["whole_export"]["warm_ms"], "whole_export_repeat_ms": r["whole_export"]["repeat_ms"], "host_octree_build_ms": r["whole_export"]["host_octree_build_ms"],
                                "roofline_frac": r["roofline"]["frac"], "hbm_frac": r["roofline"]["hbm"]["frac"], "vertices": r["mesh"]["vertices"], "triangles": r["mesh"]["triangles"],
                                "bricks_evaluated_frac": r["bricks"]["evaluated"] / max(r["bricks"]["total"], 1.0), "stage_ms": r["stage_ms"],
                                "fast_ms_per_step": r["fast"]["ms_per_step"] if "fast" in r else None, "fast_roofline_frac": r["fast"]["roofline"]["frac"] if "fast" in r else None,
                                "parity_equal": r["parity"]["equal"] if r["parity"] else None}
            except T.TangerineError as e:
                workloads[w] = {"error": str(e)}
        # MagicaVoxel export of color-cube at the two grid sizes BASELINE configs[4] names (multi-cube above 126 cells)
        try:
            tree, _ = load_workload_tree(T, "color-cube")
            model = T.Model(ctx, tree)
            vox = {}
            for g in (10.0, 25.0):
                model.export_voxels(g)
                t0 = time.perf_counter()
                size, radius, xyz = model.export_voxels(g)
                vox["grid_size_%d" % g] = {"size": list(size), "voxels": int(len(xyz)), "ms": (time.perf_counter() - t0) * 1e3}
            model.close()
            workloads["colorcube_vox"] = vox
        except T.TangerineError as e:
            workloads["colorcube_vox"] = {"error": str(e)}

        # the live viewport mesher (Sodapop): tree -> Drawable arrays at the default meshing density 20 and at 50
        try:
            tree, _ = load_workload_tree(T, "seaside_town")
            live = {}
            for density in (20.0, 50.0):
                t0 = time.perf_counter()
                model = T.Model(ctx, tree, live=True)      # the uncoalesced octree of sodapop.cpp:240, 568-571
                build_ms = (time.perf_counter() - t0) * 1e3
                grid = model.live_grid(density)            # NaiveSurfaceNetsScratch, sodapop.cpp:153-179
                mesh = model.export_mesh(grid, flags=T.MESH_NORMALS | T.MESH_LIVE_FIELD)
                parity = check_parity("live_seaside%d" % density, mesh)
                mesh.close()
                ms = 0.0
                for _ in range(3):
                    ctx.flush_l2()
                    ctx.timer_begin()
                    mesh = model.export_mesh(grid, flags=T.MESH_NORMALS | T.MESH_LIVE_FIELD | T.MESH_DEVICE_ONLY)
                    ms += ctx.timer_end() / 3
                    v, t = mesh.vertex_count, mesh.triangle_count
                    mesh.close()
                live["density_%d" % density] = {"grid": list(grid.shape), "vertices": v, "triangles": t, "ms_per_step": ms, "octree_build_ms": build_ms,
                                                "parity_equal": parity["equal"], "fixture": parity["fixture"]}
                model.close()
            workloads["live_seaside"] = live
        except T.TangerineError as e:
            workloads["live_seaside"] = {"error": str(e)}

    # ---- the reference's CPU path on this box's host cores, bounded sample (N = 1 only) ----
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        tgm, lo, hi, step32, shape = ref_args
        if os.path.exists(REF_TOOL):
            stride = args.ref_stride or max(1, shape[2] // 16)
            try:
                r = run_reference_sample(tgm, lo, hi, step32, stride, threads)
                cpu_baseline = {
                    "value": r["cells_timed"] / r["loop1_s"] * 1e-6, "unit": "Mvoxel/s", "cores": threads, "kind": "reference",
                    "sample": "reference thunks (FirstLoopInnerThunk via oracle/_ref/tangerine_ref) on %d std::threads over every %d-th z-slice of the same %dx%dx%d grid: %d slices, %.3g Mcells in %.2f s; scaled to the whole grid this is %.0f s (extrapolated); octree build %.2f s, loop 2 and the serial attribute pass not included"
                              % (threads, stride, shape[0], shape[1], shape[2], r["slices_timed"], r["cells_timed"] * 1e-6, r["loop1_s"], r["loop1_s"] * r["cells_total"] / r["cells_timed"], r["octree_build_s"]),
                }
            except (subprocess.CalledProcessError, ValueError) as e:
                cpu_baseline = {"value": None, "unit": "Mvoxel/s", "cores": threads, "kind": "reference", "sample": "failed: %s" % e}
        else:
            cpu_baseline = {"value": None, "unit": "Mvoxel/s", "cores": threads, "kind": "reference", "sample": "oracle/_ref/tangerine_ref not built"}

    config = dict(head["config"])   # the same object the reference arm prints
    run = {"attributes": "normals+colours", "culling": not args.no_cull,
           "partition": ("z-slabs %s cut from the host-side work estimate, plan_iters 0" % head.get("slabs")) if world > 1 else "one device"}
    line = {
        "metric": "mesh export throughput", "value": head["value"], "unit": "Mvoxel/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config, "run": run, "plan_iters": 0,
        "evals_per_s": head["evals_per_s"], "reference_equivalent_evals_per_s": head["reference_equivalent_evals_per_s"],
        "mesh": head["mesh"], "bricks": head["bricks"], "stage_ms_rank0": head["stage_ms"], "per_rank_ms": head.get("per_rank_ms"),
        "octree_nodes": head["octree_nodes"], "e2e": head["e2e"], "whole_export": head["whole_export"], "parity": head["parity"],
        "gpu_launches": head["gpu_launches"], "roofline": roofline, "fast": head.get("fast"), "tuned": head.get("tuned"), "workloads": workloads, "cpu_baseline": cpu_baseline, "clocks": clocks,
    }
    print(json.dumps(line))
    ctx.close()
    leave_ranks(waiters)
    return 0


if __name__ == "__main__":
    sys.exit(main())
