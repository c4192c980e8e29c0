#!/usr/bin/env python
"""bench.py -- the meshing hot path on N B200s, beside the reference's CPU path on the box's host cores.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload seaside1024] [--impl ours|reference]

A "step" is one whole mesh export of the workload: culling, brick evaluation + cell classification, vertex
numbering, quad emission, refinement + normals + colours.  With N > 1 (launched by torch.distributed.run, one
rank per GPU) the grid is cut into z-slabs, one per rank, with a one-layer halo; the only exchange is an
all-gather of the per-slab vertex / triangle counts, so the job is the same fixed grid at
every N ("strong" scaling).

JSON keys (one line on stdout, rank 0):
  value          Mvoxel/s = grid cells / device time per step, model tables resident in HBM, results left in HBM
  e2e            Mvoxel/s through the C ABI the reference would bind (tg_model_upload + tg_export_mesh with
                 host result buffers): per step the model tables go host -> device and the mesh comes back
  evals_per_s    SDF evaluations per second (unique lattice samples run + per-vertex evaluations)
  roofline       dominant kernel = MeshBricksKernel (evaluation + classification), FP32-pipe bound; `hbm` holds the
                 bandwidth-bound mesh kernels (vertex numbering / scatter / quad emission)
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


def load_workload_tree(T, name):
    """Returns (tree, path of a .tgm file of it for the reference tool)."""
    if name.startswith("synthetic:"):
        import tempfile
        tree = T.Tree.synthetic(int(name.split(":")[1]), 1234)
        path = os.path.join(tempfile.gettempdir(), "tg_%s_%d.tgm" % (name.replace(":", "_"), os.getpid()))
        tree.save(path)
        return tree, path
    path = os.path.join(MODELS, name + ".tgm")
    return T.Tree.load(path), path


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


def run_reference_sample(model_file, lo, hi, step, stride, threads):
    """The reference's FirstLoopInnerThunk / SecondLoopThunk on std::threads over every `stride`-th z-slice."""
    args = [REF_TOOL, "bench", model_file] + ["%.9g" % v for v in list(lo) + list(hi)] + ["%.9g" % step, str(threads), str(stride)]
    out = subprocess.run(args, check=True, capture_output=True, text=True).stdout
    return json.loads(out.strip().splitlines()[-1])


def reference_arm(args, workload):
    """--impl reference: the reference CPU implementation, all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    name, step, refine, desc = WORKLOADS[workload]
    if not os.path.exists(REF_TOOL):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/tangerine_ref was not built (run __graft_entry__.build() where /root/reference exists)"}))
        return 0
    import tangerine_b200 as T
    tree, model_file = load_workload_tree(T, name)
    lo, hi = tree.bounds()
    step32 = float(np.float32(step))
    threads = os.cpu_count() or 1
    grid = T.export_grid(lo, hi, np.float32(step))
    stride = args.ref_stride or max(1, grid.shape[2] // 8)
    for _ in range(args.warmup):
        run_reference_sample(model_file, lo, hi, step32, max(stride * 4, 1), threads)
    cells = 0.0
    seconds = 0.0
    last = None
    for _ in range(args.steps):
        last = run_reference_sample(model_file, lo, hi, step32, stride, threads)
        cells += last["cells_timed"]
        seconds += last["loop1_s"] + last["loop2_s"]
    value = cells / seconds * 1e-6
    sample = "every %d-th z-slice of the %dx%dx%d grid (%d slices, %.3g Mcells) per step, loop 1 (FirstLoopInnerThunk) only; octree build %.2f s excluded" % (
        stride, grid.shape[0], grid.shape[1], grid.shape[2], last["slices_timed"], last["cells_timed"] * 1e-6, last["octree_build_s"])
    line = {
        "impl": "reference", "metric": "mesh export throughput", "value": value, "unit": "Mvoxel/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": seconds / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "grid": list(grid.shape), "model": name + ".tgm"},
        "cpu_baseline": {"value": value, "unit": "Mvoxel/s", "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "Mvoxel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="seaside1024", choices=sorted(WORKLOADS))
    ap.add_argument("--refine", type=int, default=None, help="override the workload's refinement iterations")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-stride", type=int, default=0, help="z-slice stride of the CPU sample (0 = auto)")
    ap.add_argument("--no-cull", action="store_true", help="evaluate every brick like the reference does")
    ap.add_argument("--slab-align", type=int, default=1, help="z-slab cuts fall on multiples of this many cell layers (8 = whole brick rows)")
    args = ap.parse_args()

    if args.impl == "reference":
        return reference_arm(args, args.workload)

    import torch
    import torch.distributed as dist
    import tangerine_b200 as T
    from tangerine_b200.slabs import balanced_slabs, exchange_counts, layer_costs, rebalance

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: tangerine_b200 has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"   # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    name, step, refine, desc = WORKLOADS[args.workload]
    if args.refine is not None:
        refine = args.refine
    tree, model_file = load_workload_tree(T, name)
    lo, hi = tree.bounds()
    grid = T.export_grid(lo, hi, np.float32(step))
    sx, sy, sz = grid.shape
    cells_total = sx * sy * sz

    ctx = T.Context(local)
    t0 = time.perf_counter()
    model = T.Model(ctx, tree)
    model_seconds = time.perf_counter() - t0
    stats = model.stats()
    flags = T.MESH_NORMALS | T.MESH_COLORS | (T.MESH_NO_CULL if args.no_cull else 0)

    # z-slab partition.  First cut: the cull-only work estimate every rank computes identically (tg_brick_profile).
    # After the first warm-up export the measured per-layer vertex cost (sum of program FLOPs over a layer's vertices) and stage times are all-reduced and the
    # cut is redone on cost = eval_rate * brick_weight + vertex_rate * vertex_cost (same inputs on every rank, so
    # no further communication is needed to agree on it).
    if world > 1:
        profile = model.brick_profile(grid).astype(np.float64)
        slabs = balanced_slabs(profile, world, sz, args.slab_align)
        slab = slabs[rank]
    else:
        profile = None
        slabs = [(0, sz)]
        slab = None

    def barrier():
        ctx.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def all_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_sum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- device-resident throughput (`value`) -------------------------------------------------------------
    def device_step():
        mesh = model.export_mesh(grid, flags=flags | T.MESH_DEVICE_ONLY, refine=refine, slab=slab)
        t = dict(mesh.timings)
        t["vertices"], t["triangles"] = mesh.vertex_count, mesh.triangle_count
        mesh.close()
        return t

    cost = None          # modelled work per CELL layer
    # Partition planning (world > 1): untimed exports that move the cuts by the measured per-rank times.  This is
    # set-up, like the octree build -- it runs a fixed number of times whatever --warmup says -- and the partition
    # with the smallest slowest rank seen on the way is the one that is then warmed up and timed.
    plan_iters = 14 if world > 1 else 0
    best_time, best_slabs = float("inf"), slabs
    for w in range(plan_iters):
        mesh = model.export_mesh(grid, flags=flags | T.MESH_DEVICE_ONLY, refine=refine, slab=slab)
        tm = mesh.timings
        n = len(profile)
        brick_work = layer_costs(profile, sz)
        mine = np.zeros(n + 4 + world, np.float64)
        mine[:n] = mesh.layer_vertex_cost[:n]
        mine[n:n + 4] = [tm["evaluate_ms"] + tm["cull_ms"], brick_work[slab[0]:slab[1]].sum(), tm["compact_ms"] + tm["faces_ms"] + tm["attributes_ms"], mesh.layer_vertex_cost.sum()]
        mine[n + 4 + rank] = tm["total_device_ms"]
        mesh.close()
        t = torch.from_numpy(mine).cuda()
        dist.all_reduce(t)
        allv = t.cpu().numpy()
        times = [float(allv[n + 4 + r]) for r in range(world)]
        if w >= 1 and max(times) < best_time:      # the very first export also pays for first-touch allocations
            best_time, best_slabs = max(times), list(slabs)
        if cost is None:
            # first model: two rates fitted to the measured stage times of all ranks
            eval_rate = allv[n] / max(allv[n + 1], 1.0)
            vertex_rate = allv[n + 2] / max(allv[n + 3], 1.0)
            cost = eval_rate * brick_work + vertex_rate * layer_costs(allv[:n], sz) + 1e-9
        # feedback: rescale every slab's layers so that the model reproduces the time that slab just took
        for r, (k0, k1) in enumerate(slabs):
            predicted = cost[k0:k1].sum()
            if predicted > 0 and times[r] > 0:
                cost[k0:k1] *= times[r] / predicted
        if w < 2:
            slabs = balanced_slabs(cost, world, sz, args.slab_align)
        else:
            # the cost model has placed the cuts roughly; from here on they move by the measured times alone
            slabs = rebalance(slabs, times, sz, args.slab_align)
        slab = slabs[rank]
    if world > 1:
        slabs = best_slabs
        slab = slabs[rank]
    for w in range(args.warmup):
        model.export_mesh(grid, flags=flags | T.MESH_DEVICE_ONLY, refine=refine, slab=slab).close()
    sampler = ClockSampler(local)
    if rank == 0 and not os.environ.get("TG_BENCH_NO_SMI"):
        sampler.start()
    barrier()
    steps = []
    ms_local = 0.0
    for _ in range(args.steps):
        ctx.flush_l2()                 # between timed iterations, outside the timed interval
        ctx.timer_begin()              # CUDA events on the context's own stream, the one every kernel is launched on
        steps.append(device_step())
        ms_local += ctx.timer_end()
    barrier()
    ms_total = all_max(ms_local)
    ms_per_step = ms_total / args.steps
    # per-rank view of the same region: wall (events around the K steps) and the sum of the engine's stage timers
    stage_keys = ("cull_ms", "evaluate_ms", "compact_ms", "faces_ms", "attributes_ms")
    per_rank = [[ms_local / args.steps, float(np.mean([s["total_device_ms"] for s in steps]))] + [float(np.mean([s[k] for s in steps])) for k in stage_keys]]
    if world > 1:
        t = torch.zeros((world, len(per_rank[0])), dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(t, torch.tensor(per_rank[0], dtype=torch.float64, device="cuda"))
        per_rank = t.cpu().numpy().tolist()
    value = cells_total / (ms_per_step * 1e-3) * 1e-6

    # ---- end to end through the C ABI with host buffers (`e2e`) ------------------------------------------
    def e2e_step():
        model.upload()                                  # host -> device: octree table, regions, both instruction streams
        if world == 1:
            mesh = model.export_mesh(grid, flags=flags, refine=refine)      # device -> host: pinned result arrays
        else:
            mesh = model.export_mesh(grid, flags=flags | T.MESH_DEVICE_ONLY, refine=refine, slab=slab)
            # the one exchange of the path: per-slab counts -> exclusive prefix -> global vertex ids
            base, _, _, _ = exchange_counts(mesh.vertex_count, mesh.triangle_count, rank, world, device="cuda")
            mesh.download(index_base=base)              # rebase on the device, then device -> host
        d2h = mesh.vertex_count * (12 + 12 + (3 if mesh.colors is not None else 0)) + mesh.triangle_count * 12
        v, f = mesh.vertex_count, mesh.triangle_count
        mesh.close()
        return d2h, v, f

    for _ in range(max(1, args.warmup // 2)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    ctx.timer_begin()
    e2e_results = [e2e_step() for _ in range(args.steps)]
    e2e_dev_ms = ctx.timer_end()
    barrier()
    e2e_wall_ms = all_max((time.perf_counter() - t0) * 1e3)
    e2e_value = cells_total / (e2e_wall_ms / args.steps * 1e-3) * 1e-6
    clocks = sampler.stop() if rank == 0 else None
    d2h_total = all_sum(float(e2e_results[-1][0]))
    h2d_total = float(stats["device_bytes"]) * world

    # ---- whole-job tallies ------------------------------------------------------------------------------------
    last = steps[-1]
    vertices = int(all_sum(float(last["vertices"])))
    triangles = int(all_sum(float(last["triangles"])))
    samples = all_sum(float(last["samples_evaluated"]))
    flops = all_sum(float(last["algorithmic_flops"]))
    launches = int(all_sum(float(sum(s["kernel_launches"] for s in steps))))
    bricks_total = all_sum(float(last["bricks_total"]))
    bricks_eval = all_sum(float(last["bricks_evaluated"]))
    # per-vertex evaluations: R x (4-tap gradient + 1) + 4-tap normal + 1 material walk
    vertex_evals = vertices * (5 * refine + 4 + (1 if stats["has_paint"] else 0))
    evals_per_s = (samples + vertex_evals) / (ms_per_step * 1e-3)
    reference_equivalent_evals = 8.0 * cells_total + 6.0 * vertices

    def mean(key):
        return float(np.mean([s[key] for s in steps]))

    # ---- roofline of the dominant kernel (rank 0's slab; kernel time from CUDA events inside the engine) ----
    fp32_peak = ctx.fp32_peak_tflops()
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    eval_ms = mean("evaluate_ms")
    r0_flops = float(last["algorithmic_flops"])
    achieved_tflops = r0_flops / (eval_ms * 1e-3) * 1e-12 if eval_ms > 0 else 0.0
    r0_v, r0_f = last["vertices"], last["triangles"]
    slab_cells = sx * sy * ((slab[1] - slab[0]) if slab else sz)
    # SURVEY.md 8d "algorithmic bytes (mesh side)": cell->vertex map + neighbour ids + positions + indices + normals + colours
    mesh_bytes = slab_cells / 8.0 + 36.0 * r0_v + 12.0 * r0_v + 12.0 * r0_f + 12.0 * r0_v + 3.0 * r0_v
    mesh_ms = mean("compact_ms") + mean("faces_ms")
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
    roofline = {
        "kernel": "MeshBricksKernel", "bound": "fp32", "achieved": achieved_tflops, "peak": fp32_peak, "unit": "TFLOP/s",
        "frac": achieved_tflops / fp32_peak if fp32_peak else None, "traffic": traffic, "traffic_unit": "bytes per launch (ncu)", "ncu": ncu_note,
        "peak_source": "FP32 FMA-chain kernel measured in this run (MEASURED_PEAKS.json has no CUDA-core figure); theoretical 148 SM x 128 lanes x 2 x 1.965 GHz = 74.5",
        "flops_convention": "SURVEY.md 8(d): FMA = 2, sqrt/div/abs/compare = 1, summed over the samples actually evaluated (culled bricks earn nothing)",
        "kernel_ms": eval_ms, "share_of_step": eval_ms / mean("total_device_ms") if mean("total_device_ms") else None,
        "hbm": {"kernels": "dual vertex/quad scan over the bitmap (3 launches) + FinalizeMeshKernel", "bound": "hbm", "achieved": mesh_bytes / (mesh_ms * 1e-3) * 1e-9 if mesh_ms > 0 else None,
                "peak": hbm_peak, "unit": "GB/s", "frac": (mesh_bytes / (mesh_ms * 1e-3) * 1e-9 / hbm_peak) if mesh_ms > 0 else None,
                "kernel_ms": mesh_ms, "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"},
    }

    # ---- the reference's CPU path on this box's host cores, bounded sample (rank 0, N = 1 only) -----------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        if os.path.exists(REF_TOOL):
            stride = args.ref_stride or max(1, sz // 16)
            try:
                r = run_reference_sample(model_file, lo, hi, float(np.float32(step)), stride, threads)
                cpu_baseline = {
                    "value": r["cells_timed"] / (r["loop1_s"] + r["loop2_s"]) * 1e-6, "unit": "Mvoxel/s", "cores": threads, "kind": "reference",
                    "sample": "reference thunks (FirstLoopInnerThunk via oracle/_ref/tangerine_ref) on %d std::threads over every %d-th z-slice of the same %dx%dx%d grid: %d slices, %.3g Mcells in %.2f s; scaled to the whole grid this is %.0f s (extrapolated); octree build %.2f s, loop 2 and the serial attribute pass not included"
                              % (threads, stride, sx, sy, sz, r["slices_timed"], r["cells_timed"] * 1e-6, r["loop1_s"], r["loop1_s"] * r["cells_total"] / r["cells_timed"], r["octree_build_s"]),
                }
            except (subprocess.CalledProcessError, ValueError) as e:
                cpu_baseline = {"value": None, "unit": "Mvoxel/s", "cores": threads, "kind": "reference", "sample": "failed: %s" % e}
        else:
            cpu_baseline = {"value": None, "unit": "Mvoxel/s", "cores": threads, "kind": "reference", "sample": "oracle/_ref/tangerine_ref not built"}

    if rank == 0:
        line = {
            "metric": "mesh export throughput", "value": value, "unit": "Mvoxel/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "model": (name + ".tgm (CSG tree dumped from the reference's Lua front-end)") if not name.startswith("synthetic:") else "tg_make_synthetic(%s, 1234)" % name.split(":")[1], "grid": [sx, sy, sz], "refine_iterations": refine,
                       "attributes": "normals+colours", "culling": not args.no_cull, "partition": "z-slabs %s (cuts on multiples of %d layers)" % (slabs, args.slab_align),
                       "l2": "flushed before every timed step (256 MiB fill, outside the per-step event pair); per-step scratch (bitmap + prefix) also exceeds L2 at this grid"},
            "evals_per_s": evals_per_s, "reference_equivalent_evals_per_s": reference_equivalent_evals / (ms_per_step * 1e-3),
            "mesh": {"vertices": vertices, "triangles": triangles},
            "bricks": {"total": bricks_total, "evaluated": bricks_eval},
            "stage_ms_rank0": {k: mean(k) for k in ("cull_ms", "evaluate_ms", "compact_ms", "faces_ms", "attributes_ms", "total_device_ms")},
            "per_rank_ms": dict({"step_wall": [round(r[0], 4) for r in per_rank], "stage_sum": [round(r[1], 4) for r in per_rank]},
                                **{k: [round(r[2 + i], 4) for r in per_rank] for i, k in enumerate(stage_keys)}),
            "model_build_s": model_seconds, "octree_nodes": stats["octree_nodes"],
            "e2e": {"value": e2e_value, "unit": "Mvoxel/s", "ms_per_step": e2e_wall_ms / args.steps, "h2d_bytes_per_step": h2d_total, "d2h_bytes_per_step": d2h_total,
                    "timed": "host wall clock around tg_model_upload + tg_export_mesh (pinned host results)%s, max over ranks" % (" + NCCL count all-gather + tg_mesh_download (index rebase on device)" if world > 1 else ""),
                    "device_ms_per_step_rank0": e2e_dev_ms / args.steps},
            "gpu_launches": launches,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "clocks": clocks,
        }
        print(json.dumps(line))
    model.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
