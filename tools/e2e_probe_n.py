"""Where does the end-to-end time of an N-GPU export go?  (diagnostic; launch under torch.distributed.run)

Times the four phases of bench.py's e2e step separately on every rank -- model upload, slab export, count exchange,
download -- with a device synchronise after each, then all phases together without the extra synchronisation."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import tangerine_b200 as T
from tangerine_b200.slabs import balanced_slabs, exchange_counts

rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
tree = T.Tree.load("tests/golden/models/seaside_town.tgm")
lo, hi = tree.bounds()
grid = T.export_grid(lo, hi, np.float32(10 / 1022))
sz = grid.shape[2]
ctx = T.Context(local)
model = T.Model(ctx, tree)
slab = balanced_slabs(model.brick_profile(grid).astype(np.float64), world, sz, 1)[rank]
flags = T.MESH_NORMALS | T.MESH_COLORS


def sync():
    ctx.synchronize()
    torch.cuda.synchronize()


rows = []
for it in range(6):
    dist.barrier(); sync()
    t = [time.perf_counter()]
    model.upload(); sync(); t.append(time.perf_counter())
    mesh = model.export_mesh(grid, flags=flags | T.MESH_DEVICE_ONLY, slab=slab); sync(); t.append(time.perf_counter())
    base, _, _, _ = exchange_counts(mesh.vertex_count, mesh.triangle_count, rank, world, device="cuda"); sync(); t.append(time.perf_counter())
    mesh.download(index_base=base); sync(); t.append(time.perf_counter())
    nbytes = mesh.vertex_count * 27 + mesh.triangle_count * 12
    mesh.close(); t.append(time.perf_counter())
    rows.append([(t[i + 1] - t[i]) * 1e3 for i in range(5)] + [nbytes / 1e6])
mine = torch.tensor(rows[2:], dtype=torch.float64, device="cuda").mean(0)
allr = torch.zeros((world, mine.numel()), dtype=torch.float64, device="cuda")
dist.all_gather_into_tensor(allr, mine)
if rank == 0:
    print("per rank: upload  export  exchange  download  close  [ms]   result MB")
    for r, row in enumerate(allr.cpu().numpy()):
        print("  rank %d  %s   %.1f" % (r, "  ".join("%6.3f" % v for v in row[:5]), row[5]))
for it in range(5):
    dist.barrier(); sync()
    t0 = time.perf_counter()
    model.upload()
    mesh = model.export_mesh(grid, flags=flags | T.MESH_DEVICE_ONLY, slab=slab)
    base, _, _, _ = exchange_counts(mesh.vertex_count, mesh.triangle_count, rank, world, device="cuda")
    mesh.download(index_base=base)
    mesh.close()
    t1 = time.perf_counter()
    tt = torch.tensor([(t1 - t0) * 1e3], dtype=torch.float64, device="cuda")
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0 and it >= 2:
        print("whole step, max over ranks: %.3f ms" % float(tt.item()))
dist.destroy_process_group()
