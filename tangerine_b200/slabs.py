"""Host-side logic of the multi-GPU path: z-slab cuts and the one exchange of the path.

The grid shards into z-slabs, one per rank, each with a one-layer halo below it.
Every rank numbers its vertices locally in (k, j, i) order; the only collective is an all-gather of two integers
per rank (owned vertices, triangles), after which `global id = local id + sum of the lower ranks' vertices`
(tg_mesh.halo_vertices is already subtracted by the engine).  No compute here and no CUDA: the functions take
whatever process group the caller initialised (NCCL on the GPU box, gloo in the CPU tests).
"""
import numpy as np

BRICK = 8  # cell layers per brick layer: the unit of the work profiles (tg_brick_profile, tg_mesh.layer_vertex_cost)


def layer_costs(profile, sz):
    """Work per CELL layer from a profile given per cell layer (len == sz) or per brick layer (len == ceil(sz / 8),
    taken as uniform inside the brick layer)."""
    p = np.asarray(profile, np.float64)
    if len(p) != sz:
        if len(p) != (sz + BRICK - 1) // BRICK:
            raise ValueError("profile has %d entries for %d layers" % (len(p), sz))
        p = np.repeat(p / BRICK, BRICK)[:sz]
    return p + 1e-3 / BRICK


def balanced_slabs(profile, world, sz, align=BRICK):
    """Cut [0, sz) into `world` z-slabs so that each holds about the same share of `profile` (work per cell layer, or
    per brick layer b = cell layers [8b, 8b+8), see layer_costs).  Cuts fall on multiples of `align` cell layers:
    8 keeps every brick with one rank; smaller values balance finer at the price of evaluating the bricks of a cut
    row on both sides.  Deterministic, so every rank computes the same cut from the same profile without
    communicating.  Every slab gets at least `align` layers; needs sz >= world * align."""
    align = max(1, int(align))
    steps = (sz + align - 1) // align            # candidate cut positions are align * [0 .. steps]
    if steps < world:
        raise ValueError("grid has %d layers, too few for %d ranks at alignment %d" % (sz, world, align))
    per_layer = layer_costs(profile, sz)
    cum = np.concatenate([[0.0], np.cumsum(per_layer)])
    below = cum[np.minimum(np.arange(steps + 1) * align, sz)]   # work of cell layers [0, k) at every candidate cut
    cuts = [0]
    for r in range(1, world):
        target = below[-1] * r / world
        i = int(np.searchsorted(below, target))
        if i > 0 and abs(below[i - 1] - target) <= abs(below[min(i, steps)] - target):
            i -= 1
        i = max(i, cuts[-1] + 1)
        i = min(i, steps - (world - r))
        cuts.append(i)
    cuts.append(steps)
    return [(cuts[r] * align, min(cuts[r + 1] * align, sz)) for r in range(world)]


def rebalance(slabs, times, sz, align=1, relax=0.7):
    """One step of measured-time balancing: every cut between neighbouring slabs moves towards the slower side by
    the number of layers that would equalise the two, judged by their mean time per layer, damped by `relax`.
    Needs only the per-slab times of the last run (every rank holds all of them after one all-reduce), makes no
    assumption about where inside a slab the work sits, and converges where a global cost model oscillates."""
    world = len(slabs)
    if world == 1:
        return list(slabs)
    cuts = [s[0] for s in slabs] + [slabs[-1][1]]
    t = [max(float(x), 1e-9) for x in times]
    new = list(cuts)
    for b in range(1, world):
        left, right = b - 1, b
        len_l, len_r = cuts[b] - cuts[b - 1], cuts[b + 1] - cuts[b]
        rho_l, rho_r = t[left] / len_l, t[right] / len_r
        delta = relax * (t[right] - t[left]) / (rho_l + rho_r)      # > 0: the right slab is slower, the cut moves up
        delta = max(-0.5 * (len_l - align), min(0.5 * (len_r - align), delta))
        new[b] = int(round((cuts[b] + delta) / align)) * align
    for b in range(1, world):                                        # keep every slab at least `align` thick
        new[b] = max(new[b], new[b - 1] + align)
    for b in range(world - 1, 0, -1):
        new[b] = min(new[b], new[b + 1] - align if b + 1 < world else sz - align)
    new[world] = sz
    return [(new[r], new[r + 1]) for r in range(world)]


def uniform_slabs(world, sz):
    """Equal-thickness slabs (the cut used before any work profile exists)."""
    return balanced_slabs(np.ones(sz), world, sz)


def exchange_counts(vertex_count, triangle_count, rank, world, device=None, group=None):
    """The path's one collective: all-gather of (owned vertices, triangles) per slab.

    Returns (index_base, total_vertices, total_triangles, per_rank) where index_base is what this rank adds to
    its local triangle indices (tg_mesh_download's index_base) and per_rank is a (world, 2) int64 array."""
    if world == 1:
        return 0, int(vertex_count), int(triangle_count), np.array([[vertex_count, triangle_count]], np.int64)
    import torch
    import torch.distributed as dist
    mine = torch.tensor([int(vertex_count), int(triangle_count)], dtype=torch.int64, device=device)
    gathered = torch.zeros(world * 2, dtype=torch.int64, device=device)  # flat: gloo accepts only the concatenated form
    dist.all_gather_into_tensor(gathered, mine, group=group)
    per_rank = gathered.cpu().numpy().reshape(world, 2)
    base = int(per_rank[:rank, 0].sum())
    return base, int(per_rank[:, 0].sum()), int(per_rank[:, 1].sum()), per_rank


def stitch(parts):
    """Host-side concatenation of per-slab results in rank order (what rank 0 / the writer does).
    parts: list of (positions (V,3), triangles (F,3) local indices, already rebased or not, index_base)."""
    positions = np.concatenate([p[0] for p in parts]) if parts else np.zeros((0, 3), np.float32)
    triangles = np.concatenate([(p[1].astype(np.int64) + int(p[2])).astype(np.uint32) for p in parts]) if parts else np.zeros((0, 3), np.uint32)
    return positions, triangles
