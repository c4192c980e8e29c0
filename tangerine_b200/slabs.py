"""Host-side logic of the multi-GPU path: z-slab cuts and the one exchange of the path.

The grid shards into z-slabs on 8-layer (brick) boundaries, one per rank, each with a one-layer halo below it.
Every rank numbers its vertices locally in (k, j, i) order; the only collective is an all-gather of two integers
per rank (owned vertices, triangles), after which `global id = local id + sum of the lower ranks' vertices`
(tg_mesh.halo_vertices is already subtracted by the engine).  No compute here and no CUDA: the functions take
whatever process group the caller initialised (NCCL on the GPU box, gloo in the CPU tests).
"""
import numpy as np

BRICK = 8  # cell layers per brick layer: slab boundaries are multiples of this (tg_mesh_options.slab_begin/end)


def balanced_slabs(profile, world, sz):
    """Cut [0, sz) into `world` z-slabs on 8-layer boundaries so that each holds about the same share of
    `profile` (work per brick layer b = cell layers [8b, 8b+8)).  Deterministic, so every rank computes the
    same cut from the same profile without communicating.  Every slab gets at least one brick layer; needs
    len(profile) >= world."""
    nb = len(profile)
    if nb < world:
        raise ValueError("grid has %d brick layers, fewer than the %d ranks" % (nb, world))
    cost = np.asarray(profile, np.float64) + 1e-3
    cum = np.concatenate([[0.0], np.cumsum(cost)])
    cuts = [0]
    for r in range(1, world):
        target = cum[-1] * r / world
        b = int(np.searchsorted(cum, target))
        b = max(b, cuts[-1] + 1)
        b = min(b, nb - (world - r))
        cuts.append(b)
    cuts.append(nb)
    return [(cuts[r] * BRICK, min(cuts[r + 1] * BRICK, sz)) for r in range(world)]


def uniform_slabs(world, sz):
    """Equal-thickness slabs (the cut used before any work profile exists)."""
    nb = (sz + BRICK - 1) // BRICK
    return balanced_slabs(np.ones(nb), world, sz)


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
