"""Multi-GPU partition of the film (one process per GPU).

The sample tiles of SamplerIntegrator::Render (16x16 pixels, src/core/integrator.cpp:235-240) are dealt
round-robin to the ranks: tile t (row-major over the sample bounds) belongs to rank t % world.  The BVH and
every other scene table are replicated; each rank renders its tiles into its own zero-initialised film and
one reduce(sum) over the W*H*4 floats merges them on rank 0 — the distributed form of Film::MergeFilmTile
(src/core/film.cpp:117-130).  With the box filter a sample touches a neighbouring tile's pixel only when it
falls exactly on a pixel boundary, which is why the merge is a sum and not a gather.
"""
import numpy as np

TILE = 16


def tile_owner_map(sample_bounds, world):
    """owner[y, x] = rank that renders sample-space pixel (x, y); sample_bounds = (x0, y0, x1, y1)."""
    x0, y0, x1, y1 = sample_bounds
    nx = (x1 - x0 + TILE - 1) // TILE
    ys, xs = np.mgrid[y0:y1, x0:x1]
    tile = ((ys - y0) // TILE) * nx + (xs - x0) // TILE
    return (tile % world).astype(np.int32)


def owned_tile_count(sample_bounds, rank, world):
    x0, y0, x1, y1 = sample_bounds
    n = ((x1 - x0 + TILE - 1) // TILE) * ((y1 - y0 + TILE - 1) // TILE)
    return (n - rank + world - 1) // world if n > rank else 0


def reduce_film(film, dst=0):
    """Sum the per-rank films onto rank `dst` (torch.distributed; NCCL over NVLink on the GPU box, gloo in CPU tests)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(film, dst=dst, op=dist.ReduceOp.SUM)
    return film
