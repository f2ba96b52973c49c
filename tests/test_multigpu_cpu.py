"""The N > 1 path on the CPU, world_size 2 over gloo: each rank takes ITS work items from the product's own partition
function (pb2_work_items: the render kernels' decode function compiled for the host), renders them with the oracle
port's PathIntegrator::Li, deposits them like FilmTile::AddSample with the box filter, and one reduce(sum) to rank 0
must give the film of the port's single-process SamplerIntegrator::Render - the reference's MergeFilmTile semantics
(src/core/film.cpp:117-130) spread over processes.  No GPU: the product contributes the partition, the oracle the pixels."""
import os
import socket
import subprocess
import sys

import numpy as np

from conftest import ROOT

WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import pbrt_v3_b200 as pb
from pbrt_v3_b200 import multigpu
from oracle import pyoracle
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank, world = dist.get_rank(), dist.get_world_size()
XRES, YRES, SPP = 50, 35, 4                       # ragged: neither side is a multiple of the 16-pixel tile
hs = pb.HostScene.soup(3000, xres=XRES, yres=YRES, spp=SPP, maxdepth=5)
port = pyoracle.port()
sc = port.scene(hs)
items = multigpu.work_items(hs.film, hs.params, rank, world)          # (x, y, sample) this rank owns
li, pfilm = sc.li_samples(items[:, :2], items[:, 2].astype(np.int64))
film = np.zeros((YRES, XRES, 4), np.float64)
# FilmTile::AddSample with the box filter of radius 0.5 (film.h:121-161): pixels [ceil(p - 1), floor(p) + 1) per axis
d = pfilm.astype(np.float32) - np.float32(0.5)
x0, x1 = np.ceil(d[:, 0] - np.float32(0.5)).astype(int), np.floor(d[:, 0] + np.float32(0.5)).astype(int) + 1
y0, y1 = np.ceil(d[:, 1] - np.float32(0.5)).astype(int), np.floor(d[:, 1] + np.float32(0.5)).astype(int) + 1
for i in range(len(items)):
    for y in range(max(y0[i], 0), min(y1[i], YRES)):
        for x in range(max(x0[i], 0), min(x1[i], XRES)):
            film[y, x, :3] += li[i]
            film[y, x, 3] += 1
t = torch.from_numpy(film)
dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
n_all = torch.tensor([len(items)])
dist.reduce(n_all, dst=0, op=dist.ReduceOp.SUM)
if rank == 0:
    assert int(n_all) == XRES * YRES * SPP, "the two partitions must cover every (pixel, sample) exactly once"
    merged = hs.resolve(t.numpy().astype(np.float32))                   # MergeFilmTile's RGB -> XYZ + WriteImage
    want, _, _ = sc.render(n_threads=1)
    assert (t.numpy()[..., 3] >= SPP).all()
    rel = np.abs(merged - want) / np.maximum(np.abs(want), 1e-3)
    assert rel.max() <= 1e-4, float(rel.max())
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok", len(items))
'''


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_ranks_render_their_tiles_and_reduce_to_the_single_process_film(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = str(free_port())
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o


def test_work_partition_covers_every_sample_once():
    """pb2_work_items for world sizes 1, 2, 3, 8 on a frame with a crop-free ragged tile grid and with pixelbounds: the
    partitions are disjoint, cover all (pixel, sample) pairs, and a 16 x 16 tile never straddles two ranks."""
    import pbrt_v3_b200 as pb
    from pbrt_v3_b200 import multigpu
    hs = pb.HostScene.soup(100, xres=50, yres=35, spp=3)
    for world in (1, 2, 3, 8):
        seen = {}
        for r in range(world):
            items = multigpu.work_items(hs.film, hs.params, r, world)
            for x, y, s in items.tolist():
                assert (x, y, s) not in seen
                seen[(x, y, s)] = r
        assert len(seen) == 50 * 35 * 3
        for (x, y, s), r in seen.items():
            assert r == ((y // 16) * 4 + x // 16) % world
    p = hs.params_copy()
    p.pixel_bounds[0], p.pixel_bounds[1], p.pixel_bounds[2], p.pixel_bounds[3] = 10, 5, 30, 20
    total = sum(len(multigpu.work_items(hs.film, p, r, 2)) for r in range(2))
    assert total == 20 * 15 * 3


def test_dist_entry_points_without_a_device():
    """pb2_dist_init needs pb2_init (there is no CPU fallback); world 1 / no communicator leaves rank 0 of 1."""
    import ctypes as C
    import pbrt_v3_b200 as pb
    L = pb.lib()
    r, w = C.c_int(-1), C.c_int(-1)
    assert L.pb2_dist_info(C.byref(r), C.byref(w)) == 0 and (r.value, w.value) == (0, 1)
    import torch
    if not torch.cuda.is_available():
        assert L.pb2_dist_init(0, 2, None) == pb.PB2_ERR_NO_DEVICE
