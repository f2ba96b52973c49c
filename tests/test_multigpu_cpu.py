"""Host-side logic of the multi-GPU path, exercised with world_size 2 over gloo on the CPU."""
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
from pbrt_v3_b200 import multigpu
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
sb = (0, 0, 100, 70)                       # ragged: neither side is a multiple of 16
owner = multigpu.tile_owner_map(sb, 2)
rng = np.random.RandomState(5)
full = rng.rand(70, 100, 4).astype(np.float32)   # what a single-GPU render would produce
mine = np.where((owner == rank)[..., None], full, 0).astype(np.float32)
film = torch.from_numpy(mine.copy())
multigpu.reduce_film(film, dst=0)
if rank == 0:
    assert np.array_equal(film.numpy(), full), "sum of the per-rank tile films must be the single-GPU film"
    assert multigpu.owned_tile_count(sb, 0, 2) + multigpu.owned_tile_count(sb, 1, 2) == 7 * 5
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_film_reduce_world2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = str(free_port())
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o


def test_tile_owner_map_matches_kernel_partition():
    """decodeWork() in pb2_cuda.cu deals tile t to rank t % tile_count; the host map must agree and partition the film."""
    from pbrt_v3_b200 import multigpu
    sb = (-1, -1, 49, 34)
    for world in (1, 2, 3, 8):
        owner = multigpu.tile_owner_map(sb, world)
        assert owner.shape == (35, 50)
        assert owner.min() == 0 and owner.max() == min(world, 4 * 3) - 1 if world <= 12 else True
        counts = [multigpu.owned_tile_count(sb, r, world) for r in range(world)]
        assert sum(counts) == 4 * 3
        # a 16x16 tile never straddles two owners
        assert (owner[:16, :16] == owner[0, 0]).all()
