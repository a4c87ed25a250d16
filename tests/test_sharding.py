"""Multi-GPU host logic on CPU: row partition and the gather of spaxel-major tiles, exercised
with two gloo ranks (the GPU box runs the same code over NCCL)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

from popkinmocks_b200 import sharding


def test_row_partition_is_contiguous_and_balanced():
    for nx1 in (0, 1, 7, 201, 202):
        for world in (1, 2, 3, 4, 8):
            b = sharding.row_partition(nx1, world)
            sizes = np.diff(b)
            assert b[0] == 0 and b[-1] == nx1 and np.all(sizes >= 0)
            assert sizes.max() - sizes.min() <= 1
    assert sharding.my_rows(201, 8, 7) == (175, 201)
    with pytest.raises(ValueError):
        sharding.row_partition(4, 0)


def test_pixel_major_to_cube_layout():
    nx1, nx2, n = 3, 2, 5
    full = np.arange(nx1 * nx2 * n, dtype=np.float64).reshape(nx1 * nx2, n)
    cube = sharding.pixel_major_to_cube(full, nx1, nx2)
    assert cube.shape == (n, nx1, nx2)
    assert cube[4, 2, 1] == full[2 * nx2 + 1, 4]


WORKER = r"""
import os, sys
sys.path.insert(0, os.environ["PKM_ROOT"])
import numpy as np, torch, torch.distributed as dist
from popkinmocks_b200 import sharding
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
nx1, nx2, n = 5, 3, 7                      # 5 rows over 2 ranks: uneven blocks (2, 3)
r0, r1 = sharding.my_rows(nx1, world, rank)
# stand-in for the per-rank evaluate_ybar: value encodes (pixel, omega)
pix = torch.arange(r0 * nx2, r1 * nx2, dtype=torch.float64)[:, None]
tile = pix * 100.0 + torch.arange(n, dtype=torch.float64)[None, :]
full = sharding.gather_pixel_major(tile, nx1, nx2, dist=dist, dst=0)
want = (np.arange(nx1 * nx2)[None, :] * 100.0 + np.arange(n)[:, None]).reshape(n, nx1, nx2)
if rank == 0:
    cube = sharding.pixel_major_to_cube(full, nx1, nx2).numpy()
    assert np.array_equal(cube, want), "gathered cube is wrong"
    print("GATHER_OK")
else:
    assert full is None
# the overlapped tile gather used by bench.py (here without CUDA streams)
tg = sharding.TileGather(nx1, nx2, n, "cpu", dist=dist, ntile=2)
for t, (a, b, out) in enumerate(tg.my_tiles()):
    g0 = (r0 + a) * nx2
    out.copy_(torch.arange(g0, g0 + (b - a) * nx2, dtype=torch.float64)[:, None] * 100.0
              + torch.arange(n, dtype=torch.float64)[None, :])
    tg.post(t)
full = tg.finish()
if rank == 0:
    assert np.array_equal(sharding.pixel_major_to_cube(full, nx1, nx2).numpy(), want)
    print("TILE_GATHER_OK")
dist.barrier()
dist.destroy_process_group()
"""


def test_two_rank_gloo_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, PKM_ROOT=ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         env=env, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "GATHER_OK" in out.stdout and "TILE_GATHER_OK" in out.stdout
