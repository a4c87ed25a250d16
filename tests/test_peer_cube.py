"""Cross-process correctness of the default multi-GPU data plane: two ranks (one process per GPU)
evaluate their pixel shards of ONE seeded cube and move them into rank 0's cube through
`sharding.PeerCube` (CUDA IPC + strided peer copies, pkm_ybar_density_into_cube).  Rank 0 checks
the assembled cube against the single-GPU evaluation of the whole cube (bit-identical) and against
the oracle.  Needs two GPUs: skipped on a one-GPU box, run with `gpurun --gpus 2`."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT

WORKER = r"""
import os, sys
sys.path.insert(0, os.environ["PKM_ROOT"])
import numpy as np, torch, torch.distributed as dist
import popkinmocks_b200 as pkm
from popkinmocks_b200 import engine, sharding, _lib
from oracle import restatement as R

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)

ssps = pkm.milesSSPs(age_rebin=6, z_rebin=2)
nv, npix_total = 21, 333                               # uneven shards: 166 + 167 pixels, several tiles each
cube = pkm.ifu_cube.IFUCube(ssps=ssps, nx1=npix_total, nx2=1, nv=nv, vrng=(-1000.0, 1000.0))
nz, nt = ssps.par_dims
plan = engine.Plan(cube, device=local)
plan.set_workspace_limit(1 << 20)                      # force several pixel tiles per rank
synth = dict(nt=int(nt), nv=nv, total_pix=npix_total, nz=int(nz), seed=77, scale=1e-3)
r0, r1 = sharding.my_rows(npix_total, world, rank)
logp = engine.synth_log_density(pix0=r0, npix=r1 - r0, device=local, **synth).view(nt, nv, r1 - r0, 1, nz)

pc = sharding.PeerCube(ssps.n_fft, npix_total, 1, dev, dist)
flag = torch.zeros(1, device=dev)
for rep in range(2):                                   # twice: staging slots and events are reused
    if rank == 0:
        pc.cube.fill_(-7.0)
    torch.cuda.synchronize(); dist.barrier()
    pc.evaluate_rows(plan, logp, 0, r1 - r0, cube_row0=r0)
    dist.all_reduce(flag)
    torch.cuda.synchronize()
    if rank == 0:
        got = pc.cube.clone()
        whole = engine.synth_log_density(pix0=0, npix=npix_total, device=local, **synth).view(nt, nv, npix_total, 1, nz)
        want = plan.ybar_density(whole, is_log=True)
        assert torch.equal(got, want), f"rep {rep}: peer-assembled cube differs from the single-GPU cube"
        px = [0, r1 - 1, r1, npix_total - 1]
        p = np.exp(engine.synth_log_density_host(synth["nt"], nv, npix_total, px, synth["nz"], 77, 1e-3))[:, :, None]
        ref = R.ybar_density_literal(p, cube.v_edg, ssps.FXw, ssps.par_dims, ssps.n_fft, ssps.dv, ssps.delta_t,
                                     ssps.delta_z, cube.dx, cube.dy)[:, 0, :]
        err = float(np.max(np.abs(got[:, px, 0].cpu().numpy() - ref)) / np.max(np.abs(ref)))
        assert err < 1e-10, err
        print(f"PEER_CUBE_OK rep={rep} err={err:.2e}")
dist.barrier()
torch.cuda.synchronize()
pc.close()
dist.barrier()
dist.destroy_process_group()
"""


@pytest.mark.gpu
def test_two_rank_peer_cube_matches_single_gpu(gpu_ok, tmp_path):
    from popkinmocks_b200 import _lib
    if _lib.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    script = tmp_path / "peer_worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, PKM_ROOT=ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, (out.stdout[-2000:], out.stderr[-3000:])
    assert out.stdout.count("PEER_CUBE_OK") == 2
