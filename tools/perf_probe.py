"""Quick device-side timing of the two ybar paths (development aid, not the bench)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import popkinmocks_b200 as pkm
from popkinmocks_b200 import engine

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 32
nv = int(sys.argv[2]) if len(sys.argv) > 2 else 101
lmd = (float(sys.argv[3]), float(sys.argv[4])) if len(sys.argv) > 4 else (4700.0, 6500.0)
ssps = pkm.milesSSPs(lmd_min=lmd[0], lmd_max=lmd[1])
cube = pkm.ifu_cube.IFUCube(ssps=ssps, nx1=nx, nx2=nx, nv=nv, vrng=(-1000.0, 1000.0))
plan = engine.get_plan(cube)
print("n_fft", plan.n_fft, "fp64 peak TF", engine.fp64_fma_peak())
nt, nz = plan.nt, plan.nz
gen = torch.Generator(device="cuda").manual_seed(1)
p = torch.rand((nt, nv, nx, nx, nz), dtype=torch.float64, device="cuda", generator=gen)
P = torch.rand((nt, nx, nx, nz), dtype=torch.float64, device="cuda", generator=gen)
mu = 200 * (torch.rand((nt, nx, nx), dtype=torch.float64, device="cuda", generator=gen) - 0.5)
sg = 50 + 100 * torch.rand((nt, nx, nx), dtype=torch.float64, device="cuda", generator=gen)


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


vox = plan.n_fft * nx * nx
plan.set_profiling(True)
ms = timeit(lambda: plan.ybar_density(p, is_log=False))
print(f"density  {nx}x{nx} nv={nv}: {ms:.2f} ms  {vox/ms*1e3:.3e} voxel/s", {k: round(v[0]/6, 3) for k, v in plan.kernel_times().items()})
plan.set_profiling(True)
ms = timeit(lambda: plan.ybar_density(p, is_log=True))
print(f"density(log) : {ms:.2f} ms  {vox/ms*1e3:.3e} voxel/s", {k: round(v[0]/6, 3) for k, v in plan.kernel_times().items()})
plan.set_profiling(True)
ms = timeit(lambda: plan.ybar_gaussian(P, mu, sg))
print(f"gaussian {nx}x{nx}: {ms:.2f} ms  {vox/ms*1e3:.3e} voxel/s", {k: round(v[0]/6, 3) for k, v in plan.kernel_times().items()})

plan32 = engine.get_plan(cube, precision="float32")
plan32.set_profiling(True)
ms = timeit(lambda: plan32.ybar_density(p, is_log=False))
y64 = plan.ybar_density(p, is_log=False); y32 = plan32.ybar_density(p, is_log=False)
print(f"density fp32 {nx}x{nx}: {ms:.2f} ms  {vox/ms*1e3:.3e} voxel/s", {k: round(v[0]/6, 3) for k, v in plan32.kernel_times().items()},
      "rel diff vs fp64", float((y32 - y64).abs().max() / y64.abs().max()))
