"""Development aid: device time of the Gaussian path (path A) on an nx x nx cube (default 201)."""
import sys
import torch
sys.path.insert(0, ".")
import popkinmocks_b200 as pkm
from popkinmocks_b200 import engine

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 201
ssps = pkm.milesSSPs()
cube = pkm.ifu_cube.IFUCube(ssps=ssps, nx1=nx, nx2=nx, nv=101, vrng=(-1000.0, 1000.0))
plan = engine.Plan(cube)
nt, nz = plan.nt, plan.nz
gen = torch.Generator(device="cuda").manual_seed(5)
P = torch.rand((nt, nx, nx, nz), dtype=torch.float64, device="cuda", generator=gen)
P /= P.sum()
mu = 400.0 * (torch.rand((nt, nx, nx), dtype=torch.float64, device="cuda", generator=gen) - 0.5)
sg = 20.0 + 200.0 * torch.rand((nt, nx, nx), dtype=torch.float64, device="cuda", generator=gen)
out = torch.empty((plan.n_fft, nx, nx), dtype=torch.float64, device="cuda")
for _ in range(2):
    plan.ybar_gaussian(P, mu, sg, out=out)
torch.cuda.synchronize()
plan.set_profiling(True)
for _ in range(4):
    plan.ybar_gaussian(P, mu, sg, out=out)
kt = plan.kernel_times()
print({k: round(v[0] / 4, 3) for k, v in kt.items() if v[1]}, flush=True)
