"""Coefficient-kernel time on a 224x64 cube (argv[1] = 0 feeds a linear density)."""
import sys, torch
sys.path.insert(0, ".")
import popkinmocks_b200 as pkm
from popkinmocks_b200 import engine
is_log = (sys.argv[1] != "0") if len(sys.argv) > 1 else True
ssps = pkm.milesSSPs()
n1, n2 = 224, 64
cube = pkm.ifu_cube.IFUCube(ssps=ssps, nx1=n1, nx2=n2, nv=101, vrng=(-1000.0, 1000.0))
plan = engine.Plan(cube)
nt, nz = plan.nt, plan.nz
logp = engine.synth_log_density(nt, 101, n1 * n2, 0, n1 * n2, nz, 1, 1e-3).view(nt, 101, n1, n2, nz)
if not is_log:
    logp = logp.exp_()
out = torch.empty((plan.n_fft, n1, n2), dtype=torch.float64, device="cuda")
for _ in range(2):
    plan.ybar_density(logp, is_log=is_log, out=out)
torch.cuda.synchronize()
plan.set_profiling(True)
for _ in range(4):
    plan.ybar_density(logp, is_log=is_log, out=out)
kt = plan.kernel_times()
print("is_log", is_log, {k: round(v[0] / 4, 3) for k, v in kt.items() if v[1]}, flush=True)
