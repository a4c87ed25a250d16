"""Step time of the 201x201 cube against the workspace limit (= number of pixel tiles) and tile lanes."""
import os, sys, torch
sys.path.insert(0, ".")
import popkinmocks_b200 as pkm
from popkinmocks_b200 import engine
ssps = pkm.milesSSPs()
n1 = n2 = 201
cube = pkm.ifu_cube.IFUCube(ssps=ssps, nx1=n1, nx2=n2, nv=101, vrng=(-1000.0, 1000.0))
plan = engine.Plan(cube)
nt, nz = plan.nt, plan.nz
logp = engine.synth_log_density(nt, 101, n1 * n2, 0, n1 * n2, nz, 1, 1e-3).view(nt, 101, n1, n2, nz)
out = torch.empty((plan.n_fft, n1, n2), dtype=torch.float64, device="cuda")
ref = None
for lanes in (1, 2):
    os.environ["PKM_TILE_LANES"] = str(lanes)
    for gb in (1, 2, 4, 8, 24):
        plan.set_workspace_limit(gb << 30)
        for _ in range(2):
            plan.ybar_density(logp, is_log=True, out=out)
        torch.cuda.synchronize()
        if ref is None:
            ref = out.clone()
        same = bool(torch.equal(ref, out))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            plan.ybar_density(logp, is_log=True, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        plan.set_profiling(True)
        for _ in range(3):
            plan.ybar_density(logp, is_log=True, out=out)
        kt = plan.kernel_times()
        plan.set_profiling(False)
        print("lanes", lanes, gb, "GB", round(ms, 3), "ms bit-identical", same,
              {k: (round(v[0] / 3, 3), v[1] // 3) for k, v in kt.items()}, flush=True)
