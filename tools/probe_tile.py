import sys, torch, numpy as np
sys.path.insert(0, ".")
import popkinmocks_b200 as pkm
from popkinmocks_b200 import engine
ssps = pkm.milesSSPs()
for (n1, n2) in (((112, 64),) if len(sys.argv) > 1 else ((112, 64), (224, 64), (201, 201))):
    cube = pkm.ifu_cube.IFUCube(ssps=ssps, nx1=n1, nx2=n2, nv=101, vrng=(-1000.0, 1000.0))
    plan = engine.Plan(cube)
    nt, nz = plan.nt, plan.nz
    logp = engine.synth_log_density(nt, 101, n1 * n2, 0, n1 * n2, nz, 1, 1e-3).view(nt, 101, n1, n2, nz)
    out = torch.empty((plan.n_fft, n1, n2), dtype=torch.float64, device="cuda")
    for _ in range(2):
        plan.ybar_density(logp, is_log=True, out=out)
    torch.cuda.synchronize()
    plan.set_profiling(True)
    for _ in range(4):
        plan.ybar_density(logp, is_log=True, out=out)
    kt = plan.kernel_times()
    print(n1, n2, {k: (round(v[0] / 4, 3), v[1] // 4) for k, v in kt.items()}, flush=True)
    plan.close()
    del logp, out
    torch.cuda.empty_cache()
