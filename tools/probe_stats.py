"""Development aid: device time of the log-marginal reductions on the 201x201 headline array."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import popkinmocks_b200 as pkm
from popkinmocks_b200 import engine

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 201
ssps = pkm.milesSSPs()
cube = pkm.ifu_cube.IFUCube(ssps=ssps, nx1=nx, nx2=nx, nv=101, vrng=(-1000.0, 1000.0))
plan = engine.get_plan(cube)
nt, nz = plan.nt, plan.nz
logp = engine.synth_log_density(nt, 101, nx * nx, 0, nx * nx, nz, 1, 1e-3).view(nt, 101, nx, nx, nz)
log_dv = np.log(np.diff(cube.v_edg))
gb = logp.numel() * 8 / 1e9


def timeit(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


KEEPS = sys.argv[2].split(",") if len(sys.argv) > 2 else ["vx", "tz", "x", "txz", "v", "t", "z", "tvx", "tx", ""]
for keep in KEEPS:
    for lw in (None, ssps.light_weights):
        ms = timeit(lambda: plan.reduce_logp(logp, log_dv, keep, False, light_weights=lw))
        print(f"keep={keep or '-':5s} lw={int(lw is not None)}  {ms:7.3f} ms  {gb / ms:7.2f} TB/s", flush=True)
comp = pkm.components.Component(cube=cube, log_p_tvxz=logp)


def skew():
    comp.invalidate_cache()
    with np.errstate(invalid="ignore", divide="ignore"):
        comp.get_skewness("v_x")


torch.cuda.synchronize()
import time
skew()
t0 = time.perf_counter()
for _ in range(3):
    skew()
torch.cuda.synchronize()
print(f"get_skewness('v_x') wall: {(time.perf_counter() - t0) / 3 * 1e3:.2f} ms")
