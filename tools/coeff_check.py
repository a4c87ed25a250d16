"""Development aid: the octet coefficient kernel against the staged one (must be bit-identical)
and the device-side time of both, on a white-noise density."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import popkinmocks_b200 as pkm
from popkinmocks_b200 import engine

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 64
for ssp_kw, nv in (({}, 101), ({}, 41), (dict(age_rebin=6, z_rebin=2), 21), (dict(age_rebin=4, z_rebin=3), 41)):
    ssps = pkm.milesSSPs(**ssp_kw)
    cube = pkm.ifu_cube.IFUCube(ssps=ssps, nx1=nx, nx2=nx + 3, nv=nv, vrng=(-1000.0, 1000.0))
    gen = torch.Generator(device="cuda").manual_seed(1)
    shape = cube.get_distribution_shape("tvxz")
    logp = torch.log(torch.rand(shape, dtype=torch.float64, device="cuda", generator=gen))
    res = {}
    for kern in ("1", "0"):
        os.environ["PKM_COEFF_KERNEL"] = kern
        plan = engine.Plan(cube)
        plan.set_profiling(True)
        y = plan.ybar_density(logp, is_log=True)
        y = plan.ybar_density(logp, is_log=True)
        torch.cuda.synchronize()
        res[kern] = (y.clone(), {k: round(v[0] / max(v[1], 1), 3) for k, v in plan.kernel_times().items()})
        plan.close()
    same = bool(torch.equal(res["1"][0], res["0"][0]))
    err = float((res["1"][0] - res["0"][0]).abs().max() / res["0"][0].abs().max())
    print(f"nz={shape[-1]} nt={shape[0]} nv={nv} {nx}x{nx + 3}: identical={same} relerr={err:.2e}\n   oct    {res['1'][1]}\n   staged {res['0'][1]}", flush=True)
