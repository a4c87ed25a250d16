"""BASELINE.json configurations C2, C3 and C5 as runnable workloads.

Each `run_*` builds the named model with the drop-in classes, times one "step" of the hot path with
CUDA events on a device-resident input, measures the same step through the user-facing class API
with host arrays (`e2e`), and verifies the results against the oracle (`oracle/restatement.py`,
pinned to the reference by tests/test_oracle.py) on spot pixels.  `bench.py --config C2|C3|C5`
formats the returned dictionaries as the driver's JSON line; `tests/test_configs_gpu.py` runs the
same functions at the named sizes and asserts the verification.

Reference: docs/source/user/constructing_models.md:409-414 (the mixture), tests/conftest.py:24-104
(component parameters), tests/test_all.py:49-146 (what the reference checks on them).
"""
import time

import numpy as np

import popkinmocks_b200 as pkm
from popkinmocks_b200 import engine

TOL = 1e-10


# ----------------------------------------------------------------------------------------------
# models
# ----------------------------------------------------------------------------------------------
def make_cube(nx, nv=101, lmd=(4700.0, 6500.0)):
    ssps = pkm.milesSSPs(lmd_min=lmd[0], lmd_max=lmd[1])
    return pkm.ifu_cube.IFUCube(ssps=ssps, nx1=nx, nx2=nx, nv=nv, x1rng=(-1.0, 1.0), x2rng=(-1.0, 1.0),
                                vrng=(-1000.0, 1000.0))


def make_disk(cube):
    """The reference fixture's first disk (tests/conftest.py:24-49)."""
    c = pkm.components.GrowingDisk(cube=cube, rotation=0.0, center=(0.2, 0))
    c.set_p_t(lmd=2.0, phi=0.8)
    c.set_p_x_t(q_lims=(0.05, 0.5), rc_lims=(1.0, 1.0), alpha_lims=(1.2, 0.8))
    c.set_t_dep(q=0.2, alpha=3.0, t_dep_in=0.5, t_dep_out=5.0)
    c.set_mu_v(q_lims=(0.5, 0.1), rmax_lims=(0.5, 1.1), vmax_lims=(250.0, 100.0))
    c.set_sig_v(q_lims=(0.25, 1.0), alpha_lims=(3.0, 2.5), sig_v_in_lims=(70.0, 50.0),
                sig_v_out_lims=(80.0, 60.0))
    return c


def make_stream(cube):
    """The reference fixture's stream (tests/conftest.py:78-93) with the default 75 samples."""
    s = pkm.components.Stream(cube=cube, rotation=0.0, center=(0.0, 0))
    s.set_p_t(lmd=15.0, phi=0.3)
    s.set_p_x_t(theta_lims=[-np.pi / 2.0, 0.75 * np.pi], mu_r_lims=[0.2, 0.8], nsmp=75, sig=0.1)
    s.set_t_dep(t_dep=5.0)
    s.set_mu_v(mu_v_lims=[-80, 100])
    s.set_sig_v(sig_v=110.0)
    return s


def make_mixture(cube):
    """Mixture([disk, stream], [0.9, 0.1]) (docs/source/user/constructing_models.md:409-414)."""
    return pkm.components.Mixture(cube=cube, component_list=[make_disk(cube), make_stream(cube)],
                                  weights=[0.9, 0.1])


def spot_pixels(nx, n=8):
    """Deterministic spot pixels (x1, x2): the corners, the centre and a few in between."""
    c = nx // 2
    pts = [(0, 0), (nx - 1, nx - 1), (0, nx - 1), (c, c), (c, min(nx - 1, c + nx // 5)),
           (nx // 4, (3 * nx) // 4), (nx - 2, 1), ((2 * nx) // 3, nx // 3)]
    return pts[:n]


# ----------------------------------------------------------------------------------------------
# timing helpers
# ----------------------------------------------------------------------------------------------
def _timed(fn, steps, warmup):
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def _wall(fn, steps):
    import torch
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / steps


def _grid(cube):
    s = cube.ssps
    return dict(delta_t=s.delta_t, dv_bins=np.diff(cube.v_edg), dx=cube.dx, dy=cube.dy, delta_z=s.delta_z,
                light_weights=s.light_weights)


def _rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / np.max(np.abs(b)))


# ----------------------------------------------------------------------------------------------
# C2: GrowingDisk, full grid, 51x51, nv = 101 - the analytic path A, with path B beside it
# ----------------------------------------------------------------------------------------------
def run_c2(nx=51, nv=101, steps=10, warmup=3, device=0, e2e_steps=3):
    import torch
    from oracle import restatement as R
    cube = make_cube(nx, nv)
    s = cube.ssps
    nz, nt = s.par_dims
    t0 = time.perf_counter()
    disk = make_disk(cube)
    setup_s = time.perf_counter() - t0
    plan = engine.get_plan(cube, device=device)
    dev = f"cuda:{device}"
    P_txz = disk.get_p("txz", density=False)
    dP = torch.from_numpy(np.ascontiguousarray(P_txz)).to(dev)
    dmu = torch.from_numpy(np.ascontiguousarray(disk.mu_v)).to(dev)
    dsg = torch.from_numpy(np.ascontiguousarray(disk.sig_v)).to(dev)
    out = torch.empty((plan.n_fft, nx, nx), dtype=torch.float64, device=dev)
    voxels = plan.n_fft * nx * nx

    plan.set_profiling(True)
    l0 = plan.launches
    ms = _timed(lambda: plan.ybar_gaussian(dP, dmu, dsg, out=out), steps, warmup)
    kt = plan.kernel_times()
    launches = plan.launches - l0
    plan.set_profiling(False)
    y_a = out.cpu().numpy()

    # path B on the pixelated twin (5-D array produced and kept on the device)
    twin = disk.pixelate(device=device)
    out_b = torch.empty_like(out)
    ms_b = _timed(lambda: plan.ybar_density(twin.log_p_tvxz, is_log=True, out=out_b), max(2, steps // 3), 1)
    y_b = out_b.cpu().numpy()

    # ---- verification
    px = spot_pixels(nx)
    i1 = [p[0] for p in px]
    i2 = [p[1] for p in px]
    ref_a = R.ybar_gaussian(P_txz[:, i1, i2][:, None], np.asarray(disk.mu_v)[:, i1, i2][:, None],
                            np.asarray(disk.sig_v)[:, i1, i2][:, None], s.FXw, s.par_dims, s.n_fft, s.dv)[:, 0]
    err_a = _rel(y_a[:, i1, i2], ref_a)
    logp_px = twin.log_p_tvxz[:, :, i1, i2].cpu().numpy()[:, :, None]            # [nt, nv, 1, npx, nz]
    with np.errstate(over="ignore"):
        ref_b = R.ybar_density_literal(np.exp(logp_px), cube.v_edg, s.FXw, s.par_dims, s.n_fft, s.dv, s.delta_t,
                                       s.delta_z, cube.dx, cube.dy)[:, 0]
    err_b = _rel(y_b[:, i1, i2], ref_b)
    med_ab = float(np.median(np.abs(y_a / y_b - 1.0)))                            # tests/test_all.py:103-115

    # ---- the class API with host arrays
    # model parameters -> host cube: construction, device maps, kernels, D2H (nothing is cached
    # between steps: every step builds a fresh GrowingDisk)
    holder = {}

    def api_step():
        holder["disk"] = make_disk(cube)
        holder["disk"].evaluate_ybar(device=device)
    e2e_ms = _wall(api_step, e2e_steps)
    err_api = _rel(holder["disk"].ybar, y_a)
    n_rows = int(holder["disk"]._device_maps(device)["rows"].size)
    h2d_bytes = 8 * (10 * nt + cube.nx + cube.ny + 1000 + 2 * nt + n_rows * nt * nz + nz)
    # executed FP64 work per voxel: per (k, t, x) 2*nz FMAs of the metallicity contraction, a 6-operation
    # recurrence for exp(-i mu w - sig^2 w^2/2) and a complex multiply-add (SURVEY.md 8d counts 68 nt
    # flops per frequency for evaluating the exponential directly: 3.1 kFLOP/voxel algorithmic)
    flops_vox = plan.nfo * nt * (4.0 * nz + 14.0) / plan.n_fft
    fp64_peak = engine.fp64_fma_peak(device)
    k_ms = kt["gaussian"][0] / max(kt["gaussian"][1], 1)
    return {
        "workload": f"C2 GrowingDisk {nx}x{nx} px, full MILES_BASTI grid nt={nt} nz={nz}, nv={nv}, "
                    f"n_fft={plan.n_fft} (path A: ParametricComponent.evaluate_ybar, parametric.py:335-407)",
        "config": {"nx1": nx, "nx2": nx, "nv": nv, "nt": int(nt), "nz": int(nz), "n_fft": int(plan.n_fft),
                   "l2": "an L2-sized write precedes nothing here: the 13 MB operands and the 102 MB cube "
                         "partly stay in L2 between steps, as they would for a user looping over models"},
        "voxels": voxels, "ms_per_step": ms, "value": voxels / (ms * 1e-3), "launches": int(launches),
        "e2e": {"value": voxels / (e2e_ms * 1e-3), "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(h2d_bytes),
                "d2h_bytes_per_step": int(holder["disk"].ybar.nbytes), "steps": e2e_steps,
                "api": "GrowingDisk(cube); set_p_t/set_p_x_t/set_t_dep/set_mu_v/set_sig_v; evaluate_ybar() - "
                       "model parameters in, host cube out; the O(pixels) maps are generated in HBM, so the "
                       "H2D traffic is the O(nt) parameter vectors and the enrichment table rows"},
        "roofline": {"bound": "fp64", "kernel": "k_gaussian", "achieved": flops_vox * voxels / (k_ms * 1e-3) / 1e12,
                     "peak": fp64_peak, "unit": "TFLOP/s",
                     "frac": flops_vox * voxels / (k_ms * 1e-3) / 1e12 / fp64_peak, "traffic": None,
                     "flops_per_voxel": flops_vox, "peak_source": "pkm_fp64_fma_peak measured in this run",
                     "note": "executed FP64 work (1.6 kFLOP/voxel; SURVEY.md 8d's algorithmic figure is 3.1); "
                             "the metallicity contraction runs on the FP64 tensor cores (DMMA)"},
        "kernels": {k: {"ms_per_step": m / steps, "launches_per_step": c / steps} for k, (m, c) in kt.items() if c},
        "other_paths": {"density_path_B_on_pixelated_twin": {"ms": ms_b, "voxel_per_s": voxels / (ms_b * 1e-3)},
                        "host_model_setup_s": setup_s},
        "verify": {"ok": bool(err_a < TOL and err_b < TOL and med_ab < 5e-4 and err_api < 1e-12),
                   "max_rel_err": max(err_a, err_b), "tol": TOL, "pixels": len(px),
                   "path_A_vs_oracle": err_a, "path_B_vs_oracle": err_b, "median_rel_diff_A_vs_B": med_ab,
                   "class_api_vs_device": err_api,
                   "oracle": "oracle/restatement.py ybar_gaussian (parametric.py:355-373) and "
                             "ybar_density_literal (base.py:836-859); A-vs-B as tests/test_all.py:103-115"},
    }


# ----------------------------------------------------------------------------------------------
# C3: Mixture(GrowingDisk + Stream) 101x101 with get_p marginals and velocity moments
# ----------------------------------------------------------------------------------------------
C3_DISTS = ("x", "v_x", "tz_x", "t", "z")
C3_MOMENTS = ("get_mean", "get_variance", "get_skewness", "get_kurtosis")


def _host_sums(logp_dev, cube, px):
    """Reference sums over the whole array in t-slabs (base.py:1017-1041 then plain sums):
    P = exp(log p + log dV), mass and light weighted; returns the marginals the checks need."""
    s = cube.ssps
    nz, nt = s.par_dims
    dV_vz = (np.diff(cube.v_edg)[:, None, None, None] * cube.dx * cube.dy * s.delta_z[None, None, None, :])
    lw = s.light_weights                                                     # [nt, nz]
    P_t = np.zeros(nt)
    P_z = np.zeros(nz)
    L_t = np.zeros(nt)
    L_z = np.zeros(nz)
    P_x = np.zeros((cube.nx, cube.ny))
    L_x = np.zeros((cube.nx, cube.ny))
    i1 = [p[0] for p in px]
    i2 = [p[1] for p in px]
    P_tvz_px = np.zeros((nt, cube.nv, len(px), nz))
    for t in range(nt):
        P = np.exp(logp_dev[t].cpu().numpy()) * (dV_vz * s.delta_t[t])      # [nv, x1, x2, nz]
        P_t[t] = P.sum()
        pz = P.sum((0, 1, 2))
        P_z += pz
        L_t[t] = (pz * lw[t]).sum()
        L_z += pz * lw[t]
        P_x += P.sum((0, 3))
        L_x += (P * lw[t]).sum((0, 3))
        P_tvz_px[t] = P[:, i1, i2, :]
    return dict(P_t=P_t, P_z=P_z, L_t=L_t, L_z=L_z, P_x=P_x, L_x=L_x, P_tvz_px=P_tvz_px, L_norm=L_t.sum())


def verify_c3(comp, results, px):
    """`results` = {(dist, light_weighted): array, ("moment", name): array} from the class API."""
    cube = comp.cube
    s = cube.ssps
    ref = _host_sums(comp.log_p_tvxz, cube, px)
    i1 = [p[0] for p in px]
    i2 = [p[1] for p in px]
    lw = s.light_weights
    v = cube.get_variable_values("v")
    errs = {}
    for light in (False, True):
        w_tz = lw if light else np.ones_like(lw)
        norm = ref["L_norm"] if light else 1.0
        Pj = ref["P_tvz_px"] * w_tz[:, None, None, :] / norm                 # P(t,v,x,z) at the spot pixels
        Px_px = Pj.sum((0, 1, 3))
        want = {
            "t": (ref["L_t"] / norm if light else ref["P_t"]) / s.delta_t,
            "z": (ref["L_z"] / norm if light else ref["P_z"]) / s.delta_z,
            "x": ((ref["L_x"] / norm if light else ref["P_x"]) / (cube.dx * cube.dy))[i1, i2],
            "v_x": (Pj.sum((0, 3)) / Px_px / np.diff(cube.v_edg)[:, None]),
            "tz_x": np.moveaxis(Pj.sum(1), 1, -1) / Px_px / (s.delta_t[:, None, None] * s.delta_z[None, :, None]),
        }
        for dist in C3_DISTS:
            got = results[(dist, light)]
            if dist == "x":
                got = got[i1, i2]
            elif dist in ("v_x", "tz_x"):
                got = got[..., i1, i2]
            errs[f"{dist}|lw={int(light)}"] = _rel(got, want[dist])
        if not light:
            Pv = Pj.sum((0, 3)) / Px_px                                       # P(v | x) at the spot pixels
            mean = (v[:, None] * Pv).sum(0)
            cen = {j: (((v[:, None] - mean) ** j) * Pv).sum(0) for j in (2, 3, 4)}
            wantm = {"get_mean": mean, "get_variance": cen[2], "get_skewness": cen[3] / cen[2] ** 1.5,
                     "get_kurtosis": cen[4] / cen[2] ** 2}
            for name in C3_MOMENTS:
                got = results[("moment", name)][i1, i2]
                scale = max(1.0, float(np.max(np.abs(wantm[name]))))
                errs[name] = float(np.max(np.abs(got - wantm[name])) / scale)
    return errs


def c3_step(comp, plan, out, results=None):
    """One pass of the C3 workload on a device-resident pixelated mixture."""
    comp.invalidate_cache()                      # every step recomputes the marginals from the 5-D array
    plan.ybar_density(comp.log_p_tvxz, is_log=True, out=out)
    for light in (False, True):
        for dist in C3_DISTS:
            r = comp.get_p(dist, light_weighted=light)
            if results is not None:
                results[(dist, light)] = r
    for name in C3_MOMENTS:
        with np.errstate(invalid="ignore", divide="ignore"):
            r = getattr(comp, name)("v_x")
        if results is not None:
            results[("moment", name)] = r


def run_c3(nx=101, nv=101, steps=5, warmup=2, device=0, e2e_steps=2):
    import torch
    cube = make_cube(nx, nv)
    s = cube.ssps
    nz, nt = s.par_dims
    t0 = time.perf_counter()
    mix = make_mixture(cube)
    setup_s = time.perf_counter() - t0
    plan = engine.get_plan(cube, device=device)
    dev = f"cuda:{device}"
    t0 = time.perf_counter()
    comp = mix.pixelate(device=device)           # Component(cube, mixture.get_log_p('tvxz')) built in HBM
    torch.cuda.synchronize()
    pixelate_s = time.perf_counter() - t0
    out = torch.empty((plan.n_fft, nx, nx), dtype=torch.float64, device=dev)
    voxels = plan.n_fft * nx * nx
    results = {}
    c3_step(comp, plan, out, results)
    plan.set_profiling(True)
    l0 = plan.launches
    ms = _timed(lambda: c3_step(comp, plan, out), steps, warmup)
    kt = plan.kernel_times()
    launches = plan.launches - l0
    plan.set_profiling(False)
    # shares of the step: the cube, one marginal pass (vx), the derived calls
    ms_cube = _timed(lambda: plan.ybar_density(comp.log_p_tvxz, is_log=True, out=out), 3, 1)
    log_dv = np.log(np.diff(cube.v_edg))
    ms_vx = _timed(lambda: plan.reduce_logp(comp.log_p_tvxz, log_dv, "vx", False), 3, 1)
    ms_txz = _timed(lambda: plan.reduce_logp(comp.log_p_tvxz, log_dv, "txz", False), 3, 1)

    def skew():
        comp.invalidate_cache()
        with np.errstate(invalid="ignore", divide="ignore"):
            comp.get_skewness("v_x")
    ms_skew = _wall(skew, 3)
    nbytes = comp.log_p_tvxz.numel() * 8
    peaks = _peaks()
    px = spot_pixels(nx)
    errs = verify_c3(comp, results, px)
    y_dev = out.cpu().numpy()
    from oracle import restatement as R
    i1 = [p[0] for p in px]
    i2 = [p[1] for p in px]
    logp_px = comp.log_p_tvxz[:, :, i1, i2].cpu().numpy()[:, :, None]
    ref_y = R.ybar_density_literal(np.exp(logp_px), cube.v_edg, s.FXw, s.par_dims, s.n_fft, s.dv, s.delta_t,
                                   s.delta_z, cube.dx, cube.dy)[:, 0]
    errs["ybar"] = _rel(y_dev[:, i1, i2], ref_y)
    # mixture cube = weighted sum of the children's analytic cubes (mixture.py:35-48) vs the pixelated cube
    for c in mix.component_list:
        c.evaluate_ybar(device=device)
    mix.evaluate_ybar(device=device)
    med = float(np.median(np.abs(mix.ybar / y_dev - 1.0)))

    # ---- class API from a HOST array (the reference's data flow): 5.2 GB up, cube + marginals down
    host = comp.log_p_tvxz.cpu().numpy()
    hcomp = pkm.components.Component(cube=cube, log_p_tvxz=host)

    def api_step():
        hcomp.invalidate_cache()
        hcomp.evaluate_ybar(device=device)
        for light in (False, True):
            for dist in C3_DISTS:
                hcomp.get_p(dist, light_weighted=light)
        for name in C3_MOMENTS:
            with np.errstate(invalid="ignore", divide="ignore"):
                getattr(hcomp, name)("v_x")
    e2e_ms = _wall(api_step, e2e_steps)
    err_api = _rel(hcomp.ybar, y_dev)
    tol_m = {"get_mean": 1e-10, "get_variance": 1e-9, "get_skewness": 1e-7, "get_kurtosis": 1e-7}
    ok = all(e < tol_m.get(k, TOL) for k, e in errs.items()) and med < 5e-4 and err_api < 1e-12
    return {
        "workload": f"C3 Mixture(GrowingDisk+Stream, [0.9, 0.1]) {nx}x{nx} px, full grid nt={nt} nz={nz}, nv={nv}, "
                    f"n_fft={plan.n_fft}: evaluate_ybar (path B on the pixelated mixture) + get_p for "
                    f"{', '.join(C3_DISTS)} (mass and light weighted) + mean/variance/skewness/kurtosis('v_x')",
        "config": {"nx1": nx, "nx2": nx, "nv": nv, "nt": int(nt), "nz": int(nz), "n_fft": int(plan.n_fft),
                   "l2": f"the {nbytes / 1e9:.1f} GB density is far larger than the 126 MB L2; no flush needed"},
        "voxels": voxels, "ms_per_step": ms, "value": voxels / (ms * 1e-3), "launches": int(launches),
        "e2e": {"value": voxels / (e2e_ms * 1e-3), "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(2 * host.nbytes),
                "d2h_bytes_per_step": int(hcomp.ybar.nbytes + sum(np.asarray(r).nbytes for r in results.values())),
                "steps": e2e_steps,
                "api": "Component(cube, log_p_tvxz: host).evaluate_ybar() + get_p(...) x10 + 4 moments; the host "
                       "array is streamed once for the cube and uploaded once for the marginals"},
        "roofline": {"bound": "hbm", "kernel": "k_reduce_online (vx marginal)", "achieved": nbytes / (ms_vx * 1e-3) / 1e9,
                     "peak": peaks["hbm"], "unit": "GB/s", "frac": nbytes / (ms_vx * 1e-3) / 1e9 / peaks["hbm"],
                     "traffic": None, "peak_source": peaks["source"],
                     "note": "one streaming read of log p(t,v,x,z) per marginal family (8 bytes per element); "
                             "the cube itself is FP64-pipe bound as in C4"},
        "kernels": {k: {"ms_per_step": m / steps, "launches_per_step": c / steps} for k, (m, c) in kt.items() if c},
        "other_paths": {"evaluate_ybar_ms": ms_cube, "log_marginal_vx_ms": ms_vx, "log_marginal_txz_ms": ms_txz,
                        "get_skewness_v_x_ms": ms_skew,
                        "get_skewness_note": "one pass over the 5-D array (p(v,x)); p(x), the conditional and the "
                                             "moments follow from the 32 MB marginal",
                        "host_model_setup_s": setup_s, "pixelate_on_device_s": pixelate_s},
        "verify": {"ok": bool(ok), "max_rel_err": max(errs.values()), "tol": TOL, "pixels": len(px), "errors": errs,
                   "median_rel_diff_mixture_vs_pixelated": med, "class_api_vs_device": err_api,
                   "oracle": "plain float64 sums of exp(log p + log dV) over the whole array (base.py:1017-1041, "
                             "130-177, 208-260) and oracle/restatement.py ybar_density_literal on the spot pixels"},
    }


# ----------------------------------------------------------------------------------------------
# C5: evaluate_ybar + ShotNoise at several SNRs, fp64 vs fp32
# ----------------------------------------------------------------------------------------------
def run_c5(nx=101, nv=101, steps=5, warmup=2, device=0, e2e_steps=2, snrs=(50.0, 100.0, 200.0)):
    import torch
    from oracle import restatement as R
    cube = make_cube(nx, nv)
    s = cube.ssps
    nz, nt = s.par_dims
    mix = make_mixture(cube)
    plan = engine.get_plan(cube, device=device)
    plan32 = engine.get_plan(cube, device=device, precision="float32")
    dev = f"cuda:{device}"
    comp = mix.pixelate(device=device)
    out = torch.empty((plan.n_fft, nx, nx), dtype=torch.float64, device=dev)
    voxels = plan.n_fft * nx * nx
    noisy = {}

    def step(keep=False):
        plan.ybar_density(comp.log_p_tvxz, is_log=True, out=out)
        for i, snr in enumerate(snrs):
            yobs, _ = engine.noise(out, snr, 0, seed=1234 + i, want_sigma=False, device=device)
            if keep:
                noisy[snr] = yobs

    step(keep=True)
    plan.set_profiling(True)
    l0 = plan.launches
    ms = _timed(step, steps, warmup)
    kt = plan.kernel_times()
    launches = plan.launches - l0 + 3 * len(snrs) * steps        # k_max (2) + k_noise per SNR
    plan.set_profiling(False)
    ms_noise = _timed(lambda: engine.noise(out, 100.0, 0, seed=1, want_sigma=False, device=device), 5, 1)
    y64 = out.clone()
    out32 = plan32.ybar_density(comp.log_p_tvxz, is_log=True)
    ms32 = _timed(lambda: plan32.ybar_density(comp.log_p_tvxz, is_log=True, out=out32), 3, 1)
    diff32 = float((out32 - y64).abs().max() / y64.abs().max())
    del out32
    # ---- verification: sigma maps (exact formula, noise.py:58-71), noise statistics, the cube itself
    yh = y64.cpu().numpy()
    errs = {}
    for snr in snrs:
        _, sig = engine.noise(y64, snr, 0, want_sample=False, device=device)
        ref = R.sigma_shot_noise(yh, snr=snr)
        errs[f"sigma_snr{snr:g}"] = _rel(sig.cpu().numpy(), ref)
        z = ((noisy[snr] - y64) / sig).flatten()
        z = z[torch.isfinite(z)]
        errs[f"z_mean_snr{snr:g}"] = abs(float(z.mean()))
        errs[f"z_std_snr{snr:g}"] = abs(float(z.std()) - 1.0)
        # snr is reached in the brightest voxel: sigma there = ybar/snr
        errs[f"peak_snr{snr:g}"] = abs(float(y64.max() / sig.flatten()[y64.argmax()]) / snr - 1.0)
    px = spot_pixels(nx)
    i1 = [p[0] for p in px]
    i2 = [p[1] for p in px]
    logp_px = comp.log_p_tvxz[:, :, i1, i2].cpu().numpy()[:, :, None]
    ref_y = R.ybar_density_literal(np.exp(logp_px), cube.v_edg, s.FXw, s.par_dims, s.n_fft, s.dv, s.delta_t,
                                   s.delta_z, cube.dx, cube.dy)[:, 0]
    errs["ybar"] = _rel(yh[:, i1, i2], ref_y)

    # ---- class API: host density -> evaluate_ybar -> ShotNoise(...).get_noisy_data(snr) x3 on the host
    host = comp.log_p_tvxz.cpu().numpy()
    hcomp = pkm.components.Component(cube=cube, log_p_tvxz=host)

    def api_step():
        hcomp.evaluate_ybar(device=device)
        shot = pkm.noise.ShotNoise(hcomp)
        return [shot.get_noisy_data(snr=snr) for snr in snrs]
    e2e_ms = _wall(api_step, e2e_steps)
    err_api = _rel(hcomp.ybar, yh)
    peaks = _peaks()
    nb = 2 * voxels * 8                                           # read ybar, write yobs
    n_z = voxels
    ok = (errs["ybar"] < TOL and diff32 < 1e-5 and err_api < 1e-12
          and all(errs[f"sigma_snr{q:g}"] < 1e-13 for q in snrs)
          and all(errs[f"z_mean_snr{q:g}"] < 5.0 / np.sqrt(n_z) and errs[f"z_std_snr{q:g}"] < 5e-3 for q in snrs)
          and all(errs[f"peak_snr{q:g}"] < 1e-12 for q in snrs))
    return {
        "workload": f"C5 end-to-end mock: Mixture(GrowingDisk+Stream) {nx}x{nx} px, full grid, nv={nv}, "
                    f"n_fft={plan.n_fft}: evaluate_ybar (float64) + ShotNoise at snr {', '.join(f'{q:g}' for q in snrs)}; "
                    f"float32 contraction checked against float64",
        "config": {"nx1": nx, "nx2": nx, "nv": nv, "nt": int(nt), "nz": int(nz), "n_fft": int(plan.n_fft),
                   "snr": list(snrs), "l2": "the 5.2 GB density is far larger than the 126 MB L2; no flush needed"},
        "voxels": voxels, "ms_per_step": ms, "value": voxels / (ms * 1e-3), "launches": int(launches),
        "e2e": {"value": voxels / (e2e_ms * 1e-3), "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(host.nbytes),
                "d2h_bytes_per_step": int(hcomp.ybar.nbytes * (1 + len(snrs))), "steps": e2e_steps,
                "api": "Component(cube, log_p_tvxz: host).evaluate_ybar(); ShotNoise(c).get_noisy_data(snr) x3"},
        "roofline": {"bound": "hbm", "kernel": "k_noise", "achieved": nb / (ms_noise * 1e-3) / 1e9, "peak": peaks["hbm"],
                     "unit": "GB/s", "frac": nb / (ms_noise * 1e-3) / 1e9 / peaks["hbm"], "traffic": None,
                     "peak_source": peaks["source"],
                     "note": "noise pass = max-reduction (one read) + Philox/Box-Muller draw (one read, one write); "
                             "the cube itself is FP64-pipe bound as in C4"},
        "kernels": {k: {"ms_per_step": m / steps, "launches_per_step": c / steps} for k, (m, c) in kt.items() if c},
        "other_paths": {"shot_noise_one_snr_ms": ms_noise, "evaluate_ybar_float32_ms": ms32,
                        "float32_vs_float64_max_rel_diff": diff32},
        "verify": {"ok": bool(ok), "max_rel_err": errs["ybar"], "tol": TOL, "pixels": len(px), "errors": errs,
                   "float32_vs_float64": diff32, "float32_tol": 1e-5, "class_api_vs_device": err_api,
                   "oracle": "oracle/restatement.py sigma_shot_noise (noise.py:58-71), ybar_density_literal"},
    }


def _peaks():
    import json
    import os
    try:
        p = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")))
        return {"hbm": float(p["hbm_gbs"]), "source": "MEASURED_PEAKS.json"}
    except Exception:  # noqa: BLE001
        return {"hbm": 6650.0, "source": "fallback 6650 GB/s (B200_PROFILING.md)"}


RUNNERS = {"C2": run_c2, "C3": run_c3, "C5": run_c5}
