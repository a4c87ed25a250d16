"""Generate tests/golden/*.npz by executing the UNMODIFIED reference in this container.

TEST INFRASTRUCTURE.  Run from the repo root:  python oracle/make_golden.py
Needs /root/reference (read-only) and `oracle/ref_shim.py`; the GPU box never runs it.
Each file stores the INPUTS of the hot path exactly as the reference produced them
(log_p_tvxz / P_txz / mu_v / sig_v / grids) together with the reference OUTPUTS
(`ybar`, marginals, moments, noise sigma), so the oracle restatement and the CUDA
path can be diffed against the real thing anywhere.

Model parameters follow the reference's own fixtures
(/root/reference/tests/conftest.py:24-104) on smaller pixel grids.
"""
import os
import sys
import contextlib
import io

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
np.seterr(divide="ignore", invalid="ignore")


def make_cube(pkm, nx1, nx2, nv, vlim=1000.0, **ssp_kw):
    ssps = ref_shim.reference_ssps(pkm, **ssp_kw)
    return pkm.ifu_cube.IFUCube(ssps=ssps, nx1=nx1, nx2=nx2, nv=nv, vrng=(-vlim, vlim))


def disk_one(pkm, cube):
    c = pkm.components.GrowingDisk(cube=cube, rotation=0.0, center=(0.2, 0))
    c.set_p_t(lmd=2.0, phi=0.8)
    c.set_p_x_t(q_lims=(0.05, 0.5), rc_lims=(1.0, 1.0), alpha_lims=(1.2, 0.8))
    c.set_t_dep(q=0.2, alpha=3.0, t_dep_in=0.5, t_dep_out=5.0)
    c.set_mu_v(q_lims=(0.5, 0.1), rmax_lims=(0.5, 1.1), vmax_lims=(250.0, 100.0))
    c.set_sig_v(q_lims=(0.25, 1.0), alpha_lims=(3.0, 2.5), sig_v_in_lims=(70.0, 50.0),
                sig_v_out_lims=(80.0, 60.0))
    return c


def disk_two(pkm, cube):
    c = pkm.components.GrowingDisk(cube=cube, rotation=np.deg2rad(10.0), center=(0.05, -0.07))
    c.set_p_t(lmd=1.6, phi=0.3)
    c.set_p_x_t(q_lims=(1.0, 1.0), rc_lims=(1.0, 1.0), alpha_lims=(1.1, 1.1))
    c.set_t_dep(q=0.7, alpha=1.1, t_dep_in=6.0, t_dep_out=1.0)
    c.set_mu_v(q_lims=(0.5, 0.4), rmax_lims=(0.7, 0.2), vmax_lims=(-150.0, -190.0))
    c.set_sig_v(q_lims=(0.3 / 0.4, 0.4 / 0.3), alpha_lims=(2.0, 1.5),
                sig_v_in_lims=(70.0, 90.0), sig_v_out_lims=(10.0, 20.0))
    return c


def stream_one(pkm, cube):
    s = pkm.components.Stream(cube=cube, rotation=0.0, center=(0.0, 0))
    s.set_p_t(lmd=15.0, phi=0.3)
    s.set_p_x_t(theta_lims=[-np.pi / 2.0, 0.75 * np.pi], mu_r_lims=[0.2, 0.8], nsmp=1000, sig=0.1)
    s.set_t_dep(t_dep=5.0)
    s.set_mu_v(mu_v_lims=[-80, 100])
    s.set_sig_v(sig_v=110.0)
    return s


def cube_record(cube):
    s = cube.ssps
    k = np.array([0, 1, 7, s.FXw.shape[0] // 2, s.FXw.shape[0] - 1])
    return dict(
        nx1=cube.nx, nx2=cube.ny, nv=cube.nv, v_edg=cube.v_edg, dx=cube.dx, dy=cube.dy,
        n_fft=s.n_fft, dv=s.dv, par_dims=np.array(s.par_dims), delta_t=s.delta_t,
        delta_z=s.delta_z, light_weights=s.light_weights, w_first_last=np.array([s.w[0], s.w[-1]]),
        t_cents=s.par_cents[1], z_cents=s.par_cents[0], t_edges=s.par_edges[1],
        z_edges=s.par_edges[0], lmd_first_last=np.array([s.lmd[0], s.lmd[-1]]),
        FXw_rows_idx=k, FXw_rows=s.FXw[k, :], FXw_abs_sum=np.sum(np.abs(s.FXw)),
        FXw_sum=np.sum(s.FXw), Xw_sum=np.sum(s.Xw),
    )


def parametric_record(prefix, c):
    with contextlib.redirect_stdout(io.StringIO()):
        c.evaluate_ybar()
    rec = {
        prefix + "P_txz": c.get_p("txz", density=False),
        prefix + "mu_v": np.array(c.mu_v),
        prefix + "sig_v": np.array(c.sig_v),
        prefix + "ybar": c.ybar,
        prefix + "log_p_t": c.log_p_t,
        prefix + "log_p_x_t": np.array(c.log_p_x_t),
        prefix + "log_p_z_tx": c.log_p_z_tx,
        prefix + "t_dep": c.t_dep,
    }
    return rec


DISTS = ["t", "v", "x", "z", "tv", "tx", "tz", "vx", "vz", "xz", "tvx", "tvz", "txz", "vxz",
         "v_x", "tz_x", "t_x", "z_t", "x_tz", "v_tx", "tv_xz", "x_t"]
MOMENT_DISTS = ["v_x", "v_tx", "v_t", "v_tz", "v_xz", "v_z", "v", "t", "z", "t_x", "z_t", "x", "x_tz"]


def base_records(pkm, cube, log_p_tvxz, with_dists=True, skip=()):
    b = pkm.components.base.Component(cube=cube, log_p_tvxz=log_p_tvxz)
    b.evaluate_ybar()
    rec = dict(b_log_p_tvxz=log_p_tvxz, b_ybar=b.ybar)
    if with_dists:
        for d in [d for d in DISTS if d not in skip]:
            for lw in (False, True):
                for dens in (False, True):
                    rec[f"b_logp|{d}|{int(lw)}|{int(dens)}"] = b.get_log_p(d, density=dens, light_weighted=lw)
        for d in MOMENT_DISTS:
            for lw in (False, True):
                rec[f"b_mean|{d}|{int(lw)}"] = b.get_mean(d, light_weighted=lw)
                rec[f"b_var|{d}|{int(lw)}"] = b.get_variance(d, light_weighted=lw)
                rec[f"b_skew|{d}|{int(lw)}"] = b.get_skewness(d, light_weighted=lw)
                rec[f"b_kurt|{d}|{int(lw)}"] = b.get_kurtosis(d, light_weighted=lw)
        rec["b_sigma_shot_snr50"] = pkm.noise.ShotNoise(b).get_sigma(snr=50.0)
        rec["b_sigma_const_snr50"] = pkm.noise.ConstantSNR(b).get_sigma(snr=50.0)
    return rec, b


def white_noise_log_p(cube, seed):
    rng = np.random.default_rng(seed)
    shape = cube.get_distribution_shape("tvxz")
    p = rng.random(shape)
    p /= np.sum(p * cube.construct_volume_element("tvxz"))
    return np.log(p)


def save(name, rec):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **rec)
    print(f"{name}: {os.path.getsize(path)/1e6:.2f} MB, {len(rec)} arrays")


def golden_ppxf(pkm):
    """g5: the reference's pad="ppxf" option (model_grids.py:263-281): n_fft = next power of two
    above the 1021 log-lambda pixels = 1024, so the cube has 3 more samples than the spectra and
    the convolution wraps into the padding.  Paths A and B of the unmodified reference."""
    cube = make_cube(pkm, 2, 3, 21, age_rebin=6, z_rebin=2)
    cube.ssps._calculate_fourier_transform(pad="ppxf")
    rec = cube_record(cube)
    rec["n_pix"] = np.array(cube.ssps.n_pix)
    rec.update(parametric_record("d1_", disk_one(pkm, cube)))
    wrec, _ = base_records(pkm, cube, white_noise_log_p(cube, 5), with_dists=False)
    rec.update({"wn_" + k[2:]: v for k, v in wrec.items()})
    save("g5_ppxf_pad_nv21", rec)


def main():
    pkm = ref_shim.load()
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "g5":      # added in round 2: leaves g1-g4 untouched
        golden_ppxf(pkm)
        return

    # ---- g1: conftest galaxy (full 53x12 grid, nv=41 -> n_fft=1992 even), 3x2 pixels
    cube = make_cube(pkm, 3, 2, 41)
    rec = cube_record(cube)
    comps = [disk_one(pkm, cube), disk_two(pkm, cube), stream_one(pkm, cube)]
    for tag, c in zip(("d1_", "d2_", "s1_"), comps):
        rec.update(parametric_record(tag, c))
    weights = [0.65, 0.25, 0.1]
    galaxy = pkm.components.Mixture(cube=cube, component_list=comps, weights=weights)
    galaxy.evaluate_ybar()
    rec["mix_weights"] = np.array(weights)
    rec["mix_ybar"] = galaxy.ybar
    log_p = galaxy.get_log_p("tvxz", density=True, light_weighted=False, collapse_cmps=True)
    brec, _ = base_records(pkm, cube, log_p, skip=("tv_xz",))
    rec.update(brec)
    for d in ["t", "x", "tz", "v_x", "tz_x", "vx"]:
        for lw in (False, True):
            rec[f"mix_logp|{d}|{int(lw)}|1"] = galaxy.get_log_p(d, density=True, light_weighted=lw)
    save("g1_conftest_full_nv41", rec)

    # ---- g2: rebinned (6x9) grid, nv=21 -> n_fft=1021 (prime), 4x5 pixels
    cube = make_cube(pkm, 4, 5, 21, age_rebin=6, z_rebin=2)
    rec = cube_record(cube)
    d1 = disk_one(pkm, cube)
    rec.update(parametric_record("d1_", d1))
    s1 = stream_one(pkm, cube)
    rec.update(parametric_record("s1_", s1))
    brec, _ = base_records(pkm, cube, d1.get_log_p("tvxz"))
    rec.update(brec)
    wrec, _ = base_records(pkm, cube, white_noise_log_p(cube, 21), with_dists=False)
    rec.update({"wn_" + k[2:]: v for k, v in wrec.items()})
    save("g2_rebin_nv21_prime", rec)

    # ---- g3: full grid, nv=101 -> n_fft=4907 = 7*701 (odd), 1x2 pixels (headline shape)
    cube = make_cube(pkm, 1, 2, 101)
    rec = cube_record(cube)
    d1 = disk_one(pkm, cube)
    rec.update(parametric_record("d1_", d1))
    wrec, _ = base_records(pkm, cube, white_noise_log_p(cube, 20241019), with_dists=False)
    rec.update({"wn_" + k[2:]: v for k, v in wrec.items()})
    save("g3_full_nv101_odd", rec)

    # ---- g4: even nv (nv=40 -> no zero-centred bin: path B raises IndexError), path A only
    cube = make_cube(pkm, 2, 2, 40, age_rebin=4, z_rebin=3)
    rec = cube_record(cube)
    rec.update(parametric_record("d2_", disk_two(pkm, cube)))
    b = pkm.components.base.Component(cube=cube, log_p_tvxz=white_noise_log_p(cube, 4))
    try:
        b.evaluate_ybar()
        rec["b_error"] = np.array("none")
    except Exception as exc:  # noqa: BLE001
        rec["b_error"] = np.array(type(exc).__name__)
    save("g4_even_nv40", rec)
    golden_ppxf(pkm)


if __name__ == "__main__":
    main()
