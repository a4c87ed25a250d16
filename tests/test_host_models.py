"""Host-side classes (no GPU): grids, SSP operand, parametric models, error behaviour.

Every array is compared with what the unmodified reference produced for the same
configuration (tests/golden/*.npz)."""
import numpy as np
import pytest

from conftest import GOLDEN_CUBES, build_cube, golden_cube, load_golden, rel_err

import popkinmocks_b200 as pkm


@pytest.mark.parametrize("name", list(GOLDEN_CUBES))
def test_cube_and_ssp_operand_match_reference(name):
    g = load_golden(name)
    cube = golden_cube(name)
    s = cube.ssps
    assert s.n_fft == int(g["n_fft"]) and tuple(s.par_dims) == tuple(g["par_dims"])
    assert s.dv == float(g["dv"]) and cube.dx == float(g["dx"]) and cube.dy == float(g["dy"])
    for mine, key in [(cube.v_edg, "v_edg"), (s.delta_t, "delta_t"), (s.delta_z, "delta_z"),
                      (s.light_weights, "light_weights"), (s.par_cents[1], "t_cents"),
                      (s.par_cents[0], "z_cents"), (s.par_edges[1], "t_edges"), (s.par_edges[0], "z_edges")]:
        assert np.array_equal(mine, g[key]), key
    # the constant operand S-hat of the hot path, bit for bit
    assert np.array_equal(s.FXw[g["FXw_rows_idx"]], g["FXw_rows"])
    assert np.sum(np.abs(s.FXw)) == float(g["FXw_abs_sum"])
    assert np.sum(s.FXw) == complex(g["FXw_sum"])
    assert cube.get_distribution_shape("tvxz") == (s.par_dims[1], cube.nv, cube.nx, cube.ny, s.par_dims[0])
    for var in ("t", "v", "x1", "x2", "z"):
        assert cube.get_variable_edges(var).size == cube.get_variable_values(var).size + 1


def _disk_one(cube):
    c = pkm.components.GrowingDisk(cube=cube, rotation=0.0, center=(0.2, 0))
    c.set_p_t(lmd=2.0, phi=0.8)
    c.set_p_x_t(q_lims=(0.05, 0.5), rc_lims=(1.0, 1.0), alpha_lims=(1.2, 0.8))
    c.set_t_dep(q=0.2, alpha=3.0, t_dep_in=0.5, t_dep_out=5.0)
    c.set_mu_v(q_lims=(0.5, 0.1), rmax_lims=(0.5, 1.1), vmax_lims=(250.0, 100.0))
    c.set_sig_v(q_lims=(0.25, 1.0), alpha_lims=(3.0, 2.5), sig_v_in_lims=(70.0, 50.0),
                sig_v_out_lims=(80.0, 60.0))
    return c


def _disk_two(cube):
    c = pkm.components.GrowingDisk(cube=cube, rotation=np.deg2rad(10.0), center=(0.05, -0.07))
    c.set_p_t(lmd=1.6, phi=0.3)
    c.set_p_x_t(q_lims=(1.0, 1.0), rc_lims=(1.0, 1.0), alpha_lims=(1.1, 1.1))
    c.set_t_dep(q=0.7, alpha=1.1, t_dep_in=6.0, t_dep_out=1.0)
    c.set_mu_v(q_lims=(0.5, 0.4), rmax_lims=(0.7, 0.2), vmax_lims=(-150.0, -190.0))
    c.set_sig_v(q_lims=(0.3 / 0.4, 0.4 / 0.3), alpha_lims=(2.0, 1.5), sig_v_in_lims=(70.0, 90.0),
                sig_v_out_lims=(10.0, 20.0))
    return c


def _stream_one(cube):
    s = pkm.components.Stream(cube=cube, rotation=0.0, center=(0.0, 0))
    s.set_p_t(lmd=15.0, phi=0.3)
    s.set_p_x_t(theta_lims=[-np.pi / 2.0, 0.75 * np.pi], mu_r_lims=[0.2, 0.8], nsmp=1000, sig=0.1)
    s.set_t_dep(t_dep=5.0)
    s.set_mu_v(mu_v_lims=[-80, 100])
    s.set_sig_v(sig_v=110.0)
    return s


BUILDERS = {"d1_": _disk_one, "d2_": _disk_two, "s1_": _stream_one}


@pytest.mark.parametrize("name,tags", [("g1_conftest_full_nv41", ["d1_", "d2_", "s1_"]),
                                       ("g2_rebin_nv21_prime", ["d1_", "s1_"]), ("g4_even_nv40", ["d2_"])])
def test_parametric_inputs_of_the_hot_path_match_reference(name, tags):
    g = load_golden(name)
    cube = golden_cube(name)
    for tag in tags:
        c = BUILDERS[tag](cube)
        for attr in ("log_p_t", "log_p_x_t", "log_p_z_tx", "t_dep", "mu_v", "sig_v"):
            mine, ref = np.asarray(getattr(c, attr)), g[tag + attr]
            assert mine.shape == ref.shape, (tag, attr)
            ok = np.isfinite(ref)
            assert np.array_equal(np.isfinite(mine), ok), (tag, attr)
            assert np.allclose(mine[ok], ref[ok], rtol=1e-13, atol=1e-13), (tag, attr)
        P = c.get_p("txz", density=False)
        assert rel_err(P, g[tag + "P_txz"]) < 1e-13


def test_parametric_marginals_are_normalised_and_consistent():
    cube = golden_cube("g2_rebin_nv21_prime")
    c = _disk_one(cube)
    for lw in (False, True):
        for dist in ("t", "x", "z", "tx", "tz", "xz", "txz"):
            assert np.isclose(np.sum(c.get_p(dist, density=False, light_weighted=lw)), 1.0, atol=1e-10), dist
        # conditionals: hard-coded factor == joint / marginal
        a = c.get_log_p("z_tx", density=True, light_weighted=lw)
        assert a.shape == (cube.ssps.par_dims[0], cube.ssps.par_dims[1], cube.nx, cube.ny)
    joint = c._log_marginal("tx", density=False) [:, None] + np.moveaxis(c._log_P_v_tx(), 0, 1)
    assert np.allclose(c.get_log_p("tvx", density=False), joint, atol=1e-12)
    # analytic velocity moments: a Gaussian law has zero skewness and kurtosis 3 given (t,x)
    assert np.allclose(c.get_mean("v_tx"), c.mu_v)
    assert np.allclose(c.get_variance("v_tx"), c.sig_v ** 2)
    assert np.allclose(c.get_skewness("v_tx"), 0.0)
    assert np.allclose(c.get_kurtosis("v_tx"), 3.0)
    ev = c.get_mean("v")
    assert np.isclose(ev, np.sum(c.mu_v * c.get_p("tx", density=False)))
    assert c.get_mean("v_x").shape == (cube.nx, cube.ny)
    assert c.get_variance("v_tz").shape == (cube.ssps.par_dims[1], cube.ssps.par_dims[0])
    assert np.all(c.get_variance("v_x") > 0)


def test_analytic_velocity_moments_converge():
    """Analytic mixture skewness of v vs the discrete sum on a fine velocity grid (nv=201)."""
    from oracle import restatement as R
    ssps = pkm.milesSSPs(age_rebin=4, z_rebin=2)
    cube = pkm.ifu_cube.IFUCube(ssps=ssps, nx1=3, nx2=2, nv=201, vrng=(-1000, 1000))
    comps = [BUILDERS[t](cube) for t in ("d1_", "d2_", "s1_")]
    mix = pkm.components.Mixture(cube=cube, component_list=comps, weights=[0.65, 0.25, 0.1])
    v = cube.get_variable_values("v")
    with np.errstate(all="ignore"):
        for lw in (False, True):
            for dist in ("v_tx", "v_x", "v_t", "v_tz", "v_xz", "v_z"):
                analytic = mix.get_skewness(dist, light_weighted=lw)
                discrete = R.standardised_moment_1d(mix.get_p(dist, density=False, light_weighted=lw), v, 3)
                assert np.nanmedian(np.abs((analytic - discrete) / analytic)) < 0.01, (dist, lw)
                kurt = mix.get_kurtosis(dist, light_weighted=lw)
                kd = R.standardised_moment_1d(mix.get_p(dist, density=False, light_weighted=lw), v, 4)
                assert np.nanmedian(np.abs((kurt - kd) / kurt)) < 0.01, (dist, lw)


def test_constructor_and_argument_errors_match_reference():
    cube = golden_cube("g2_rebin_nv21_prime")
    shape = cube.get_distribution_shape("tvxz")
    with pytest.raises(ValueError, match="should have shape"):
        pkm.components.Component(cube=cube, log_p_tvxz=np.zeros(shape[:-1]))
    comp = pkm.components.Component(cube=cube, log_p_tvxz=np.zeros(shape))
    with pytest.raises(ValueError, match="Unknown option for batch"):
        comp.evaluate_ybar(batch="rows")
    with pytest.raises(ValueError, match="Unknown option for batch"):
        _disk_one(cube).evaluate_ybar(batch="rows")
    with pytest.raises(ValueError, match="bimodal"):
        pkm.components.GrowingDisk(cube=cube).set_p_t(lmd=1.0, phi=0.5)
    # even nv: no zero-centred velocity bin -> IndexError, exactly as the reference
    even = golden_cube("g4_even_nv40")
    with pytest.raises(IndexError):
        pkm.components.Component(cube=even, log_p_tvxz=np.zeros(even.get_distribution_shape("tvxz"))).evaluate_ybar()
    # velocity step different from the SSP step -> ValueError naming the requirement
    bad = build_cube("g2_rebin_nv21_prime")
    bad.v_edg = bad.v_edg * 1.5
    with pytest.raises(ValueError, match="uniformly spaced"):
        pkm.components.Component(cube=bad, log_p_tvxz=np.zeros(shape)).evaluate_ybar()


def test_plans_are_not_pickled():
    import dill
    cube = build_cube("g4_even_nv40")
    cube._pkm_plans["fake"] = object()
    clone = dill.loads(dill.dumps(cube))
    assert clone._pkm_plans == {} and clone.ssps.n_fft == cube.ssps.n_fft


def test_mixture_distributions_match_reference_on_host():
    """Mixture.get_log_p composes the children's host-side marginals: checked against the
    reference's outputs for the conftest galaxy (no GPU involved)."""
    g = load_golden("g1_conftest_full_nv41")
    cube = golden_cube("g1_conftest_full_nv41")
    comps = [BUILDERS[t](cube) for t in ("d1_", "d2_", "s1_")]
    mix = pkm.components.Mixture(cube=cube, component_list=comps, weights=g["mix_weights"])
    keys = [k for k in g.files if k.startswith("mix_logp|")]
    assert len(keys) == 12
    for k in keys:
        _, dist, lw, _ = k.split("|")
        out = mix.get_log_p(dist, density=True, light_weighted=bool(int(lw)))
        ok = np.isfinite(g[k])
        assert out.shape == g[k].shape
        assert np.allclose(out[ok], g[k][ok], rtol=0, atol=1e-10), k
    # the pixelated copy of the mixture is what the reference feeds to the FFT path
    log_p = mix.get_log_p("tvxz", density=True, light_weighted=False, collapse_cmps=True)
    ok = np.isfinite(g["b_log_p_tvxz"])
    assert np.array_equal(np.isfinite(log_p), ok)
    assert np.allclose(log_p[ok], g["b_log_p_tvxz"][ok], rtol=0, atol=1e-10)
    for lw in (False, True):
        assert np.isclose(np.sum(mix.get_p("tvxz", density=False, light_weighted=lw)), 1.0, atol=1e-6)


def test_save_numpy_and_dill_dump_formats(tmp_path):
    """The reference's two on-disk formats (base.py:1176-1225): the dill pickle round-trips without
    device handles, the v0.0 npz carries the same keys and the density rearranged to f_xvtz."""
    import dill
    g = load_golden("g1_conftest_full_nv41")
    cube = golden_cube("g1_conftest_full_nv41")
    comps = [BUILDERS[t](cube) for t in ("d1_", "s1_")]
    mix = pkm.components.Mixture(cube=cube, component_list=comps, weights=[0.7, 0.3])
    mix.ybar = 0.7 * g["d1_ybar"] + 0.3 * g["s1_ybar"]        # stands in for evaluate_ybar (GPU)
    direc = str(tmp_path / "out") + "/"                        # the reference concatenates direc + fname
    mix.save_numpy("mix.npz", yobs=mix.ybar + 1.0, direc=direc)
    z = np.load(direc + "mix.npz")
    assert sorted(z.files) == sorted(["yobs", "nx1", "nx2", "x1rng", "x2rng", "S", "w", "z_bin_edges",
                                      "t_bin_edges", "ybar", "v_edg", "f_xvtz"])
    p = mix.get_p("tvxz", density=True, collapse_cmps=True)
    nt, nv, nx1, nx2, nz = p.shape
    assert z["f_xvtz"].shape == (nx1, nx2, nv, nz, nt)
    assert np.array_equal(z["f_xvtz"], np.transpose(p, (2, 3, 1, 4, 0)))
    assert z["S"].shape == (cube.ssps.Xw.shape[0], nz, nt) and np.array_equal(z["w"], cube.ssps.w)
    assert int(z["nx1"]) == nx1 and int(z["nx2"]) == nx2 and np.array_equal(z["ybar"], mix.ybar)
    assert z["z_bin_edges"].size == nz + 1 and z["t_bin_edges"].size == nt + 1
    # a single parametric component goes through the same writer (no collapse_cmps argument)
    comps[0].ybar = g["d1_ybar"]
    comps[0].save_numpy("disk.npz", direc=direc)
    zd = np.load(direc + "disk.npz", allow_pickle=True)
    assert np.array_equal(zd["f_xvtz"], np.transpose(comps[0].get_p("tvxz"), (2, 3, 1, 4, 0)))
    # dill
    cube._pkm_plans["fake"] = object()
    comps[0].dill_dump("disk.dill", direc=direc)
    cube._pkm_plans.pop("fake")
    with open(direc + "disk.dill", "rb") as fh:
        clone = dill.load(fh)
    assert clone.cube._pkm_plans == {} and np.array_equal(clone.ybar, comps[0].ybar)
    assert np.array_equal(clone.mu_v, comps[0].mu_v)
