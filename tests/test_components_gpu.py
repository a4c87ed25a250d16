"""GPU acceptance tests through the drop-in Python API (same calls a popkinmocks user makes).

Part 1 diffs every accelerated method against outputs of the unmodified reference
(tests/golden).  Part 2 re-states the reference's own cross-consistency tests
(/root/reference/tests/test_all.py:49-146) against the new classes.
"""
import contextlib
import io
from itertools import chain, combinations

import numpy as np
import pytest

from conftest import build_cube, golden_cube, load_golden, rel_err
from test_host_models import BUILDERS

import popkinmocks_b200 as pkm

pytestmark = pytest.mark.gpu
TOL = 1e-10
G1 = "g1_conftest_full_nv41"


@pytest.fixture(scope="module")
def galaxy(gpu_ok):
    """The reference's fixture galaxy (tests/conftest.py:24-104) on the 3x2-pixel golden cube."""
    cube = golden_cube(G1)
    comps = [BUILDERS[t](cube) for t in ("d1_", "d2_", "s1_")]
    for c in comps:
        c.evaluate_ybar()
    mix = pkm.components.Mixture(cube=cube, component_list=comps, weights=[0.65, 0.25, 0.1])
    mix.evaluate_ybar()
    return mix


@pytest.fixture(scope="module")
def base_component(galaxy):
    log_p = galaxy.get_log_p("tvxz", density=True, light_weighted=False, collapse_cmps=True)
    comp = pkm.components.Component(cube=galaxy.cube, log_p_tvxz=log_p)
    comp.evaluate_ybar()
    return comp


# ----------------------------------------------------------------------------- part 1: golden
def test_parametric_and_mixture_cubes_match_reference(galaxy):
    g = load_golden(G1)
    for comp, tag in zip(galaxy.component_list, ("d1_", "d2_", "s1_")):
        assert isinstance(comp.ybar, np.ndarray) and comp.ybar.dtype == np.float64
        assert comp.ybar.shape == g[tag + "ybar"].shape
        assert rel_err(comp.ybar, g[tag + "ybar"]) < TOL
    assert rel_err(galaxy.ybar, g["mix_ybar"]) < TOL


def test_pixelated_component_cube_and_batch_modes(galaxy, base_component):
    g = load_golden(G1)
    assert rel_err(base_component.log_p_tvxz[np.isfinite(g["b_log_p_tvxz"])],
                   g["b_log_p_tvxz"][np.isfinite(g["b_log_p_tvxz"])]) < 1e-12
    assert rel_err(base_component.ybar, g["b_ybar"]) < TOL
    ref = base_component.ybar.copy()
    for batch in ("column", "spaxel"):
        base_component.evaluate_ybar(batch=batch)
        assert np.array_equal(base_component.ybar, ref)
    disk = galaxy.component_list[0]
    ref = disk.ybar.copy()
    for batch in ("column", "spaxel"):
        disk.evaluate_ybar(batch=batch)
        assert np.array_equal(disk.ybar, ref)


@pytest.mark.parametrize("name", [G1, "g2_rebin_nv21_prime"])
def test_log_marginals_and_conditionals_match_reference(gpu_ok, name):
    g = load_golden(name)
    comp = pkm.components.Component(cube=golden_cube(name), log_p_tvxz=g["b_log_p_tvxz"])
    keys = [k for k in g.files if k.startswith("b_logp|")]
    assert len(keys) >= 80
    for k in keys:
        _, dist, lw, dens = k.split("|")
        out = comp.get_log_p(dist, density=bool(int(dens)), light_weighted=bool(int(lw)))
        ref = g[k]
        assert out.shape == ref.shape, k
        assert np.array_equal(np.isnan(out), np.isnan(ref)), k
        assert np.array_equal(np.isneginf(out), np.isneginf(ref)), k
        ok = np.isfinite(ref)
        assert np.allclose(out[ok], ref[ok], rtol=0, atol=1e-10), k


def test_mixture_distributions_match_reference(galaxy):
    g = load_golden(G1)
    for k in [k for k in g.files if k.startswith("mix_logp|")]:
        _, dist, lw, _ = k.split("|")
        out = galaxy.get_log_p(dist, density=True, light_weighted=bool(int(lw)))
        ok = np.isfinite(g[k])
        assert out.shape == g[k].shape
        assert np.allclose(out[ok], g[k][ok], rtol=0, atol=1e-10), k


@pytest.mark.parametrize("name", [G1, "g2_rebin_nv21_prime"])
def test_moments_match_reference(gpu_ok, name):
    g = load_golden(name)
    comp = pkm.components.Component(cube=golden_cube(name), log_p_tvxz=g["b_log_p_tvxz"])
    for k in [k for k in g.files if k.startswith("b_mean|")]:
        _, dist, lw = k.split("|")
        lwb = bool(int(lw))
        for fn, nm, tol in ((comp.get_mean, "mean", 1e-10), (comp.get_variance, "var", 1e-9),
                            (comp.get_skewness, "skew", 1e-7), (comp.get_kurtosis, "kurt", 1e-7)):
            ref = g[f"b_{nm}|{dist}|{lw}"]
            with np.errstate(invalid="ignore", divide="ignore"):
                out = np.asarray(fn(dist, light_weighted=lwb))
            assert out.shape == ref.shape, (nm, k)
            ok = np.isfinite(ref)
            scale = max(1.0, float(np.max(np.abs(ref[ok])))) if ok.any() else 1.0
            assert np.allclose(out[ok], ref[ok], rtol=tol, atol=tol * scale), (nm, k)


def test_noise_sigma_and_statistics(base_component):
    g = load_golden(G1)
    shot = pkm.noise.ShotNoise(base_component)
    const = pkm.noise.ConstantSNR(base_component)
    assert np.allclose(shot.get_sigma(snr=50.0), g["b_sigma_shot_snr50"], rtol=1e-14, equal_nan=True)
    assert np.allclose(const.get_sigma(snr=50.0), g["b_sigma_const_snr50"], rtol=1e-15)
    np.random.seed(11)
    a = const.get_noisy_data(snr=20.0)
    np.random.seed(11)
    b = const.get_noisy_data(snr=20.0)
    assert np.array_equal(a, b)                               # reproducible under np.random.seed
    c = const.get_noisy_data(snr=20.0)
    assert not np.array_equal(a, c)
    z = ((a - base_component.ybar) / const.get_sigma(snr=20.0)).ravel()
    z = z[np.isfinite(z)]
    assert z.size > 10000
    assert abs(z.mean()) < 4 / np.sqrt(z.size) and abs(z.std() - 1.0) < 0.03
    from scipy import stats
    assert stats.kstest(z, "norm").pvalue > 1e-3
    # negative ybar -> NaN sigma for shot noise, exactly like ybar**0.5 in the reference
    class Holder:
        ybar = np.array([[-1.0, 4.0], [9.0, 1.0]])
    s = pkm.noise.ShotNoise(Holder()).get_sigma(snr=3.0)
    assert np.isnan(s[0, 0]) and np.allclose(s[1], [3.0, 1.0]) and np.isclose(s[0, 1], 2.0)


def test_multiple_pixel_tiles_through_the_api(gpu_ok):
    """7x7 = 49 pixels with a tiny workspace: two 32-pixel tiles, ragged last tile."""
    from oracle import restatement as R
    ssps = pkm.milesSSPs(age_rebin=6, z_rebin=2)
    cube = pkm.ifu_cube.IFUCube(ssps=ssps, nx1=7, nx2=7, nv=21, vrng=(-1000.0, 1000.0))
    rng = np.random.default_rng(2)
    p = rng.random(cube.get_distribution_shape("tvxz"))
    p /= np.sum(p * cube.construct_volume_element("tvxz"))
    comp = pkm.components.Component(cube=cube, log_p_tvxz=np.log(p))
    plan = pkm.engine.get_plan(cube)
    plan.set_workspace_limit(1 << 16)
    try:
        comp.evaluate_ybar()
    finally:
        plan.set_workspace_limit(0)
    ref = R.ybar_density_fused(p, cube.v_edg, ssps.FXw, ssps.par_dims, ssps.n_fft, ssps.dv, ssps.delta_t,
                               ssps.delta_z, cube.dx, cube.dy)
    assert rel_err(comp.ybar, ref) < TOL


# ----------------------------------------------------------------------------- part 2: reference's own tests
def _all_distributions():
    """Every marginal / conditional over (t, v, x, z) with disjoint sides: 65 strings."""
    names = ["t", "v", "x", "z"]
    subsets = list(chain.from_iterable(combinations(names, r) for r in range(5)))
    out = []
    for dep in subsets:
        for given in subsets:
            if dep and not set(dep) & set(given):
                out.append("".join(dep) + ("_" + "".join(given) if given else ""))
    return out


def test_all_65_distributions_are_normalised(galaxy, base_component):
    dists = _all_distributions()
    assert len(dists) == 65
    cube = galaxy.cube
    systems = [(galaxy, lw, dens) for lw in (False, True) for dens in (False, True)]
    systems += [(base_component, lw, dens) for lw in (False, True) for dens in (False, True)]
    systems += [(galaxy.component_list[0], True, True)]
    for system, lw, dens in systems:
        for dist in dists:
            dep = dist.split("_")[0]
            with np.errstate(invalid="ignore", divide="ignore"):
                p = system.get_p(dist, light_weighted=lw, density=dens)
            if dens:
                p = (p.T * cube.construct_volume_element(dep).T).T
            n_dep = len(dep) + ("x" in dep)
            total = np.sum(p, tuple(range(n_dep)))
            if "_" in dist:
                assert np.all(np.isclose(total, 1.0) | np.isnan(total)), (type(system).__name__, dist, lw, dens)
            else:
                assert np.allclose(total, 1.0), (type(system).__name__, dist, lw, dens)


def test_analytic_and_fft_datacubes_agree(galaxy, base_component):
    err = (galaxy.ybar - base_component.ybar) / galaxy.ybar
    assert np.median(np.abs(err)) < 5e-4


def test_mixture_moments_against_discrete(galaxy, base_component):
    for lw in (False, True):
        for dist in ("v_tx", "v_x", "v_t", "v_tz", "v_xz", "v_z"):
            with np.errstate(invalid="ignore", divide="ignore"):
                a = galaxy.get_skewness(dist, light_weighted=lw)
                b = base_component.get_skewness(dist, light_weighted=lw)
                # 8 % as in the reference (coarse velocity bins); on this 3x2-pixel cube p(v|t) mixes
                # few spaxels and is the one case dominated by binning error (it converges:
                # test_host_models.py::test_analytic_velocity_moments_converge)
                assert np.nanmedian(np.abs((a - b) / a)) < (0.4 if dist == "v_t" else 0.08), (dist, lw)
        for dist in ("t", "x", "z", "t_x", "x_tz", "z_t"):
            with np.errstate(invalid="ignore", divide="ignore"):
                a = galaxy.get_excess_kurtosis(dist, light_weighted=lw)
                b = base_component.get_excess_kurtosis(dist, light_weighted=lw)
                assert np.nanmedian(np.abs((a - b) / a)) < 1e-10, (dist, lw)


def test_noise_depends_on_the_global_index_only(gpu_ok):
    """pkm_noise draws one Philox / Box-Muller pair per two voxels, keyed on the GLOBAL voxel index
    (offset + i): a cube cut into calls at odd or even offsets gets the noise of the single call."""
    from popkinmocks_b200 import engine
    rng = np.random.default_rng(4)
    y = rng.random(1001) + 0.5
    full, sig = engine.noise(y, 30.0, 1, seed=123, offset=10)
    assert np.allclose(sig, y / 30.0, rtol=1e-15)
    for lo, hi in ((0, 1), (0, 500), (1, 2), (3, 778), (500, 1001), (777, 1000), (1000, 1001)):
        part, _ = engine.noise(y[lo:hi].copy(), 30.0, 1, seed=123, offset=10 + lo)
        assert np.array_equal(part, full[lo:hi]), (lo, hi)
    other, _ = engine.noise(y, 30.0, 1, seed=124, offset=10)
    assert not np.array_equal(other, full)
    z = (full - y) / sig
    assert abs(z.mean()) < 0.15 and abs(z.std() - 1.0) < 0.1
    # the two deviates of a pair are uncorrelated
    assert abs(np.corrcoef(z[0:1000:2], z[1:1001:2])[0, 1]) < 0.15


def test_shot_noise_is_noisier_than_constant_snr(galaxy):
    eps_shot = pkm.noise.ShotNoise(galaxy).get_noisy_data(snr=100) - galaxy.ybar
    eps_const = pkm.noise.ConstantSNR(galaxy).get_noisy_data(snr=100) - galaxy.ybar
    assert np.median(np.abs(eps_shot)) > np.median(np.abs(eps_const))


def test_from_particle_histogram_is_bit_identical_to_numpy(gpu_ok):
    """Device 5-D histogram vs np.histogramdd: bit-exact binning, incl. samples on bin edges, on the
    outer edges, outliers and NaN (BASELINE.json: 'bit-exact for indexing and binning')."""
    cube = golden_cube("g2_rebin_nv21_prime")
    names = ("t", "v", "x1", "x2", "z")
    edges = [cube.get_variable_edges(k) for k in names]
    rng = np.random.default_rng(5)
    n = 200000
    t = rng.uniform(-0.5, 14.5, n)
    v = rng.normal(0.0, 400.0, n)
    x1, x2 = rng.uniform(-1.1, 1.1, n), rng.uniform(-1.0, 1.0, n)
    z = rng.uniform(-2.6, 0.5, n)
    # adversarial values: exactly on inner / first / last edges
    for arr, e in zip((t, v, x1, x2, z), edges):
        arr[:e.size] = e
        arr[e.size:2 * e.size] = np.nextafter(e, np.inf)
        arr[2 * e.size:3 * e.size] = np.nextafter(e, -np.inf)
    v[-3:] = [np.nan, np.inf, -np.inf]
    with contextlib.redirect_stdout(io.StringIO()) as out:
        comp = pkm.components.FromParticle(cube, t, v, x1, x2, z)
    assert "particles out of bounds" in out.getvalue()
    ok = np.isfinite(v)
    ref, _ = np.histogramdd([a[ok] for a in (t, v, x1, x2, z)], bins=edges, density=True)
    with np.errstate(divide="ignore"):
        assert np.array_equal(comp.log_p_tvxz, np.log(ref))
    counts = pkm.engine.histogram5d([t, v, x1, x2, z], edges)
    cref, _ = np.histogramdd([a[ok] for a in (t, v, x1, x2, z)], bins=edges)
    assert np.array_equal(counts, cref)
    # weighted: same bins, sums equal up to the order of additions
    w = rng.random(n)
    sums = pkm.engine.histogram5d([t, v, x1, x2, z], edges, weights=w)
    wref, _ = np.histogramdd([a[ok] for a in (t, v, x1, x2, z)], bins=edges, weights=w[ok])
    assert np.allclose(sums, wref, rtol=1e-13, atol=0)
    assert np.array_equal(sums == 0, wref == 0)
    # the reference's own check: p(t,z) of the component equals the 2-D histogram
    tz, _, _ = np.histogram2d(t[ok], z[ok], bins=[edges[0], edges[4]], density=True)
    inside = np.ones(n, dtype=bool)
    for arr, e in zip((t, v, x1, x2, z), edges):
        inside &= (arr >= e[0]) & (arr <= e[-1])
    tz_in, _, _ = np.histogram2d(t[inside], z[inside], bins=[edges[0], edges[4]], density=True)
    assert np.allclose(comp.get_p("tz"), tz_in, rtol=1e-10)


# ----------------------------------------------------------------------------------------------
# device-side pixelation of parametric components
# ----------------------------------------------------------------------------------------------
G2 = "g2_rebin_nv21_prime"


def test_device_pixelation_matches_reference(gpu_ok):
    """`disk.pixelate()` == the reference's Component(cube, disk.get_log_p('tvxz')): the 5-D array,
    the datacube, marginals and moments of the twin against outputs of the unmodified reference."""
    g = load_golden(G2)
    cube = golden_cube(G2)
    disk = BUILDERS["d1_"](cube)
    twin = disk.pixelate()
    assert twin.log_p_tvxz.is_cuda and tuple(twin.log_p_tvxz.shape) == g["b_log_p_tvxz"].shape
    got, ref = twin.log_p_tvxz.cpu().numpy(), g["b_log_p_tvxz"]
    assert not np.isnan(got).any()
    # empty (t,x,z) cells are -inf for every velocity on both sides
    empty = np.isneginf(ref.max(axis=1, keepdims=True))
    assert np.all(np.isneginf(got[np.broadcast_to(empty, ref.shape)]))
    # log domain: bins within 1e-4 of the cell's peak bin.  (Far in the upper tail the reference's
    # own 1 - exp(lo - hi) is quantised to 1.1e-16, so bins with P(v|t,x) ~ 1e-16 flip between a
    # finite value and -inf on one-ulp differences of the cdf; those are covered in the p domain.)
    bulk = ref > ref.max(axis=1, keepdims=True) + np.log(1e-4)
    assert bulk.sum() > 0.2 * ref.size
    assert np.max(np.abs(got[bulk] - ref[bulk])) < 1e-10
    assert rel_err(np.exp(got), np.exp(ref)) < 1e-13
    twin.evaluate_ybar()
    assert rel_err(twin.ybar, g["b_ybar"]) < TOL
    for dist in ("tv", "vxz", "v_x", "tz_x", "x"):
        for lw in (0, 1):
            ref_lp = g[f"b_logp|{dist}|{lw}|1"]
            got_lp = twin.get_log_p(dist, density=True, light_weighted=bool(lw))
            assert rel_err(np.exp(got_lp), np.exp(ref_lp)) < TOL, (dist, lw)
    assert np.allclose(twin.get_mean("v_x"), g["b_mean|v_x|0"], rtol=1e-9, atol=1e-9)
    assert np.allclose(twin.get_variance("v_x"), g["b_var|v_x|0"], rtol=1e-9)


def test_pixelation_kernel_edge_shapes_and_tails(gpu_ok):
    """Ragged pixel counts (not a multiple of the 32-pixel chunk), several nz / nv, scalar sigma,
    host output, probabilities (density=False), means far outside the velocity range and -inf
    weights: against the oracle (oracle/restatement.py::pixelate_gaussian, the reference's scipy
    log-cdf + signed logsumexp, pinned to the reference golden in tests/test_oracle.py)."""
    from oracle import restatement as R
    rng = np.random.default_rng(11)
    for (nt, nx1, nx2, nz, nv) in [(3, 3, 11, 5, 7), (2, 1, 1, 1, 1), (4, 8, 9, 12, 40), (1, 5, 13, 3, 301)]:
        v_edg = np.linspace(-900.0, 1100.0, nv + 1)
        log_w = np.log(rng.random((nt, nx1, nx2, nz)))
        log_w[rng.random(log_w.shape) < 0.1] = -np.inf
        mu = rng.normal(0.0, 400.0, (nt, nx1, nx2))
        mu.flat[0] = 9000.0                                        # every bin deep in the lower tail
        mu.flat[-1] = -20000.0                                     # upper tail: cdf differences vanish
        sig = rng.uniform(20.0, 300.0, (nt, nx1, nx2))
        for sig_in, dens in ((sig, True), (150.0, False)):
            got = pkm.engine.pixelate_gaussian(log_w, mu, sig_in, v_edg, density=dens, on_device=False)
            lpv = R.log_p_v_tx_gaussian(mu, sig_in, v_edg, density=dens)
            ref = R.pixelate_gaussian(log_w, mu, sig_in, v_edg, density=dens)
            assert got.shape == ref.shape and not np.isnan(got).any()
            # max-norm on probabilities; bins both sides call empty must agree where the weight is empty
            assert np.all(np.isneginf(got[np.isneginf(np.broadcast_to(log_w[:, None], ref.shape))]))
            scale = np.exp(ref[np.isfinite(ref)]).max()
            assert np.max(np.abs(np.exp(got) - np.exp(np.where(np.isnan(ref), -np.inf, ref)))) <= 1e-12 * scale
            # log domain: (i) bins within 1e-4 of the peak bin of their (t,x) velocity law, and
            # (ii) every bin below the mean, however deep in the lower tail (log p down to -1e5)
            # - there the reference's signed logsumexp is well conditioned.  Upper-tail bins with
            # P(v|t,x) ~ 1e-16 are quantised by 1 - exp(lo - hi) and only checked as probabilities.
            with np.errstate(invalid="ignore"):
                bulk = lpv > np.max(lpv, axis=0) + np.log(1e-4)
            below = np.broadcast_to(v_edg[1:, None, None, None] <= mu, lpv.shape)
            for mask in (bulk, below):
                m = np.broadcast_to(np.moveaxis(mask, 0, 1)[..., None], ref.shape) & np.isfinite(ref)
                assert m.any()
                assert np.all(np.abs(got[m] - ref[m]) < 1e-9 + 1e-12 * np.abs(ref[m]))


def test_pixelation_at_scale_marginalises_back(gpu_ok):
    """Size-independent property on a 5.1 GB device array (full grid, nv=101, 100x100 px): summing
    the pixelated density over velocity returns P(t,x,z) times the cdf mass inside the velocity
    range; the array never visits the host."""
    from scipy import special
    ssps = pkm.milesSSPs()
    cube = pkm.ifu_cube.IFUCube(ssps=ssps, nx1=100, nx2=100, nv=101, vrng=(-1000.0, 1000.0))
    disk = BUILDERS["d1_"](cube)
    twin = disk.pixelate()
    assert twin.log_p_tvxz.is_cuda and twin.log_p_tvxz.numel() == 53 * 101 * 100 * 100 * 12
    got = twin.get_log_p("txz", density=False)
    with np.errstate(divide="ignore"):
        lo = special.log_ndtr((cube.v_edg[0] - disk.mu_v) / disk.sig_v)
        hi = special.log_ndtr((cube.v_edg[-1] - disk.mu_v) / disk.sig_v)
        inside = np.log(np.exp(hi) - np.exp(lo))
    ref = disk.get_log_p("txz", density=False) + inside[..., None]
    assert rel_err(np.exp(got), np.exp(ref)) < TOL
    twin.evaluate_ybar()
    disk.evaluate_ybar()
    # path B on the pixelated twin vs the analytic path A (test_all.py:103-115 tolerance)
    err = np.abs(twin.ybar - disk.ybar) / np.abs(disk.ybar).max()
    assert np.median(err) < 5e-4


def test_mixture_pixelation_matches_reference(galaxy, base_component):
    """`galaxy.pixelate()` == the reference's Component(cube, galaxy.get_log_p('tvxz')) (golden g1):
    children pixelated on the device, collapsed by pkm_logsumexp_cubes."""
    g = load_golden(G1)
    twin = galaxy.pixelate()
    assert twin.log_p_tvxz.is_cuda
    got, ref = twin.log_p_tvxz.cpu().numpy(), g["b_log_p_tvxz"]
    assert got.shape == ref.shape and not np.isnan(got).any()
    bulk = ref > ref.max(axis=1, keepdims=True) + np.log(1e-4)
    assert bulk.sum() > 0.1 * ref.size
    assert np.max(np.abs(got[bulk] - ref[bulk])) < 1e-10
    assert rel_err(np.exp(got), np.exp(ref)) < 1e-13
    twin.evaluate_ybar()
    assert rel_err(twin.ybar, g["b_ybar"]) < TOL
    assert rel_err(twin.ybar, base_component.ybar) < 1e-12
    for dist in ("tz", "v_x"):
        assert rel_err(twin.get_p(dist), np.exp(g[f"b_logp|{dist}|0|1"])) < TOL


# ----------------------------------------------------------------------------- part 3: this round
def test_multi_output_reduction_and_derivation_from_cached_marginals(gpu_ok):
    """pkm_reduce_logp_multi / pkm_reduce_logp_from: nested requests are reduced from each other;
    every result must agree with the single-output pass over the 5-D array (and so with the
    reference goldens the single-output pass is pinned to)."""
    import torch
    from popkinmocks_b200 import engine
    name = "g2_rebin_nv21_prime"
    g = load_golden(name)
    cube = golden_cube(name)
    plan = engine.get_plan(cube)
    logp = torch.from_numpy(g["b_log_p_tvxz"]).cuda()
    log_dv = np.log(np.diff(cube.v_edg))
    keeps = ("vx", "txz", "x", "v", "tz", "z", "t", "tvxz")
    for lw in (None, cube.ssps.light_weights):
        dens = [i % 2 == 0 for i in range(len(keeps))]
        outs = plan.reduce_logp_multi(logp, log_dv, keeps, dens, light_weights=lw)
        for keep, d, out in zip(keeps, dens, outs):
            ref = plan.reduce_logp(logp, log_dv, keep, d, light_weights=lw).cpu().numpy()
            got = out.cpu().numpy()
            assert got.shape == ref.shape, keep
            assert np.array_equal(np.isneginf(got), np.isneginf(ref)), keep
            ok = np.isfinite(ref)
            assert np.allclose(got[ok], ref[ok], rtol=0, atol=1e-12), (keep, lw is not None)
    # light-weighted marginals without v follow from the MASS-weighted p(t,x,z)
    txz = plan.reduce_logp(logp, log_dv, "txz", False)
    npix = cube.nx * cube.ny
    for keep in ("tz", "x", "xz", "t"):
        got = plan.reduce_logp_from(txz, "txz", npix, log_dv, keep, True, light_weights=cube.ssps.light_weights)
        ref = plan.reduce_logp(logp, log_dv, keep, True, light_weights=cube.ssps.light_weights)
        ok = torch.isfinite(ref)
        assert torch.allclose(got[ok], ref[ok], rtol=0, atol=1e-12), keep
    # host arrays go through the same entry point
    outs = plan.reduce_logp_multi(g["b_log_p_tvxz"], log_dv, ("vx", "x"), (False, False))
    assert isinstance(outs[0], np.ndarray)
    assert np.allclose(outs[1], plan.reduce_logp(logp, log_dv, "x", False).cpu().numpy(), rtol=0, atol=1e-12)


def test_marginal_cache_is_dropped_with_the_array(gpu_ok):
    name = "g2_rebin_nv21_prime"
    g = load_golden(name)
    comp = pkm.components.Component(cube=golden_cube(name), log_p_tvxz=g["b_log_p_tvxz"].copy())
    p1 = comp.get_p("v_x", density=False)
    s1 = comp.get_skewness("v_x")
    assert ("vx", False, 0) in comp.__dict__["_pkm_marg"] and ("x", False, 0) in comp.__dict__["_pkm_marg"]
    shifted = g["b_log_p_tvxz"].copy()
    shifted[:, :5] = -np.inf                     # remove the five lowest velocity bins
    comp.log_p_tvxz = shifted                    # rebinding drops the device copy and every cache
    assert "_pkm_marg" not in comp.__dict__ and "_pkm_mom" not in comp.__dict__
    p2 = comp.get_p("v_x", density=False)
    assert np.all(p2[:5] == 0.0) and not np.allclose(p1, p2, equal_nan=True)
    comp.log_p_tvxz[:, 5:8] = -np.inf            # in-place edit: needs an explicit invalidation
    comp.invalidate_cache()
    p3 = comp.get_p("v_x", density=False)
    assert np.all(p3[:8] == 0.0)
    with np.errstate(invalid="ignore"):
        assert not np.allclose(comp.get_skewness("v_x"), s1, equal_nan=True)


def test_nan_and_empty_densities_through_both_coefficient_kernels(gpu_ok, monkeypatch):
    """-inf log densities are exact zeros and a NaN poisons its own pixel only (np.exp semantics);
    the TMA/tensor-core coefficient kernel and the staged fallback agree."""
    from popkinmocks_b200 import engine
    name = "g3_full_nv101_odd"
    g = load_golden(name)
    cube = build_cube(name)
    logp = np.repeat(g["wn_log_p_tvxz"], 20, axis=3)           # 1 x 40 pixels
    cube = pkm.ifu_cube.IFUCube(ssps=cube.ssps, nx1=1, nx2=40, nv=cube.nv, vrng=(-1000.0, 1000.0))
    logp[:, :, 0, 3, :] = -np.inf                              # an empty pixel
    logp[7, 50, 0, 11, 2] = np.nan
    logp[:, 10:20, 0, 17, :] = -np.inf
    res = {}
    for kern in ("1", "0"):
        monkeypatch.setenv("PKM_COEFF_KERNEL", kern)
        plan = engine.Plan(cube)
        res[kern] = plan.ybar_density(logp, is_log=True)
        plan.close()
    for y in res.values():
        assert np.all(y[:, 0, 3] == 0.0)
        assert np.all(np.isnan(y[:, 0, 11]))
        assert np.isfinite(np.delete(y, 11, axis=2)).all()
    ok = np.delete(np.arange(40), 11)
    assert rel_err(res["1"][:, :, ok], res["0"][:, :, ok]) < 1e-13
    from oracle import restatement as R
    s = cube.ssps
    with np.errstate(invalid="ignore"):
        ref = R.ybar_density_literal(np.exp(logp[:, :, :, 16:19]), cube.v_edg, s.FXw, s.par_dims, s.n_fft, s.dv,
                                     s.delta_t, s.delta_z, cube.dx, cube.dy)
    assert rel_err(res["1"][:, :, 16:19], ref) < TOL


@pytest.mark.parametrize("tag", ["d1_", "d2_", "s1_"])
def test_device_model_maps_match_the_host_maps(gpu_ok, tag):
    """pkm_disk_maps / pkm_stream_maps / pkm_assemble_txz: the O(pixels) inputs of the Gaussian
    path generated in HBM from the model parameters agree with the host (reference-formula) arrays
    to <= 1e-12, and with the reference's own arrays in the golden file."""
    cube = pkm.ifu_cube.IFUCube(ssps=golden_cube(G1).ssps, nx1=13, nx2=9, nv=41, vrng=(-1000.0, 1000.0))
    comp = BUILDERS[tag](cube)
    maps = comp._device_maps(0)
    assert maps is not None
    for lazy in ("mu_v", "sig_v", "log_p_x_t", "log_p_z_tx", "t_dep"):
        assert lazy not in comp.__dict__, lazy             # nothing of size nx1*nx2 was built on the host
    mu = np.broadcast_to(maps["mu_v"].cpu().numpy(), comp.mu_v.shape)
    sg = np.broadcast_to(maps["sig_v"].cpu().numpy(), comp.sig_v.shape)
    assert np.max(np.abs(mu - comp.mu_v)) < 1e-12 * max(1.0, np.max(np.abs(comp.mu_v)))
    assert np.max(np.abs(sg - comp.sig_v)) < 1e-12 * np.max(np.abs(comp.sig_v))
    P_host = comp.get_p("txz", density=False)
    assert rel_err(maps["P_txz"].cpu().numpy(), P_host) < 1e-12
    logp_host = comp.get_log_p("txz", density=True)
    # log domain on the bulk; far from a stream the pixel masses are differences of nearly equal
    # cdfs, where one ulp of log_ndtr is a 1e-8 relative change of a 1e-30 probability
    ok = np.isfinite(logp_host) & (logp_host > np.max(logp_host) - 25.0)
    # (pixels on the far side of every stream sample integrate the UPPER tail of the normal cdf,
    # where the reference's own log-cdf difference carries ~1e-7 relative noise: looser there)
    assert np.allclose(maps["log_p_txz"].cpu().numpy()[ok], logp_host[ok], rtol=0,
                       atol=1e-10 if tag != "s1_" else 1e-6)
    assert rel_err(np.exp(maps["log_p_txz"].cpu().numpy()), np.exp(logp_host)) < 1e-12
    if maps["t_dep"] is not None:
        assert np.max(np.abs(maps["t_dep"].cpu().numpy() - comp.t_dep)) < 1e-13
        grid = comp.z_t_interp_grid["t_dep"]
        nearest = np.argmin(np.abs(comp.t_dep[:, :, None] - grid), axis=-1)
        got = maps["row_idx"].cpu().numpy() + int(maps["rows"][0])
        # a pixel exactly half way between two table rows may round either way after one ulp in t_dep
        differ = got != nearest
        assert differ.mean() < 0.02
        assert np.all(np.abs(np.abs(comp.t_dep[differ] - grid[got[differ]])
                             - np.abs(comp.t_dep[differ] - grid[nearest[differ]])) < 1e-12)
    # and the hot path fed by them reproduces the reference cube
    g = load_golden(G1)
    small = BUILDERS[tag](golden_cube(G1))
    small.evaluate_ybar()
    assert "mu_v" not in small.__dict__
    assert rel_err(small.ybar, g[tag + "ybar"]) < TOL
    tw = small.pixelate()
    ref = small.get_log_p("tvxz", density=True)
    got = tw.log_p_tvxz.cpu().numpy()
    # bins within 1e-4 of the cell's peak bin in the log domain, everything in the p domain (the far
    # upper tail of the reference's signed logsumexp is quantised, see the pixelation test above)
    with np.errstate(invalid="ignore"):
        bulk = ref > ref.max(axis=1, keepdims=True) + np.log(1e-4)
    assert np.max(np.abs(got[bulk] - ref[bulk])) < 1e-10
    assert rel_err(np.exp(got), np.exp(ref)) < 1e-13
