"""The CPU oracle (oracle/restatement.py) against outputs of the reference itself.

tests/golden/*.npz were written by oracle/make_golden.py executing the UNMODIFIED reference
under oracle/ref_shim.py.  These tests pin every oracle function; they run without a GPU.
"""
import numpy as np
import pytest
from scipy import stats

from conftest import load_golden, rel_err
from oracle import restatement as R

GOLDENS = ["g1_conftest_full_nv41", "g2_rebin_nv21_prime", "g3_full_nv101_odd", "g4_even_nv40"]


def _fxw(name):
    """Full FXw is not stored in the golden files (size); rebuild it with the product's host
    classes, which test_host_models.py pins bit-for-bit against the stored FXw rows."""
    from conftest import golden_cube
    return golden_cube(name).ssps.FXw


def _grid(g):
    return dict(delta_t=g["delta_t"], dv_bins=np.diff(g["v_edg"]), dx=float(g["dx"]), dy=float(g["dy"]),
                delta_z=g["delta_z"], light_weights=g["light_weights"])


@pytest.mark.parametrize("name,tags", [("g1_conftest_full_nv41", ["d1_", "d2_", "s1_"]),
                                       ("g2_rebin_nv21_prime", ["d1_", "s1_"]),
                                       ("g3_full_nv101_odd", ["d1_"]), ("g4_even_nv40", ["d2_"]),
                                       ("g5_ppxf_pad_nv21", ["d1_"])])
def test_gaussian_path(name, tags):
    g = load_golden(name)
    for tag in tags:
        y = R.ybar_gaussian(g[tag + "P_txz"], g[tag + "mu_v"], g[tag + "sig_v"], _fxw(name),
                            tuple(g["par_dims"]), int(g["n_fft"]), float(g["dv"]))
        assert rel_err(y, g[tag + "ybar"]) < 1e-13


@pytest.mark.parametrize("name,tag", [("g1_conftest_full_nv41", "b_"), ("g2_rebin_nv21_prime", "b_"),
                                      ("g2_rebin_nv21_prime", "wn_"), ("g3_full_nv101_odd", "wn_"),
                                      ("g5_ppxf_pad_nv21", "wn_")])
def test_density_path_literal_and_fused(name, tag):
    g = load_golden(name)
    with np.errstate(over="ignore"):
        p = np.exp(g[tag + "log_p_tvxz"])
    args = (p, g["v_edg"], _fxw(name), tuple(g["par_dims"]), int(g["n_fft"]), float(g["dv"]),
            g["delta_t"], g["delta_z"], float(g["dx"]), float(g["dy"]))
    ref = g[tag + "ybar"]
    assert rel_err(R.ybar_density_literal(*args, pixel_block=1), ref) < 1e-14
    assert rel_err(R.ybar_density_fused(*args), ref) < 1e-13


def test_density_path_even_nv_raises():
    g = load_golden("g4_even_nv40")
    assert str(g["b_error"]) == "IndexError"
    with pytest.raises(IndexError):
        R.validate_v_edg(g["v_edg"], float(g["dv"]))


def test_mixture_sum():
    g = load_golden("g1_conftest_full_nv41")
    y = R.ybar_mixture([g["d1_ybar"], g["d2_ybar"], g["s1_ybar"]], g["mix_weights"])
    assert np.array_equal(y, g["mix_ybar"])


@pytest.mark.parametrize("name", ["g1_conftest_full_nv41", "g2_rebin_nv21_prime"])
def test_log_distributions(name):
    g = load_golden(name)
    grid = _grid(g)
    keys = [k for k in g.files if k.startswith("b_logp|")]
    assert len(keys) >= 80
    for k in keys:
        _, dist, lw, dens = k.split("|")
        out = R.log_distribution(g["b_log_p_tvxz"], dist, grid, density=bool(int(dens)),
                                 light_weighted=bool(int(lw)))
        ref = g[k]
        assert out.shape == ref.shape, k
        both = np.isfinite(ref)
        assert np.array_equal(np.isnan(out), np.isnan(ref)), k
        assert np.allclose(out[both], ref[both], rtol=0, atol=1e-10), k


@pytest.mark.parametrize("name", ["g1_conftest_full_nv41", "g2_rebin_nv21_prime"])
def test_moments(name):
    g = load_golden(name)
    grid = _grid(g)
    v = (g["v_edg"][:-1] + g["v_edg"][1:]) / 2
    values = {"t": g["t_cents"], "v": v, "z": g["z_cents"]}
    for k in [k for k in g.files if k.startswith("b_mean|")]:
        _, dist, lw = k.split("|")
        var = dist.split("_")[0]
        if var == "x":
            continue
        P = np.exp(R.log_distribution(g["b_log_p_tvxz"], dist, grid, density=False, light_weighted=bool(int(lw))))
        ok = np.isfinite(g[k])
        assert np.allclose(R.mean_1d(P, values[var])[ok], g[k][ok], rtol=1e-11, atol=1e-11), k
        for j, nm in ((3, "skew"), (4, "kurt")):
            ref = g[f"b_{nm}|{dist}|{lw}"]
            ok = np.isfinite(ref)
            out = R.standardised_moment_1d(P, values[var], j)
            assert np.allclose(out[ok], ref[ok], rtol=1e-8, atol=1e-8), (nm, k)


def test_gaussian_pixelation():
    """oracle pixelate_gaussian vs the reference's disk.get_log_p('tvxz') (golden g2)."""
    g = load_golden("g2_rebin_nv21_prime")
    grid = _grid(g)
    log_dV = np.log(R.volume_element("txz", grid["delta_t"], grid["dv_bins"], grid["dx"], grid["dy"],
                                     grid["delta_z"]))
    with np.errstate(divide="ignore"):
        log_P_txz = np.log(g["d1_P_txz"])
    ref = g["b_log_p_tvxz"]
    got = R.pixelate_gaussian(log_P_txz - log_dV, g["d1_mu_v"], g["d1_sig_v"], g["v_edg"], density=True)
    assert got.shape == ref.shape
    assert np.array_equal(np.isfinite(got), np.isfinite(ref))
    fin = np.isfinite(ref)
    assert np.max(np.abs(got[fin] - ref[fin])) < 1e-11
    # probabilities: log P(t,v,x,z) sums back to P(t,x,z) times the cdf mass inside the velocity range
    logP = R.pixelate_gaussian(log_P_txz, g["d1_mu_v"], g["d1_sig_v"], g["v_edg"], density=False)
    inside = stats.norm(g["d1_mu_v"], g["d1_sig_v"]).cdf(g["v_edg"][-1]) - \
        stats.norm(g["d1_mu_v"], g["d1_sig_v"]).cdf(g["v_edg"][0])
    assert np.allclose(np.exp(logP).sum(1), g["d1_P_txz"] * inside[..., None], rtol=1e-12, atol=0)


def test_noise_sigma():
    g = load_golden("g1_conftest_full_nv41")
    assert np.allclose(R.sigma_shot_noise(g["b_ybar"], 50.0), g["b_sigma_shot_snr50"], rtol=1e-15, equal_nan=True)
    assert np.array_equal(R.sigma_constant_snr(g["b_ybar"], 50.0), g["b_sigma_const_snr50"])
