"""CPU restatement (NumPy/SciPy, float64) of popkinmocks' datacube hot path.

TEST INFRASTRUCTURE.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
cpu_baseline / `--impl reference` legs may import this module, and only as the
checker or the reported CPU baseline; the product package never imports it.

Parity status: PINNED.  The reference publishes no golden vectors
(SURVEY.md section 8c), so the functions below are pinned against outputs of the
reference itself, executed in the build container under `oracle/ref_shim.py`:
`oracle/make_golden.py` wrote `tests/golden/*.npz` (inputs + reference outputs)
and `tests/test_oracle.py` checks every function here against them.

Third-party arithmetic the reference relies on (not vendored, unpinned in its
requirements.txt; versions in the build container: numpy 2.3.5, scipy 1.18.1):
numpy.fft.rfft/irfft (pocketfft), scipy.interpolate.interp1d(kind="cubic")
(= make_interp_spline(k=3), not-a-knot), scipy.special.logsumexp.  The same calls
are used here, so these functions need numpy+scipy only and travel to the GPU box.
"""
import numpy as np
from scipy import interpolate, special, stats


# ----------------------------------------------------------------------------
# path B: Component.evaluate_ybar        /root/reference/popkinmocks/components/base.py:789-882
# ----------------------------------------------------------------------------
def validate_v_edg(v_edg, dv_ssps):
    """Velocity-grid checks, verbatim semantics of base.py:810-835.

    Returns v0_idx.  Raises ValueError (bad spacing / order / centring) or
    IndexError (no velocity bin centred on zero, base.py:822-823)."""
    v_edg = np.asarray(v_edg)
    msg = "Velocity array must be "
    ok_spacing = np.allclose(v_edg[1:] - v_edg[:-1], dv_ssps)
    if not ok_spacing:
        msg += "(i) uniformly spaced with dv = SSP resolution"
    v = (v_edg[:-1] + v_edg[1:]) / 2.0
    ok_order = np.all(np.sort(v) == v)
    if not ok_order:
        msg += "(ii) ascending"
    v0_idx = np.where(np.isclose(v, 0.0))[0][0]
    ok_centre = v0_idx == ((v.size - 1) / 2 if v.size % 2 == 1 else v.size / 2)
    if not ok_centre:
        msg += "(iii) zero-centered."
    if not (ok_spacing and ok_order and ok_centre):
        raise ValueError(msg)
    return int(v0_idx)


def ybar_density_literal(p_tvxz, v_edg, FXw, par_dims, n_fft, dv, delta_t, delta_z,
                         dx, dy, pixel_block=None):
    """Path B exactly as the reference orders it (base.py:836-859).

    roll -> rfft(v)*dv -> cubic interp1d in Fourier space -> times S-hat ->
    irfft PER (t,z) -> sum over (t,z) with dt*dz -> times dx*dy.
    `pixel_block` (int) bounds memory by looping over blocks of x1 rows; pixels
    are independent so this does not change any value (base.py:860-879 does the same).
    Returns ybar [n_fft, nx1, nx2] float64.
    """
    p_tvxz = np.asarray(p_tvxz, dtype=np.float64)
    nt, nv, nx1, nx2, nz = p_tvxz.shape
    v0_idx = validate_v_edg(v_edg, dv)
    F_s = np.moveaxis(np.reshape(FXw, (-1,) + tuple(par_dims)), -1, 0)   # [nt, nfo, nz]
    nfo = F_s.shape[1]
    na = np.newaxis
    dtz = delta_t[:, na, na, na, na] * delta_z[na, na, na, na, :]
    ybar = np.empty((n_fft, nx1, nx2), dtype=np.float64)
    step = nx1 if pixel_block is None else int(pixel_block)
    for r0 in range(0, nx1, step):
        p = np.roll(p_tvxz[:, :, r0:r0 + step], nv - v0_idx, axis=1)
        F_p = np.fft.rfft(p, axis=1) * dv
        resample = interpolate.interp1d(np.linspace(0, 1, F_p.shape[1]), F_p, axis=1,
                                        kind="cubic", bounds_error=True)
        F_p = resample(np.linspace(0, 1, nfo))
        y = np.fft.irfft(F_s[:, :, na, na, :] * F_p, n_fft, axis=1)
        ybar[:, r0:r0 + step] = np.sum(y * dtz, (0, 4)) * dx * dy
    return ybar


def spline_resampling_matrix(nfi, nfo):
    """Dense matrix W [nfo, nfi] of the reference's Fourier-space resampling.

    interp1d(linspace(0,1,nfi), ., kind="cubic")(linspace(0,1,nfo)) is linear in
    its data; W is that operator (base.py:846-853)."""
    eye = np.eye(nfi)
    f = interpolate.interp1d(np.linspace(0, 1, nfi), eye, axis=0, kind="cubic",
                             bounds_error=True)
    return f(np.linspace(0, 1, nfo))


def ybar_density_fused(p_tvxz, v_edg, FXw, par_dims, n_fft, dv, delta_t, delta_z, dx, dy):
    """Path B restructured the way the CUDA kernels compute it.

    Sum over (t,z) in Fourier space, ONE irfft per pixel, spline as the linear
    operator W composed with the rolled nv-point DFT.  Equal to the literal form
    by linearity (to ~1e-15 max-norm)."""
    p_tvxz = np.asarray(p_tvxz, dtype=np.float64)
    nt, nv, nx1, nx2, nz = p_tvxz.shape
    v0_idx = validate_v_edg(v_edg, dv)
    S = np.reshape(FXw, (-1,) + tuple(par_dims))                         # [nfo, nz, nt]
    nfo, nfi = S.shape[0], nv // 2 + 1
    Sp = S * delta_z[None, :, None] * delta_t[None, None, :]
    u = (np.arange(nv) - v0_idx) % nv
    j = np.arange(nfi)
    D = dv * np.exp(-2j * np.pi * ((j[:, None] * u[None, :]) % nv) / nv)  # [nfi, nv]
    M = spline_resampling_matrix(nfi, nfo) @ D                           # [nfo, nv]
    Y = np.einsum("kzt,kv,tvxyz->kxy", Sp, M, p_tvxz, optimize=True) * dx * dy
    return np.fft.irfft(Y, n_fft, axis=0)


# ----------------------------------------------------------------------------
# path A: ParametricComponent.evaluate_ybar   .../components/parametric.py:335-407
# ----------------------------------------------------------------------------
def gaussian_omega(nfo, dv):
    """omega_k = linspace(0, pi, nfo)/dv  (parametric.py:360-362) - NOT 2*pi*k/(n_fft*dv)."""
    omega = np.linspace(0, np.pi, nfo)
    omega /= dv
    return omega


def ybar_gaussian(P_txz, mu_v, sig_v, FXw, par_dims, n_fft, dv, pixel_block=None):
    """Path A (parametric.py:355-373): analytic Fourier transform of a Gaussian LOSVD.

    P_txz [nt,nx1,nx2,nz] volume-weighted probabilities, mu_v/sig_v [nt,nx1,nx2].
    Returns ybar [n_fft, nx1, nx2]."""
    P_txz = np.asarray(P_txz, dtype=np.float64)
    nt, nx1, nx2, nz = P_txz.shape
    mu_v = np.broadcast_to(mu_v, (nt, nx1, nx2))
    sig_v = np.broadcast_to(sig_v, (nt, nx1, nx2))
    F_s = np.reshape(FXw, (-1,) + tuple(par_dims))                       # [nfo, nz, nt]
    nfo = F_s.shape[0]
    omega = gaussian_omega(nfo, dv)[:, None, None, None]
    ybar = np.empty((n_fft, nx1, nx2), dtype=np.float64)
    step = nx1 if pixel_block is None else int(pixel_block)
    for r0 in range(0, nx1, step):
        sl = slice(r0, r0 + step)
        F_v = np.exp(-1j * mu_v[:, sl] * omega - 0.5 * (sig_v[:, sl] * omega) ** 2)
        F_y = np.einsum("txyz,wtxy,wzt->wxy", P_txz[:, sl], F_v, F_s, optimize=True)
        ybar[:, sl] = np.fft.irfft(F_y, n_fft, axis=0)
    return ybar


# ----------------------------------------------------------------------------
# Mixture.evaluate_ybar                     .../components/mixture.py:35-48
# ----------------------------------------------------------------------------
def ybar_mixture(ybars, weights):
    out = np.zeros_like(ybars[0])
    for y, w in zip(ybars, weights):
        out += w * y
    return out


# ----------------------------------------------------------------------------
# get_p marginals / conditionals of the pixelated density   base.py:33-177, 884-1174
# ----------------------------------------------------------------------------
_AXIS_OF = {"t": (0,), "v": (1,), "x": (2, 3), "z": (4,)}


def volume_element(which, delta_t, dv_bins, dx, dy, delta_z):
    """dV for the dependent variables of a distribution string (ifu_cube.py:52-95)."""
    dep, _, given = which.partition("_")
    axes = dep.replace("x", "xy")
    widths = {"t": delta_t, "v": dv_bins, "x": np.array([dx]), "y": np.array([dy]),
              "z": delta_z}
    dvol = np.ones((1,) * len(axes))
    for pos, a in enumerate(axes):
        shape = [1] * len(axes)
        shape[pos] = -1
        dvol = dvol * widths[a].reshape(shape)
    return dvol.reshape(dvol.shape + (1,) * len(given.replace("x", "xy")))


def log_P_tvxz(log_p_tvxz, grid, light_weighted=False, density=False):
    """base.py:1017-1041.  `grid` = dict(delta_t, dv_bins, dx, dy, delta_z, light_weights)."""
    log_dV = np.log(volume_element("tvxz", grid["delta_t"], grid["dv_bins"], grid["dx"],
                                   grid["dy"], grid["delta_z"]))
    out = np.array(log_p_tvxz, dtype=np.float64, copy=True)
    out += log_dV
    if light_weighted:
        out += np.log(grid["light_weights"][:, None, None, None, :])
        out -= special.logsumexp(out)
    if density:
        out -= log_dV
    return out


def log_marginal(log_p_tvxz, which, grid, density=True, light_weighted=False):
    """log p(which) for an un-conditioned, alphabetically ordered `which` (base.py:884-1174)."""
    logP = log_P_tvxz(log_p_tvxz, grid, light_weighted=light_weighted, density=False)
    drop = tuple(ax for name, axs in _AXIS_OF.items() if name not in which for ax in axs)
    out = special.logsumexp(logP, drop) if drop else logP
    if density:
        out = out - np.log(volume_element(which, grid["delta_t"], grid["dv_bins"],
                                          grid["dx"], grid["dy"], grid["delta_z"]))
    return out


def log_distribution(log_p_tvxz, which, grid, density=True, light_weighted=False):
    """Marginal or conditional (`dep_given`) log distribution (base.py:67-177)."""
    if "_" not in which:
        return log_marginal(log_p_tvxz, which, grid, density, light_weighted)
    dep, given = which.split("_")
    joint = "".join(sorted(dep + given))
    lj = log_marginal(log_p_tvxz, joint, grid, False, light_weighted)
    lm = log_marginal(log_p_tvxz, given, grid, False, light_weighted)
    joint_axes = joint.replace("x", "xy")
    given_axes = given.replace("x", "xy")
    src = [joint_axes.find(g) for g in given_axes]
    dst = list(range(-len(given_axes), 0))
    with np.errstate(invalid="ignore"):
        out = np.moveaxis(lj, src, dst) - lm
    if density:
        out = out - np.log(volume_element(which, grid["delta_t"], grid["dv_bins"],
                                          grid["dx"], grid["dy"], grid["delta_z"]))
    return out


# ----------------------------------------------------------------------------
# moments of 1-D (conditional) distributions                 base.py:179-452
# ----------------------------------------------------------------------------
def mean_1d(P, values):
    """sum_i var_i P(var_i | ...) along axis 0 (base.py:208-213), var in (t, v, z)."""
    return np.sum((values.T * P.T).T, 0)


def noncentral_moment_1d(P, values, j, mu):
    """sum_i (var_i - mu)^j P(var_i | ...) along axis 0 (base.py:255-260)."""
    mu = np.asarray(mu)
    return np.sum((np.power(values - mu[np.newaxis].T, j) * P.T).T, 0)


def standardised_moment_1d(P, values, j):
    """central moment j / dispersion^j (base.py:354-377): skewness j=3, kurtosis j=4."""
    mu = mean_1d(P, values)
    with np.errstate(invalid="ignore", divide="ignore"):
        return noncentral_moment_1d(P, values, j, mu) / noncentral_moment_1d(P, values, 2, mu) ** (0.5 * j)


# ----------------------------------------------------------------------------
# analytic pixelation of a Gaussian-LOSVD component
#                       /root/reference/popkinmocks/components/parametric.py:180-211, 567-596
# ----------------------------------------------------------------------------
def log_p_v_tx_gaussian(mu_v, sig_v, v_edg, density=True):
    """log p(v-bin | t,x): normal log-cdf at the bin edges, signed logsumexp of neighbouring
    edges (parametric.py:193-200), minus log dv for a density (parametric.py:208-210).
    mu_v, sig_v broadcastable to [t, x1, x2]; returns [v, t, x1, x2]."""
    v_edg = np.asarray(v_edg, dtype=np.float64)[:, None, None, None]
    with np.errstate(divide="ignore", invalid="ignore"):
        norm = stats.norm(loc=mu_v, scale=sig_v)
        log_cdf_0 = norm.logcdf(v_edg[:-1])
        log_cdf_1 = norm.logcdf(v_edg[1:])
        out = special.logsumexp(np.array([log_cdf_1, log_cdf_0]).T, -1, b=[1, -1]).T
    if density:
        out = (out.T - np.log(np.diff(v_edg[:, 0, 0, 0]))).T
    return out


def pixelate_gaussian(log_p_txz, mu_v, sig_v, v_edg, density=True):
    """log p(t,v,x,z) = log p(t,x,z) + log p(v|t,x)  (parametric.py:589-592, mass weighted).
    `log_p_txz` [t, x1, x2, z] is the density or the volume-weighted probability, matching
    `density`."""
    log_p_v_tx = log_p_v_tx_gaussian(mu_v, sig_v, v_edg, density=density)
    return np.moveaxis(log_p_v_tx[..., None], 0, 1) + np.asarray(log_p_txz)[:, None]


# ----------------------------------------------------------------------------
# noise                                                     /root/reference/popkinmocks/noise.py:41-71
# ----------------------------------------------------------------------------
def sigma_constant_snr(ybar, snr=100.0):
    return ybar / snr


def sigma_shot_noise(ybar, snr=100.0):
    with np.errstate(invalid="ignore"):
        return (np.max(ybar) ** 0.5 / snr) * ybar ** 0.5
