"""`Component`: a galaxy component given by its pixelated log density log p(t, v, x1, x2, z).

Drop-in for the reference class (/root/reference/popkinmocks/components/base.py):
same constructor, `get_p / get_log_p`, `get_mean / get_variance / get_skewness / get_kurtosis /
...`, `evaluate_ybar(batch=)`.  The arithmetic is delegated to the CUDA library:

  * `evaluate_ybar`            -> pkm_ybar_density   (base.py:789-882)
  * every `_get_log_p_<dist>`  -> pkm_reduce_logp    (base.py:884-1174): one streaming
    log-sum-exp reduction of the 5-D array instead of three full-size temporaries per call
  * moments                    -> pkm_moments        (base.py:179-452)

Validation and error behaviour stay on the host, verbatim in meaning: shape `ValueError`
(base.py:23-31), velocity-grid `ValueError` / `IndexError` (base.py:810-835),
`ValueError("Unknown option for batch")` (base.py:881).
"""
import os

import numpy as np

from .. import engine

_MARGINALS = ("t", "tv", "tx", "tz", "tvx", "tvz", "txz", "tvxz", "v", "vx", "vz", "vxz", "x", "xz", "z")


def _double_factorial_odd(n):
    """(n)!! for odd n >= -1, with (-1)!! = 1 (the convention the moment formulas need)."""
    out = 1.0
    while n > 1:
        out *= n
        n -= 2
    return out


def check_velocity_grid(cube):
    """Velocity-grid requirements of the FFT datacube path (base.py:810-835).

    Returns the index of the zero-centred bin; raises IndexError when no bin is centred on
    zero and ValueError when the spacing, ordering or centring is wrong."""
    v_edg = cube.v_edg
    problems = []
    if not np.allclose(np.diff(v_edg), cube.ssps.dv):
        problems.append("(i) uniformly spaced with dv = SSP resolution")
    v = (v_edg[:-1] + v_edg[1:]) / 2.0
    if not np.all(np.sort(v) == v):
        problems.append("(ii) ascending")
    v0_idx = np.where(np.isclose(v, 0.0))[0][0]      # IndexError if there is no such bin
    expected = (v.size - 1) / 2 if v.size % 2 == 1 else v.size / 2
    if v0_idx != expected:
        problems.append("(iii) zero-centered.")
    if problems:
        raise ValueError("Velocity array must be " + "".join(problems))
    return int(v0_idx)


class Component(object):
    """A galaxy component specified by the (log) joint density p(t,v,x,z).

    Args:
        cube: a `popkinmocks_b200.ifu_cube.IFUCube`
        log_p_tvxz (array): 5-D natural log of the mass-weighted probability density
            log p(t, v, x1, x2, z)
    """

    def __init__(self, cube=None, log_p_tvxz=None):
        self.cube = cube
        shape_tvxz = self.cube.get_distribution_shape("tvxz")
        if tuple(log_p_tvxz.shape) != shape_tvxz:
            raise ValueError(f"`log_p_tvxz` has shape {tuple(log_p_tvxz.shape)} "
                             f"but should have shape {shape_tvxz}")
        self.log_p_tvxz = log_p_tvxz

    # ------------------------------------------------------------------ distributions
    def get_p(self, which_dist, density=True, light_weighted=False):
        """Marginal or conditional distribution named by `which_dist` ('tv', 'tz_x', ...).

        Variables are listed alphabetically on each side of the underscore; `density=False`
        returns volume-element weighted probabilities; `light_weighted` switches from mass
        to light weighting.  Array axes follow the order of the string ('x' is two axes)."""
        return np.exp(self.get_log_p(which_dist, density=density, light_weighted=light_weighted))

    def get_log_p(self, which_dist, density=True, light_weighted=False):
        if "_" in which_dist:
            return self._get_log_conditional_distribution(which_dist, density=density,
                                                          light_weighted=light_weighted)
        return self._get_log_marginal_distribution(which_dist, density=density,
                                                   light_weighted=light_weighted)

    def _get_log_marginal_distribution(self, which_dist, density=True, light_weighted=False):
        fn = getattr(self, "_get_log_p_" + which_dist)
        return fn(density=density, light_weighted=light_weighted)

    def _get_log_conditional_distribution(self, which_dist, light_weighted=False, density=True):
        """log p(dep | given) = log P(joint) - log P(given) [- log dV]  (base.py:130-177).

        A hard-coded `_get_log_p_<dep>_<given>` method takes precedence."""
        dependent, given = which_dist.split("_")
        hard_coded = getattr(self, "_get_log_p_" + which_dist, None)
        if hard_coded is not None:
            try:
                return hard_coded(density=density, light_weighted=light_weighted)
            except NotImplementedError:
                pass
        joint = "".join(sorted(dependent + given))
        log_joint = self.get_log_p(joint, density=False, light_weighted=light_weighted)
        log_given = self.get_log_p(given, density=False, light_weighted=light_weighted)
        joint_axes = joint.replace("x", "xy")
        given_axes = given.replace("x", "xy")
        source = [joint_axes.find(a) for a in given_axes]
        target = list(range(-len(given_axes), 0))
        with np.errstate(invalid="ignore"):
            out = np.moveaxis(log_joint, source, target) - log_given   # nan where P(given) = 0
        if density:
            out = out - np.log(self.cube.construct_volume_element(which_dist))
        return out

    # device-resident copy of log_p_tvxz, shared by evaluate_ybar / get_p / moments so that a
    # sequence of calls uploads the 5-D array once (the reference re-derives it per call)
    _DEVICE_CACHE_LIMIT = 64 << 30

    def _device_log_p(self, device=0):
        cached = self.__dict__.get("_pkm_dev")
        src = self.log_p_tvxz
        if cached is not None and cached[0] is src and cached[1] == device:
            return cached[2]
        if engine._is_cuda_tensor(src):
            return src
        engine._lib.require_gpu()
        if src.nbytes > self._DEVICE_CACHE_LIMIT:
            return src                       # too large to keep resident: stream it per call
        import torch
        dev = torch.from_numpy(np.ascontiguousarray(src, dtype=np.float64)).to(f"cuda:{device}")
        self.__dict__["_pkm_dev"] = (src, device, dev)
        return dev

    def __getstate__(self):
        state = dict(self.__dict__)
        state.pop("_pkm_dev", None)          # never pickle device handles (dill_dump must keep working)
        return state

    def _log_marginal(self, keep, density=True, light_weighted=False):
        """log-marginal over the axes not in `keep`, on the GPU (pkm_reduce_logp)."""
        cube = self.cube
        plan = engine.get_plan(cube)
        log_dv = np.log(cube.v_edg[1:] - cube.v_edg[:-1])
        lw = cube.ssps.light_weights if light_weighted else None
        out = plan.reduce_logp(self._device_log_p(), log_dv, keep, density, light_weights=lw)
        return out.cpu().numpy() if engine._is_torch(out) else out

    # ------------------------------------------------------------------ moments
    def _moment_axis(self, which_dist):
        variable = which_dist.split("_")[0]
        assert len(variable) == 1
        return variable

    def _raw_moments(self, which_dist, light_weighted):
        """[mean, mu2, mu3, mu4] of a 1-D (conditional) distribution of t, v or z on the GPU."""
        variable = self._moment_axis(which_dist)
        P = self.get_p(which_dist, light_weighted=light_weighted, density=False)
        values = self.cube.get_variable_values(variable)
        return engine.moments(P, values)

    def get_mean(self, which_dist, light_weighted=False):
        """Mean of a 1-D distribution, e.g. 'v' -> E(v), 'v_x' -> E(v|x).

        For 'x' the result has a leading axis of size 2 (x1, x2) (base.py:179-231)."""
        variable = self._moment_axis(which_dist)
        hard_coded = getattr(self, "get_mean_" + which_dist, None)
        if hard_coded is not None:
            return hard_coded(light_weighted=light_weighted)
        if variable in ("t", "v", "z"):
            return self._raw_moments(which_dist, light_weighted)[0]
        assert variable == "x"
        p = self.get_p(which_dist, light_weighted=light_weighted, density=False)
        p1, p2 = np.sum(p, 1), np.sum(p, 0)
        return np.array([np.sum(self.cube.x * p1.T, -1).T, np.sum(self.cube.y * p2.T, -1).T])

    def get_noncentral_moment(self, which_dist, j, mu, light_weighted=False):
        """sum_i (var_i - mu)^j P(var_i | ...) (base.py:233-282); `mu` broadcasts against the
        conditioners."""
        variable = self._moment_axis(which_dist)
        hard_coded = getattr(self, f"get_noncentral_moment_{which_dist}", None)
        if hard_coded is not None:
            return hard_coded(j, mu, light_weighted=light_weighted)
        p = self.get_p(which_dist, light_weighted=light_weighted, density=False)
        if variable in ("t", "v", "z"):
            values = self.cube.get_variable_values(variable)
            return np.sum((np.power(values - np.asarray(mu)[np.newaxis].T, j) * p.T).T, 0)
        assert variable == "x"
        p1, p2 = np.sum(p, 1), np.sum(p, 0)
        mu1, mu2 = mu
        d1 = np.power(self.cube.x - np.asarray(mu1)[np.newaxis].T, j).T
        d2 = np.power(self.cube.y - np.asarray(mu2)[np.newaxis].T, j).T
        return np.array([np.sum(d1 * p1, 0), np.sum(d2 * p2, 0)])

    def get_central_moment(self, which_dist, j, light_weighted=False):
        variable = self._moment_axis(which_dist)
        uses_hard_coded = (hasattr(self, "get_mean_" + which_dist)
                           or hasattr(self, f"get_noncentral_moment_{which_dist}"))
        if variable in ("t", "v", "z") and not uses_hard_coded and 2 <= j <= 4:
            # one pass on the device gives the mean and the central moments 2..4
            return self._raw_moments(which_dist, light_weighted)[j - 1]
        mu = self.get_mean(which_dist, light_weighted=light_weighted)
        if variable == "x":
            mu = np.array([mu[0], mu[1]])
        return self.get_noncentral_moment(which_dist, j, mu, light_weighted=light_weighted)

    def get_variance(self, which_dist, light_weighted=False):
        return self.get_central_moment(which_dist, 2, light_weighted=light_weighted)

    def get_dispersion(self, which_dist, light_weighted=False):
        return self.get_variance(which_dist, light_weighted=light_weighted) ** 0.5

    def get_normalised_central_moment(self, which_dist, j, light_weighted=False):
        central = self.get_central_moment(which_dist, j, light_weighted=light_weighted)
        dispersion = self.get_dispersion(which_dist, light_weighted=light_weighted)
        with np.errstate(invalid="ignore", divide="ignore"):
            return central / dispersion ** j

    def get_normalised_central_moment_normal(self, j):
        """Normalised central moment j of a Gaussian: (j-1)!! for even j, else 0."""
        return _double_factorial_odd(j - 1) if j % 2 == 0 else 0.0

    def get_excess_central_moment(self, which_dist, j, light_weighted=False):
        return (self.get_normalised_central_moment(which_dist, j, light_weighted=light_weighted)
                - self.get_normalised_central_moment_normal(j))

    def get_skewness(self, which_dist, light_weighted=False):
        return self.get_normalised_central_moment(which_dist, 3, light_weighted=light_weighted)

    def get_kurtosis(self, which_dist, light_weighted=False):
        return self.get_normalised_central_moment(which_dist, 4, light_weighted=light_weighted)

    def get_excess_kurtosis(self, which_dist, light_weighted=False):
        return self.get_excess_central_moment(which_dist, 4, light_weighted=light_weighted)

    # ------------------------------------------------------------------ the hot path
    def evaluate_ybar(self, batch="none", device=0, dtype="float64"):
        """Evaluate the datacube  ybar(omega, x) = sum_{t,z} p(t,x,z) [S_tz * p(v|t,x,z)](omega).

        Stores `self.ybar` [n_fft, nx1, nx2] (float64 NumPy).  `batch` ('none', 'column',
        'spaxel') is accepted for compatibility: the device path tiles pixels internally so
        all three give the same result (base.py:789-882).  `dtype="float32"` (extension) runs
        the contraction in float32 arithmetic (~1e-7 relative; default float64 matches the
        reference to 1e-10)."""
        if batch not in ("none", "column", "spaxel"):
            raise ValueError("Unknown option for batch")
        check_velocity_grid(self.cube)
        plan = engine.get_plan(self.cube, device=device, precision=dtype)
        # the reference evaluates get_p('tvxz') = exp(log p + log dV - log dV); exp is fused
        out = plan.ybar_density(self._device_log_p(device), is_log=True)
        self.ybar = out.cpu().numpy() if engine._is_torch(out) else out

    # ------------------------------------------------------------------ persistence
    def dill_dump(self, fname, direc=None):
        """Pickle this component with dill into `direc + fname`, creating `direc` if needed
        (base.py:1176-1190).  Device plans and device copies are dropped from the pickle."""
        import dill
        with open(self._out_path(fname, direc), "wb") as fh:
            dill.dump(self, fh)

    @staticmethod
    def _out_path(fname, direc):
        # the reference concatenates `direc + fname` (base.py:1189, 1213); `direc=None` (its
        # default, on which it raises) writes to the working directory here
        if direc is None:
            return fname
        if not os.path.isdir(direc):
            os.mkdir(direc)
        return direc + fname

    def save_numpy(self, fname, yobs=None, direc=None):
        """The v0.0 `npz` format (base.py:1192-1225, deprecated there in favour of `dill_dump`):
        cube geometry, SSP templates, `ybar`, and the density rearranged to `f_xvtz`
        [x1, x2, v, z, t].  A `Mixture` stores the component-collapsed density."""
        cube = self.cube
        try:
            p_tvxz = self.get_p("tvxz", collapse_cmps=True, density=True)     # Mixture
        except TypeError:
            p_tvxz = self.get_p("tvxz", density=True)
        f_xvtz = np.moveaxis(p_tvxz, [0, 1, 2, 3, 4], [4, 2, 0, 1, 3])
        np.savez(self._out_path(fname, direc), yobs=yobs, nx1=cube.nx, nx2=cube.ny, x1rng=cube.x1rng,
                 x2rng=cube.x2rng, S=cube.ssps.Xw.reshape((-1,) + tuple(cube.ssps.par_dims)),
                 w=cube.ssps.w, z_bin_edges=cube.ssps.par_edges[0], t_bin_edges=cube.ssps.par_edges[1],
                 ybar=self.ybar, v_edg=cube.v_edg, f_xvtz=f_xvtz)


def _make_marginal(keep):
    def _get_log_p(self, density=True, light_weighted=False):
        return self._log_marginal(keep, density=density, light_weighted=light_weighted)
    _get_log_p.__name__ = "_get_log_p_" + keep
    _get_log_p.__doc__ = f"log p({','.join(keep)}) of the pixelated density (base.py:884-1174)."
    return _get_log_p


for _keep in _MARGINALS:
    setattr(Component, "_get_log_p_" + _keep, _make_marginal(_keep))
