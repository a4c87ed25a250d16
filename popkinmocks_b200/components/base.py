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
        out = self._device_distribution(which_dist, density, light_weighted, exponentiate=True)
        if out is not None:
            return out
        return np.exp(self.get_log_p(which_dist, density=density, light_weighted=light_weighted))

    def get_log_p(self, which_dist, density=True, light_weighted=False):
        out = self._device_distribution(which_dist, density, light_weighted, exponentiate=False)
        if out is not None:
            return out
        if "_" in which_dist:
            return self._get_log_conditional_distribution(which_dist, density=density,
                                                          light_weighted=light_weighted)
        return self._get_log_marginal_distribution(which_dist, density=density,
                                                   light_weighted=light_weighted)

    def _device_distribution(self, which_dist, density, light_weighted, exponentiate):
        """get_p / get_log_p of the pixelated density assembled on the device (pkm_conditional):
        log P(joint) - log P(given) - log dV and the exponential are one kernel on the cached log
        marginals, and only the result crosses PCIe.  None = not applicable (subclasses with
        their own `_log_marginal`, hard-coded conditionals, arrays streamed from the host)."""
        if type(self)._log_marginal is not Component._log_marginal:
            return None
        dependent, _, given = which_dist.partition("_")
        if given and getattr(self, "_get_log_p_" + which_dist, None) is not None:
            return None
        joint = "".join(sorted(dependent + given))
        if "".join(sorted(dependent)) != dependent or "".join(sorted(given)) != given or joint == "tvxz":
            return None                                  # the 5-D array itself goes through pkm_reduce_logp
        log_joint = self._log_marginal_dev(joint, light_weighted)
        if not engine._is_torch(log_joint):
            return None
        log_given = self._log_marginal_dev(given, light_weighted) if given else None
        cube = self.cube
        plan = engine.get_plan(cube, device=int(log_joint.device.index or 0))
        log_dv = np.log(cube.v_edg[1:] - cube.v_edg[:-1])
        return plan.conditional(log_joint, joint, log_given, given, cube.nx * cube.ny, log_dv, density,
                                exponentiate, (cube.nx, cube.ny))

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

    # ------------------------------------------------------------------ device state
    # `log_p_tvxz` is a property so that rebinding it drops everything derived from the old array:
    # the device-resident copy (shared by evaluate_ybar / get_p / moments so that a sequence of
    # calls uploads the 5-D array once - the reference re-derives it per call), the cached log
    # marginals and the cached moments.  An IN-PLACE edit of the array cannot be detected: call
    # `invalidate_cache()` after one.
    _DEVICE_CACHE_LIMIT = 64 << 30
    _MARGINAL_CACHE_ITEM = 1 << 30       # marginals above this size are not kept on the device
    _MARGINAL_CACHE_TOTAL = 4 << 30

    @property
    def log_p_tvxz(self):
        d = self.__dict__
        if "_log_p_tvxz" in d:
            return d["_log_p_tvxz"]
        if "log_p_tvxz" in d:                # instance state written before this was a property
            return d["log_p_tvxz"]
        raise AttributeError(f"{type(self).__name__} has no pixelated density `log_p_tvxz`")

    @log_p_tvxz.setter
    def log_p_tvxz(self, value):
        self.invalidate_cache()
        self.__dict__["_log_p_tvxz"] = value

    def invalidate_cache(self):
        """Forget the device copy of `log_p_tvxz` and every cached marginal / moment (call this
        after modifying the array in place)."""
        for key in ("_pkm_dev", "_pkm_marg", "_pkm_mom"):
            self.__dict__.pop(key, None)

    release_device = invalidate_cache

    def _device_index(self, device=None):
        """The GPU this component computes on: where `log_p_tvxz` lives if it is a CUDA tensor,
        else `device`, else the device of the cached copy, else 0."""
        src = self.log_p_tvxz
        if engine._is_cuda_tensor(src):
            return int(src.device.index or 0)
        if device is not None:
            return int(device)
        cached = self.__dict__.get("_pkm_dev")
        return int(cached[1]) if cached is not None else 0

    def _device_log_p(self, device=None):
        device = self._device_index(device)
        cached = self.__dict__.get("_pkm_dev")
        src = self.log_p_tvxz
        if cached is not None and cached[0] is src and cached[1] == device:
            return cached[2]
        if engine._is_cuda_tensor(src):
            return src
        engine._lib.require_gpu()
        if src.nbytes > self._DEVICE_CACHE_LIMIT:
            return src                       # too large to keep resident: stream it per call
        dev = engine.to_device(src, device)  # pinned, chunked upload (no pageable full-array copy)
        self.__dict__["_pkm_dev"] = (src, device, dev)
        return dev

    def __getstate__(self):
        state = dict(self.__dict__)
        for key in ("_pkm_dev", "_pkm_marg", "_pkm_mom", "_pkm_ybar_pin"):
            state.pop(key, None)             # never pickle device handles (dill_dump must keep working)
        if "ybar" in state and state["ybar"] is not None:
            state["ybar"] = np.array(state["ybar"])     # a plain copy of the page-locked view
        return state

    # Device-resident log PROBABILITY marginals (density=False), keyed by (axes, light_weighted).
    # A marginal whose axes are a subset of a cached one is reduced from that small array
    # (pkm_reduce_logp_from) instead of the 5-D array: `get_skewness("v_x")` reads the 5-D array
    # once (for p(v,x)); p(x), the conditional and all four moments follow from the 32 MB result.
    def _log_marginal_dev(self, keep, light_weighted=False, device=None):
        device = self._device_index(device)
        cache = self.__dict__.setdefault("_pkm_marg", {})
        key = (keep, bool(light_weighted), device)
        if key in cache:
            return cache[key]
        cube = self.cube
        plan = engine.get_plan(cube, device=device)
        log_dv = np.log(cube.v_edg[1:] - cube.v_edg[:-1])
        lw = cube.ssps.light_weights
        npix = cube.nx * cube.ny
        # smallest cached superset: same weighting, or mass weighted with t and z kept (re-weightable)
        best = None
        for (k2, lw2, dev2), arr in cache.items():
            if dev2 != device or not set(keep) <= set(k2) or k2 == "tvxz":
                continue
            if lw2 != bool(light_weighted) and not (light_weighted and not lw2 and "t" in k2 and "z" in k2):
                continue
            if best is None or arr.numel() < best[2].numel():
                best = (k2, lw2, arr)
        if best is not None:
            k2, lw2, arr = best
            out = plan.reduce_logp_from(arr, k2, npix, log_dv, keep, False,
                                        light_weights=lw if (light_weighted and not lw2) else None)
        else:
            src = self._device_log_p(device)
            out = plan.reduce_logp(src, log_dv, keep, False, light_weights=lw if light_weighted else None)
            if not engine._is_torch(out):        # host array too large to keep resident: no caching
                return out
        nbytes = out.numel() * 8
        total = sum(a.numel() * 8 for a in cache.values())
        if nbytes <= self._MARGINAL_CACHE_ITEM and total + nbytes <= self._MARGINAL_CACHE_TOTAL:
            cache[key] = out
        return out

    def _log_marginal(self, keep, density=True, light_weighted=False):
        """log-marginal over the axes not in `keep`, on the GPU (pkm_reduce_logp)."""
        out = self._log_marginal_dev(keep, light_weighted)
        if not engine._is_torch(out):
            if density:
                out = out - np.log(self.cube.construct_volume_element(keep))
            return out
        if density:
            cube = self.cube
            plan = engine.get_plan(cube, device=int(out.device.index or 0))
            log_dv = np.log(cube.v_edg[1:] - cube.v_edg[:-1])
            out = plan.reduce_logp_from(out, keep, cube.nx * cube.ny, log_dv, keep, True)
        return out.cpu().numpy()

    # ------------------------------------------------------------------ moments
    def _moment_axis(self, which_dist):
        variable = which_dist.split("_")[0]
        assert len(variable) == 1
        return variable

    def _raw_moments(self, which_dist, light_weighted):
        """[mean, mu2, mu3, mu4] of a 1-D (conditional) distribution of t, v or z.

        One kernel on the device-resident log marginals (pkm_moments_logp): the conditional
        P = exp(log P(joint) - log P(given)) is formed on the fly and never visits the host."""
        cache = self.__dict__.setdefault("_pkm_mom", {})
        key = (which_dist, bool(light_weighted))
        if key in cache:
            return cache[key]
        variable = self._moment_axis(which_dist)
        given = which_dist.split("_")[1] if "_" in which_dist else ""
        joint = "".join(sorted(variable + given))
        log_joint = self._log_marginal_dev(joint, light_weighted)
        if not engine._is_torch(log_joint):        # streamed (host) arrays: the probability route
            P = self.get_p(which_dist, light_weighted=light_weighted, density=False)
            out = engine.moments(P, self.cube.get_variable_values(variable))
        else:
            log_given = self._log_marginal_dev(given, light_weighted) if given else None
            shape = self.cube.get_distribution_shape(joint)
            axis = joint.replace("x", "xy").find(variable)
            nouter = int(np.prod(shape[:axis], dtype=np.int64))
            ninner = int(np.prod(shape[axis + 1:], dtype=np.int64))
            out = engine.moments_logp(log_joint, log_given, self.cube.get_variable_values(variable),
                                      nouter, shape[axis], ninner)
            out = out.reshape((4,) + tuple(shape[:axis]) + tuple(shape[axis + 1:]))
        cache[key] = out
        return out

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
    def evaluate_ybar(self, batch="none", device=None, dtype="float64"):
        """Evaluate the datacube  ybar(omega, x) = sum_{t,z} p(t,x,z) [S_tz * p(v|t,x,z)](omega).

        Stores `self.ybar` [n_fft, nx1, nx2] (float64 NumPy).  `batch` ('none', 'column',
        'spaxel') is accepted for compatibility: the device path tiles pixels internally so
        all three give the same result (base.py:789-882).  `dtype="float32"` (extension) runs
        the contraction in float32 arithmetic (~1e-7 relative; default float64 matches the
        reference to 1e-10).

        A host `log_p_tvxz` is page-locked in place on first use and STREAMED through the
        library's pipelined staging (H2D of tile i+1, kernels of tile i, D2H of tile i-1): no
        device copy of the 5-D array is made for this call.  If a device copy already exists
        (because `get_p` / the moments made one) it is used instead.  The result lands in a
        page-locked buffer from torch's caching allocator, viewed as a NumPy array."""
        if batch not in ("none", "column", "spaxel"):
            raise ValueError("Unknown option for batch")
        check_velocity_grid(self.cube)
        device = self._device_index(device)
        plan = engine.get_plan(self.cube, device=device, precision=dtype)
        src = self.log_p_tvxz
        cached = self.__dict__.get("_pkm_dev")
        if engine._is_cuda_tensor(src) or (cached is not None and cached[0] is src and cached[1] == device):
            dev_src = src if engine._is_cuda_tensor(src) else cached[2]
            out = self._new_host_cube(plan.n_fft)
            plan.ybar_density(dev_src, is_log=True, out=out)
            self.ybar = out
            return
        # the reference evaluates get_p('tvxz') = exp(log p + log dV - log dV); exp is fused
        host = np.ascontiguousarray(src, dtype=np.float64)
        engine.pin_host_array(host)
        out = self._new_host_cube(plan.n_fft)
        plan.ybar_density(host, is_log=True, out=out)
        self.ybar = out

    def _new_host_cube(self, n_fft):
        """A page-locked [n_fft, nx1, nx2] result buffer (torch's caching host allocator), viewed as
        a NumPy array: device-to-host copies into it run at PCIe rate.  The previous result is
        dropped first so that its block can be reused (it is not, and stays valid, if the caller
        still holds the old array)."""
        self.__dict__.pop("ybar", None)
        self.__dict__.pop("_pkm_ybar_pin", None)
        keep, out = engine.pinned_empty((n_fft, self.cube.nx, self.cube.ny))
        self.__dict__["_pkm_ybar_pin"] = keep
        return out

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
