"""`ParametricComponent`: factorised densities with a Gaussian line-of-sight velocity law.

    p(t, v, x, z) = p(t) p(x|t) p(z|t,x) N(v; mu_v(t,x), sig_v(t,x))

Host-side model construction (off the hot path) behind the reference's
API (/root/reference/popkinmocks/components/parametric.py): `set_p_t`, the chemical
enrichment model for p(z|t,x), every `_get_log_p_<dist>` marginal, the analytic velocity
moments, and `evaluate_ybar`, which hands `P_txz`, `mu_v`, `sig_v` to the CUDA Gaussian-LOSVD
kernel (pkm_ybar_gaussian <- parametric.py:335-407).

The marginals are organised differently from the reference: all of them are reductions of two
arrays, log P(t,x,z) (no velocity) and log P(t,v,x,z) = log P(t,x,z) + log P(v|t,x), built from
the stored factors.  Distributions without `v` never integrate the velocity law (its tails
outside the cube's velocity range are not renormalised, exactly as in the reference).
"""
from abc import abstractmethod

import numpy as np
from scipy import special, stats

from . import base
from .. import _lib, engine


def _log_diff_exp(log_hi, log_lo):
    """log(exp(log_hi) - exp(log_lo)) elementwise, for log_hi >= log_lo.

    Same arithmetic as the signed `logsumexp(..., b=[1, -1])` the reference uses (shift by the
    maximum, subtract, log) written out directly: ~8x faster on the [z, t, x1, x2] arrays of the
    enrichment model, where it dominated model set-up."""
    log_hi = np.asarray(log_hi, dtype=np.float64)
    log_lo = np.asarray(log_lo, dtype=np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        delta = np.where(np.isneginf(log_hi), 0.0, log_lo - log_hi)   # (-inf) - (-inf) -> empty bin
        out = log_hi + np.log(1.0 - np.exp(delta))
    return out


class ParametricComponent(base.Component):
    """Abstract parametric component; subclasses provide p(x|t), t_dep(x), mu_v and sig_v.

    Args:
        cube: a `popkinmocks_b200.ifu_cube.IFUCube`
        center (x1, x2): centre of the component
        rotation: anti-clockwise rotation of the component's axes (radians)
    """

    def __init__(self, cube=None, center=(0, 0), rotation=0.0):
        self.cube = cube
        self.center = center
        self.rotation = rotation
        dx1 = self.cube.xx - center[0]
        dx2 = self.cube.yy - center[1]
        c, s = np.cos(rotation), np.sin(rotation)
        # coordinates in the component's own frame; the reference contracts the FIRST index of
        # [[c, s], [-s, c]] with the position vector (parametric.py:43-49)
        self.xxp = c * dx1 - s * dx2
        self.yyp = s * dx1 + c * dx2
        self.get_z_interpolation_grid()

    # ------------------------------------------------------------------ lazy host arrays / device maps
    # The O(pixels) arrays the reference builds in `set_*` (log_p_x_t, p_x_t, t_dep, log_p_z_tx, p_z_tx,
    # mu_v, sig_v) stay available as attributes but are computed on first ACCESS: `set_*` only records
    # the parameters.  `evaluate_ybar` and `pixelate` never touch them - they build the same maps in
    # HBM from the parameters (`_build_device_maps`, pkm_disk_maps / pkm_stream_maps /
    # pkm_assemble_txz), so nothing of size nx1*nx2 is computed on the host or crosses PCIe.
    def _set_lazy(self, names, method):
        lazy = self.__dict__.setdefault("_lazy", {})
        for name in names:
            self.__dict__.pop(name, None)
            lazy[name] = method
        self.__dict__.pop("_pkm_maps", None)          # parameters changed: device maps are stale

    def __getattr__(self, name):
        lazy = self.__dict__.get("_lazy")
        if lazy is not None and name in lazy:
            getattr(self, lazy[name])()                # fills self.__dict__[name] (and its siblings)
            return self.__dict__[name]
        raise AttributeError(f"{type(self).__name__!s} object has no attribute {name!r}")

    def __getstate__(self):
        state = super().__getstate__()
        state.pop("_pkm_maps", None)                   # device tensors are not pickled
        return state

    def _build_device_maps(self, device):
        """Subclasses with device kernels return dict(log_p_x, px_per_age, mu_v, sig_v, rows,
        row_idx) of CUDA tensors; None means 'use the host arrays'."""
        return None

    def _device_maps(self, device=0):
        """P(t,x,z) [nt,nx1,nx2,nz], the log density log p(t,x,z), mu_v, sig_v on `device`, built
        there from the model parameters; None if this class has no device generator."""
        _lib.require_gpu()
        cache = self.__dict__.setdefault("_pkm_maps", {})
        if device in cache:
            return cache[device]
        import torch
        built = self._build_device_maps(device)
        if built is None:
            return None
        cube = self.cube
        nz, nt = cube.ssps.par_dims
        npix = cube.nx * cube.ny
        dev = f"cuda:{device}"
        rows = np.asarray(built["rows"])
        # log P(z | t, depletion-time row), volume weighted, [row][t][z]
        log_p_z_t_row = self._enrichment_table(rows)                         # [z, t, row] log density
        log_dz = np.log(cube.construct_volume_element("z"))
        log_Pz = np.ascontiguousarray(np.moveaxis(log_p_z_t_row, (0, 1, 2), (2, 1, 0)) + log_dz)
        log_dt = np.log(cube.construct_volume_element("t"))
        log_P_t = _lib.as_f64(self.log_p_t + log_dt)
        P = torch.empty((nt, cube.nx, cube.ny, nz), dtype=torch.float64, device=dev)
        logp = torch.empty_like(P)
        row_idx = built.get("row_idx")
        _lib.check(_lib.load().pkm_assemble_txz(
            _lib.ptr(log_P_t), _lib.ptr(_lib.as_f64(log_dt)), _lib.ptr(built["log_p_x"]), int(built["px_per_age"]),
            float(np.log(cube.dx) + np.log(cube.dy)), _lib.ptr(log_Pz), int(rows.size), _lib.ptr(_lib.as_f64(log_dz)),
            _lib.ptr(row_idx), int(nt), int(npix), int(nz), _lib.ptr(P), _lib.ptr(logp), int(device), None))
        out = dict(P_txz=P, log_p_txz=logp, mu_v=built["mu_v"], sig_v=built["sig_v"],
                   log_p_x=built["log_p_x"], t_dep=built.get("t_dep"), row_idx=row_idx, rows=rows)
        cache[device] = out
        return out

    # ------------------------------------------------------------------ star formation history
    def get_beta_a_b_from_lmd_phi(self, lmd, phi):
        """Beta shape parameters from total concentration `lmd` and mean `phi`."""
        return lmd * phi, lmd * (1.0 - phi)

    def set_p_t(self, lmd=None, phi=None, cdf_start_end=(0.05, 0.95)):
        """Beta-distributed star formation history p(t) (parametric.py:89-144).

        Also records the ages `t_start`, `t_end` between which the time-varying parameters of
        the other factors are interpolated."""
        a, b = self.get_beta_a_b_from_lmd_phi(lmd, phi)
        if not (a > 1.0 or b > 1.0):
            raise ValueError("SFH is bimodal - increase lmd?'")
        t_edges = self.cube.ssps.par_edges[1]
        sfh = stats.beta(a, b, loc=t_edges[0], scale=t_edges[-1] - t_edges[0])
        t_start, t_end = sfh.ppf(cdf_start_end)
        self.t_pars = dict(lmd=lmd, phi=phi, t_start=t_start, t_end=t_end, delta_t=t_end - t_start,
                           idx_start_end=np.digitize([t_start, t_end], t_edges))
        log_cdf = sfh.logcdf(t_edges)
        log_P_t = _log_diff_exp(log_cdf[1:], log_cdf[:-1])
        self.log_p_t = log_P_t - np.log(self.cube.construct_volume_element("t"))
        self.p_t = np.exp(self.log_p_t)

    def linear_interpolate_t(self, f_end, f_start):
        """A parameter varying linearly in age between f_start (at t_start) ... note the
        reference's argument order: value `f_end` applies before `t_start`, `f_start` after
        `t_end` (parametric.py:68-87)."""
        t = self.cube.ssps.par_cents[1]
        slope = (f_start - f_end) / self.t_pars["delta_t"]
        f = f_end + slope * (t - self.t_pars["t_start"])
        f[t < self.t_pars["t_start"]] = f_end
        f[t > self.t_pars["t_end"]] = f_start
        return f

    # ------------------------------------------------------------------ abstract factors
    @abstractmethod
    def set_p_x_t(self):
        pass

    @abstractmethod
    def set_mu_v(self):
        pass

    @abstractmethod
    def set_sig_v(self):
        pass

    @abstractmethod
    def set_t_dep(self):
        pass

    # ------------------------------------------------------------------ chemical enrichment
    def get_z_interpolation_grid(self, t_dep_lim=(0.1, 10.0), n_t_dep=1000, n_z=1000):
        """Tabulate mean metallicity against age for a grid of depletion times.

        Zhu, van de Venn, Leaman et al. 2020 (eqs 3-10): t(z) = t_H - t_dep z / (1 - a z^b);
        inverted by interpolation onto the SSP ages (parametric.py:213-247)."""
        self.ahat = 10.0 ** (-0.689)
        self.bhat = 1.899 - 1.0
        key = ("_pkm_z_grid", tuple(t_dep_lim), n_t_dep, n_z)
        cached = self.cube.__dict__.get(key)
        if cached is not None:                         # same cube, same table: components share it
            self.t, self.z_t_interp_grid = cached
            return
        z_max = 0.9999 * self.ahat ** (-1.0 / self.bhat)
        t_hubble = self.cube.ssps.par_edges[1][-1]
        t_dep = np.logspace(np.log10(t_dep_lim[0]), np.log10(t_dep_lim[1]), n_t_dep)
        z = np.linspace(z_max, 0.0, n_z)
        t_of_z = t_hubble - t_dep[:, None] * z[None, :] / (1.0 - self.ahat * z[None, :] ** self.bhat)
        self.t = t_of_z
        t_ssps = self.cube.ssps.par_cents[1]
        z_of_t = np.empty((n_t_dep, t_ssps.size))
        for i in range(n_t_dep):
            z_of_t[i] = np.interp(t_ssps, t_of_z[i], z)
        self.z_t_interp_grid = dict(t=t_ssps, z=z_of_t, t_dep=t_dep)
        self.cube.__dict__[key] = (self.t, self.z_t_interp_grid)

    def _enrichment_table(self, rows):
        """log p(z | t, row) [z, t, len(rows)] for rows of the tabulated depletion-time grid
        (parametric.py:253-289): linear metallicity is Gaussian about the tabulated mean with
        variance a z^b, binned onto the SSP metallicity edges and normalised over the grid."""
        grid = self.z_t_interp_grid
        z_mu = grid["z"][rows].T                                          # [t, row]
        z_sd = (self.ahat * z_mu ** self.bhat) ** 0.5
        lin_z_edges = 10.0 ** self.cube.ssps.par_edges[0]
        # = stats.norm(z_mu, z_sd).logcdf(edges): the same log_ndtr call without the frozen-
        # distribution overhead
        with np.errstate(divide="ignore", invalid="ignore"):
            log_cdf = special.log_ndtr((lin_z_edges[:, None, None] - z_mu) / z_sd)
        log_P = _log_diff_exp(log_cdf[1:], log_cdf[:-1])                  # [z, t, row]
        log_dz = np.log(self.cube.construct_volume_element("z"))
        log_norm = special.logsumexp(log_P.T + log_dz, -1).T
        return log_P - log_norm

    def evaluate_chemical_enrichment_model_base(self, t_dep):
        """log p(z|t,x) for a map of depletion times t_dep[x1,x2] (parametric.py:253-289)."""
        grid = self.z_t_interp_grid
        nearest = np.argmin(np.abs(t_dep[:, :, None] - grid["t_dep"]), axis=-1)
        # p(z|t,x) depends on x only through the tabulated depletion time nearest to t_dep[x]:
        # evaluate the model once per distinct table row and gather (same arithmetic per element
        # as the reference, which works on the full [z+1, t, x1, x2] arrays - bit-identical, and
        # 10-100x less work for maps with few distinct depletion times)
        rows, inverse = np.unique(nearest.ravel(), return_inverse=True)
        inverse = inverse.reshape(nearest.shape)
        log_p_z_t_row = self._enrichment_table(rows)
        return log_p_z_t_row[:, :, inverse], np.exp(log_p_z_t_row)[:, :, inverse]   # [z, t, x1, x2]

    def set_p_z_tx(self):
        self.__dict__["log_p_z_tx"], self.__dict__["p_z_tx"] = \
            self.evaluate_chemical_enrichment_model_base(self.t_dep)

    def evaluate_chemical_enrichment_model(self, t_dep):
        """p(z|t) for one spatially constant depletion time."""
        uniform = np.full((self.cube.nx, self.cube.ny), t_dep)
        _, p_z_tx = self.evaluate_chemical_enrichment_model_base(uniform)
        return p_z_tx[:, :, 0, 0]

    # ------------------------------------------------------------------ building blocks
    def _log_P_txz_mass(self):
        """log P(t,x,z): volume-weighted, mass-weighted, shape [t, x1, x2, z]."""
        cube = self.cube
        log_P_t = self.log_p_t + np.log(cube.construct_volume_element("t"))
        log_P_x_t = self.log_p_x_t + (np.log(cube.dx) + np.log(cube.dy))          # [x1, x2, t]
        log_P_z_tx = (self.log_p_z_tx.T + np.log(cube.construct_volume_element("z"))).T  # [z,t,x1,x2]
        log_P_tx = np.moveaxis(log_P_x_t + log_P_t, -1, 0)
        return log_P_tx[..., None] + np.moveaxis(log_P_z_tx, 0, -1)

    def _log_P_v_tx(self):
        """log P(v-bin | t,x) from Gaussian cdf differences, shape [v, t, x1, x2]."""
        v_edg = self.cube.v_edg[:, None, None, None]
        with np.errstate(divide="ignore", invalid="ignore"):
            log_cdf = special.log_ndtr((v_edg - self.mu_v) / self.sig_v)   # == stats.norm.logcdf
        return _log_diff_exp(log_cdf[1:], log_cdf[:-1])

    def _log_P_joint(self, with_v, light_weighted):
        """log P over (t,x,z) or (t,v,x,z), light-weighted and renormalised if asked."""
        log_P = self._log_P_txz_mass()
        lw_axes = (slice(None), None, None, slice(None))
        if with_v:
            log_P = log_P[:, None] + np.moveaxis(self._log_P_v_tx(), 0, 1)[..., None]
            lw_axes = (slice(None), None, None, None, slice(None))
        if light_weighted:
            log_P = log_P + np.log(self.cube.ssps.light_weights[lw_axes])
            log_P = log_P - special.logsumexp(log_P)
        return log_P

    def _log_marginal(self, keep, density=True, light_weighted=False):
        with_v = "v" in keep
        log_P = self._log_P_joint(with_v, light_weighted)
        axes = {"t": (0,), "v": (1,), "x": (2, 3), "z": (4,)} if with_v else \
               {"t": (0,), "x": (1, 2), "z": (3,)}
        drop = tuple(ax for name, axs in axes.items() if name not in keep for ax in axs)
        if drop:
            log_P = special.logsumexp(log_P, drop)
        if density:
            log_P = log_P - np.log(self.cube.construct_volume_element(keep))
        return log_P

    def _raw_moments(self, which_dist, light_weighted):
        """[mean, mu2, mu3, mu4] of a 1-D (conditional) distribution of t or z: the analytic
        marginals live on the host, so the probabilities are handed to pkm_moments."""
        variable = self._moment_axis(which_dist)
        P = self.get_p(which_dist, light_weighted=light_weighted, density=False)
        return engine.moments(P, self.cube.get_variable_values(variable))

    # hard-coded conditionals kept for API parity: mass-weighted they are the stored factors
    def _get_log_p_x_t(self, density=True, light_weighted=False):
        if light_weighted:
            raise NotImplementedError  # falls back to joint - marginal in get_log_p
        out = self.log_p_x_t.copy()
        if not density:
            out += np.log(self.cube.dx) + np.log(self.cube.dy)
        return out

    def _get_log_p_z_tx(self, density=True, light_weighted=False):
        if light_weighted:
            raise NotImplementedError
        out = self.log_p_z_tx.copy()
        if not density:
            out = (out.T + np.log(self.cube.construct_volume_element("z"))).T
        return out

    def _get_log_p_v_tx(self, density=True, light_weighted=False):
        if light_weighted:
            raise NotImplementedError
        out = self._log_P_v_tx()
        if density:
            out = (out.T - np.log(self.cube.construct_volume_element("v"))).T
        return out

    # ------------------------------------------------------------------ the hot path
    def evaluate_ybar(self, batch="none", device=0, dtype="float64"):
        """Datacube of a Gaussian-LOSVD component (parametric.py:335-407).

        Y-hat[k,x] = sum_t exp(-i mu_v w_k - (sig_v w_k)^2/2) sum_z P(t,x,z) S-hat[k,z,t];
        one inverse real FFT per pixel; all on the device.  `batch` is accepted for
        compatibility (the device path tiles pixels internally).  `dtype="float32"` (extension)
        runs the metallicity contraction and the Gaussian kernel in float32 arithmetic
        (~1e-6 relative; default float64 matches the reference to 1e-10)."""
        if batch not in ("none", "column", "spaxel"):
            raise ValueError("Unknown option for batch")
        plan = engine.get_plan(self.cube, device=device, precision=dtype)
        maps = self._device_maps(device)
        out = self._new_host_cube(plan.n_fft)
        if maps is not None:                           # P(t,x,z), mu_v, sig_v generated in HBM
            plan.ybar_gaussian(maps["P_txz"], maps["mu_v"], maps["sig_v"], out=out)
        else:
            plan.ybar_gaussian(self.get_p("txz", density=False), self.mu_v, self.sig_v, out=out)
        self.ybar = out

    def pixelate(self, device=0):
        """The pixelated twin of this component, `Component(cube, self.get_log_p("tvxz"))`, with
        the 5-D array produced and kept on the device.

        The reference builds log p(t,v,x,z) on the host (parametric.py:567-596: normal-cdf
        differences over the velocity bins, parametric.py:180-211) - 20.8 GB at 201x201 pixels.
        Here log p(t,x,z), mu_v and sig_v are generated in HBM from the model parameters (or, for
        classes without a device generator, cross PCIe); pkm_pixelate_gaussian writes the 5-D
        array into HBM, where `evaluate_ybar`, `get_p` and the moments of the returned
        `Component` read it without a host round trip."""
        maps = self._device_maps(device)
        if maps is not None:
            log_p_txz, mu_v, sig_v = maps["log_p_txz"], maps["mu_v"], maps["sig_v"]
        else:
            log_p_txz = self._log_marginal("txz", density=True, light_weighted=False)
            mu_v, sig_v = self.mu_v, self.sig_v
        log_p = engine.pixelate_gaussian(log_p_txz, mu_v, sig_v, self.cube.v_edg,
                                         density=True, device=device, on_device=True)
        return base.Component(cube=self.cube, log_p_tvxz=log_p)

    # ------------------------------------------------------------------ analytic velocity moments
    # E and central moments of v given (t,x) are those of the Gaussian law and do not depend on
    # z or on the weighting; every other conditioning set averages them over P(t,x | given)
    # (parametric.py:757-1228).
    _V_WEIGHTS = {"tx": None, "txz": None, "x": "t_x", "t": "x_t", "": "tx", "tz": "x_tz",
                  "xz": "t_xz", "z": "tx_z"}

    def _tx_weights(self, given, light_weighted):
        """P(t,x | given) as an array broadcastable to [t, x1, x2, (z)] and the axes to sum."""
        key = self._V_WEIGHTS[given]
        if key is None:
            return 1.0, ()
        P = self.get_p(key if given else "tx", light_weighted=light_weighted, density=False)
        if given == "":
            return P, (0, 1, 2)
        if given == "x":
            return P, (0,)                                  # [t, x1, x2]
        if given == "t":
            return np.moveaxis(P, -1, 0), (1, 2)            # [x1,x2,t] -> [t,x1,x2]
        if given == "tz":
            return np.moveaxis(P, -2, 0), (1, 2)            # [x1,x2,t,z] -> [t,x1,x2,z]
        if given == "xz":
            return P, (0,)                                  # [t,x1,x2,z]
        if given == "z":
            return P, (0, 1, 2)                             # [t,x1,x2,z]
        raise ValueError(given)

    def _gauss_central_moment(self, k):
        if k % 2:
            return np.zeros_like(self.sig_v)
        return self.sig_v ** k * base._double_factorial_odd(k - 1)

    @staticmethod
    def _align_to_tx(given, arr):
        """Reshape an array laid out like the conditioners `given` to broadcast over [t,x1,x2,(z)]."""
        arr = np.asarray(arr)
        if given in ("", "tx", "txz"):
            return arr
        if given in ("x", "xz"):
            return arr[None]
        if given == "t":
            return arr[:, None, None]
        if given == "tz":
            return arr[:, None, None, :]
        if given == "z":
            return arr[None, None, None, :]
        raise ValueError(given)

    def _mu_v_aligned(self, given):
        mu_v = np.asarray(self.mu_v)
        if "z" not in given:
            return mu_v
        mu_v = mu_v[..., None]
        if given == "txz":
            mu_v = np.broadcast_to(mu_v, mu_v.shape[:-1] + (self.cube.ssps.delta_z.size,))
        return mu_v

    def _mean_v(self, given, light_weighted=False):
        W, axes = self._tx_weights(given, light_weighted)
        out = self._mu_v_aligned(given) * W
        return np.sum(out, axes) if axes else out

    def _noncentral_moment_v(self, given, j, mu, light_weighted=False):
        """E[(v - mu)^j | given] = sum_k C(j,k) E_tx[(mu_v - mu)^(j-k) m_k(sig_v)]."""
        W, axes = self._tx_weights(given, light_weighted)
        delta = self._mu_v_aligned(given) - self._align_to_tx(given, mu)
        total = 0.0
        for k in range(j + 1):
            mk = self._gauss_central_moment(k)
            if "z" in given:
                mk = mk[..., None]
            total = total + special.comb(j, k) * delta ** (j - k) * mk
        total = total * W
        return np.sum(total, axes) if axes else total

    def get_mean_v_tx(self, light_weighted=False):
        return self.mu_v

    def get_central_moment_v_tx(self, j, light_weighted=False):
        return self._gauss_central_moment(j)


def _install_v_moment_methods():
    def mean_factory(given):
        def get_mean(self, light_weighted=False):
            return self._mean_v(given, light_weighted=light_weighted)
        return get_mean

    def moment_factory(given):
        def get_noncentral_moment(self, j, mu, light_weighted=False):
            return self._noncentral_moment_v(given, j, mu, light_weighted=light_weighted)
        return get_noncentral_moment

    for given in ("x", "t", "", "txz", "tz", "xz", "z"):
        suffix = "v" + ("_" + given if given else "")
        setattr(ParametricComponent, "get_mean_" + suffix, mean_factory(given))
        setattr(ParametricComponent, "get_noncentral_moment_" + suffix, moment_factory(given))
    setattr(ParametricComponent, "get_noncentral_moment_v_tx", moment_factory("tx"))


def _install_marginals():
    def factory(keep):
        def _get_log_p(self, density=True, light_weighted=False):
            return self._log_marginal(keep, density=density, light_weighted=light_weighted)
        _get_log_p.__name__ = "_get_log_p_" + keep
        return _get_log_p

    for keep in base._MARGINALS:
        setattr(ParametricComponent, "_get_log_p_" + keep, factory(keep))


_install_marginals()
_install_v_moment_methods()
