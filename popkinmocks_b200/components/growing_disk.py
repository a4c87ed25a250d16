"""`GrowingDisk`: a disk whose flattening, size, rotation curve and dispersion evolve with age.

Host-side parameter maps (off the hot path) behind the reference's API
(/root/reference/popkinmocks/components/growing_disk.py:6-239).  Elliptical radius in the
component frame: r = sqrt(x'^2 + (y'/q)^2).
"""
import numpy as np
from scipy import special

from . import parametric


def _per_age(values):
    """[nt] -> [nt, 1, 1] for broadcasting against [x1, x2] maps."""
    return values[:, None, None]


class GrowingDisk(parametric.ParametricComponent):

    def _elliptical_radius(self, q):
        return (self.xxp ** 2 + (self.yyp / q) ** 2) ** 0.5

    def set_p_x_t(self, q_lims=(0.5, 0.1), rc_lims=(0.5, 0.1), alpha_lims=(0.5, 2.0)):
        """Cored power-law surface density rho(r) ~ (r + rc)^-alpha, normalised per age.

        Each `*_lims` pair gives the parameter for (young, old) stars, interpolated linearly
        in age between them (growing_disk.py:30-71).  `log_p_x_t` / `p_x_t` [x1, x2, t] are
        computed on first access."""
        q_lims, rc_lims, alpha_lims = (np.array(v) for v in (q_lims, rc_lims, alpha_lims))
        assert np.all(q_lims > 0.0)
        assert np.all(rc_lims > 0.0)
        assert np.all(alpha_lims >= 0.0)
        self.x_t_pars = dict(q_lims=q_lims, rc_lims=rc_lims, alpha_lims=alpha_lims)
        self._set_lazy(("log_p_x_t", "p_x_t"), "_host_p_x_t")

    def _host_p_x_t(self):
        pars = self.x_t_pars
        q = _per_age(self.linear_interpolate_t(*pars["q_lims"]))
        rc = _per_age(self.linear_interpolate_t(*pars["rc_lims"]))
        alpha = _per_age(self.linear_interpolate_t(*pars["alpha_lims"]))
        log_rho = -alpha * np.log(self._elliptical_radius(q) + rc)            # [t, x1, x2]
        log_pixel = np.log(self.cube.dx) + np.log(self.cube.dy)
        log_norm = special.logsumexp(log_rho + log_pixel, (1, 2))
        self.__dict__["log_p_x_t"] = np.moveaxis(log_rho - _per_age(log_norm), 0, -1)      # [x1, x2, t]
        self.__dict__["p_x_t"] = np.exp(self.__dict__["log_p_x_t"])

    def set_t_dep(self, q=0.1, alpha=1.5, t_dep_in=0.5, t_dep_out=6.0, t_dep_min=0.1,
                  t_dep_max=13.0):
        """Depletion-time map: power law in radius from `t_dep_in` towards `t_dep_out`,
        saturating at `t_dep_out` (growing_disk.py:73-120).  `t_dep`, `log_p_z_tx` and `p_z_tx`
        are computed on first access."""
        assert q > 0.0 and alpha >= 0.0
        assert t_dep_min >= 0.1 and t_dep_max <= 13.0
        assert t_dep_min <= t_dep_in <= t_dep_max
        assert t_dep_min <= t_dep_out <= t_dep_max
        self.t_dep_pars = dict(q=q, alpha=alpha, t_dep_in=t_dep_in, t_dep_out=t_dep_out)
        self._set_lazy(("t_dep",), "_host_t_dep")
        self._set_lazy(("log_p_z_tx", "p_z_tx"), "set_p_z_tx")

    def _host_t_dep(self):
        pars = self.t_dep_pars
        t_dep_in, t_dep_out = pars["t_dep_in"], pars["t_dep_out"]
        scale = np.max(np.abs(self.xxp))
        t_dep = t_dep_in + (t_dep_out - t_dep_in) * (self._elliptical_radius(pars["q"]) / scale) ** pars["alpha"]
        beyond = t_dep > t_dep_out if t_dep_in <= t_dep_out else t_dep < t_dep_out
        t_dep[beyond] = t_dep_out
        self.__dict__["t_dep"] = t_dep

    def set_mu_v(self, q_lims=(0.5, 0.5), rmax_lims=(0.1, 1.0), vmax_lims=(50.0, 250.0)):
        """Rotation maps mu_v(t,x) = K r / (r + rc)^3 cos(theta), peaking at `vmax` at radius
        `rmax` (rc = 2 rmax), on elliptical radii with flattening q (growing_disk.py:122-170).
        `mu_v` [t, x1, x2] is computed on first access."""
        q_lims, rmax_lims, vmax_lims = (np.array(v) for v in (q_lims, rmax_lims, vmax_lims))
        assert np.all(q_lims > 0.0)
        signs = np.sign(vmax_lims)
        assert np.all(np.isin(signs, [0, 1])) or np.all(np.isin(signs, [0, -1]))
        self.mu_v_pars = dict(q_lims=q_lims, rmax_lims=rmax_lims, vmax_lims=vmax_lims)
        self._set_lazy(("mu_v",), "_host_mu_v")

    def _mu_v_age_pars(self):
        pars = self.mu_v_pars
        q = self.linear_interpolate_t(*pars["q_lims"])
        rmax = self.linear_interpolate_t(*pars["rmax_lims"])
        vmax = self.linear_interpolate_t(*pars["vmax_lims"])
        rc = 2.0 * rmax
        K = vmax * (rmax + rc) ** 3 / rmax
        return q, rc, K

    def _host_mu_v(self):
        q, rc, K = (_per_age(a) for a in self._mu_v_age_pars())
        theta = np.arctan2(self.yyp / q, self.xxp)
        r = self._elliptical_radius(q)
        self.__dict__["mu_v"] = K * r / (r + rc) ** 3 * np.cos(theta)

    def set_sig_v(self, q_lims=(0.5, 0.1), alpha_lims=(1.5, 2.5), sig_v_in_lims=(50.0, 250.0),
                  sig_v_out_lims=(10.0, 50), sigma_max=500.0, sigma_min=10.0):
        """Dispersion maps: power law in radius from `sig_v_in` towards `sig_v_out`, saturating
        at `sig_v_out` (growing_disk.py:172-239).  `sig_v` [t, x1, x2] is computed on first access."""
        q_lims, alpha_lims, sig_v_in_lims, sig_v_out_lims = (
            np.array(v) for v in (q_lims, alpha_lims, sig_v_in_lims, sig_v_out_lims))
        assert np.all(q_lims > 0.0)
        assert np.all(alpha_lims >= 0.0)
        assert np.all(sig_v_in_lims > 0.0)
        assert np.all(sig_v_out_lims > 0.0)
        self.sig_v_pars = dict(q_lims=q_lims, alpha_lims=alpha_lims, sig_v_in_lims=sig_v_in_lims,
                               sig_v_out_lims=sig_v_out_lims)
        self._set_lazy(("sig_v",), "_host_sig_v")

    def _host_sig_v(self):
        pars = self.sig_v_pars
        q = _per_age(self.linear_interpolate_t(*pars["q_lims"]))
        alpha = _per_age(self.linear_interpolate_t(*pars["alpha_lims"]))
        sig_in = _per_age(self.linear_interpolate_t(*pars["sig_v_in_lims"]))
        sig_out = _per_age(self.linear_interpolate_t(*pars["sig_v_out_lims"]))
        scale = np.max(np.abs(self.xxp))
        sigma = sig_in + (sig_out - sig_in) * (self._elliptical_radius(q) / scale) ** alpha
        rising = sig_out > sig_in                                             # [t, 1, 1]
        beyond = np.where(rising, sigma > sig_out, sigma < sig_out)
        self.__dict__["sig_v"] = np.where(beyond, np.broadcast_to(sig_out, sigma.shape), sigma)

    # ------------------------------------------------------------------ the same maps, in HBM
    def _build_device_maps(self, device):
        """p(x|t), mu_v, sig_v, t_dep and the depletion-time row of every pixel generated on the
        device from the O(nt) parameter vectors (pkm_disk_maps)."""
        for attr in ("x_t_pars", "t_dep_pars", "mu_v_pars", "sig_v_pars", "t_pars"):
            if attr not in self.__dict__:
                return None                            # incomplete model: let the host path raise
        import torch
        from .. import _lib
        cube = self.cube
        nt = cube.ssps.par_dims[1]
        lin = self.linear_interpolate_t
        q_m, rc_m, K_m = self._mu_v_age_pars()
        xp, sp = self.x_t_pars, self.sig_v_pars
        age_pars = _lib.as_f64(np.stack([
            lin(*xp["q_lims"]), lin(*xp["rc_lims"]), lin(*xp["alpha_lims"]), q_m, rc_m, K_m,
            lin(*sp["q_lims"]), lin(*sp["alpha_lims"]), lin(*sp["sig_v_in_lims"]), lin(*sp["sig_v_out_lims"])]))
        scale = float(np.max(np.abs(self.xxp)))
        tp = self.t_dep_pars
        grid = _lib.as_f64(self.z_t_interp_grid["t_dep"])
        dev = f"cuda:{device}"
        shape = (nt, cube.nx, cube.ny)
        log_p_x = torch.empty(shape, dtype=torch.float64, device=dev)
        mu_v = torch.empty(shape, dtype=torch.float64, device=dev)
        sig_v = torch.empty(shape, dtype=torch.float64, device=dev)
        t_dep = torch.empty((cube.nx, cube.ny), dtype=torch.float64, device=dev)
        row_idx = torch.empty((cube.nx, cube.ny), dtype=torch.int32, device=dev)
        tdp = _lib.as_f64([tp["q"], tp["alpha"], tp["t_dep_in"], tp["t_dep_out"]])
        x1, x2 = _lib.as_f64(cube.x), _lib.as_f64(cube.y)
        _lib.check(_lib.load().pkm_disk_maps(
            _lib.ptr(x1), _lib.ptr(x2), cube.nx, cube.ny, float(self.center[0]), float(self.center[1]),
            float(self.rotation), _lib.ptr(age_pars), int(nt), scale, _lib.ptr(tdp), _lib.ptr(grid), int(grid.size),
            _lib.ptr(log_p_x), _lib.ptr(mu_v), _lib.ptr(sig_v), _lib.ptr(t_dep), _lib.ptr(row_idx),
            float(np.log(cube.dx) + np.log(cube.dy)), int(device), None))
        # the depletion time runs between its two limits: only those table rows are ever needed
        ends = np.array([tp["t_dep_in"], tp["t_dep_out"]])
        near = np.argmin(np.abs(ends[:, None] - grid), axis=-1)
        r_lo, r_hi = int(near.min()), int(near.max())
        row_idx -= r_lo
        return dict(log_p_x=log_p_x, px_per_age=1, mu_v=mu_v, sig_v=sig_v, t_dep=t_dep, row_idx=row_idx,
                    rows=np.arange(r_lo, r_hi + 1))
