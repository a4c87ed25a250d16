"""`Stream`: a thin stellar stream along an arc, with age-independent spatial density.

Host-side parameter maps behind the reference's API
(/root/reference/popkinmocks/components/stream.py:6-138).
"""
import numpy as np
from scipy import special, stats

from . import parametric


class Stream(parametric.ParametricComponent):

    def set_p_x_t(self, theta_lims=[0.0, np.pi / 2.0], mu_r_lims=[0.7, 0.1], sig=0.03, nsmp=75):
        """Spatial density: an equally weighted sum of `nsmp` isotropic Gaussians (width `sig`)
        along the curve r(theta) linear between `mu_r_lims` over `theta_lims`, integrated
        exactly over each pixel and normalised over the cube (stream.py:28-90).  `log_p_x_t`
        [x1, x2, t] is computed on first access."""
        assert np.min(theta_lims) >= -np.pi, "Angles must be in -pi<theta<pi'"
        assert np.max(theta_lims) <= np.pi, "Angles must be in -pi<theta<pi'"
        self.theta_lims = theta_lims
        self.nsmp = nsmp
        self.p_x_pars = dict(theta_lims=theta_lims, mu_r_lims=mu_r_lims, sig=sig, nsmp=nsmp)
        self._set_lazy(("log_p_x_t",), "_host_p_x_t")

    def _arc(self):
        """Sample angles and radii along the stream."""
        pars = self.p_x_pars
        th0, th1 = pars["theta_lims"]
        theta = np.linspace(th0, th1, pars["nsmp"])
        mu_r = pars["mu_r_lims"]
        radius = mu_r[0] + (mu_r[1] - mu_r[0]) * ((theta - th0) / (th1 - th0))
        return theta, radius

    def _host_p_x_t(self):
        cube = self.cube
        sig = self.p_x_pars["sig"]
        theta, radius = self._arc()

        def log_pixel_mass(coord, half_width, centres):
            # stats.norm(centres, sig).logcdf(.) is log_ndtr((. - centres)/sig); evaluated once per
            # distinct coordinate value (an unrotated cube has nx1 of them, not nx1*nx2) and
            # gathered - the same arithmetic per element as the reference's full [x1, x2, nsmp] arrays
            values, inverse = np.unique(coord.ravel(), return_inverse=True)
            with np.errstate(divide="ignore"):
                hi = special.log_ndtr((values[:, None] + half_width - centres) / sig)
                lo = special.log_ndtr((values[:, None] - half_width - centres) / sig)
            return parametric._log_diff_exp(hi, lo)[inverse.reshape(coord.shape)]   # [x1, x2, nsmp]

        log_p = (log_pixel_mass(self.xxp, cube.dx / 2.0, radius * np.cos(theta))
                 + log_pixel_mass(self.yyp, cube.dy / 2.0, radius * np.sin(theta)))
        log_p = special.logsumexp(log_p, -1)
        log_p -= special.logsumexp(log_p)
        log_p -= np.log(cube.dx) + np.log(cube.dy)
        nt = len(cube.get_variable_values("t"))
        self.__dict__["log_p_x_t"] = np.moveaxis(np.broadcast_to(log_p, (nt,) + log_p.shape), 0, -1)

    def set_t_dep(self, t_dep=3.0):
        """Spatially constant depletion time (stream.py:92-103)."""
        self.t_dep_value = float(t_dep)
        self._set_lazy(("t_dep",), "_host_t_dep")
        self._set_lazy(("log_p_z_tx", "p_z_tx"), "set_p_z_tx")

    def _host_t_dep(self):
        self.__dict__["t_dep"] = self.t_dep_value * np.ones((self.cube.nx, self.cube.ny))

    def set_mu_v(self, mu_v_lims=[-100, 100]):
        """Mean velocity varying linearly with position angle along the stream, constant
        beyond its ends; identical for all ages (a broadcast view, stream.py:105-127)."""
        self.mu_v_lims = mu_v_lims
        self._set_lazy(("mu_v",), "_host_mu_v")

    def _mu_v_ends(self):
        if self.theta_lims[0] < self.theta_lims[1]:
            v_lo, v_hi = self.mu_v_lims
        else:
            v_hi, v_lo = self.mu_v_lims
        return float(np.min(self.theta_lims)), float(np.max(self.theta_lims)), float(v_lo), float(v_hi)

    def _host_mu_v(self):
        theta = np.arctan2(self.yyp, self.xxp)
        th_lo, th_hi, v_lo, v_hi = self._mu_v_ends()
        mu_v = np.zeros_like(theta)
        mu_v[theta <= th_lo] = v_lo
        mu_v[theta >= th_hi] = v_hi
        inside = (theta > th_lo) & (theta < th_hi)
        mu_v[inside] = (theta[inside] - th_lo) / (th_hi - th_lo) * (v_hi - v_lo)
        mu_v[inside] += v_lo
        nt = len(self.cube.get_variable_values("t"))
        self.__dict__["mu_v"] = np.broadcast_to(mu_v, (nt,) + mu_v.shape)

    def set_sig_v(self, sig_v=100.0):
        """Constant velocity dispersion (stream.py:129-138)."""
        self.sig_v_value = float(sig_v)
        self._set_lazy(("sig_v",), "_host_sig_v")

    def _host_sig_v(self):
        nt = self.cube.ssps.par_dims[1]
        self.__dict__["sig_v"] = self.sig_v_value * np.ones((nt, self.cube.nx, self.cube.ny))

    # ------------------------------------------------------------------ the same maps, in HBM
    def _build_device_maps(self, device):
        """p(x), mu_v generated on the device (pkm_stream_maps); sig_v and the depletion time are
        constants, so mu_v / sig_v are stride-0 views over the ages and one enrichment row is used."""
        for attr in ("p_x_pars", "t_dep_value", "mu_v_lims", "sig_v_value", "t_pars"):
            if attr not in self.__dict__:
                return None
        import torch
        from .. import _lib
        cube = self.cube
        nt = cube.ssps.par_dims[1]
        theta, radius = self._arc()
        cxs, cys = _lib.as_f64(radius * np.cos(theta)), _lib.as_f64(radius * np.sin(theta))
        th_lo, th_hi, v_lo, v_hi = self._mu_v_ends()
        dev = f"cuda:{device}"
        log_p_x = torch.empty((cube.nx, cube.ny), dtype=torch.float64, device=dev)
        mu = torch.empty((cube.nx, cube.ny), dtype=torch.float64, device=dev)
        x1, x2 = _lib.as_f64(cube.x), _lib.as_f64(cube.y)
        log_dxdy = float(np.log(cube.dx) + np.log(cube.dy))
        _lib.check(_lib.load().pkm_stream_maps(
            _lib.ptr(x1), _lib.ptr(x2), cube.nx, cube.ny, float(self.center[0]), float(self.center[1]),
            float(self.rotation), _lib.ptr(cxs), _lib.ptr(cys), int(cxs.size), float(self.p_x_pars["sig"]),
            float(cube.dx / 2.0), float(cube.dy / 2.0), th_lo, th_hi, v_lo, v_hi, _lib.ptr(log_p_x), _lib.ptr(mu),
            log_dxdy, int(device), None))
        log_p_x -= log_dxdy                            # pixel mass -> density
        grid = self.z_t_interp_grid["t_dep"]
        row = int(np.argmin(np.abs(self.t_dep_value - grid)))
        sig = torch.full((1, 1, 1), self.sig_v_value, dtype=torch.float64, device=dev)
        return dict(log_p_x=log_p_x, px_per_age=0, mu_v=mu[None].expand(nt, cube.nx, cube.ny),
                    sig_v=sig.expand(nt, cube.nx, cube.ny), row_idx=None, rows=np.array([row]))
