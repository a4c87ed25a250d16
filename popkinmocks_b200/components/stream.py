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
        exactly over each pixel and normalised over the cube (stream.py:28-90)."""
        assert np.min(theta_lims) >= -np.pi, "Angles must be in -pi<theta<pi'"
        assert np.max(theta_lims) <= np.pi, "Angles must be in -pi<theta<pi'"
        self.theta_lims = theta_lims
        self.nsmp = nsmp
        self.p_x_pars = dict(theta_lims=theta_lims, mu_r_lims=mu_r_lims, sig=sig, nsmp=nsmp)
        cube = self.cube
        th0, th1 = theta_lims
        theta = np.linspace(th0, th1, nsmp)
        radius = mu_r_lims[0] + (mu_r_lims[1] - mu_r_lims[0]) * ((theta - th0) / (th1 - th0))

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
        self.log_p_x_t = np.moveaxis(np.broadcast_to(log_p, (nt,) + log_p.shape), 0, -1)

    def set_t_dep(self, t_dep=3.0):
        """Spatially constant depletion time (stream.py:92-103)."""
        self.t_dep = t_dep * np.ones((self.cube.nx, self.cube.ny))
        self.set_p_z_tx()

    def set_mu_v(self, mu_v_lims=[-100, 100]):
        """Mean velocity varying linearly with position angle along the stream, constant
        beyond its ends; identical for all ages (a broadcast view, stream.py:105-127)."""
        theta = np.arctan2(self.yyp, self.xxp)
        if self.theta_lims[0] < self.theta_lims[1]:
            v_lo, v_hi = mu_v_lims
        else:
            v_hi, v_lo = mu_v_lims
        th_lo, th_hi = np.min(self.theta_lims), np.max(self.theta_lims)
        mu_v = np.zeros_like(theta)
        mu_v[theta <= th_lo] = v_lo
        mu_v[theta >= th_hi] = v_hi
        inside = (theta > th_lo) & (theta < th_hi)
        mu_v[inside] = (theta[inside] - th_lo) / (th_hi - th_lo) * (v_hi - v_lo)
        mu_v[inside] += v_lo
        self.mu_v_lims = mu_v_lims
        nt = len(self.cube.get_variable_values("t"))
        self.mu_v = np.broadcast_to(mu_v, (nt,) + mu_v.shape)

    def set_sig_v(self, sig_v=100.0):
        """Constant velocity dispersion (stream.py:129-138)."""
        nt = self.cube.ssps.par_dims[1]
        self.sig_v = sig_v * np.ones((nt, self.cube.nx, self.cube.ny))
