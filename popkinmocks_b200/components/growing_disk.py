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
        in age between them (growing_disk.py:30-71)."""
        q_lims, rc_lims, alpha_lims = (np.array(v) for v in (q_lims, rc_lims, alpha_lims))
        assert np.all(q_lims > 0.0)
        assert np.all(rc_lims > 0.0)
        assert np.all(alpha_lims >= 0.0)
        self.x_t_pars = dict(q_lims=q_lims, rc_lims=rc_lims, alpha_lims=alpha_lims)
        q = _per_age(self.linear_interpolate_t(*q_lims))
        rc = _per_age(self.linear_interpolate_t(*rc_lims))
        alpha = _per_age(self.linear_interpolate_t(*alpha_lims))
        log_rho = -alpha * np.log(self._elliptical_radius(q) + rc)            # [t, x1, x2]
        log_pixel = np.log(self.cube.dx) + np.log(self.cube.dy)
        log_norm = special.logsumexp(log_rho + log_pixel, (1, 2))
        self.log_p_x_t = np.moveaxis(log_rho - _per_age(log_norm), 0, -1)      # [x1, x2, t]
        self.p_x_t = np.exp(self.log_p_x_t)

    def set_t_dep(self, q=0.1, alpha=1.5, t_dep_in=0.5, t_dep_out=6.0, t_dep_min=0.1,
                  t_dep_max=13.0):
        """Depletion-time map: power law in radius from `t_dep_in` towards `t_dep_out`,
        saturating at `t_dep_out` (growing_disk.py:73-120)."""
        assert q > 0.0 and alpha >= 0.0
        assert t_dep_min >= 0.1 and t_dep_max <= 13.0
        assert t_dep_min <= t_dep_in <= t_dep_max
        assert t_dep_min <= t_dep_out <= t_dep_max
        scale = np.max(np.abs(self.xxp))
        t_dep = t_dep_in + (t_dep_out - t_dep_in) * (self._elliptical_radius(q) / scale) ** alpha
        beyond = t_dep > t_dep_out if t_dep_in <= t_dep_out else t_dep < t_dep_out
        t_dep[beyond] = t_dep_out
        self.t_dep_pars = dict(q=q, alpha=alpha, t_dep_in=t_dep_in, t_dep_out=t_dep_out)
        self.t_dep = t_dep
        self.set_p_z_tx()

    def set_mu_v(self, q_lims=(0.5, 0.5), rmax_lims=(0.1, 1.0), vmax_lims=(50.0, 250.0)):
        """Rotation maps mu_v(t,x) = K r / (r + rc)^3 cos(theta), peaking at `vmax` at radius
        `rmax` (rc = 2 rmax), on elliptical radii with flattening q (growing_disk.py:122-170)."""
        q_lims, rmax_lims, vmax_lims = (np.array(v) for v in (q_lims, rmax_lims, vmax_lims))
        assert np.all(q_lims > 0.0)
        signs = np.sign(vmax_lims)
        assert np.all(np.isin(signs, [0, 1])) or np.all(np.isin(signs, [0, -1]))
        self.mu_v_pars = dict(q_lims=q_lims, rmax_lims=rmax_lims, vmax_lims=vmax_lims)
        q = self.linear_interpolate_t(*q_lims)
        rmax = self.linear_interpolate_t(*rmax_lims)
        vmax = self.linear_interpolate_t(*vmax_lims)
        rc = 2.0 * rmax
        K = vmax * (rmax + rc) ** 3 / rmax
        q, rc, K = _per_age(q), _per_age(rc), _per_age(K)
        theta = np.arctan2(self.yyp / q, self.xxp)
        r = self._elliptical_radius(q)
        self.mu_v = K * r / (r + rc) ** 3 * np.cos(theta)

    def set_sig_v(self, q_lims=(0.5, 0.1), alpha_lims=(1.5, 2.5), sig_v_in_lims=(50.0, 250.0),
                  sig_v_out_lims=(10.0, 50), sigma_max=500.0, sigma_min=10.0):
        """Dispersion maps: power law in radius from `sig_v_in` towards `sig_v_out`, saturating
        at `sig_v_out` (growing_disk.py:172-239)."""
        q_lims, alpha_lims, sig_v_in_lims, sig_v_out_lims = (
            np.array(v) for v in (q_lims, alpha_lims, sig_v_in_lims, sig_v_out_lims))
        assert np.all(q_lims > 0.0)
        assert np.all(alpha_lims >= 0.0)
        assert np.all(sig_v_in_lims > 0.0)
        assert np.all(sig_v_out_lims > 0.0)
        self.sig_v_pars = dict(q_lims=q_lims, alpha_lims=alpha_lims, sig_v_in_lims=sig_v_in_lims,
                               sig_v_out_lims=sig_v_out_lims)
        q = _per_age(self.linear_interpolate_t(*q_lims))
        alpha = _per_age(self.linear_interpolate_t(*alpha_lims))
        sig_in = _per_age(self.linear_interpolate_t(*sig_v_in_lims))
        sig_out = _per_age(self.linear_interpolate_t(*sig_v_out_lims))
        scale = np.max(np.abs(self.xxp))
        sigma = sig_in + (sig_out - sig_in) * (self._elliptical_radius(q) / scale) ** alpha
        rising = sig_out > sig_in                                             # [t, 1, 1]
        beyond = np.where(rising, sigma > sig_out, sigma < sig_out)
        self.sig_v = np.where(beyond, np.broadcast_to(sig_out, sigma.shape), sigma)
