"""`Mixture`: p(t,v,x,z) = sum_i w_i p_i(t,v,x,z) over previously built components.

Behind the reference's API (/root/reference/popkinmocks/components/mixture.py).  The datacube
is linear in the density, so `evaluate_ybar` is one weighted sum of the children's cubes
(pkm_axpy_cubes <- mixture.py:35-48); the distribution and moment methods are host-side
compositions of the children's (accelerated) results.
"""
import numpy as np
from scipy import special

from . import base
from .. import engine


class Mixture(base.Component):
    """Args:
        cube: the shared `IFUCube`
        component_list (list): components defined on that cube
        weights (array): positive, summing to one
    """

    def __init__(self, cube=None, component_list=None, weights=None):
        self.cube = cube
        assert len(component_list) == len(weights)
        self.n_cmps = len(component_list)
        self.component_list = component_list
        weights = np.array(weights)
        assert np.all(weights > 0.0)
        assert np.isclose(np.sum(weights), 1.0)
        self.weights = weights

    def evaluate_ybar(self, device=0):
        """ybar = sum_i w_i ybar_i of the ALREADY evaluated children (takes no `batch`)."""
        self.ybar = engine.axpy_cubes([c.ybar for c in self.component_list], self.weights,
                                      device=device)

    def pixelate(self, device=0):
        """`Component(cube, self.get_log_p("tvxz", collapse_cmps=True))` built and kept on the
        device: the children are pixelated there (`ParametricComponent.pixelate`, or the stored
        array of an already pixelated child) and collapsed by pkm_logsumexp_cubes, so the
        5-D arrays of a large cube never exist on the host."""
        import torch
        terms = []
        for cmp in self.component_list:
            if hasattr(cmp, "pixelate"):
                terms.append(cmp.pixelate(device=device).log_p_tvxz)
            else:
                log_p = cmp.log_p_tvxz
                if not engine._is_cuda_tensor(log_p):
                    log_p = torch.from_numpy(np.ascontiguousarray(log_p, dtype=np.float64)).to(f"cuda:{device}")
                terms.append(log_p)
        log_p = engine.logsumexp_cubes(terms, self.weights, device=device)
        return base.Component(cube=self.cube, log_p_tvxz=log_p)

    # ------------------------------------------------------------------ distributions
    def get_p(self, which_dist, collapse_cmps=True, density=True, light_weighted=False):
        return np.exp(self.get_log_p(which_dist, collapse_cmps=collapse_cmps, density=density,
                                     light_weighted=light_weighted))

    def get_log_p(self, which_dist, collapse_cmps=True, density=True, light_weighted=False):
        """As `Component.get_log_p`; with `collapse_cmps=False` the leading axis indexes the
        mixture components (each term includes its weight)."""
        if "_" in which_dist:
            log_p = self._get_log_conditional_distribution(which_dist, density=density,
                                                           light_weighted=light_weighted)
        elif light_weighted:
            log_p = self._get_log_marginal_distribution_light_wtd(which_dist, density=density)
        else:
            log_p = self._get_log_marginal_distribution_mass_wtd(which_dist, density=density)
        if collapse_cmps:
            log_p = special.logsumexp(log_p, 0)
        return log_p

    def _get_log_marginal_distribution_mass_wtd(self, which_dist, density=True):
        terms = [np.log(w) + getattr(cmp, "_get_log_p_" + which_dist)(density=density, light_weighted=False)
                 for cmp, w in zip(self.component_list, self.weights)]
        return np.stack(terms, 0)

    def _get_log_marginal_distribution_light_wtd(self, which_dist, density=True):
        """Light weighting needs the (t,z) dependence of every term: start from the weighted
        joint over (t,[v],x,z), renormalise over all components, then reduce (mixture.py:166-205)."""
        lw = self.cube.ssps.light_weights
        if "v" in which_dist:
            log_P = self._get_log_marginal_distribution_mass_wtd("tvxz", density=False)
            log_P = log_P + np.log(lw[None, :, None, None, None, :])
        else:
            log_P = self._get_log_marginal_distribution_mass_wtd("txz", density=False)
            log_P = log_P + np.log(lw[None, :, None, None, :])
        log_P = log_P - special.logsumexp(log_P)
        if "t" not in which_dist:
            log_P = special.logsumexp(log_P, 1)
        if "x" not in which_dist:      # x before z: negative axes shift once z is removed
            log_P = special.logsumexp(log_P, (-3, -2))
        if "z" not in which_dist:
            log_P = special.logsumexp(log_P, -1)
        if density:
            log_P = log_P - np.log(self.cube.construct_volume_element(which_dist))
        return log_P

    def _get_log_conditional_distribution(self, which_dist, light_weighted=False, density=True):
        dependent, given = which_dist.split("_")
        joint = "".join(sorted(dependent + given))
        log_joint = self.get_log_p(joint, collapse_cmps=False, density=False, light_weighted=light_weighted)
        log_given = self.get_log_p(given, collapse_cmps=True, density=False, light_weighted=light_weighted)
        joint_axes = "i" + joint.replace("x", "xy")     # leading component axis
        given_axes = given.replace("x", "xy")
        source = [joint_axes.find(a) for a in given_axes]
        target = list(range(-len(given_axes), 0))
        with np.errstate(invalid="ignore"):
            out = np.moveaxis(log_joint, source, target) - log_given
        if density:
            out = out - np.log(self.cube.construct_volume_element(which_dist))
        return out

    # ------------------------------------------------------------------ moments
    def _get_light_reweightings(self):
        """Ratio of each component's light fraction to its mass fraction."""
        lw = self.cube.ssps.light_weights
        p_tz = self.get_p("tz", light_weighted=False, density=False)
        p_tz_i = np.array([c.get_p("tz", light_weighted=False, density=False) for c in self.component_list])
        return np.sum(lw * p_tz_i, (1, 2)) / np.sum(lw * p_tz)

    def _component_posteriors(self, conditioners, light_weighted, variable):
        """w_i p_i(conditioners), component axis first (mixture.py:268-286)."""
        w_p = np.array([c.get_p(conditioners, light_weighted=light_weighted, density=True)
                        for c in self.component_list])
        w_p = (w_p.T * self.weights).T
        return w_p[:, None] if variable == "x" else w_p

    def get_mean(self, which_dist, light_weighted=False):
        mu_i = np.array([c.get_mean(which_dist, light_weighted=light_weighted) for c in self.component_list])
        rw = self._get_light_reweightings() if light_weighted else np.ones(self.n_cmps)
        if "_" not in which_dist:
            return np.sum(self.weights * rw * mu_i.T, -1).T
        variable, conditioners = which_dist.split("_")
        w_p = self._component_posteriors(conditioners, light_weighted, variable)
        denom = self.get_p(conditioners, light_weighted=light_weighted, collapse_cmps=True, density=True)
        return np.sum(w_p.T * rw * mu_i.T, -1).T / denom

    def get_central_moment(self, which_dist, j, light_weighted=False):
        mu = self.get_mean(which_dist, light_weighted=light_weighted)
        mom_i = np.array([c.get_noncentral_moment(which_dist, j, mu, light_weighted=light_weighted)
                          for c in self.component_list])
        if "_" not in which_dist:
            mom_i = np.moveaxis(mom_i, 0, -1)
            rw = self._get_light_reweightings() if light_weighted else 1.0
            return np.sum(self.weights * rw * mom_i, -1)
        variable, conditioners = which_dist.split("_")
        w_p = self.get_p(conditioners, light_weighted=light_weighted, collapse_cmps=False, density=True)
        if variable == "x":
            w_p = w_p[:, None]
        return np.sum(w_p * mom_i, 0) / np.sum(w_p, 0)

    def get_noncentral_moment(self, which_dist, j, mu, light_weighted=False):
        raise NotImplementedError("Mixture exposes central moments only (as the reference does)")
