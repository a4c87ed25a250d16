"""Observation noise models: yobs(x, w) = ybar(x, w) + eps(x, w).

Behind the reference's API (/root/reference/popkinmocks/noise.py:6-71): `ConstantSNR` and
`ShotNoise`, each with `get_sigma(snr=)`, `sample(snr=)` and `get_noisy_data(snr=)`.  The
sigma maps and the Gaussian draws are produced on the device (pkm_noise: max-reduction,
elementwise sigma, Philox4x32-10 + Box-Muller).  The reference draws from NumPy's global
legacy stream through SciPy; here the Philox key is drawn from that same global stream, so
`np.random.seed(...)` still makes runs reproducible (the realisations themselves differ -
parity is statistical).
"""
from abc import abstractmethod

import numpy as np

from . import engine


class Noise(object):
    """Args:
        component: a component with a pre-evaluated `ybar`
    """

    def __init__(self, component):
        self.component = component

    def get_noisy_data(self, **kwargs):
        return self.component.ybar + self.sample(**kwargs)

    @abstractmethod
    def sample(self):
        pass


class Diagonal(Noise):
    """Independent Gaussian noise per voxel, eps ~ N(0, sigma(x, w))."""

    _mode = None

    def get_sigma(self, snr=100.0):
        _, sigma = engine.noise(self.component.ybar, snr, self._mode, want_sample=False)
        return sigma

    def sample(self, snr=100.0):
        seed = int(np.random.randint(0, 2 ** 31 - 1)) | (int(np.random.randint(0, 2 ** 31 - 1)) << 31)
        yobs, _ = engine.noise(self.component.ybar, snr, self._mode, seed=seed, want_sigma=False)
        return yobs - self.component.ybar

    def get_noisy_data(self, snr=100.0):
        seed = int(np.random.randint(0, 2 ** 31 - 1)) | (int(np.random.randint(0, 2 ** 31 - 1)) << 31)
        yobs, _ = engine.noise(self.component.ybar, snr, self._mode, seed=seed, want_sigma=False)
        return yobs


class ConstantSNR(Diagonal):
    """sigma(x, w) = ybar(x, w) / snr  (noise.py:41-54)."""
    _mode = 1


class ShotNoise(Diagonal):
    """sigma(x, w) = K sqrt(ybar(x, w)) with K = sqrt(max ybar) / snr, i.e. `snr` is reached in
    the brightest voxel (noise.py:57-71).  Negative ybar gives NaN, as in the reference."""
    _mode = 0
