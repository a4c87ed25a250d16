"""pytest configuration: the `gpu` marker and shared golden-fixture helpers."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


# golden file -> kwargs that rebuild the same cube with the product's host classes
GOLDEN_CUBES = {
    "g1_conftest_full_nv41": dict(nx1=3, nx2=2, nv=41, ssp={}),
    "g2_rebin_nv21_prime": dict(nx1=4, nx2=5, nv=21, ssp=dict(age_rebin=6, z_rebin=2)),
    "g3_full_nv101_odd": dict(nx1=1, nx2=2, nv=101, ssp={}),
    "g4_even_nv40": dict(nx1=2, nx2=2, nv=40, ssp=dict(age_rebin=4, z_rebin=3)),
    # the reference's pad="ppxf" option: n_fft = 1024 > n_pix = 1021 (model_grids.py:263-281)
    "g5_ppxf_pad_nv21": dict(nx1=2, nx2=3, nv=21, ssp=dict(age_rebin=6, z_rebin=2), pad="ppxf"),
}

_cache = {}


def load_golden(name):
    if ("g", name) not in _cache:
        _cache[("g", name)] = np.load(os.path.join(GOLDEN, name + ".npz"))
    return _cache[("g", name)]


def build_cube(name):
    """A fresh popkinmocks_b200 cube with the golden file's configuration."""
    import popkinmocks_b200 as pkm
    cfg = GOLDEN_CUBES[name]
    ssps = pkm.milesSSPs(**cfg["ssp"])
    cube = pkm.ifu_cube.IFUCube(ssps=ssps, nx1=cfg["nx1"], nx2=cfg["nx2"], nv=cfg["nv"],
                                vrng=(-1000.0, 1000.0))
    if cfg.get("pad"):
        cube.ssps._calculate_fourier_transform(pad=cfg["pad"])
    return cube


def golden_cube(name):
    if ("c", name) not in _cache:
        _cache[("c", name)] = build_cube(name)
    return _cache[("c", name)]


def rel_err(a, b):
    """max-norm relative error, the parity metric of SURVEY.md section 8c."""
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


@pytest.fixture(scope="session")
def gpu_ok():
    from popkinmocks_b200 import _lib
    if _lib.device_count() <= 0:
        pytest.fail("no CUDA device visible: gpu-marked tests must run on the B200 box")
    return True
