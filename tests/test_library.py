"""The C-ABI shared library: loads, exports every symbol of include/pkm_b200.h, host-only
helpers agree with SciPy, and compute entry points fail loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
from scipy import interpolate

from conftest import ROOT

from popkinmocks_b200 import _lib


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "pkm_b200.h")).read()
    return sorted(set(re.findall(r"PKM_API\s+[\w\s\*]+?\b(pkm_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/pkm_b200.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), "ctypes table and header disagree"
    assert lib.pkm_abi_version() == 1


@pytest.mark.parametrize("nv,nfo", [(41, 997), (21, 511), (101, 2454), (7, 64), (127, 5590)])
def test_spline_tables_equal_scipy_cubic_interp1d(nv, nfo):
    nfi = nv // 2 + 1
    A_inv, idx, w = _lib.spline_tables(nfi, nfo)
    W = np.zeros((nfo, nfi))
    for q in range(4):
        W += w[:, q:q + 1] * A_inv[idx + q]
    ref = interpolate.interp1d(np.linspace(0, 1, nfi), np.eye(nfi), axis=0, kind="cubic")(np.linspace(0, 1, nfo))
    assert np.max(np.abs(W - ref)) < 5e-15
    assert np.allclose(w.sum(1), 1.0, atol=1e-15)          # partition of unity (used by the kernel)
    assert np.array_equal(W[0], np.eye(nfi)[0]) or np.allclose(W[0], np.eye(nfi)[0], atol=1e-15)


@pytest.mark.parametrize("nv", [7, 21, 40, 41, 101])
def test_fold_tables_compose_roll_rfft_and_spline_solve(nv):
    v0 = nv // 2
    nfi = nv // 2 + 1
    A_inv, _, _ = _lib.spline_tables(nfi, 8)
    CR, CI, ip, im = _lib.fold_tables(nv, v0, 25.0)
    p = np.random.default_rng(nv).random(nv)
    c_ref = A_inv @ (np.fft.rfft(np.roll(p, nv - v0)) * 25.0)
    e = p[ip] + p[im]
    o = (p[ip] - p[im])[1:1 + CI.shape[1]]
    c = CR @ e + 1j * (CI @ o)
    assert np.max(np.abs(c - c_ref)) / np.max(np.abs(c_ref)) < 1e-14
    assert np.array_equal(ip, (np.arange(nfi) + v0) % nv) and np.array_equal(im, (v0 - np.arange(nfi)) % nv)


def test_invalid_arguments_return_status_not_exceptions():
    lib = _lib.load()
    rc = lib.pkm_spline_tables(3, 10, None, None, None)      # nfi < 4: cubic spline undefined
    assert rc == _lib.PKM_ERR_INVALID
    assert b"at least 4" in lib.pkm_last_error()


@pytest.mark.skipif(_lib.device_count() > 0, reason="checks behaviour WITHOUT a GPU")
def test_no_cpu_fallback_without_a_gpu():
    from conftest import golden_cube
    import popkinmocks_b200 as pkm
    cube = golden_cube("g2_rebin_nv21_prime")
    comp = pkm.components.Component(cube=cube, log_p_tvxz=np.zeros(cube.get_distribution_shape("tvxz")))
    with pytest.raises(_lib.PkmError, match="no CUDA device"):
        comp.evaluate_ybar()
    with pytest.raises(_lib.PkmError):
        comp.get_p("t")
    out = ctypes.c_double()
    assert _lib.load().pkm_fp64_fma_peak(0, 1, ctypes.byref(out)) == _lib.PKM_ERR_CUDA
    # device-side pixelation of a parametric component, through the class and through the C ABI
    from test_host_models import BUILDERS
    with pytest.raises(_lib.PkmError, match="no CUDA device"):
        BUILDERS["d1_"](cube).pixelate()
    from popkinmocks_b200 import engine
    with pytest.raises(_lib.PkmError, match="no CUDA device"):
        engine.logsumexp_cubes([np.zeros(4), np.zeros(4)], [0.5, 0.5])
    lw, mu, sig = np.zeros((2, 3, 4)), np.zeros((2, 3)), np.ones((2, 3))
    v_edg, res = np.linspace(-1.0, 1.0, 6), np.empty((2, 5, 3, 4))
    rc = _lib.load().pkm_pixelate_gaussian(_lib.ptr(lw), _lib.ptr(mu), _lib.ptr(sig), _lib.ptr(v_edg),
                                           2, 5, 3, 4, 1, _lib.ptr(res), 0, None)
    assert rc == _lib.PKM_ERR_CUDA
