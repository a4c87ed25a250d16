"""GPU parity tests: the CUDA hot path (through the C ABI) against reference outputs.

tests/golden/*.npz were produced by executing the unmodified reference
(oracle/make_golden.py); the NumPy restatement (oracle/restatement.py) is used for inputs
the golden files do not cover.  Tolerances: fp64 path <= 1e-10 max-norm relative
(BASELINE.json north_star); indexing / layout exact.
"""
import numpy as np
import pytest

from conftest import build_cube, golden_cube, load_golden, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _plan(name):
    from popkinmocks_b200 import engine
    return engine.get_plan(golden_cube(name))


@pytest.mark.parametrize("name,tags", [
    ("g1_conftest_full_nv41", ["d1_", "d2_", "s1_"]),
    ("g2_rebin_nv21_prime", ["d1_", "s1_"]),
    ("g3_full_nv101_odd", ["d1_"]),
    ("g4_even_nv40", ["d2_"]),
    ("g5_ppxf_pad_nv21", ["d1_"]),          # pad="ppxf": n_fft = 1024 > n_pix = 1021
])
def test_gaussian_path_matches_reference(gpu_ok, name, tags):
    g = load_golden(name)
    plan = _plan(name)
    for tag in tags:
        y = plan.ybar_gaussian(g[tag + "P_txz"], g[tag + "mu_v"], g[tag + "sig_v"])
        ref = g[tag + "ybar"]
        assert y.shape == ref.shape
        assert rel_err(y, ref) < TOL, (name, tag, rel_err(y, ref))


@pytest.mark.parametrize("name,tags", [
    ("g1_conftest_full_nv41", ["b_"]),
    ("g2_rebin_nv21_prime", ["b_", "wn_"]),
    ("g3_full_nv101_odd", ["wn_"]),
    ("g5_ppxf_pad_nv21", ["wn_"]),          # pad="ppxf" (model_grids.py:263-281)
])
def test_density_path_matches_reference(gpu_ok, name, tags):
    g = load_golden(name)
    plan = _plan(name)
    for tag in tags:
        y = plan.ybar_density(g[tag + "log_p_tvxz"], is_log=True)
        ref = g[tag + "ybar"]
        assert y.shape == ref.shape
        assert rel_err(y, ref) < TOL, (name, tag, rel_err(y, ref))
        # linear-density input gives the same cube
        with np.errstate(over="ignore"):
            y2 = plan.ybar_density(np.exp(g[tag + "log_p_tvxz"]), is_log=False)
        assert rel_err(y2, ref) < TOL


@pytest.mark.parametrize("env", [
    {"PKM_COEFF_KERNEL": "0"},                                  # staged kernel, 2-D tensor copies
    {"PKM_COEFF_TMA_MODE": "1"},                                # staged kernel, 1-D bulk copies
    {"PKM_COEFF_TMA_MODE": "0"},                                # staged kernel, plain loads
    {"PKM_COEFF_KERNEL": "0", "PKM_COEFF_GROUPS": "1"},         # staged kernel, one warp group
    {"PKM_COEFF_FLOW": "0"},                                    # octet kernel with CTA barriers + TMA tensor store
    {"PKM_COEFF_APF": "0"},                                     # barrier-free kernel with a producer warp, one A-fragment set
])
def test_coefficient_kernel_fallbacks_match_reference(gpu_ok, monkeypatch, env):
    """The default path uses the octet coefficient kernel; the staged kernel and its load modes
    (taken for shapes the tensor maps cannot describe) are forced here through the developer
    switches and held to the same parity bar, on white-noise and pixelated-disk densities."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    for name, tag in (("g2_rebin_nv21_prime", "wn_"), ("g1_conftest_full_nv41", "b_"), ("g3_full_nv101_odd", "wn_")):
        g = load_golden(name)
        y = _plan(name).ybar_density(g[tag + "log_p_tvxz"], is_log=True)
        assert rel_err(y, g[tag + "ybar"]) < TOL, (env, name, rel_err(y, g[tag + "ybar"]))


@pytest.mark.parametrize("env", [{}, {"PKM_COEFF_APF": "0"}, {"PKM_COEFF_FLOW": "0"}, {"PKM_COEFF_KERNEL": "0"},
                                 {"PKM_COEFF_GENERIC": "1"}])
def test_density_special_values(gpu_ok, monkeypatch, env):
    """log p = -inf and log p < -708 are exact zeros of the density, NaN poisons its own pixel only
    (np.exp semantics, base.py:841): the rare-value branch of the fused table exponential, through
    every coefficient kernel."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    g = load_golden("g3_full_nv101_odd")
    plan = _plan("g3_full_nv101_odd")
    logp = np.array(g["wn_log_p_tvxz"], copy=True)
    rng = np.random.default_rng(3)
    logp[rng.random(logp.shape) < 0.05] = -np.inf
    logp[rng.random(logp.shape) < 0.05] = -800.0
    logp[0, 3] = -np.inf                                    # a whole velocity plane of one age
    ref = plan.ybar_density(np.exp(logp), is_log=False)
    y = plan.ybar_density(logp, is_log=True)
    assert np.isfinite(y).all() and rel_err(y, ref) < TOL
    nx1, nx2 = y.shape[1:]
    bad = logp.copy()
    bad[5, 7, nx1 - 1, nx2 - 1, 2] = np.nan
    yb = plan.ybar_density(bad, is_log=True)
    assert np.isnan(yb[:, nx1 - 1, nx2 - 1]).all()
    keep = np.ones((nx1, nx2), bool)
    keep[nx1 - 1, nx2 - 1] = False
    assert np.array_equal(yb[:, keep], y[:, keep])


def test_density_even_nv_raises_index_error(gpu_ok):
    g = load_golden("g4_even_nv40")
    assert str(g["b_error"]) == "IndexError"
    plan = _plan("g4_even_nv40")
    cube = golden_cube("g4_even_nv40")
    p = np.zeros(cube.get_distribution_shape("tvxz"))
    with pytest.raises(IndexError):
        plan.ybar_density(p, is_log=False)


@pytest.mark.parametrize("name", ["g1_conftest_full_nv41", "g2_rebin_nv21_prime", "g3_full_nv101_odd"])
def test_irfft_matches_numpy(gpu_ok, name):
    plan = _plan(name)
    rng = np.random.default_rng(7)
    Y = rng.standard_normal((5, plan.nfo)) + 1j * rng.standard_normal((5, plan.nfo))
    ref = np.fft.irfft(Y, plan.n_fft, axis=1)
    assert rel_err(plan.irfft(Y), ref) < 1e-13
    assert rel_err(plan.irfft(Y, naive=True), ref) < 1e-12


@pytest.mark.parametrize("lmd,n_fft,M", [
    ((5000, 5027), 17, 32), ((5000, 5032), 20, 32),      # M = 2 * 16: the fused middle pass is the whole transform
    ((5000, 5500), 300, 512), ((5000, 5480), 288, 512),  # M = 2 * 16^2
])
def test_irfft_halves_kernel_sizes(gpu_ok, lmd, n_fft, M):
    """The M = 2 * 16^p inverse-FFT kernel (k_irfft_czt_h; M = 8192 is the headline n_fft = 4907,
    covered by test_irfft_matches_numpy on g3) at its other sizes, odd and even n_fft."""
    import popkinmocks_b200 as pkm
    ssps = pkm.milesSSPs(lmd_min=lmd[0], lmd_max=lmd[1], age_rebin=8, z_rebin=4)
    cube = pkm.ifu_cube.IFUCube(ssps=ssps, nx1=2, nx2=2, nv=21, vrng=(-1000.0, 1000.0))
    plan = pkm.engine.get_plan(cube)
    assert plan.n_fft == n_fft
    rng = np.random.default_rng(n_fft)
    Y = rng.standard_normal((300, plan.nfo)) + 1j * rng.standard_normal((300, plan.nfo))
    ref = np.fft.irfft(Y, n_fft, axis=1)
    assert rel_err(plan.irfft(Y), ref) < 1e-13
    assert rel_err(plan.irfft(Y, scale=0.25), 0.25 * ref) < 1e-13


def test_irfft_generic_kernel_still_matches(gpu_ok, monkeypatch):
    """PKM_IRFFT_HALVES=0 keeps the general radix-16/4/2 kernel reachable at the headline size."""
    plan = _plan("g3_full_nv101_odd")
    rng = np.random.default_rng(11)
    Y = rng.standard_normal((3, plan.nfo)) + 1j * rng.standard_normal((3, plan.nfo))
    ref = np.fft.irfft(Y, plan.n_fft, axis=1)
    fast = plan.irfft(Y)
    monkeypatch.setenv("PKM_IRFFT_HALVES", "0")
    slow = plan.irfft(Y)
    assert rel_err(fast, ref) < 1e-13 and rel_err(slow, ref) < 1e-13


def test_row_ranges_and_pixel_major_layout(gpu_ok):
    from popkinmocks_b200 import _lib
    g = load_golden("g2_rebin_nv21_prime")
    plan = _plan("g2_rebin_nv21_prime")
    ref = g["wn_ybar"]                                   # [n_fft, 4, 5]
    logp = g["wn_log_p_tvxz"]
    part = plan.ybar_density(logp, rows=(1, 3))
    assert part.shape == (plan.n_fft, 2, 5)
    assert rel_err(part, ref[:, 1:3]) < TOL
    pm = plan.ybar_density(logp, rows=(2, 4), layout=_lib.LAYOUT_PIXEL_MAJOR)
    assert pm.shape == (10, plan.n_fft)
    assert rel_err(pm.reshape(2, 5, -1), np.moveaxis(ref[:, 2:4], 0, -1)) < TOL
    pa = plan.ybar_gaussian(g["d1_P_txz"], g["d1_mu_v"], g["d1_sig_v"], rows=(3, 4))
    assert rel_err(pa, g["d1_ybar"][:, 3:4]) < TOL


def test_small_workspace_tiles_pixels(gpu_ok):
    g = load_golden("g2_rebin_nv21_prime")
    plan = _plan("g2_rebin_nv21_prime")
    plan.set_workspace_limit(1 << 16)      # forces 16-pixel tiles -> 2 tiles for 20 pixels
    try:
        y = plan.ybar_density(g["wn_log_p_tvxz"])
        ya = plan.ybar_gaussian(g["s1_P_txz"], g["s1_mu_v"], g["s1_sig_v"])
    finally:
        plan.set_workspace_limit(0)
    assert rel_err(y, g["wn_ybar"]) < TOL
    assert rel_err(ya, g["s1_ybar"]) < TOL


def test_device_resident_inputs(gpu_ok):
    import torch
    g = load_golden("g2_rebin_nv21_prime")
    plan = _plan("g2_rebin_nv21_prime")
    d = torch.from_numpy(g["wn_log_p_tvxz"]).cuda()
    y = plan.ybar_density(d)
    torch.cuda.synchronize()
    assert y.is_cuda and tuple(y.shape) == g["wn_ybar"].shape
    assert rel_err(y.cpu().numpy(), g["wn_ybar"]) < TOL
    P, mu, sg = (torch.from_numpy(np.ascontiguousarray(g["s1_" + k])).cuda() for k in ("P_txz", "mu_v", "sig_v"))
    ya = plan.ybar_gaussian(P, mu, sg)
    torch.cuda.synchronize()
    assert rel_err(ya.cpu().numpy(), g["s1_ybar"]) < TOL


def test_mixture_axpy(gpu_ok):
    from popkinmocks_b200 import engine
    g = load_golden("g1_conftest_full_nv41")
    out = engine.axpy_cubes([g["d1_ybar"], g["d2_ybar"], g["s1_ybar"]], g["mix_weights"])
    assert rel_err(out, g["mix_ybar"]) < 1e-15


def test_headline_shape_properties(gpu_ok):
    """Full MILES grid, nv=101 (n_fft=4907 = 7*701), 48x40 pixels = several pixel tiles and a
    ragged last one.  The reference cannot run this size (25 MB/pixel); checked through
    size-independent properties: spot pixels against the oracle's literal algorithm, linearity
    in the density, and independence of the pixel tiling / row sharding."""
    import torch
    from oracle import restatement as R
    import popkinmocks_b200 as pkm
    ssps = pkm.milesSSPs()
    nx1, nx2, nv = 48, 40, 101
    cube = pkm.ifu_cube.IFUCube(ssps=ssps, nx1=nx1, nx2=nx2, nv=nv, vrng=(-1000.0, 1000.0))
    plan = pkm.engine.get_plan(cube)
    nz, nt = ssps.par_dims
    gen = torch.Generator(device="cuda").manual_seed(5)
    p1 = torch.rand((nt, nv, nx1, nx2, nz), dtype=torch.float64, device="cuda", generator=gen)
    p2 = torch.rand((nt, nv, nx1, nx2, nz), dtype=torch.float64, device="cuda", generator=gen) ** 3
    y1 = plan.ybar_density(p1, is_log=False)
    y2 = plan.ybar_density(p2, is_log=False)
    # (i) spot pixels vs the oracle (reference order: irfft per (t,z))
    y1h = y1.cpu().numpy()
    for (i, j) in [(0, 0), (17, 31), (47, 39), (23, 8)]:
        sub = p1[:, :, i:i + 1, j:j + 1, :].cpu().numpy()
        ref = R.ybar_density_literal(sub, cube.v_edg, ssps.FXw, ssps.par_dims, ssps.n_fft, ssps.dv,
                                     ssps.delta_t, ssps.delta_z, cube.dx, cube.dy)
        assert rel_err(y1h[:, i, j], ref[:, 0, 0]) < TOL, (i, j)
    # (ii) linearity
    y12 = plan.ybar_density(0.25 * p1 + 3.0 * p2, is_log=False)
    lin = 0.25 * y1 + 3.0 * y2
    assert float((y12 - lin).abs().max() / lin.abs().max()) < 1e-12
    # (iii) log input == linear input; tiling and row sharding do not change any value
    ylog = plan.ybar_density(torch.log(p1), is_log=True)
    assert float((ylog - y1).abs().max() / y1.abs().max()) < 1e-12
    plan.set_workspace_limit(200 << 20)          # ~330-pixel tiles
    try:
        ytile = plan.ybar_density(p1, is_log=False)
    finally:
        plan.set_workspace_limit(0)
    assert torch.equal(ytile, y1)
    rows = plan.ybar_density(p1, is_log=False, rows=(11, 29))
    assert torch.equal(rows, y1[:, 11:29])
    # Gaussian path on the same cube: spot pixels vs the oracle
    P = torch.rand((nt, nx1, nx2, nz), dtype=torch.float64, device="cuda", generator=gen)
    P /= P.sum()
    mu = 400.0 * (torch.rand((nt, nx1, nx2), dtype=torch.float64, device="cuda", generator=gen) - 0.5)
    sg = 20.0 + 200.0 * torch.rand((nt, nx1, nx2), dtype=torch.float64, device="cuda", generator=gen)
    ya = plan.ybar_gaussian(P, mu, sg).cpu().numpy()
    for (i, j) in [(0, 0), (31, 17), (47, 39)]:
        ref = R.ybar_gaussian(P[:, i:i + 1, j:j + 1].cpu().numpy(), mu[:, i:i + 1, j:j + 1].cpu().numpy(),
                              sg[:, i:i + 1, j:j + 1].cpu().numpy(), ssps.FXw, ssps.par_dims, ssps.n_fft, ssps.dv)
        assert rel_err(ya[:, i, j], ref[:, 0, 0]) < TOL, (i, j)


@pytest.mark.parametrize("name,tags", [("g1_conftest_full_nv41", ["b_"]), ("g2_rebin_nv21_prime", ["b_", "wn_"]),
                                       ("g3_full_nv101_odd", ["wn_"])])
def test_float32_option_within_1e5(gpu_ok, name, tags):
    """Optional float32 arithmetic of the contraction: BASELINE.json tolerance 1e-5 max-norm."""
    from popkinmocks_b200 import engine
    g = load_golden(name)
    plan32 = engine.get_plan(golden_cube(name), precision="float32")
    plan64 = engine.get_plan(golden_cube(name))
    assert plan32 is not plan64
    for tag in tags:
        y32 = plan32.ybar_density(g[tag + "log_p_tvxz"], is_log=True)
        err = rel_err(y32, g[tag + "ybar"])
        assert err < 1e-5, (name, tag, err)
        assert err > 1e-12      # it really is the float32 path


def test_output_embedded_in_a_larger_cube(gpu_ok):
    """pkm_ybar_density_into_cube: rows written at an offset of a larger [n_fft, nx1, nx2] buffer
    (the multi-GPU path stores into rank 0's cube this way, through a peer pointer)."""
    import ctypes as C
    import torch
    from popkinmocks_b200 import _lib
    g = load_golden("g2_rebin_nv21_prime")
    plan = _plan("g2_rebin_nv21_prime")
    ref = g["wn_ybar"]                                   # [n_fft, 4, 5]
    d = torch.from_numpy(g["wn_log_p_tvxz"]).cuda()
    big = torch.full((plan.n_fft, 9, 5), -1.0, dtype=torch.float64, device="cuda")
    _lib.check(_lib.load().pkm_ybar_density_into_cube(plan._h, _lib.ptr(d), 1, 4, 5, 1, 4, _lib.ptr(big), 9, 5, None))
    torch.cuda.synchronize()
    out = big.cpu().numpy()
    assert rel_err(out[:, 5:8], ref[:, 1:4]) < TOL
    assert np.all(out[:, :5] == -1.0) and np.all(out[:, 8:] == -1.0)
    # IPC handle round trip of a raw allocation (same process: export only)
    h = (C.c_uint8 * 64)()
    base = torch.empty(1 << 21, dtype=torch.uint8, device="cuda")
    assert _lib.load().pkm_ipc_export(_lib.ptr(base), h) == 0 and any(h)


@pytest.mark.parametrize("ssp_kw,nv,nx", [
    (dict(lmd_min=3540.5, lmd_max=7409.6), 101, (2, 3)),   # full-resolution MILES range: n_fft = 11179 = 7*1597, M = 32768 (global-scratch FFT)
    (dict(age_rebin=8, z_rebin=4), 7, (3, 3)),             # smallest velocity grid the cubic spline admits (nfi = 4)
    (dict(age_rebin=8, z_rebin=12), 127, (1, 5)),          # largest nv of the tensor-core coefficient kernels (nfi = 64), one metallicity bin
    (dict(age_rebin=8, z_rebin=4), 201, (2, 3)),           # IFUCube's default-sized velocity grid (nfi = 101): generic coefficient kernel
    (dict(age_rebin=53, z_rebin=3), 33, (1, 1)),           # one age bin, one pixel
])
def test_edge_configurations_against_oracle(gpu_ok, ssp_kw, nv, nx):
    """Shapes the golden files do not cover, checked against the oracle (which
    tests/test_oracle.py pins to the reference)."""
    from oracle import restatement as R
    import popkinmocks_b200 as pkm
    ssps = pkm.milesSSPs(**ssp_kw)
    cube = pkm.ifu_cube.IFUCube(ssps=ssps, nx1=nx[0], nx2=nx[1], nv=nv, vrng=(-1000.0, 1000.0))
    plan = pkm.engine.get_plan(cube)
    nz, nt = ssps.par_dims
    rng = np.random.default_rng(nv)
    p = rng.random(cube.get_distribution_shape("tvxz"))
    p /= np.sum(p * cube.construct_volume_element("tvxz"))
    y = plan.ybar_density(np.log(p), is_log=True)
    ref = R.ybar_density_literal(p, cube.v_edg, ssps.FXw, ssps.par_dims, ssps.n_fft, ssps.dv, ssps.delta_t,
                                 ssps.delta_z, cube.dx, cube.dy, pixel_block=1)
    assert y.shape == ref.shape == (ssps.n_fft, nx[0], nx[1])
    assert rel_err(y, ref) < TOL, (ssps.n_fft, rel_err(y, ref))
    P = rng.random((nt, nx[0], nx[1], nz))
    P /= P.sum()
    mu = 500.0 * (rng.random((nt, nx[0], nx[1])) - 0.5)
    sg = 15.0 + 250.0 * rng.random((nt, nx[0], nx[1]))
    ya = plan.ybar_gaussian(P, mu, sg)
    ra = R.ybar_gaussian(P, mu, sg, ssps.FXw, ssps.par_dims, ssps.n_fft, ssps.dv)
    assert rel_err(ya, ra) < TOL, (ssps.n_fft, rel_err(ya, ra))


def test_full_size_201_cube_properties(gpu_ok):
    """BASELINE.json's headline size: 201x201 pixels, full grid, nv=101 (20.8 GB density, generated on
    the device).  Size-independent checks: spot pixels against the oracle, a checksum of checksums
    (the pixel-sum of the cube equals the cube of the pixel-summed density, by linearity), and the
    cube assembled from 8 pixel shards is identical to the single-call cube."""
    import torch
    from oracle import restatement as R
    import popkinmocks_b200 as pkm
    free, _ = torch.cuda.mem_get_info()
    if free < 60 << 30:
        pytest.skip("needs ~45 GB of free device memory")
    ssps = pkm.milesSSPs()
    nx, nv = 201, 101
    cube = pkm.ifu_cube.IFUCube(ssps=ssps, nx1=nx, nx2=nx, nv=nv, vrng=(-1000.0, 1000.0))
    plan = pkm.engine.get_plan(cube)
    nz, nt = ssps.par_dims
    gen = torch.Generator(device="cuda").manual_seed(201)
    p = torch.empty((nt, nv, nx, nx, nz), dtype=torch.float64, device="cuda")
    for t in range(nt):
        p[t] = torch.rand((nv, nx, nx, nz), dtype=torch.float64, device="cuda", generator=gen)
    y = plan.ybar_density(p, is_log=False)
    assert tuple(y.shape) == (ssps.n_fft, nx, nx)
    for (i, j) in [(0, 0), (200, 200), (100, 57), (13, 188)]:
        sub = p[:, :, i:i + 1, j:j + 1, :].cpu().numpy()
        ref = R.ybar_density_literal(sub, cube.v_edg, ssps.FXw, ssps.par_dims, ssps.n_fft, ssps.dv,
                                     ssps.delta_t, ssps.delta_z, cube.dx, cube.dy)
        assert rel_err(y[:, i, j].cpu().numpy(), ref[:, 0, 0]) < TOL, (i, j)
    # checksum of checksums
    psum = p.sum(dim=(2, 3), keepdim=True)
    ysum = plan.ybar_density(psum, is_log=False)[:, 0, 0]
    tot = y.sum(dim=(1, 2))
    assert float((tot - ysum).abs().max() / ysum.abs().max()) < 1e-11
    # 8 contiguous pixel shards written into one cube (the multi-GPU decomposition on one device)
    from popkinmocks_b200 import _lib, sharding
    flat = p.reshape(nt, nv, nx * nx, 1, nz)
    big = torch.zeros((ssps.n_fft, nx * nx, 1), dtype=torch.float64, device="cuda")
    bounds = sharding.row_partition(nx * nx, 8)
    for r in range(8):
        _lib.check(_lib.load().pkm_ybar_density_into_cube(plan._h, _lib.ptr(flat), 0, nx * nx, 1, bounds[r], bounds[r + 1],
                                                          _lib.ptr(big), nx * nx, bounds[r], None))
    torch.cuda.synchronize()
    assert torch.equal(big.reshape(ssps.n_fft, nx, nx), y)


def test_synthetic_density_is_shard_independent(gpu_ok):
    """pkm_synth_log_density: the values of a pixel depend only on its GLOBAL index, and the NumPy
    replica used by bench.py's verification reproduces them (integer hash exact, log to 1 ulp)."""
    import torch
    from popkinmocks_b200 import engine
    kw = dict(nt=5, nv=7, total_pix=1000, nz=3, seed=20241019, scale=2.5e-4)
    whole = engine.synth_log_density(pix0=0, npix=1000, **kw)
    part = engine.synth_log_density(pix0=613, npix=200, **kw)
    assert torch.equal(whole[:, :, 613:813], part)
    px = [0, 1, 613, 812, 999]
    host = engine.synth_log_density_host(kw["nt"], kw["nv"], 1000, px, kw["nz"], kw["seed"], kw["scale"])
    dev = whole[:, :, px].cpu().numpy()
    assert np.max(np.abs(host - dev)) < 1e-14
    u = np.exp(dev) / kw["scale"]
    assert 0.0 < u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.05


def test_generic_coefficient_kernel_matches_the_tensor_core_kernels(gpu_ok, monkeypatch):
    """PKM_COEFF_GENERIC=1 forces the plain-FMA coefficient kernel (the path of nv > 127) on a
    shape the DMMA kernels also handle: same cube to rounding, in float64 and float32 mode."""
    from popkinmocks_b200 import engine
    name = "g2_rebin_nv21_prime"
    g = load_golden(name)
    cube = build_cube(name)
    for precision, tol in (("float64", 1e-13), ("float32", 1e-6)):
        ref = engine.Plan(cube, precision=precision).ybar_density(g["wn_log_p_tvxz"], is_log=True)
        monkeypatch.setenv("PKM_COEFF_GENERIC", "1")
        got = engine.Plan(cube, precision=precision).ybar_density(g["wn_log_p_tvxz"], is_log=True)
        monkeypatch.delenv("PKM_COEFF_GENERIC")
        assert rel_err(got, ref) < tol, precision
    assert rel_err(got, g["wn_ybar"]) < 1e-5


@pytest.mark.parametrize("name,tags", [("g1_conftest_full_nv41", ["d1_", "d2_", "s1_"]), ("g2_rebin_nv21_prime", ["d1_", "s1_"]),
                                       ("g3_full_nv101_odd", ["d1_"]), ("g4_even_nv40", ["d2_"]),
                                       ("g5_ppxf_pad_nv21", ["d1_"])])
def test_float32_gaussian_path_within_1e5(gpu_ok, name, tags):
    """Optional float32 arithmetic of the Gaussian-LOSVD path (k_gaussian_f32): BASELINE.json
    tolerance 1e-5 max-norm against the reference's float64 cube."""
    from popkinmocks_b200 import engine
    g = load_golden(name)
    plan32 = engine.get_plan(golden_cube(name), precision="float32")
    for tag in tags:
        y32 = plan32.ybar_gaussian(g[tag + "P_txz"], g[tag + "mu_v"], g[tag + "sig_v"])
        err = rel_err(y32, g[tag + "ybar"])
        assert err < 1e-5, (name, tag, err)
        assert err > 1e-12      # it really is the float32 path


@pytest.mark.parametrize("name", ["g2_rebin_nv21_prime", "g5_ppxf_pad_nv21", "g3_full_nv101_odd"])
def test_ssp_operand_pipeline_on_device(gpu_ok, name):
    """pkm_ssp_operand: log-resampling + real DFT of all SSPs on the GPU agree with
    the host path (SciPy interp1d + pocketfft, bit-identical to the reference) to rounding, for
    pad=False (prime and odd composite n_fft) and pad="ppxf"; a cube built on that operand
    reproduces the reference datacubes."""
    import popkinmocks_b200 as pkm
    from conftest import GOLDEN_CUBES
    cfg = GOLDEN_CUBES[name]
    g = load_golden(name)
    host = golden_cube(name).ssps
    ssps = pkm.milesSSPs(**cfg["ssp"])
    cube = pkm.ifu_cube.IFUCube(ssps=ssps, nx1=cfg["nx1"], nx2=cfg["nx2"], nv=cfg["nv"], vrng=(-1000.0, 1000.0),
                                operand_device=0)
    if cfg.get("pad"):
        ssps.prepare_operand_on_device(dv=ssps.dv, pad=cfg["pad"], device=0)
    assert ssps.n_fft == host.n_fft and ssps.FXw.shape == host.FXw.shape and ssps.Xw.shape == host.Xw.shape
    assert rel_err(ssps.Xw, host.Xw) < 1e-14
    assert rel_err(ssps.FXw, host.FXw) < 1e-13
    assert np.array_equal(ssps.light_weights, host.light_weights)
    plan = pkm.engine.Plan(cube)
    tag = "wn_"
    y = plan.ybar_density(g[tag + "log_p_tvxz"], is_log=True)
    assert rel_err(y, g[tag + "ybar"]) < TOL
    ya = plan.ybar_gaussian(g["d1_P_txz"], g["d1_mu_v"], g["d1_sig_v"])
    assert rel_err(ya, g["d1_ybar"]) < TOL
    plan.close()
