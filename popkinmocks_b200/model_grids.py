"""SSP model grid: the constant operand S-hat of the datacube hot path.

Host-side, off the hot path ("stays NumPy so S-hat is
bit-identical to the oracle").  Keeps the public surface of the reference class
`popkinmocks.model_grids.milesSSPs` (/root/reference/popkinmocks/model_grids.py:82-315):
constructor signature, and the attributes the components read
(`X, lmd, par_dims, par_edges, par_cents, delta_t, delta_z`, and after an
`IFUCube` is built on it `w, Xw, dv, dw, n_fft, n_pix, pad, FXw, light_weights`).

The arithmetic that defines S-hat is performed with the same library calls as the
reference so the result is identical to the last bit:
  * log-resampling: cubic `interp1d` of X.T over log(lambda), zero outside,
    sampled on `arange(log lmd_min, log lmd_max, dv/299792)`  (model_grids.py:240-261)
  * Fourier transform: `np.fft.rfft(Xw.T, n_fft).T`                (model_grids.py:263-281)
"""
import numpy as np
from scipy import interpolate

from . import read_miles

SPEED_OF_LIGHT = 299792.0  # km/s, the reference's constant (model_grids.py:249)


def _edges_from_centres(c):
    """Bin edges half way between centres, end bins mirrored (model_grids.py:181-191)."""
    c = np.asarray(c, dtype=np.float64)
    step = np.diff(c)
    lo = np.concatenate(([c[0] - step[0] / 2.0], c[1:] - step / 2.0))
    hi = np.concatenate((c[:-1] + step / 2.0, [c[-1] + step[-1] / 2.0]))
    assert np.allclose(lo[1:], hi[:-1])
    if np.abs(lo[0]) < 0.011:
        lo[0] = 0.0
    return np.concatenate((lo, [hi[-1]]))


def _coarsen_edges(edges, factor):
    coarse = edges[::factor]
    if coarse[-1] != edges[-1]:
        coarse = np.concatenate((coarse, [edges[-1]]))
    return coarse


class modelGrid:
    """Bookkeeping for a regular parameter grid (reference: model_grids.py:8-79)."""

    par_mask_function = None

    def _check_par_input(self):
        if len(self.par_dims) != self.npars:
            raise ValueError("par_dims is wrong length")
        if np.array(self.par_lims).shape != (self.npars, 2):
            raise ValueError("par_lims is wrong shape")
        if np.prod(self.par_dims) > 1e5:
            raise ValueError("total size of par grid > 1e5")

    def _get_par_grid(self, par_edges=None):
        if par_edges is None:
            par_edges = [np.linspace(lo, hi, n + 1)
                         for n, (lo, hi) in zip(self.par_dims, self.par_lims)]
        self.par_edges = list(par_edges)
        self.par_widths = [np.diff(e) for e in self.par_edges]
        self.par_cents = [(e[:-1] + e[1:]) / 2.0 for e in self.par_edges]
        grid = np.array(np.meshgrid(*self.par_cents, indexing="ij"))
        index = np.array(np.meshgrid(*[np.arange(len(e) - 1) for e in self.par_edges],
                                     indexing="ij"))
        self.par_mask_nd = np.zeros(grid.shape[1:], dtype=bool)
        self.pars = grid.reshape(self.npars, -1)
        self.par_idx = index.reshape(self.npars, -1)


class milesSSPs(modelGrid):
    """MILES SSP models (Vazdekis et al. 2010, BaSTI isochrones, Chabrier IMF).

    Args (same as the reference, model_grids.py:98-100):
        lmd_min, lmd_max : wavelength window (Angstrom)
        age_lim, z_lim   : (min, max) ages (Gyr) / metallicities [M/H] to keep
        age_rebin, z_rebin : integer down-sampling factors of the (t, z) grid
        mod_dir : packed library name or FITS directory (extension; default is the
            shipped MILES_BASTI_CH_baseFe library)
    """

    def __init__(self, lmd_min=4700, lmd_max=6500, age_lim=None, z_lim=None,
                 age_rebin=1, z_rebin=1, mod_dir="MILES_BASTI_CH_baseFe"):
        lib = read_miles.milesSSPs(mod_dir=mod_dir, age_lim=age_lim, z_lim=z_lim,
                                   thin_age=1, thin_z=1)
        lib.truncate_wavelengths(lmd_min=lmd_min, lmd_max=lmd_max)
        self.ssps = lib
        self.par_string = ("1) metallicity of stellar population",
                           "2) age of stellar population")
        self.par_pltsym = ["[Z/H]", "Age [Gyr]"]
        self.npars = 2
        self.n = lib.X.shape[0]
        self.lmd = lib.lmd
        self.lmd_min = lib.lmd[0]
        self.lmd_max = lib.lmd[-1]
        self.delta_lmd = (self.lmd_max - self.lmd_min) / (self.n - 1)
        # parameter axes are ordered (metallicity, age): columns of X are z-major
        self.par_dims = (lib.nz, lib.nt)
        self.par_lims = ((0, 1), (0, 1))
        self._check_par_input()
        self._get_par_grid(par_edges=[_edges_from_centres(lib.z_unq),
                                      _edges_from_centres(lib.t_unq)])
        self.p = self.pars.shape[-1]
        self.min_n_p = np.min([self.n, self.p])
        self.X = lib.X
        self.override_ticks = True
        self.set_tick_positions()
        self.delta_z = np.diff(self.par_edges[0])
        self.delta_t = np.diff(self.par_edges[1])
        self._rebin_par_grid(age_rebin=age_rebin, z_rebin=z_rebin)

    def _get_par_edge_lims(self, x_cnt):
        return _edges_from_centres(x_cnt)

    def _rebin_par_grid(self, age_rebin=1, z_rebin=1):
        """Merge (z, t) cells in blocks, averaging spectra with weights dz*dt.

        Reference behaviour: model_grids.py:142-179 (a trailing partial block is kept).
        """
        z_edges = _coarsen_edges(self.par_edges[0], z_rebin)
        t_edges = _coarsen_edges(self.par_edges[1], age_rebin)
        nz_new, nt_new = z_edges.size - 1, t_edges.size - 1
        cube = self.X.reshape((-1,) + tuple(self.par_dims))
        area = self.delta_z[:, None] * self.delta_t[None, :]
        weighted = cube * area
        merged = np.zeros((cube.shape[0], nz_new, nt_new))
        for a in range(nz_new):
            zs = slice(a * z_rebin, None if a == nz_new - 1 else (a + 1) * z_rebin)
            for b in range(nt_new):
                ts = slice(b * age_rebin, None if b == nt_new - 1 else (b + 1) * age_rebin)
                merged[:, a, b] = np.sum(weighted[:, zs, ts], (1, 2))
                merged[:, a, b] /= np.sum(area[zs, ts])
        self.par_edges = [z_edges, t_edges]
        self.par_cents = [(z_edges[1:] + z_edges[:-1]) / 2.0,
                          (t_edges[1:] + t_edges[:-1]) / 2.0]
        self.par_dims = (nz_new, nt_new)
        self.delta_z = z_edges[1:] - z_edges[:-1]
        self.delta_t = t_edges[1:] - t_edges[:-1]
        self.X = merged.reshape(merged.shape[0], -1)

    def set_tick_positions(self, t_ticks=None, z_ticks=None):
        """Tick bookkeeping kept for API compatibility (model_grids.py:193-238)."""
        if t_ticks is None:
            t_ticks = getattr(self, "t_ticks", [0.1, 1, 5, 13])
        if z_ticks is None:
            z_ticks = getattr(self, "z_ticks", [-2, -1, 0])
        t_ticks = np.array(t_ticks)
        z_ticks = np.array(z_ticks)
        z_e, t_e = self.par_edges
        self.t_ticks = t_ticks[(t_ticks > t_e[0]) & (t_ticks < t_e[-1])]
        self.z_ticks = z_ticks[(z_ticks > z_e[0]) & (z_ticks < z_e[-1])]
        self.img_t_ticks = np.interp(self.t_ticks, t_e, np.linspace(0, 1, self.par_dims[1] + 1))
        self.img_z_ticks = np.interp(self.z_ticks, z_e, np.linspace(0, 1, self.par_dims[0] + 1))

    def _logarithmically_resample(self, dv=5.0, interp_kind="cubic"):
        """Resample the SSPs uniformly in log-wavelength with velocity step `dv` km/s."""
        self.speed_of_light = SPEED_OF_LIGHT
        self.dv = dv
        self.dw = dv / SPEED_OF_LIGHT
        log_lmd = np.log(self.lmd)
        resampler = interpolate.interp1d(log_lmd, self.X.T, kind=interp_kind,
                                         bounds_error=False, fill_value=0.0)
        self.w = np.arange(np.min(log_lmd), np.max(log_lmd), self.dw)
        self.Xw = resampler(self.w).T

    def _calculate_fourier_transform(self, pad=False):
        """FXw[k, z*nt+t] = rfft of the log-resampled spectra, length n_fft."""
        self.n_pix = len(self.w)
        if pad is False or pad == False:  # noqa: E712  (reference accepts False only)
            n_fft = self.n_pix
        elif pad == "ppxf":
            n_fft = 2 ** int(np.ceil(np.log2(self.Xw.shape[0])))
        else:
            raise ValueError("Unknown choice for pad")
        self.pad = pad
        self.n_fft = n_fft
        self.FXw = np.fft.rfft(self.Xw.T, n_fft).T

    def prepare_operand_on_device(self, dv=5.0, pad=False, device=0):
        """`_logarithmically_resample(dv)` + `_calculate_fourier_transform(pad)` with the
        O(n_lambda x n_ssp) spline evaluation and the O(n^2 x n_ssp) real DFT on the GPU
        (pkm_ssp_operand; SURVEY.md section 8f rank 4, model_grids.py:240-281).

        The host keeps what is O(n_lambda): the banded not-a-knot solve for the B-spline
        coefficients and the four basis values of every output pixel (SciPy - the same
        `make_interp_spline` that `interp1d(kind="cubic")` uses).  `Xw` / `FXw` agree with the host
        path to ~1e-14 (direct DFT instead of pocketfft), not bit for bit - the default host path
        keeps the operand identical to the reference's."""
        from . import _lib
        self.speed_of_light = SPEED_OF_LIGHT
        self.dv = dv
        self.dw = dv / SPEED_OF_LIGHT
        log_lmd = np.log(self.lmd)
        self.w = np.arange(np.min(log_lmd), np.max(log_lmd), self.dw)
        spl = interpolate.make_interp_spline(log_lmd, self.X, k=3, check_finite=False)       # [n_lmd, n_ssp]
        design = interpolate.BSpline.design_matrix(self.w, spl.t, 3).tocsr()                 # 4 values per row
        n_pix = self.w.size
        counts = np.diff(design.indptr)
        if not np.all(counts == 4):
            raise ValueError("unexpected B-spline design matrix structure")
        tap_idx = np.ascontiguousarray(design.indices.reshape(n_pix, 4)[:, 0], dtype=np.int32)
        if not np.all(np.diff(design.indices.reshape(n_pix, 4), axis=1) == 1):
            raise ValueError("unexpected B-spline design matrix structure")
        tap_w = _lib.as_f64(design.data.reshape(n_pix, 4))
        self.n_pix = n_pix
        if pad is False or pad == False:  # noqa: E712
            n_fft = n_pix
        elif pad == "ppxf":
            n_fft = 2 ** int(np.ceil(np.log2(n_pix)))
        else:
            raise ValueError("Unknown choice for pad")
        self.pad, self.n_fft = pad, n_fft
        coef = _lib.as_f64(spl.c)
        nssp = coef.shape[1]
        Xw = np.empty((n_pix, nssp))
        FXw = np.empty((n_fft // 2 + 1, nssp), dtype=np.complex128)
        _lib.require_gpu()
        _lib.check(_lib.load().pkm_ssp_operand(_lib.ptr(coef), int(coef.shape[0]), _lib.ptr(tap_idx), _lib.ptr(tap_w),
                                               int(n_pix), int(nssp), int(n_fft), _lib.ptr(Xw), _lib.ptr(FXw),
                                               int(device), None))
        self.Xw, self.FXw = Xw, FXw

    def get_light_weights(self):
        """light_weights[t, z] = sum over wavelength of the (linear-lambda) SSP flux."""
        self.light_weights = np.sum(self.X, 0).reshape(self.par_dims).T

    def get_ssp(self, ssp_id, spacing="wavelength"):
        id_z, id_t = self.par_idx[:, ssp_id]
        if spacing == "wavelength":
            spectrum = self.X[:, ssp_id]
        elif spacing == "log-wavelength":
            spectrum = self.Xw[:, ssp_id]
        else:
            raise ValueError("Unknown choice for spacing")
        return self.par_cents[1][id_t], self.par_cents[0][id_z], spectrum
