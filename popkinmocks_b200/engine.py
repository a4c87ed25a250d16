"""Host-side driver of the CUDA hot path: plans and marshalling.

A `Plan` wraps one `pkm_plan*` (include/pkm_b200.h): the device-resident constant
operands of one (IFUCube, milesSSPs) pair - S-hat' = FXw*dz*dt, the folded
DFT x spline matrices, the 4-tap spline tables, chirp-z FFT tables - plus its
workspace.  Plans are cached on the cube (`cube._pkm_plans`), one per device.

Everything numerical happens inside libpkm_b200.so; this module only validates,
shapes and hands over pointers.  NumPy arrays are host buffers (the library
stages them), torch CUDA tensors are passed through as device pointers.
"""
import ctypes as C

import numpy as np

from . import _lib


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _is_cuda_tensor(x):
    return _is_torch(x) and x.is_cuda


def find_v0_idx(cube):
    """Index of the velocity bin centred on zero, or -1 (base.py:820-823 raises IndexError)."""
    v = (cube.v_edg[:-1] + cube.v_edg[1:]) / 2.0
    hit = np.where(np.isclose(v, 0.0))[0]
    return int(hit[0]) if hit.size else -1


class Plan:
    """Device-resident constants of one cube on one GPU."""

    def __init__(self, cube, device=0, precision="float64"):
        _lib.require_gpu()
        if precision not in ("float64", "float32"):
            raise ValueError("precision must be 'float64' or 'float32'")
        self.precision = precision
        ssps = cube.ssps
        nz, nt = ssps.par_dims
        self.nt, self.nz, self.nv = int(nt), int(nz), int(cube.nv)
        self.n_fft = int(ssps.n_fft)
        self.nfo = self.n_fft // 2 + 1
        self.device = int(device)
        desc = _lib.CubeDesc(nt=self.nt, nz=self.nz, nv=self.nv, n_fft=self.n_fft,
                             v0_idx=find_v0_idx(cube), device=self.device,
                             precision=1 if precision == "float32" else 0,
                             reserved=0, dv=float(ssps.dv), dx=float(cube.dx), dy=float(cube.dy))
        FXw = np.ascontiguousarray(ssps.FXw, dtype=np.complex128)
        if FXw.shape != (self.nfo, self.nz * self.nt):
            raise ValueError(f"ssps.FXw has shape {FXw.shape}, expected {(self.nfo, self.nz * self.nt)}")
        dt = _lib.as_f64(ssps.delta_t)
        dz = _lib.as_f64(ssps.delta_z)
        handle = C.c_void_p()
        _lib.check(_lib.load().pkm_plan_create(C.byref(desc), _lib.ptr(FXw), _lib.ptr(dt),
                                               _lib.ptr(dz), C.byref(handle)))
        self._h = handle
        # omega_k of the Gaussian path: linspace(0, pi, nfo)/dv (parametric.py:360-362)
        self.omega = np.linspace(0, np.pi, self.nfo)
        self.omega /= ssps.dv

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            _lib.load().pkm_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass

    @property
    def launches(self):
        return int(_lib.load().pkm_plan_launch_count(self._h))

    KERNEL_CLASSES = ("coeff", "contract", "gaussian", "irfft", "transpose")

    def set_profiling(self, enable):
        """Bracket every hot-path launch with CUDA events (resets the accumulated times)."""
        _lib.check(_lib.load().pkm_plan_set_profiling(self._h, int(bool(enable))))

    def kernel_times(self):
        """{class: (total ms, launches)} since profiling was enabled (waits for the launches)."""
        ms = np.zeros(len(self.KERNEL_CLASSES))
        cnt = np.zeros(len(self.KERNEL_CLASSES), dtype=np.int64)
        _lib.check(_lib.load().pkm_plan_kernel_times(self._h, _lib.ptr(ms), _lib.ptr(cnt)))
        return {k: (float(m), int(c)) for k, m, c in zip(self.KERNEL_CLASSES, ms, cnt)}

    def set_workspace_limit(self, nbytes):
        _lib.check(_lib.load().pkm_plan_set_workspace_limit(self._h, int(nbytes)))

    # ------------------------------------------------------------------
    def _alloc_out(self, like_device, rows, nx2, layout):
        shape = (self.n_fft, rows, nx2) if layout == _lib.LAYOUT_CUBE else (rows * nx2, self.n_fft)
        if like_device:
            import torch
            return torch.empty(shape, dtype=torch.float64, device=f"cuda:{self.device}")
        return np.empty(shape, dtype=np.float64)

    def ybar_density(self, dens, is_log=True, rows=None, out=None, layout=_lib.LAYOUT_CUBE,
                     stream=None):
        """Component.evaluate_ybar (base.py:789-882) for x1 rows `rows=(begin, end)`.

        dens: [nt, nv, nx1, nx2, nz] float64 - log p if `is_log` else p; NumPy (host) or a
        torch CUDA tensor.  Returns ybar for the rows, [n_fft, nrows, nx2] (cube layout) or
        [nrows*nx2, n_fft] (pixel-major), on the same side as the input unless `out` is given."""
        if tuple(dens.shape[:2]) != (self.nt, self.nv) or dens.shape[-1] != self.nz or len(dens.shape) != 5:
            raise ValueError(f"density has shape {tuple(dens.shape)}, expected (nt={self.nt}, "
                             f"nv={self.nv}, nx1, nx2, nz={self.nz})")
        nx1, nx2 = int(dens.shape[2]), int(dens.shape[3])
        r0, r1 = (0, nx1) if rows is None else (int(rows[0]), int(rows[1]))
        on_dev = _is_cuda_tensor(dens)
        if on_dev:
            import torch
            if dens.dtype != torch.float64 or not dens.is_contiguous():
                dens = dens.to(torch.float64).contiguous()
        else:
            dens = _lib.as_f64(dens)
        if out is None:
            out = self._alloc_out(on_dev, r1 - r0, nx2, layout)
        _lib.check(_lib.load().pkm_ybar_density(self._h, _lib.ptr(dens), int(bool(is_log)), nx1, nx2,
                                                r0, r1, _lib.ptr(out), layout,
                                                _lib.stream_ptr(stream)))
        return out

    def ybar_gaussian(self, P_txz, mu_v, sig_v, rows=None, out=None, layout=_lib.LAYOUT_CUBE,
                      stream=None):
        """ParametricComponent.evaluate_ybar (parametric.py:335-407) for x1 rows `rows`.

        P_txz [nt, nx1, nx2, nz]; mu_v, sig_v broadcastable to [nt, nx1, nx2] (NumPy broadcast
        views are passed by stride, not copied)."""
        if len(P_txz.shape) != 4 or P_txz.shape[0] != self.nt or P_txz.shape[-1] != self.nz:
            raise ValueError(f"P_txz has shape {tuple(P_txz.shape)}, expected (nt={self.nt}, nx1, "
                             f"nx2, nz={self.nz})")
        nx1, nx2 = int(P_txz.shape[1]), int(P_txz.shape[2])
        r0, r1 = (0, nx1) if rows is None else (int(rows[0]), int(rows[1]))
        on_dev = _is_cuda_tensor(P_txz)
        if on_dev:
            import torch
            P_txz = P_txz.to(torch.float64).contiguous()
            mu_v = torch.broadcast_to(mu_v.to(torch.float64), (self.nt, nx1, nx2))
            sig_v = torch.broadcast_to(sig_v.to(torch.float64), (self.nt, nx1, nx2))
            mstr = [int(s) for s in mu_v.stride()]
            sstr = [int(s) for s in sig_v.stride()]
        else:
            P_txz = _lib.as_f64(P_txz)
            mu_v = np.broadcast_to(np.asarray(mu_v, dtype=np.float64), (self.nt, nx1, nx2))
            sig_v = np.broadcast_to(np.asarray(sig_v, dtype=np.float64), (self.nt, nx1, nx2))
            mstr = [s // 8 for s in mu_v.strides]
            sstr = [s // 8 for s in sig_v.strides]
        if out is None:
            out = self._alloc_out(on_dev, r1 - r0, nx2, layout)
        ms = (C.c_int64 * 3)(*mstr)
        ss = (C.c_int64 * 3)(*sstr)
        _lib.check(_lib.load().pkm_ybar_gaussian(self._h, _lib.ptr(P_txz), _lib.ptr(mu_v), ms,
                                                 _lib.ptr(sig_v), ss, _lib.ptr(self.omega), nx1, nx2,
                                                 r0, r1, _lib.ptr(out), layout,
                                                 _lib.stream_ptr(stream)))
        return out

    def irfft(self, Yhat, scale=1.0, naive=False):
        """Batched inverse real FFT of length n_fft: Yhat [npix, nfo] complex128 -> [npix, n_fft]."""
        Yhat = np.ascontiguousarray(Yhat, dtype=np.complex128)
        npix = Yhat.shape[0]
        out = np.empty((npix, self.n_fft))
        fn = _lib.load().pkm_irfft_naive if naive else _lib.load().pkm_irfft_batch
        _lib.check(fn(self._h, _lib.ptr(Yhat), npix, float(scale), _lib.ptr(out), None))
        return out

    def reduce_logp(self, log_p_tvxz, log_dv, keep, density, light_weights=None):
        """log-marginal of the pixelated density over the axes NOT in `keep` (subset of 'tvxz').

        Returns the array with kept axes in (t, v, x1, x2, z) order (base.py:884-1174)."""
        on_dev = _is_cuda_tensor(log_p_tvxz)
        if not on_dev:
            log_p_tvxz = _lib.as_f64(log_p_tvxz)
        nt, nv, nx1, nx2, nz = [int(s) for s in log_p_tvxz.shape]
        mask = sum(bit for ax, bit in zip("tvxz", (1, 2, 4, 8)) if ax in keep)
        shape = []
        if "t" in keep:
            shape.append(nt)
        if "v" in keep:
            shape.append(nv)
        if "x" in keep:
            shape += [nx1, nx2]
        if "z" in keep:
            shape.append(nz)
        if on_dev:
            import torch
            out = torch.empty(tuple(shape), dtype=torch.float64, device=log_p_tvxz.device)
        else:
            out = np.empty(tuple(shape), dtype=np.float64)
        lw = None if light_weights is None else _lib.as_f64(light_weights)
        ldv = _lib.as_f64(log_dv)
        _lib.check(_lib.load().pkm_reduce_logp(self._h, _lib.ptr(log_p_tvxz), nx1 * nx2, _lib.ptr(ldv),
                                               _lib.ptr(lw), mask, int(bool(density)), _lib.ptr(out),
                                               None))
        return out

    @staticmethod
    def _mask(keep):
        return sum(bit for ax, bit in zip("tvxz", (1, 2, 4, 8)) if ax in keep)

    def _marginal_shape(self, keep, nx1, nx2):
        shape = []
        for ax in "tvxz":
            if ax in keep:
                shape += {"t": [self.nt], "v": [self.nv], "x": [nx1, nx2], "z": [self.nz]}[ax]
        return tuple(shape)

    def reduce_logp_multi(self, log_p_tvxz, log_dv, keeps, densities, light_weights=None):
        """Several log-marginals in one call (pkm_reduce_logp_multi): nested requests are reduced
        from each other, so e.g. keeps=("vx", "x") reads the 5-D array once.  Returns a list."""
        on_dev = _is_cuda_tensor(log_p_tvxz)
        if not on_dev:
            log_p_tvxz = _lib.as_f64(log_p_tvxz)
        nt, nv, nx1, nx2, nz = [int(s) for s in log_p_tvxz.shape]
        n = len(keeps)
        outs = []
        for keep in keeps:
            shape = self._marginal_shape(keep, nx1, nx2)
            if on_dev:
                import torch
                outs.append(torch.empty(shape, dtype=torch.float64, device=log_p_tvxz.device))
            else:
                outs.append(np.empty(shape, dtype=np.float64))
        masks = (C.c_uint32 * n)(*[self._mask(k) for k in keeps])
        dens = (C.c_int32 * n)(*[int(bool(d)) for d in densities])
        ptrs = (C.c_void_p * n)(*[_lib.ptr(o) for o in outs])
        lw = None if light_weights is None else _lib.as_f64(light_weights)
        ldv = _lib.as_f64(log_dv)
        _lib.check(_lib.load().pkm_reduce_logp_multi(self._h, _lib.ptr(log_p_tvxz), nx1 * nx2, _lib.ptr(ldv),
                                                     _lib.ptr(lw), n, masks, dens, ptrs, None))
        return outs

    def conditional(self, log_joint, joint, log_given, given, npix, log_dv, density, exponentiate, xshape):
        """A (conditional) distribution from device-resident log-probability marginals
        (pkm_conditional): axes = dependents (joint minus given) then conditioners, each group in
        t, v, x, z order.  Returns a NumPy array (the result of get_p / get_log_p)."""
        import torch
        dims = {"t": [self.nt], "v": [self.nv], "x": list(xshape), "z": [self.nz]}
        dep = [a for a in "tvxz" if a in joint and a not in given]
        shape = sum((dims[a] for a in dep), []) + sum((dims[a] for a in "tvxz" if a in given), [])
        out = torch.empty(tuple(shape), dtype=torch.float64, device=log_joint.device)
        ldv = _lib.as_f64(log_dv)
        lg = None if log_given is None else log_given.contiguous()
        _lib.check(_lib.load().pkm_conditional(self._h, _lib.ptr(log_joint.contiguous()), self._mask(joint),
                                               _lib.ptr(lg), self._mask(given) if lg is not None else 0, int(npix),
                                               _lib.ptr(ldv), int(bool(density)), int(bool(exponentiate)),
                                               _lib.ptr(out), None))
        return out.cpu().numpy()

    def reduce_logp_from(self, src, src_keep, npix, log_dv, keep, density, light_weights=None):
        """A log-marginal reduced from the device-resident log-PROBABILITY marginal `src` over the
        axes `src_keep` (pkm_reduce_logp_from); `light_weights` re-weights a mass-weighted source
        that keeps t and z.  Returns a CUDA tensor with the axes of `keep`."""
        import torch
        dims = {"t": [self.nt], "v": [self.nv], "z": [self.nz]}
        xshape = None
        if "x" in src_keep:
            xi = src_keep.index("x")
            xshape = list(src.shape[xi:xi + 2])
        shape = []
        for ax in "tvxz":
            if ax in keep:
                shape += xshape if ax == "x" else dims[ax]
        out = torch.empty(tuple(shape), dtype=torch.float64, device=src.device)
        lw = None if light_weights is None else _lib.as_f64(light_weights)
        ldv = _lib.as_f64(log_dv)
        _lib.check(_lib.load().pkm_reduce_logp_from(self._h, _lib.ptr(src.contiguous()), self._mask(src_keep),
                                                    int(npix), _lib.ptr(ldv), _lib.ptr(lw), self._mask(keep),
                                                    int(bool(density)), _lib.ptr(out), None))
        return out


# ----------------------------------------------------------------------------------------------
# host buffers: page-locked in place on first use, released when the array is garbage collected
# ----------------------------------------------------------------------------------------------
_PINNED = {}


def pin_host_array(arr):
    """Page-lock a C-contiguous NumPy array in place (pkm_host_register) so that the library's
    pipelined staging and `to_device` read it at PCIe rate.  Idempotent; the registration is
    undone when the array is collected.  Returns True if the array is (now) pinned."""
    if not isinstance(arr, np.ndarray) or not arr.flags.c_contiguous or arr.nbytes < (1 << 20):
        return False
    addr = arr.ctypes.data
    if addr in _PINNED:
        return True
    base = arr
    while isinstance(base.base, np.ndarray):
        base = base.base
    try:
        _lib.check(_lib.load().pkm_host_register(C.c_void_p(addr), int(arr.nbytes)))
    except _lib.PkmError:
        return False                         # e.g. read-only mapping: the pageable path still works
    _PINNED[addr] = True

    def _release(a=addr):
        if _PINNED.pop(a, None):
            try:
                _lib.load().pkm_host_unregister(C.c_void_p(a))
            except Exception:  # noqa: BLE001 - interpreter shutdown
                pass
    import weakref
    try:
        weakref.finalize(base, _release)
    except TypeError:
        pass
    return True


def to_device(arr, device=0):
    """float64 CUDA tensor copy of a host array: the source is page-locked in place and copied by
    one asynchronous DMA (no pageable staging of the whole array)."""
    import torch
    host = np.ascontiguousarray(arr, dtype=np.float64)
    pin_host_array(host)
    out = torch.empty(host.shape, dtype=torch.float64, device=f"cuda:{device}")
    out.copy_(torch.from_numpy(host), non_blocking=True)
    torch.cuda.current_stream(out.device).synchronize()
    return out


def pinned_empty(shape):
    """(tensor, ndarray view) of a page-locked float64 host buffer from torch's caching allocator."""
    import torch
    t = torch.empty(tuple(shape), dtype=torch.float64, pin_memory=True)
    return t, t.numpy()


def get_plan(cube, device=0, precision="float64"):
    """The cube's plan for (`device`, `precision`), created on first use.

    precision="float32" selects float32 arithmetic in the density-path contraction (inputs and
    outputs stay float64; BASELINE.json tolerance 1e-5); everything else is always float64."""
    cache = cube.__dict__.setdefault("_pkm_plans", {})
    key = int(device) if precision == "float64" else (int(device), precision)
    if key not in cache:
        cache[key] = Plan(cube, device=int(device), precision=precision)
    return cache[key]


def axpy_cubes(ybars, weights, device=0):
    """sum_i w_i * ybar_i (Mixture.evaluate_ybar, mixture.py:35-48)."""
    _lib.require_gpu()
    on_dev = _is_cuda_tensor(ybars[0])
    if on_dev:
        import torch
        ybars = [y.contiguous() for y in ybars]
        out = torch.empty_like(ybars[0])
    else:
        ybars = [_lib.as_f64(y) for y in ybars]
        out = pinned_empty(ybars[0].shape)[1]          # page-locked: the D2H of the sum runs at PCIe rate
    n = int(np.prod(ybars[0].shape))
    ptrs = (C.c_void_p * len(ybars))(*[_lib.ptr(y) for y in ybars])
    w = _lib.as_f64(weights)
    _lib.check(_lib.load().pkm_axpy_cubes(len(ybars), ptrs, _lib.ptr(w), n, _lib.ptr(out), int(device), None))
    return out


def logsumexp_cubes(log_ps, weights, device=0):
    """log sum_i w_i exp(log_p_i): the collapsed log density of a Mixture (mixture.py:50-80).
    `log_ps`: equally shaped NumPy arrays or CUDA tensors; the result lives on the same side."""
    _lib.require_gpu()
    if _is_cuda_tensor(log_ps[0]):
        import torch
        log_ps = [a.contiguous() for a in log_ps]
        out = torch.empty_like(log_ps[0])
    else:
        log_ps = [_lib.as_f64(a) for a in log_ps]
        out = np.empty_like(log_ps[0])
    n = int(np.prod(log_ps[0].shape))
    ptrs = (C.c_void_p * len(log_ps))(*[_lib.ptr(a) for a in log_ps])
    log_w = _lib.as_f64(np.log(np.asarray(weights, dtype=np.float64)))
    _lib.check(_lib.load().pkm_logsumexp_cubes(len(log_ps), ptrs, _lib.ptr(log_w), n, _lib.ptr(out),
                                               int(device), None))
    return out


def moments(P, values, device=0):
    """Mean and central moments 2..4 along axis 0 of P[nvar, ...] (base.py:208-260).

    Returns an array [4, ...]: mean, mu2, mu3, mu4."""
    _lib.require_gpu()
    P = _lib.as_f64(P)
    values = _lib.as_f64(values)
    nvar = P.shape[0]
    ncond = int(np.prod(P.shape[1:])) if P.ndim > 1 else 1
    out = np.empty((4,) + P.shape[1:])
    _lib.check(_lib.load().pkm_moments(_lib.ptr(P), _lib.ptr(values), nvar, ncond, _lib.ptr(out),
                                       int(device), None))
    return out


def moments_logp(log_joint, log_given, values, nouter, nvar, ninner):
    """Mean and central moments 2..4 of P = exp(log_joint - log_given) along the middle axis of
    log_joint [nouter, nvar, ninner] (CUDA tensors; pkm_moments_logp).  Returns a NumPy array
    [4, nouter * ninner]."""
    import torch
    _lib.require_gpu()
    dev = log_joint.device
    vals = torch.as_tensor(np.ascontiguousarray(values, dtype=np.float64), device=dev)
    out = torch.empty((4, int(nouter) * int(ninner)), dtype=torch.float64, device=dev)
    lg = None if log_given is None else log_given.contiguous()
    _lib.check(_lib.load().pkm_moments_logp(_lib.ptr(log_joint.contiguous()), _lib.ptr(lg), _lib.ptr(vals),
                                            int(nouter), int(nvar), int(ninner), _lib.ptr(out),
                                            int(dev.index or 0), None))
    return out.cpu().numpy()


def noise(ybar, snr, mode, seed=0, offset=0, ybar_max=-1.0, want_sigma=True, want_sample=True,
          device=0):
    """ShotNoise (mode 0) / ConstantSNR (mode 1) sigma map and one noisy realisation (noise.py:16-71)."""
    _lib.require_gpu()
    on_dev = _is_cuda_tensor(ybar)
    if on_dev:
        import torch
        ybar = ybar.contiguous()
        yobs = torch.empty_like(ybar) if want_sample else None
        sigma = torch.empty_like(ybar) if want_sigma else None
    else:
        ybar = _lib.as_f64(ybar)
        yobs = np.empty_like(ybar) if want_sample else None
        sigma = np.empty_like(ybar) if want_sigma else None
    n = int(np.prod(ybar.shape))
    _lib.check(_lib.load().pkm_noise(_lib.ptr(ybar), n, float(snr), int(mode), int(seed), int(offset),
                                     float(ybar_max), _lib.ptr(yobs), _lib.ptr(sigma), int(device), None))
    return yobs, sigma


def fp64_fma_peak(device=0, repeats=5):
    """Measured FP64 FMA throughput (TFLOP/s) - roofline denominator of the FP64-bound kernels."""
    out = C.c_double(0.0)
    _lib.check(_lib.load().pkm_fp64_fma_peak(int(device), int(repeats), C.byref(out)))
    return out.value


def histogram5d(samples, edges, weights=None, device=0):
    """Counts (or weighted sums) of particles on a 5-D grid with numpy.histogramdd bin semantics.

    samples: 5 arrays (t, v, x1, x2, z); edges: 5 ascending edge arrays.  Returns a float64
    array [nt, nv, nx1, nx2, nz] (host).  Bin indices are bit-exact w.r.t. NumPy."""
    _lib.require_gpu()
    samples = [_lib.as_f64(s).ravel() for s in samples]
    n = samples[0].size
    if any(s.size != n for s in samples):
        raise ValueError("sample arrays must have equal length")
    edges = [_lib.as_f64(e) for e in edges]
    nbin = (C.c_int32 * 5)(*[e.size - 1 for e in edges])
    out = np.empty(tuple(e.size - 1 for e in edges), dtype=np.float64)
    sp = (C.c_void_p * 5)(*[_lib.ptr(s) for s in samples])
    ep = (C.c_void_p * 5)(*[_lib.ptr(e) for e in edges])
    w = None if weights is None else _lib.as_f64(weights).ravel()
    if w is not None and w.size != n:
        raise ValueError("weights must have one entry per particle")
    _lib.check(_lib.load().pkm_histogram5d(sp, _lib.ptr(w), n, ep, nbin, _lib.ptr(out), int(device), None))
    return out


def pixelate_gaussian(log_w_txz, mu_v, sig_v, v_edg, density=True, device=0, on_device=True, out=None):
    """log p(t,v,x,z) of a Gaussian-LOSVD component, pixelated on the device.

    log_w_txz [nt, nx1, nx2, nz] = log p(t,x,z); mu_v, sig_v broadcastable to [nt, nx1, nx2];
    v_edg [nv+1].  Adds the log of the normal-cdf difference over every velocity bin (minus
    log dv if `density`), parametric.py:180-211,567-596.  Returns a CUDA tensor
    [nt, nv, nx1, nx2, nz] (`on_device`, the 5-D array then never exists on the host) or a
    NumPy array; `out` (contiguous float64, that shape, either side) is filled if given."""
    _lib.require_gpu()
    import torch
    nt, nx1, nx2, nz = log_w_txz.shape
    v_edg = _lib.as_f64(v_edg)
    nv = v_edg.size - 1

    def dense(a, shape):
        if _is_cuda_tensor(a):
            return a.to(torch.float64).expand(shape).contiguous()
        return _lib.as_f64(np.broadcast_to(np.asarray(a, dtype=np.float64), shape))

    log_w = dense(log_w_txz, (nt, nx1, nx2, nz))
    mu = dense(mu_v, (nt, nx1, nx2))
    sig = dense(sig_v, (nt, nx1, nx2))
    if out is not None:
        if tuple(out.shape) != (nt, nv, nx1, nx2, nz):
            raise ValueError(f"`out` has shape {tuple(out.shape)}, expected {(nt, nv, nx1, nx2, nz)}")
    elif on_device:
        out = torch.empty((nt, nv, nx1, nx2, nz), dtype=torch.float64, device=f"cuda:{device}")
    else:
        out = np.empty((nt, nv, nx1, nx2, nz), dtype=np.float64)
    _lib.check(_lib.load().pkm_pixelate_gaussian(
        _lib.ptr(log_w), _lib.ptr(mu), _lib.ptr(sig), _lib.ptr(v_edg), nt, nv, nx1 * nx2, nz,
        int(bool(density)), _lib.ptr(out), int(device), None))
    return out


_SM64 = (1 << 64) - 1


def synth_log_density(nt, nv, total_pix, pix0, npix, nz, seed, scale, device=0, out=None):
    """Pixels [pix0, pix0+npix) of the synthetic white-noise log density of the benchmark cube,
    generated on the device (pkm_synth_log_density).  Returns a CUDA tensor [nt, nv, npix, nz]."""
    _lib.require_gpu()
    import torch
    if out is None:
        out = torch.empty((nt, nv, npix, nz), dtype=torch.float64, device=f"cuda:{device}")
    _lib.check(_lib.load().pkm_synth_log_density(_lib.ptr(out), nt, nv, int(total_pix), int(pix0), int(npix),
                                                 nz, int(seed) & _SM64, float(scale), int(device), None))
    return out


def synth_log_density_host(nt, nv, total_pix, pixels, nz, seed, scale):
    """NumPy replica of pkm_synth_log_density for a list of global pixel indices:
    [nt, nv, len(pixels), nz].  Same integer hash; `log` may differ from CUDA's in the last bit."""
    pixels = np.asarray(pixels, dtype=np.uint64)
    t = np.arange(nt, dtype=np.uint64)[:, None, None, None]
    v = np.arange(nv, dtype=np.uint64)[None, :, None, None]
    z = np.arange(nz, dtype=np.uint64)[None, None, None, :]
    g = ((t * np.uint64(nv) + v) * np.uint64(total_pix) + pixels[None, None, :, None]) * np.uint64(nz) + z
    with np.errstate(over="ignore"):
        x = g + np.uint64(int(seed) & _SM64) + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    u = ((x >> np.uint64(11)).astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)
    return np.log(u * scale)
