"""IFU datacube discretisation (host side; carrier type of the hot path).

Keeps the public surface of `popkinmocks.ifu_cube.IFUCube`
(/root/reference/popkinmocks/ifu_cube.py:5-179, 234-260): constructor signature,
`x, y, xx, yy, dx, dy, nx, ny, nv, v_edg, x1rng, x2rng, ssps`, and the helpers
`construct_volume_element`, `get_variable_values/edges/size/extent`,
`get_distribution_shape`.  As in the reference, building a cube MUTATES the
`ssps` it is given (log-resampling at the cube's dv, FFT, light weights;
ifu_cube.py:48-50), so two cubes with different `nv` must not share one `ssps`.

Plotting wrappers (ifu_cube.py:181-397) are out of scope.

The cube also owns the cache of device plans (`_pkm_plans`): one `pkm_plan`
per (dtype, device), holding the device copy of S-hat', spline/FFT tables and
workspace, created lazily by `popkinmocks_b200.engine`.
"""
import numpy as np

_AXIS_LABELS = {"x1": "$x_1$", "x2": "$x_2$", "t": "$t$ [Gyr]", "z": "$z$ [M/H]",
                "v": "$v$ [km/s]"}


class IFUCube(object):
    """The integral field unit (IFU) datacube.

    Args:
        ssps : SSP templates, a `popkinmocks_b200.milesSSPs`
        nx1, nx2 (int): number of spatial pixels
        nv (int): number of velocity bins
        x1rng, x2rng (tuple): spatial extents (arbitrary units)
        vrng (tuple): velocity extent, km/s
        interp_kind (string): interpolation used to log-resample the SSPs
        operand_device (int, optional, extension): build the SSP operand (log-resampling + Fourier
            transform of all SSPs) on this GPU instead of with SciPy/NumPy on the host
    """

    def __init__(self, ssps=None, nx1=300, nx2=299, nv=200, x1rng=(-1, 1),
                 x2rng=(-1, 1), vrng=(-1000, 1000), interp_kind="cubic", operand_device=None):
        self.ssps = ssps
        self.nx, self.ny, self.nv = nx1, nx2, nv
        self.x1rng, self.x2rng = x1rng, x2rng
        e1 = np.linspace(*x1rng, nx1 + 1)
        e2 = np.linspace(*x2rng, nx2 + 1)
        self.dx = e1[1] - e1[0]
        self.dy = e2[1] - e2[0]
        self.x = e1[:-1] + self.dx / 2.0
        self.y = e2[:-1] + self.dy / 2.0
        self.xx, self.yy = np.meshgrid(self.x, self.y, indexing="ij")
        self.v_edg = np.linspace(*vrng, nv + 1)
        dv = self.v_edg[1] - self.v_edg[0]
        if operand_device is None:
            self.ssps._logarithmically_resample(dv=dv, interp_kind=interp_kind)
            self.ssps._calculate_fourier_transform()
        else:
            # extension: log-resampling and Fourier transform of the SSPs on the GPU (cubic only)
            if interp_kind != "cubic":
                raise ValueError("operand_device supports interp_kind='cubic' only")
            self.ssps.prepare_operand_on_device(dv=dv, device=int(operand_device))
        self.ssps.get_light_weights()
        self._pkm_plans = {}

    # device plans hold raw handles: never pickle them (dill_dump must keep working)
    def __getstate__(self):
        state = dict(self.__dict__)
        state["_pkm_plans"] = {}
        return state

    def _axis_widths(self, axis):
        if axis == "t":
            return self.ssps.delta_t
        if axis == "z":
            return self.ssps.delta_z
        if axis == "v":
            return self.v_edg[1:] - self.v_edg[:-1]
        if axis == "x":
            return np.array([self.dx])
        if axis == "y":
            return np.array([self.dy])
        raise ValueError(f"Unknown variable: {axis}")

    def construct_volume_element(self, which_dist):
        """Volume element of the dependent variables of `which_dist`.

        e.g. 'vz' -> dv*dz with shape (nv, nz); 't_x' -> dt with shape (nt, 1, 1).
        The spatial factors enter as length-1 axes (reference: ifu_cube.py:52-95).
        """
        dependent, _, given = which_dist.partition("_")
        axes = dependent.replace("x", "xy")
        dvol = np.ones((1,) * len(axes))
        for pos, axis in enumerate(axes):
            shape = [1] * len(axes)
            shape[pos] = -1
            dvol = dvol * self._axis_widths(axis).reshape(shape)
        n_given = len(given.replace("x", "xy"))
        return dvol.reshape(dvol.shape + (1,) * n_given)

    def get_variable_values(self, which_variable):
        """Bin centres of one of (t, v, x1, x2, z)."""
        if which_variable == "t":
            return self.ssps.par_cents[1]
        if which_variable == "z":
            return self.ssps.par_cents[0]
        if which_variable == "v":
            return (self.v_edg[:-1] + self.v_edg[1:]) / 2.0
        if which_variable == "x1":
            return self.x
        if which_variable == "x2":
            return self.y
        raise ValueError(f"Unknown variable: {which_variable}")

    def get_variable_edges(self, which_variable):
        """Bin edges of one of (t, v, x1, x2, z)."""
        if which_variable == "t":
            return self.ssps.par_edges[1]
        if which_variable == "z":
            return self.ssps.par_edges[0]
        if which_variable == "v":
            return self.v_edg
        if which_variable == "x1":
            return np.concatenate([self.x - self.dx / 2, [self.x[-1] + self.dx / 2]])
        if which_variable == "x2":
            return np.concatenate([self.y - self.dy / 2, [self.y[-1] + self.dy / 2]])
        raise ValueError(f"Unknown variable: {which_variable}")

    def get_variable_size(self, which_variable):
        return len(self.get_variable_values(which_variable))

    def get_variable_extent(self, which_variable):
        """Outermost bin edges (start, end) of a variable."""
        if which_variable == "x1":
            return self.x1rng
        if which_variable == "x2":
            return self.x2rng
        if which_variable in ("t", "z", "v"):
            e = self.get_variable_edges(which_variable)
            return (e[0], e[-1])
        raise ValueError("Unknown `which_variable`")

    def get_distribution_shape(self, which_dist):
        """Array shape of a distribution string such as 'tvxz' or 'v_tx'."""
        shape = []
        for var in which_dist.replace("_", ""):
            if var == "x":
                shape += [self.get_variable_size("x1"), self.get_variable_size("x2")]
            else:
                shape += [self.get_variable_size(var)]
        return tuple(shape)

    def get_axis_label(self, which_dist):
        if which_dist not in _AXIS_LABELS:
            raise ValueError("Unknown `which_dist`")
        return _AXIS_LABELS[which_dist]

    def imshow(self, *args, **kwargs):
        raise NotImplementedError("plotting is outside the scope of popkinmocks_b200")

    plot = imshow
    plot_spectrum = imshow
