"""`FromParticle`: a component built by binning simulation particles onto the cube's grid.

Behind the reference's API (/root/reference/popkinmocks/components/particle.py:5-56).  The
5-D histogram is counted on the device (pkm_histogram5d) with `numpy.histogramdd` bin
semantics - searchsorted(side="right"), right edge of the last bin inclusive, outliers dropped -
so bin indices are bit-exact; the `density=True` normalisation then applies NumPy's own
sequence of operations to the counts on the host, which makes the unweighted result
bit-identical to `np.histogramdd(..., density=True)`.
"""
import numpy as np

from . import base
from .. import engine


def density_from_counts(counts, edges):
    """The `density=True` normalisation of numpy.histogramdd, operation for operation
    (numpy/lib/_histograms_impl.py: total first, then one division per axis, then the total)."""
    hist = counts
    total = hist.sum()
    ndim = hist.ndim
    for axis, e in enumerate(edges):
        shape = np.ones(ndim, int)
        shape[axis] = hist.shape[axis]
        hist = hist / np.diff(e).reshape(shape)
    hist /= total
    return hist


class FromParticle(base.Component):
    """Args:
        cube: the `IFUCube` whose bin edges are used
        t, v, x1, x2, z (arrays): particle ages [Gyr], velocities [km/s], positions, metallicities
        mass_weights (array): optional particle masses
    """

    def __init__(self, cube, t, v, x1, x2, z, mass_weights=None):
        names = ["t", "v", "x1", "x2", "z"]
        samples = [np.asarray(a) for a in (t, v, x1, x2, z)]
        edges = [cube.get_variable_edges(n) for n in names]
        inside = np.ones(samples[0].shape, dtype=bool)
        for data, e in zip(samples, edges):
            inside &= (data >= e[0]) & (data <= e[-1])
        lines = [f"Note: {inside.size - int(inside.sum())}/{inside.size} particles out of bounds.",
                 "N lost\t: low/high"]
        for name, data, e in zip(names, samples, edges):
            lines.append(f"   {name}\t: {int(np.sum(data < e[0]))}/{int(np.sum(data > e[-1]))}")
        print("\n".join(lines))
        counts = engine.histogram5d(samples, edges, weights=mass_weights)
        with np.errstate(invalid="ignore", divide="ignore"):
            p_tvxz = density_from_counts(counts, edges)
        with np.errstate(divide="ignore"):
            log_p_tvxz = np.log(p_tvxz)
        super().__init__(cube=cube, log_p_tvxz=log_p_tvxz)
