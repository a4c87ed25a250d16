"""`FromParticle`: a component built by binning simulation particles onto the cube's grid.

Behind the reference's API (/root/reference/popkinmocks/components/particle.py:5-56).  The
5-D histogram uses `numpy.histogramdd` semantics exactly (right edge of the last bin
inclusive, `density=True`, optional mass weights); a device histogram is the first "next" row
of SURVEY.md section 8f.
"""
import numpy as np

from . import base


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
        p_tvxz, _ = np.histogramdd(samples, bins=edges, density=True, weights=mass_weights)
        with np.errstate(divide="ignore"):
            log_p_tvxz = np.log(p_tvxz)
        super().__init__(cube=cube, log_p_tvxz=log_p_tvxz)
