"""Load the UNMODIFIED reference popkinmocks from /root/reference in this container.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this file.  It
exists so that `oracle/make_golden.py` can execute the reference itself and pin
both the NumPy restatement (`oracle/restatement.py`) and the CUDA path against
real reference outputs.  `/root/reference` does not exist on the GPU box, so
nothing that runs there may import this module.

The reference cannot be imported as is:
  * `matplotlib` and `astropy` are not installed (imported at
    popkinmocks/ifu_cube.py:2, popkinmocks/model_grids.py:4,
    popkinmocks/read_miles.py:3);
  * `milesSSPs()` with default limits crashes because `all_spectra.npy` is a
    missing large blob (popkinmocks/read_miles.py:143-164).
The shim stubs the two plotting/IO modules in `sys.modules` (the FITS stub is a
minimal primary-HDU reader: 2880-byte header block, BITPIX=-32 big-endian f32)
and callers construct `milesSSPs(age_lim=(-3, 30))`, which is identical to the
intended default grid (popkinmocks/read_miles.py:80-88).
"""
import os
import sys
import types
import warnings

import numpy as np

REFERENCE_ROOT = os.environ.get("PKM_REFERENCE_ROOT", "/root/reference")


class _HDU:
    def __init__(self, data):
        self.data = data


def _fits_open(path):
    with open(path, "rb") as f:
        raw = f.read()
    cards = {}
    off = 0
    done = False
    while not done:
        block = raw[off:off + 2880]
        off += 2880
        for i in range(0, 2880, 80):
            card = block[i:i + 80].decode("ascii")
            key = card[:8].strip()
            if key == "END":
                done = True
                break
            if card[8:10] == "= ":
                cards[key] = card[10:].split("/")[0].strip()
    assert int(cards["BITPIX"]) == -32 and int(cards["NAXIS"]) == 1
    n = int(cards["NAXIS1"])
    data = np.frombuffer(raw, dtype=">f4", count=n, offset=off).astype(np.float32)
    return [_HDU(data)]


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "popkinmocks"))


def load():
    """Return the reference `popkinmocks` module (imported under the shim)."""
    if "popkinmocks" in sys.modules:
        return sys.modules["popkinmocks"]
    if not available():
        raise RuntimeError(f"reference tree not present at {REFERENCE_ROOT}")
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    if "astropy" not in sys.modules:
        astropy = types.ModuleType("astropy")
        io = types.ModuleType("astropy.io")
        fits = types.ModuleType("astropy.io.fits")
        fits.open = _fits_open
        io.fits = fits
        astropy.io = io
        sys.modules["astropy"] = astropy
        sys.modules["astropy.io"] = io
        sys.modules["astropy.io.fits"] = fits
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", SyntaxWarning)
        import popkinmocks  # noqa: E402

    return popkinmocks


def reference_ssps(pkm=None, **kwargs):
    """`pkm.milesSSPs` on the default grid, dodging the missing-blob crash."""
    pkm = pkm or load()
    kwargs.setdefault("age_lim", (-3, 30))
    return pkm.milesSSPs(**kwargs)
