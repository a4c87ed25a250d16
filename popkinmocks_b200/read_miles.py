"""MILES SSP library I/O (host side, off the hot path).

Mirrors the public surface of the reference reader
(/root/reference/popkinmocks/read_miles.py:8-200): an object with `X`
[n_lambda, nz*nt] (metallicity-major, age-minor column order, read_miles.py:157-161),
`lmd`, `z_unq`, `t_unq`, `nz`, `nt`, and `truncate_wavelengths`.

Two sources are supported:
  * the packed library shipped in `popkinmocks_b200/data/<mod_dir>.npz` (the 636
    MILES_BASTI_CH_baseFe spectra exactly as stored in the FITS primary HDUs,
    float32; made by `python -m popkinmocks_b200.read_miles --pack <fits_dir>`);
  * a directory of MILES web-tool FITS files (`Mch1.30Zm0.25T00.0300_iTp0.00_baseFe.fits`
    naming), parsed with a dependency-free primary-HDU reader (astropy is not
    needed).
"""
import os
import re

import numpy as np

_DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
_NAME_RE = re.compile(r"^(?P<imf>[A-Za-z]+[0-9.]+)Z(?P<zs>[mp])(?P<z>[0-9.]+)T(?P<t>[0-9.]+)_i(?P<tail>.*)$")


def read_fits_primary(path):
    """Return the 1-D primary-HDU image of a FITS file as float32 plus its header dict."""
    with open(path, "rb") as fh:
        raw = fh.read()
    header = {}
    pos = 0
    finished = False
    while not finished:
        if pos + 2880 > len(raw):
            raise ValueError(f"{path}: FITS header has no END card")
        for c in range(pos, pos + 2880, 80):
            card = raw[c:c + 80].decode("ascii", errors="replace")
            name = card[:8].rstrip()
            if name == "END":
                finished = True
                break
            if card[8:10] == "= ":
                header[name] = card[10:].split("/", 1)[0].strip().strip("'").strip()
        pos += 2880
    bitpix = int(header["BITPIX"])
    if int(header["NAXIS"]) != 1:
        raise ValueError(f"{path}: expected a 1-D spectrum")
    n = int(header["NAXIS1"])
    dtypes = {-32: ">f4", -64: ">f8", 16: ">i2", 32: ">i4"}
    if bitpix not in dtypes:
        raise ValueError(f"{path}: unsupported BITPIX {bitpix}")
    data = np.frombuffer(raw, dtype=dtypes[bitpix], count=n, offset=pos)
    return data.astype(np.float32 if bitpix == -32 else np.float64), header


def parse_model_filename(name):
    """(metallicity [M/H], age Gyr) encoded in a MILES model file name, or None."""
    m = _NAME_RE.match(name)
    if m is None:
        return None
    z = float(m.group("z"))
    if m.group("zs") == "m":
        z = -z
    return z, float(m.group("t"))


def load_fits_directory(fits_dir):
    """Read every model in `fits_dir` -> (X[n_lambda, nz*nt] f32, z_unq, t_unq, crval, cdelt)."""
    entries = []
    for name in sorted(os.listdir(fits_dir)):
        if not name.endswith(".fits"):
            continue
        zt = parse_model_filename(name)
        if zt is None:
            raise ValueError(f"cannot parse MILES model name: {name}")
        entries.append((zt[0], zt[1], name))
    if not entries:
        raise ValueError(f"no FITS models in {fits_dir}")
    z_unq = np.unique([e[0] for e in entries])
    t_unq = np.unique([e[1] for e in entries])
    if len(entries) != z_unq.size * t_unq.size:
        raise ValueError("Not age metallicity grid")
    lookup = {(e[0], e[1]): e[2] for e in entries}
    X = None
    crval = cdelt = None
    for iz, z in enumerate(z_unq):
        for it, t in enumerate(t_unq):
            spec, hdr = read_fits_primary(os.path.join(fits_dir, lookup[(z, t)]))
            if X is None:
                X = np.zeros((spec.size, z_unq.size * t_unq.size), dtype=np.float32)
                crval, cdelt = float(hdr.get("CRVAL1", 3540.5)), float(hdr.get("CDELT1", 0.9))
            X[:, iz * t_unq.size + it] = spec
    return X, z_unq, t_unq, crval, cdelt


def pack_fits_directory(fits_dir, out_path):
    X, z_unq, t_unq, crval, cdelt = load_fits_directory(fits_dir)
    np.savez_compressed(out_path, X=X, z_unq=z_unq, t_unq=t_unq,
                        crval=np.float64(crval), cdelt=np.float64(cdelt))
    return out_path


class milesSSPs:
    """Grid of MILES SSP spectra on their native linear wavelength axis.

    Args follow the reference (read_miles.py:9-16).  `mod_dir` may name a packed
    library in the package data directory or be a path to a directory of FITS files.
    """

    def __init__(self, mod_dir="MILES_BASTI_CH_baseFe", age_lim=None, z_lim=None,
                 thin_age=1, thin_z=1):
        if mod_dir is None:
            self.list_mod_directories()
            return
        self.mod_dir = mod_dir
        self.age_lim = age_lim
        self.z_lim = z_lim
        self.thin_age = thin_age
        self.thin_z = thin_z
        self._load_library()
        self._select()
        self.get_lmd()
        if self.X.shape[0] != self.lmd.size:
            raise ValueError("len(wavelength array) != len(spectrum)")

    def list_mod_directories(self):
        print("Initilise with 'mod_dir' set to:")
        names = sorted(n[:-4] for n in os.listdir(_DATA_DIR) if n.endswith(".npz"))
        for i, d in enumerate(names):
            print(f"{i+1}) {d}")

    def _load_library(self):
        name = self.mod_dir.rstrip("/")
        packed = os.path.join(_DATA_DIR, os.path.basename(name) + ".npz")
        if os.path.isdir(name):
            X, z_unq, t_unq, crval, cdelt = load_fits_directory(name)
        elif os.path.isfile(packed):
            with np.load(packed) as lib:
                X, z_unq, t_unq = lib["X"], lib["z_unq"], lib["t_unq"]
        else:
            raise FileNotFoundError(
                f"no SSP library '{self.mod_dir}': neither a FITS directory nor {packed}")
        self._lib_X = X
        self._lib_z = np.asarray(z_unq, dtype=np.float64)
        self._lib_t = np.asarray(t_unq, dtype=np.float64)

    def _select(self):
        # age / metallicity window (reference defaults: read_miles.py:80-88)
        t_lo, t_hi = (-3, 30) if self.age_lim is None else self.age_lim
        z_lo, z_hi = (-3, 3) if self.z_lim is None else self.z_lim
        keep_t = np.where((self._lib_t >= t_lo) & (self._lib_t <= t_hi))[0]
        keep_z = np.where((self._lib_z >= z_lo) & (self._lib_z <= z_hi))[0]
        # thinning keeps every n-th unique value (read_miles.py:102-105)
        keep_t = keep_t[:: self.thin_age]
        keep_z = keep_z[:: self.thin_z]
        if keep_t.size == 0 or keep_z.size == 0:
            raise ValueError("Not age metallicity grid")
        nt_lib = self._lib_t.size
        cols = (keep_z[:, None] * nt_lib + keep_t[None, :]).ravel()
        self.X = np.ascontiguousarray(self._lib_X[:, cols], dtype=np.float64)
        self.z_unq = self._lib_z[keep_z]
        self.t_unq = self._lib_t[keep_t]
        self.nz = self.z_unq.size
        self.nt = self.t_unq.size
        zz, tt = np.meshgrid(self.z_unq, self.t_unq, indexing="ij")
        self.z = zz.ravel()
        self.t = tt.ravel()

    def get_lmd(self):
        # fixed MILES sampling (read_miles.py:166-167)
        self.lmd = np.arange(3540.5, 7409.6 + 0.9, 0.9)

    def truncate_wavelengths(self, lmd_min=None, lmd_max=None):
        if lmd_min is None:
            lmd_min = np.min(self.lmd)
        if lmd_max is None:
            lmd_max = np.max(self.lmd)
        if lmd_min < np.min(self.lmd) or lmd_max > np.max(self.lmd):
            msg = "Requested wavelength range outside of SSP limits:, {0}-{1}"
            raise ValueError(msg.format(np.min(self.lmd), np.max(self.lmd)))
        sel = np.where((self.lmd >= lmd_min) & (self.lmd <= lmd_max))[0]
        self.X = self.X[sel, :]
        self.lmd = self.lmd[sel]

    def reset(self):
        self._select()
        self.get_lmd()


if __name__ == "__main__":
    import argparse

    ap = argparse.ArgumentParser(description="pack a directory of MILES FITS models")
    ap.add_argument("--pack", required=True, help="directory of *.fits models")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    out = a.out or os.path.join(_DATA_DIR, os.path.basename(a.pack.rstrip("/")) + ".npz")
    print(pack_fits_directory(a.pack, out))
