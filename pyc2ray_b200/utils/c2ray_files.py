"""On-disk formats either side of the hot path: the C2Ray binary grids pyc2ray writes and resumes from,
the halo-source catalogues it reads, and the redshift bookkeeping of the 244 Mpc run.

The reference does this I/O through tools21cm (a pyproject dependency without a pinned version, absent from this image
and from /root/reference) and h5py 3.8.0.  What is restated here is anchored on the reference's own call sites:

* grids: ``t2c.save_cbin(filename, data, bits, order)`` / ``t2c.read_cbin(filename, bits, order)`` as called at
  c2ray_244paper.py:282-283 (write_output: xfrac 64 bit, IonRates 32 bit, order 'F') and :327,331 (resume).  The layout
  is tools21cm's published "cbin" one: three int32 mesh dimensions, then the raw values.
* sources: the HDF5 catalogue layout is defined in the reference itself (utils/source_converter.py:52-61: datasets
  ``sources_positions`` (numsrc, 3) and ``sources_mass`` in solar masses); the mass -> photon-rate conversion is
  c2ray_244paper.py:221,229.  The legacy text catalogues go through ``t2c.SourceFile``, whose column conventions are
  tools21cm's and are not restated.
* densities: ``t2c.DensityFile(...).cgs_density`` (c2ray_244paper.py:270) converts CubeP3M grid masses with tools21cm's
  simulation constants; only the file layout (three int32, float32 values, Fortran order) is handled here and the
  conversion factor is an argument.
"""
import glob
import os

import numpy as np

__all__ = ["save_cbin", "read_cbin", "write_output_cbin", "read_output_cbin", "sources_from_catalogue",
           "read_sources_hdf5", "write_sources_hdf5", "read_density_cbin", "get_source_redshifts",
           "get_redshifts_from_output", "find_bins", "M_P", "MSUN2G"]

M_P = 1.672661e-24      # c2ray_244paper.py:22
MSUN2G = 1.98892e33     # c2ray_base.py:80


def save_cbin(filename, data, bits=32, order="C"):
    """Three int32 mesh sizes followed by the values as float32 / float64 in the given memory order."""
    if bits not in (32, 64):
        raise ValueError("bits must be 32 or 64")
    data = np.asarray(data)
    with open(filename, "wb") as f:
        np.array(data.shape, dtype="int32").tofile(f)
        data.flatten(order=order).astype(np.float32 if bits == 32 else np.float64).tofile(f)


def read_cbin(filename, bits=32, order="C", dimensions=3):
    if bits not in (32, 64):
        raise ValueError("bits must be 32 or 64")
    with open(filename, "rb") as f:
        mesh = np.fromfile(f, count=dimensions, dtype="int32")
        if mesh.size != dimensions or (mesh <= 0).any():
            raise ValueError(f"{filename}: not a cbin file (mesh header {mesh})")
        n = int(np.prod(mesh.astype(np.int64)))
        data = np.fromfile(f, dtype=np.float32 if bits == 32 else np.float64, count=n)
    if data.size != n:
        raise ValueError(f"{filename}: {data.size} values, header promises {n}")
    return data.reshape(tuple(int(m) for m in mesh), order=order)


def write_output_cbin(results_basename, z, xh, phi_ion):
    """C2Ray_244Paper.write_output (c2ray_244paper.py:273-283): xfrac_<z>.dat in 64 bit, IonRates_<z>.dat in 32 bit."""
    suffix = f"_{z:.3f}.dat"
    save_cbin(results_basename + "xfrac" + suffix, xh, bits=64, order="F")
    save_cbin(results_basename + "IonRates" + suffix, phi_ion, bits=32, order="F")
    return results_basename + "xfrac" + suffix, results_basename + "IonRates" + suffix


def read_output_cbin(results_basename, z):
    """The resume path (c2ray_244paper.py:327,331).  Returns (xh float64, phi_ion float32), Fortran-ordered."""
    suffix = f"_{z:.3f}.dat"
    return (read_cbin(results_basename + "xfrac" + suffix, bits=64, order="F"),
            read_cbin(results_basename + "IonRates" + suffix, bits=32, order="F"))


def sources_from_catalogue(positions, mass_msun, fgamma_hm, Ob0, Om0, ts_seconds, S_star_ref=1e48):
    """(srcpos (3, numsrc), normflux) from halo positions (numsrc, 3; 1-indexed mesh cells) and masses in solar
    masses: normflux = mass * msun2g * fgamma_hm * Ob0 / (m_p * ts * Om0) / S_star_ref (c2ray_244paper.py:221,229)."""
    positions = np.asarray(positions)
    if positions.ndim != 2 or positions.shape[1] != 3:
        raise ValueError("positions must have shape (numsrc, 3)")
    mass2phot = MSUN2G * fgamma_hm * Ob0 / (M_P * ts_seconds * Om0)
    srcpos = positions.T
    normflux = np.asarray(mass_msun, dtype=np.float64) * mass2phot / S_star_ref
    return srcpos, normflux


def _h5py():
    try:
        import h5py
        return h5py
    except ImportError as e:  # not a silent fallback: the HDF5 catalogues cannot be read without it
        raise ImportError("reading / writing HDF5 source catalogues needs h5py (the reference pins h5py==3.8.0)") from e


def read_sources_hdf5(file, fgamma_hm, Ob0, Om0, ts_seconds, S_star_ref=1e48):
    """C2Ray_244Paper.read_sources for ``*.hdf5`` catalogues (c2ray_244paper.py:223-229)."""
    h5py = _h5py()
    with h5py.File(file, "r") as f:
        pos = f["sources_positions"][:]
        mass = f["sources_mass"][:]
    return sources_from_catalogue(pos, mass, fgamma_hm, Ob0, Om0, ts_seconds, S_star_ref)


def write_sources_hdf5(file, positions, mass_msun, z=None, masstype="hm"):
    """The catalogue layout of utils/source_converter.py:52-61 (zero-mass halos dropped, :48)."""
    h5py = _h5py()
    positions, mass_msun = np.asarray(positions), np.asarray(mass_msun)
    keep = mass_msun != 0
    with h5py.File(file, "w") as f:
        f.create_dataset("sources_positions", data=positions[keep])
        d = f.create_dataset("sources_mass", data=mass_msun[keep])
        if z is not None:
            f.attrs["z"] = z
        f.attrs["masstype"] = masstype
        f.attrs["filename"] = file
        d.attrs["unit"] = "Solar Mass"


def read_density_cbin(filename, to_cgs=1.0):
    """Coarse-grained CubeP3M density file ``<z>n_all.dat``: three int32 mesh sizes, float32 grid masses in Fortran
    order.  ``to_cgs`` is tools21cm's grid-mass -> g/cm^3 factor for the simulation (DensityFile.cgs_density); divide the
    result by mean_molecular * m_p and scale by (1+z)^3 as c2ray_244paper.py:270 does."""
    return read_cbin(filename, bits=32, order="F").astype(np.float64) * to_cgs


def _in_range(redshifts, z_low, z_high, bracket):
    """other_utils.py:103-128"""
    redshifts = np.sort(np.array(redshifts, dtype=float))
    if bracket:
        if z_low < redshifts.min() or z_high > redshifts.max():
            raise Exception("No redshifts to bracket range.")
        z_low = redshifts[redshifts <= z_low][-1]
        z_high = redshifts[redshifts >= z_high][0]
    if z_low is None:
        z_low = redshifts.min() - 1 if redshifts.size else 0.0
    if z_high is None:
        z_high = redshifts.max() + 1 if redshifts.size else 0.0
    return redshifts[(redshifts >= z_low) & (redshifts <= z_high)]


def get_source_redshifts(source_dir, z_low=None, z_high=None, bracket=False):
    """Redshifts of the ``<z>-coarsest_wsubgrid_sources.dat`` catalogues in a directory (other_utils.py:66-93)."""
    tag = "-coarsest_wsubgrid_sources"
    zs = []
    for f in glob.glob(os.path.join(source_dir, "*" + tag + ".dat")):
        name = os.path.basename(f)
        try:
            zs.append(float(name[:name.rfind(tag)]))
        except ValueError:
            pass
    return _in_range(zs, z_low, z_high, bracket)


def get_redshifts_from_output(output_dir, z_low=None, z_high=None, bracket=False):
    """Redshifts for which ``xfrac_<z>.dat`` exists (the resume logic of c2ray_244paper.py:308)."""
    zs = []
    for f in glob.glob(os.path.join(output_dir, "xfrac_*.dat")):
        name = os.path.basename(f)
        try:
            zs.append(float(name[len("xfrac_"):-len(".dat")]))
        except ValueError:
            pass
    return _in_range(zs, z_low, z_high, bracket)


def find_bins(value, bin_edges):
    """(left, right) neighbours of ``value`` in the sorted ``bin_edges``; None beyond the ends (other_utils.py:9-63,
    scalar branch)."""
    edges = np.sort(np.asarray(bin_edges, dtype=float))
    idx = int(np.digitize(value, edges))
    if 0 < idx < len(edges):
        return edges[idx - 1], edges[idx]
    if idx == 0:
        return None, edges[0]
    return edges[idx - 1], None
