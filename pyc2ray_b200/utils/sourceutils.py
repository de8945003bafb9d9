"""Source-list wire format of the ASORA boundary (reference: pyc2ray/utils/sourceutils.py)."""
import numpy as np

__all__ = ["format_sources", "generate_test_sources", "generate_test_sourcefile", "read_test_sources"]


def format_sources(source_pos, source_flux):
    """(3,Ns) 1-indexed positions -> int32[3*Ns] interleaved xyz, 0-indexed; flux -> float64[Ns].

    sourceutils.py:7-33: ``ravel((pos-1).astype(int32), order='F')`` of a (3,Ns) array gives
    ``[x0,y0,z0,x1,y1,z1,...]``.
    """
    pos0 = (np.asarray(source_pos) - 1).astype("int32")
    return np.ravel(pos0, order="F"), np.asarray(source_flux).astype("float64")


def generate_test_sources(N, numsrc, seed=100):
    """Seeded random 1-indexed positions, shape (3,numsrc), drawn exactly as
    generate_test_sourcefile does (sourceutils.py:55-58)."""
    rng = np.random.RandomState(seed)
    srcpos = 1 + rng.randint(0, N, size=3 * numsrc)
    return np.ascontiguousarray(srcpos.reshape((numsrc, 3), order="C").T)


def generate_test_sourcefile(filename, N, numsrc, strength, seed=100):
    """Write a C2Ray-formatted source file (sourceutils.py:35-68)."""
    pos = generate_test_sources(N, numsrc, seed).T
    out = np.hstack((pos, strength * np.ones((numsrc, 1)), np.zeros((numsrc, 1))))
    with open(filename, "w") as f:
        f.write(f"{numsrc:n}\n")
        np.savetxt(f, out, ("%i %i %i %.0e %.1f"))


def read_test_sources(file, numsrc, S_star_ref=1e48):
    """Read ``x y z flux dummy`` rows after a one-line header (sourceutils.py:70-112).

    Returns src_pos (3,numsrc) and src_flux (numsrc) in units of S_star_ref.
    """
    inp = np.loadtxt(file, skiprows=1, usecols=(0, 1, 2, 3), ndmin=2)
    if numsrc > inp.shape[0]:
        raise ValueError(f"Number of sources given ({numsrc:n}) is larger than that of the file ({inp.shape[0]:n})")
    return np.transpose(inp[:numsrc, 0:3]), inp[:numsrc, 3] / S_star_ref
