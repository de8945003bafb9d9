"""On-disk formats: byte layout of the C2Ray binary grids as the reference's call sites use them
(c2ray_244paper.py:282-283,327,331), the HDF5 catalogue conversion (c2ray_244paper.py:221-229) and the redshift
bookkeeping against a fixture produced by the reference's own utils/other_utils.py (tests/golden/make_golden.py)."""
import json
import os
import struct

import numpy as np
import pytest

from pyc2ray_b200.utils import c2ray_files as cf
from tests.fields import GOLDEN


def test_cbin_byte_layout_known_answer(tmp_path):
    a = np.arange(24, dtype=np.float64).reshape(2, 3, 4) / 8.0
    f32, f64 = str(tmp_path / "a32.dat"), str(tmp_path / "a64.dat")
    cf.save_cbin(f32, a, bits=32, order="F")
    cf.save_cbin(f64, a, bits=64, order="F")
    raw = open(f32, "rb").read()
    assert len(raw) == 12 + 24 * 4 and struct.unpack("<3i", raw[:12]) == (2, 3, 4)
    # Fortran order: the first index runs fastest -> a[0,0,0], a[1,0,0], a[0,1,0], ...
    assert struct.unpack("<4f", raw[12:28]) == (a[0, 0, 0], a[1, 0, 0], a[0, 1, 0], a[1, 1, 0])
    raw = open(f64, "rb").read()
    assert len(raw) == 12 + 24 * 8 and struct.unpack("<2d", raw[12:28]) == (a[0, 0, 0], a[1, 0, 0])
    np.testing.assert_array_equal(cf.read_cbin(f64, bits=64, order="F"), a)
    np.testing.assert_array_equal(cf.read_cbin(f32, bits=32, order="F"), a.astype(np.float32))
    # C order is the library default (t2c.save_cbin(..., order='C'))
    cf.save_cbin(f32, a)
    np.testing.assert_array_equal(cf.read_cbin(f32), a.astype(np.float32))
    with pytest.raises(ValueError):
        cf.read_cbin(f32, bits=64)  # too few values for the header
    with pytest.raises(ValueError):
        cf.save_cbin(f32, a, bits=16)


def test_write_output_and_resume_roundtrip(tmp_path):
    rng = np.random.default_rng(1)
    xh = np.asfortranarray(rng.uniform(1e-4, 1.0, size=(6, 6, 6)))
    phi = np.asfortranarray(10 ** rng.uniform(-20, -10, size=(6, 6, 6)))
    base = str(tmp_path) + "/"
    fx, fi = cf.write_output_cbin(base, 9.9384, xh, phi)
    assert os.path.basename(fx) == "xfrac_9.938.dat" and os.path.basename(fi) == "IonRates_9.938.dat"
    x2, p2 = cf.read_output_cbin(base, 9.938)
    assert x2.flags.f_contiguous and x2.dtype == np.float64 and p2.dtype == np.float32
    np.testing.assert_array_equal(x2, xh)                       # 64 bit: exact
    np.testing.assert_allclose(p2, phi, rtol=6e-8, atol=0)       # 32 bit, as the reference stores the rates
    np.testing.assert_array_equal(cf.get_redshifts_from_output(base), [9.938])


def test_catalogue_to_sources():
    pos = np.array([[1, 2, 3], [250, 249, 1]])
    mass = np.array([1e9, 3e10])
    ts = 10e6 * 3.15576e7
    srcpos, flux = cf.sources_from_catalogue(pos, mass, fgamma_hm=30.0, Ob0=0.044, Om0=0.27, ts_seconds=ts)
    assert srcpos.shape == (3, 2) and (srcpos[:, 1] == [250, 249, 1]).all()
    # c2ray_244paper.py:221: photons/s per solar mass = msun2g fgamma Ob0 / (m_p ts Om0)
    expect = mass * 1.98892e33 * 30.0 * 0.044 / (1.672661e-24 * ts * 0.27) / 1e48
    np.testing.assert_allclose(flux, expect, rtol=1e-15)
    assert 1e3 < flux[0] < 1e5  # a 1e9 Msun halo: ~2e52 photons/s, in units of 1e48
    with pytest.raises(ValueError):
        cf.sources_from_catalogue(pos.T, mass, 30.0, 0.044, 0.27, ts)


def test_hdf5_catalogue_needs_h5py_or_roundtrips(tmp_path):
    try:
        import h5py  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError, match="h5py"):
            cf.read_sources_hdf5(str(tmp_path / "x.hdf5"), 30.0, 0.044, 0.27, 1.0)
        return
    f = str(tmp_path / "10.478-coarsest_wsubgrid_sources.hdf5")
    cf.write_sources_hdf5(f, np.array([[1, 2, 3], [4, 5, 6], [7, 8, 9]]), np.array([1e9, 0.0, 2e9]), z=10.478)
    srcpos, flux = cf.read_sources_hdf5(f, 30.0, 0.044, 0.27, 3.15576e14)
    assert srcpos.shape == (3, 2) and flux.shape == (2,)  # the zero-mass halo is dropped (source_converter.py:48)


def test_density_cbin(tmp_path):
    g = np.asfortranarray(np.random.default_rng(2).uniform(0.5, 2.0, size=(4, 4, 4)).astype(np.float32))
    f = str(tmp_path / "9.938n_all.dat")
    cf.save_cbin(f, g, bits=32, order="F")
    np.testing.assert_allclose(cf.read_density_cbin(f, to_cgs=2.5e-30), g.astype(np.float64) * 2.5e-30, rtol=0)


def test_redshift_bookkeeping_matches_reference_utils(tmp_path):
    g = json.load(open(os.path.join(GOLDEN, "ref_other_utils.json")))
    res, src = tmp_path / "results", tmp_path / "sources"
    res.mkdir(), src.mkdir()
    for f in g["run_files"]:
        (res / f).touch()
    for f in g["src_files"]:
        (src / f).touch()
    assert list(cf.get_redshifts_from_output(str(res))) == g["from_output"]
    assert list(cf.get_source_redshifts(str(src))) == g["source_redshifts"]
    assert list(cf.get_source_redshifts(str(src), 9.0, 11.0)) == g["source_redshifts_9_11"]
    assert list(cf.get_source_redshifts(str(src), 9.0, 11.0, True)) == g["source_redshifts_bracket"]
    for v, (lo, hi) in g["find_bins"].items():
        assert cf.find_bins(float(v), g["edges"]) == (lo, hi)
    # beyond the last edge the reference indexes out of bounds (other_utils.py:50-52); here: (last, None)
    assert cf.find_bins(30.0, g["edges"]) == (21.062, None)
