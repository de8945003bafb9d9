"""CPU tests (no GPU): the C-ABI library loads and exports what include/asora_b200.h declares, the Python
boundary mirrors the reference's signatures and error behaviour, and the multi-rank host logic
(source sharding + phi_ion reduction) works over gloo with world_size 2."""
import ctypes
import inspect
import os
import re
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "asora_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(asora_\w+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from pyc2ray_b200.lib import _cabi
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(_cabi.L, s), f"{s} declared in include/asora_b200.h but not exported"
    assert sorted(_cabi.SIGNATURES) == syms, "ctypes signature table and header disagree"
    assert b"sm_100a" in _cabi.L.asora_version()


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """The boundary is a C ABI: the header compiles as C99 and a C program links against the library and calls an
    entry point that needs no device (no torch / C++ types in the signatures)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    from pyc2ray_b200.lib import _cabi
    hdr = os.path.join(ROOT, "include", "asora_b200.h")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", hdr], check=True)
    src = tmp_path / "use.c"
    src.write_text('#include "asora_b200.h"\n#include <stdio.h>\n'
                   'int main(void) { printf("%lld %s\\n", (long long)asora_cells_per_source(256, 30.0), asora_version());\n'
                   '  return asora_set_heating(1) == 0; /* no device: must fail, not crash */ }\n')
    exe = tmp_path / "use"
    libdir = os.path.dirname(_cabi.LIB_PATH)
    subprocess.run([gcc, "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe), "-L", libdir,
                    "-l:libasora_b200.so", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert out.stdout.split()[0] == "193025" and "sm_100a" in out.stdout


def test_cells_per_source_host_arithmetic():
    import oracle
    from pyc2ray_b200.lib import _cabi
    for N, R in ((256, 10), (256, 10.76), (256, 30), (250, 30), (128, 1e9), (15, 7.3), (64, 20.0)):
        assert _cabi.L.asora_cells_per_source(N, R) == oracle.cells_per_source(N, R)


def test_no_cpu_fallback_and_error_reporting():
    """Without a device every compute entry point fails loudly with a RuntimeError (never a silent CPU path,
    never an abort), and the CPU-only halves of the reference API say so."""
    import torch
    import pyc2ray_b200 as p
    from pyc2ray_b200.lib import libasora, libc2ray
    with pytest.raises(NotImplementedError):
        libc2ray.raytracing.do_all_sources()
    with pytest.raises(NotImplementedError):
        p.evolve3D(1.0, 1.0, np.ones(1), np.ones((3, 1)), False, 0, 0, 0, *([np.ones((2, 2, 2))] * 3), None, None, 0, 0,
                   1, 1e-4, 0, 0, 0, 0, 0, 0)
    with pytest.raises(RuntimeError, match="not initialized"):
        p.evolve3D(1.0, 1.0, np.ones(1), np.ones((3, 1)), True, 0, 0, 0, *([np.ones((2, 2, 2))] * 3), np.ones(3), np.ones(3),
                   0, 0, 1, 1e-4, 0, 0, 0, 0, 0, 0)
    with pytest.raises(RuntimeError, match="not initialized"):
        p.photo_table_to_device(np.ones(4), np.ones(4))
    with pytest.raises(RuntimeError, match="not initialized"):
        p.device_close()
    with pytest.raises(RuntimeError, match="not initialized"):
        libasora.density_to_device(np.zeros(8), 2)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="libasora_b200"):
            p.device_init(8, 1)
        assert not p.cuda_is_init()
        with pytest.raises(RuntimeError, match="libasora_b200"):
            libc2ray.chemistry.global_pass(1.0, *([np.ones((2, 2, 2))] * 6), 1, 1, 1, 1, 1)


def test_boundary_signatures_match_reference():
    """Positional signatures of the drop-in boundary (src/asora/python_module.cu:21-148, pyc2ray/evolve.py:38-46,
    249-258, pyc2ray/raytracing.py:34-43, pyc2ray/asora_core.py)."""
    import pyc2ray_b200 as p
    from pyc2ray_b200.lib import libasora
    sig = lambda f: list(inspect.signature(f).parameters)
    # the twelve positional parameters of the reference; the multi-GPU extras (group, download) are keyword-only
    pos = [n for n, q in inspect.signature(libasora.do_all_sources).parameters.items() if q.kind != q.KEYWORD_ONLY]
    assert pos == ["R", "coldensh_out", "sig", "dr", "ndens", "xh_av", "phi_ion", "NumSrc", "m1", "minlogtau", "dlogtau", "NumTau"]
    assert sig(libasora.do_all_sources)[12:] == ["group", "download", "xh_from"]
    assert sig(libasora.device_init) == ["N", "num_src_par"]
    assert sig(libasora.density_to_device) == ["ndens", "N"]
    assert sig(libasora.photo_table_to_device) == ["thin_table", "thick_table", "NumTau"]
    assert sig(libasora.source_data_to_device) == ["pos", "flux", "NumSrc"]
    assert sig(p.evolve3D) == ["dt", "dr", "src_flux", "src_pos", "use_gpu", "max_subbox", "subboxsize", "loss_fraction",
                               "temp", "ndens", "xh", "photo_thin_table", "photo_thick_table", "minlogtau", "dlogtau",
                               "R_max_LLS", "convergence_fraction", "sig", "bh00", "albpow", "colh0", "temph0", "abu_c",
                               "logfile", "quiet"]
    assert sig(p.evolve3D_MPI)[:12] == ["dt", "dr", "src_flux", "src_pos", "use_gpu", "max_subbox", "subboxsize",
                                        "loss_fraction", "use_mpi", "comm", "rank", "nprocs"]
    assert sig(p.do_raytracing)[:9] == ["dr", "src_flux", "src_pos", "use_gpu", "max_subbox", "subboxsize", "loss_fraction",
                                        "ndens", "xh_av"]
    assert sig(p.hydrogenODE) == ["dt", "ndens", "temp", "xh", "phi_ion", "bh00", "albpow", "colh0", "abu_c"]
    assert sig(p.device_init) == ["N", "source_batch_size"]


def test_argument_checks_of_the_shim():
    from pyc2ray_b200.lib import libasora
    with pytest.raises(TypeError, match="coldensh_out must be Array of type double"):  # python_module.cu:53-57
        libasora.do_all_sources(1.0, np.zeros(1, dtype=np.float32), 1.0, 1.0, None, np.zeros(8), np.zeros(8), 1, 2, -20.0, 0.1, 10)
    with pytest.raises(TypeError):
        libasora.source_data_to_device(np.zeros(3, dtype=np.int64), np.ones(1), 1)
    with pytest.raises(ValueError):
        libasora.density_to_device(np.zeros(7), 2)


def test_shard_bounds_follow_reference_split():
    """evolve.py:362-367: contiguous blocks of NumSrc//nprocs, remainder to the last rank."""
    from pyc2ray_b200.parallel import shard_bounds
    for ns, nprocs in ((10, 2), (10, 3), (7, 8), (100000, 8), (8, 8), (9, 4)):
        covered = []
        for r in range(nprocs):
            a, b = shard_bounds(ns, r, nprocs)
            assert b - a == ns // nprocs or r == nprocs - 1
            covered += list(range(a, b))
        assert covered == list(range(ns))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _rank_main(rank, world, port, outdir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import oracle
    from pyc2ray_b200.parallel import shard_bounds, allreduce_sum_
    from pyc2ray_b200.utils.sourceutils import format_sources
    from tests.fields import make_case
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    c = make_case("multi_n32")
    a, b = shard_bounds(c["flux"].size, rank, world)
    pos_flat, flux_flat = format_sources(c["srcpos"][:, a:b], c["flux"][a:b])
    phi, _, _ = oracle.asora_do_all_sources(c["R"], c["sig"], c["dr"], c["ndens"].ravel(), c["xh"].ravel(), pos_flat, flux_flat,
                                            c["N"], c["thin"], c["thick"], c["minlogtau"], c["dlogtau"], c["NumTau"])
    t = torch.from_numpy(phi)
    t2 = t.clone()
    allreduce_sum_(t)
    np.save(os.path.join(outdir, f"phi_{rank}.npy"), t.numpy())
    # the reduce-scatter / all-gather exchange (evolve3D_dist, decomposition "rsag"): my chunk of the sum, then the
    # chunks of all ranks put together again
    from pyc2ray_b200.parallel import reduce_scatter_sum_, allgather_chunks_
    mine = reduce_scatter_sum_(t2, rank, world)
    n = t2.numel() // world
    assert mine.data_ptr() == t2[rank * n:(rank + 1) * n].data_ptr()
    assert torch.equal(mine, t[rank * n:(rank + 1) * n])
    t2[:rank * n] = -1.0            # only the rank's own chunk may be used
    t2[(rank + 1) * n:] = -1.0
    allgather_chunks_(t2, rank, world)
    assert torch.equal(t2, t)
    dist.destroy_process_group()


def test_two_rank_source_sharding_over_gloo(tmp_path):
    """world_size-2 run of the N>1 host logic on CPU: each rank ray-traces its shard (the oracle stands in for
    the GPU sweep), the rate grids are summed with the same all-reduce helper the NCCL path uses, and every
    rank ends with the single-process result."""
    import torch.multiprocessing as mp
    import oracle
    from tests.fields import make_case
    port = _free_port()
    mp.spawn(_rank_main, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    c = make_case("multi_n32")
    ref, _, _ = oracle.asora_do_all_sources(c["R"], c["sig"], c["dr"], c["ndens"].ravel(), c["xh"].ravel(), c["pos_flat"],
                                            c["flux_flat"], c["N"], c["thin"], c["thick"], c["minlogtau"], c["dlogtau"],
                                            c["NumTau"])
    p0 = np.load(tmp_path / "phi_0.npy")
    p1 = np.load(tmp_path / "phi_1.npy")
    np.testing.assert_array_equal(p0, p1)
    np.testing.assert_allclose(p0, ref, rtol=1e-13, atol=0)


def test_cosmology_bookkeeping():
    """FlatLambdaCDM stand-in for astropy: the test-case time axis of the reference (zred_0 = 9, slices 10 Myr
    apart) must give the redshift the reference's outputs are named after (xfrac_8.835.pkl) and time steps
    equal to the requested spacing (c2ray_test.py:122-146, c2ray_base.py:147-168)."""
    from pyc2ray_b200.cosmology import FlatLambdaCDM
    c = FlatLambdaCDM(100.0, 0.27, 2.726, Ob0=0.044)
    t9 = c.age(9.0)
    z1 = c.z_at_age(t9 + 1e7 * 3.15576e7)
    assert f"{z1:.3f}" == "8.835"
    dt = c.lookback_time(9.0) - c.lookback_time(z1)
    assert abs(dt / (1e7 * 3.15576e7) - 1.0) < 1e-9
    # matter + Lambda analytic age as a sanity bound (radiation shortens it by < 0.5 % at z = 9)
    H0 = 100 * 1e5 / 3.0856775814913673e24
    an = 2 / (3 * H0 * np.sqrt(0.73)) * np.arcsinh(np.sqrt(0.73 / 0.27) * 0.1 ** 1.5)
    assert 0.995 < t9 / an < 1.0
    assert abs(c.scale_factor(9.0) - 0.1) < 1e-15


def test_slab_edges():
    from pyc2ray_b200.parallel import slab_edges
    rng = np.random.RandomState(0)
    x = rng.randint(0, 256, size=10000)
    edges, h = slab_edges(x, 256, 8, 10.76)
    assert h == 11 and edges[0] == 0 and edges[-1] == 256 and len(edges) == 9
    counts = [np.sum((x >= edges[r]) & (x < edges[r + 1])) for r in range(8)]
    assert sum(counts) == 10000 and max(counts) - min(counts) < 300
    assert all(edges[r + 1] - edges[r] >= h for r in range(8))
    e30, h30 = slab_edges(x, 256, 8, 30.0)                   # 8 slabs of ~32 planes, 31-plane halos inside the neighbours
    assert h30 == 31 and all(31 <= e30[r + 1] - e30[r] <= 256 - 62 for r in range(8))
    assert slab_edges(x, 256, 8, 32.0)[0] is None            # a 33-plane halo does not fit a 32-plane neighbour
    assert slab_edges(x, 256, 2, 64.0)[0] is None            # two ranks: a slab and its two halos would meet themselves
    assert slab_edges(x, 256, 1, 5.0)[0] is None
    clustered = np.full(1000, 17)
    e2, h2 = slab_edges(clustered, 128, 2, 4.5)               # all sources in one plane: widths are enforced
    assert e2 is not None and all(h2 <= e2[r + 1] - e2[r] <= 128 - 2 * h2 for r in range(2))


def _slab_rank_main(rank, world, port, outdir, R=3.3):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import oracle
    from pyc2ray_b200.parallel import slab_edges, SlabHalo
    from pyc2ray_b200.utils.sourceutils import format_sources
    from tests.fields import make_case
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    c = make_case("multi_n32")
    N = c["N"]
    x0 = c["srcpos"][0] - 1
    edges, h = slab_edges(x0, N, world, R)
    halo = SlabHalo(edges, h, N, rank, world)
    mine = (x0 >= edges[rank]) & (x0 < edges[rank + 1])
    pos_flat, flux_flat = format_sources(c["srcpos"][:, mine], c["flux"][mine])
    phi, _, _ = oracle.asora_do_all_sources(R, c["sig"], c["dr"], c["ndens"].ravel(), c["xh"].ravel(), pos_flat, flux_flat, N,
                                            c["thin"], c["thick"], c["minlogtau"], c["dlogtau"], c["NumTau"])
    t = torch.from_numpy(phi)
    halo.reduce_phi_(t)
    # xh halo gather: every rank marks its own planes with its rank id
    xh = torch.full((N ** 3,), -1.0, dtype=torch.float64)
    o, cnt = halo.own_cells()
    xh[o:o + cnt] = float(rank)
    halo.gather_xh_(xh)
    np.save(os.path.join(outdir, f"xh_{rank}.npy"), xh.numpy().copy())
    halo.assemble_(t)
    np.save(os.path.join(outdir, f"phi_{rank}.npy"), t.numpy())
    np.save(os.path.join(outdir, f"edges_{rank}.npy"), np.array(edges + [h]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,R", [(2, 3.3), (3, 3.3), (3, 7.3)])   # the last: slabs thinner than two halos
def test_slab_decomposition_over_gloo(tmp_path, world, R):
    """Slab-sharded sources + halo reduction of the rates + assembly == single-process result; the xh_av halo gather
    delivers the neighbours' planes.  CPU ranks over gloo, the oracle standing in for the GPU sweep."""
    import torch.multiprocessing as mp
    import oracle
    from tests.fields import make_case
    port = _free_port()
    mp.spawn(_slab_rank_main, args=(world, port, str(tmp_path), R), nprocs=world, join=True)
    c = make_case("multi_n32")
    N = c["N"]
    ref, _, _ = oracle.asora_do_all_sources(R, c["sig"], c["dr"], c["ndens"].ravel(), c["xh"].ravel(), c["pos_flat"],
                                            c["flux_flat"], N, c["thin"], c["thick"], c["minlogtau"], c["dlogtau"], c["NumTau"])
    e = np.load(tmp_path / "edges_0.npy")
    edges, h = list(e[:-1]), int(e[-1])
    for r in range(world):
        np.testing.assert_allclose(np.load(tmp_path / f"phi_{r}.npy"), ref, rtol=1e-12, atol=0)
        xh = np.load(tmp_path / f"xh_{r}.npy").reshape(N, N, N)
        assert (xh[edges[r]:edges[r + 1]] == r).all()
        assert (xh[[(edges[r] - k) % N for k in range(1, h + 1)]] == (r - 1) % world).all()
        assert (xh[[(edges[r + 1] + k) % N for k in range(h)]] == (r + 1) % world).all()


def test_cosmology_matches_astropy_printed_age():
    """test/paper_tests/test2_Ifront_cosmo/make_plot.ipynb cell 5 prints astropy's FlatLambdaCDM(70, 0.27, 2.726,
    Ob0=0.043).age(9) in Myr and lambda = t_i / t_rec; the stand-in must reproduce both."""
    import json
    from pyc2ray_b200.cosmology import FlatLambdaCDM
    k = json.load(open(os.path.join(ROOT, "tests", "golden", "kat.json")))["test2_cosmo_Ifront"]
    c = FlatLambdaCDM(k["cosmology"]["H0"], k["cosmology"]["Om0"], k["cosmology"]["Tcmb0"], Ob0=k["cosmology"]["Ob0"])
    myr = 1e6 * 365.25 * 86400.0
    ti = c.age(9.0) / myr
    assert abs(ti / k["age_z9_Myr"] - 1.0) < 1e-10
    t_rec = 1.0 / (2.59e-13 * 1.87e-4 * 3.15576e7 * 1e6)
    assert abs(ti / t_rec / k["lambda"] - 1.0) < 1e-10
