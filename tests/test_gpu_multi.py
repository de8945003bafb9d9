"""Multi-GPU test (needs >= 2 CUDA devices, skipped otherwise): source-sharded evolve3D over NCCL must give
every rank the single-GPU result."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, outdir, decomposition="list", R=None):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), ASORA_QUIET="1")
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import pyc2ray_b200 as p
    from tests.fields import make_case
    c = make_case("multi_n32")
    N = c["N"]
    flux = c["flux"] * 1e6
    temp = np.full((N, N, N), 1e4)
    xh0 = np.full((N, N, N), 2e-4)
    chem = (2.59e-13, -0.7, 1.3e-8 * 0.83 / 13.598 ** 2, 13.598 / 8.617e-5, 7.1e-7)
    p.device_init(N, 8)
    p.photo_table_to_device(c["thin"], c["thick"])
    x, phi = p.evolve3D_dist(1e6 * 3.15576e7, c["dr"], flux, c["srcpos"], temp, c["ndens"], xh0, c["thin"], c["thick"],
                             c["minlogtau"], c["dlogtau"], R or c["R"], 1e-4, c["sig"], *chem, logfile=None, quiet=True,
                             decomposition=decomposition)
    np.save(os.path.join(outdir, f"x_{rank}.npy"), x)
    np.save(os.path.join(outdir, f"phi_{rank}.npy"), phi)
    if rank == 0:
        x1, phi1 = p.evolve3D(1e6 * 3.15576e7, c["dr"], flux, c["srcpos"], True, 0, 0, 0, temp, c["ndens"], xh0, c["thin"],
                              c["thick"], c["minlogtau"], c["dlogtau"], R or c["R"], 1e-4, c["sig"], *chem, logfile=None, quiet=True)
        np.save(os.path.join(outdir, "x_single.npy"), x1)
        np.save(os.path.join(outdir, "phi_single.npy"), phi1)
    p.device_close()
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_evolve_matches_single_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    x0, x1, xs = (np.load(tmp_path / f) for f in ("x_0.npy", "x_1.npy", "x_single.npy"))
    p0, p1, ps = (np.load(tmp_path / f) for f in ("phi_0.npy", "phi_1.npy", "phi_single.npy"))
    np.testing.assert_array_equal(x0, x1)      # identical all-reduced phi -> identical deterministic chemistry
    np.testing.assert_array_equal(p0, p1)
    np.testing.assert_allclose(x0, xs, rtol=1e-10, atol=1e-16)
    np.testing.assert_allclose(p0, ps, rtol=1e-10, atol=1e-12 * ps.max())


def test_two_gpu_slab_decomposition_matches_single_gpu(tmp_path):
    """Position-sharded sources with halo exchanges (no N^3 collective inside the convergence loop) must reproduce
    the single-GPU time step."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path), "slab", 3.3), nprocs=2, join=True)
    x0, x1, xs = (np.load(tmp_path / f) for f in ("x_0.npy", "x_1.npy", "x_single.npy"))
    p0, p1, ps = (np.load(tmp_path / f) for f in ("phi_0.npy", "phi_1.npy", "phi_single.npy"))
    np.testing.assert_array_equal(x0, x1)
    np.testing.assert_array_equal(p0, p1)
    assert xs.mean() > 3e-4
    np.testing.assert_allclose(x0, xs, rtol=1e-10, atol=1e-16)
    np.testing.assert_allclose(p0, ps, rtol=1e-10, atol=1e-12 * ps.max())


def test_two_gpu_reduce_scatter_decomposition_matches_single_gpu(tmp_path):
    """List-order sharding with reduce-scatter -> chemistry on the rank's own cells -> all-gather."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path), "rsag"), nprocs=2, join=True)
    x0, x1, xs = (np.load(tmp_path / f) for f in ("x_0.npy", "x_1.npy", "x_single.npy"))
    p0, p1, ps = (np.load(tmp_path / f) for f in ("phi_0.npy", "phi_1.npy", "phi_single.npy"))
    np.testing.assert_array_equal(x0, x1)
    np.testing.assert_array_equal(p0, p1)
    np.testing.assert_allclose(x0, xs, rtol=1e-10, atol=1e-16)
    np.testing.assert_allclose(p0, ps, rtol=1e-10, atol=1e-12 * ps.max())
