"""Multi-GPU test (needs >= 2 CUDA devices, skipped otherwise): source-sharded evolve3D over NCCL must give
every rank the single-GPU result."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, outdir, decomposition="list", R=None, peer_halo="1"):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), ASORA_QUIET="1", ASORA_PEER_HALO=peer_halo)
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


@pytest.mark.parametrize("peer_halo", ["1", "0"])
def test_two_gpu_slab_decomposition_matches_single_gpu(tmp_path, peer_halo):
    """Position-sharded sources with halo exchanges (no N^3 collective inside the convergence loop) must reproduce
    the single-GPU time step; halo planes read from the neighbour's GPU memory (CUDA IPC over NVLink) or sent by NCCL."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path), "slab", 3.3, peer_halo), nprocs=2, join=True)
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


def _peer_worker(rank, world, port):
    """SlabHalo in peer-memory mode against its NCCL mode on random grids: identical results."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), ASORA_QUIET="1")
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import pyc2ray_b200 as p
    from pyc2ray_b200.lib import _cabi
    from pyc2ray_b200.parallel import SlabHalo, device_tensor
    N, h = 48, 7
    p.device_init(N, 8)
    edges = [(r * N) // world for r in range(world + 1)]
    phi = device_tensor(_cabi.L.asora_device_buffer(_cabi.BUF_PHI_ION), N ** 3)
    xav = device_tensor(_cabi.L.asora_device_buffer(_cabi.BUF_XH_AV), N ** 3)
    out = {}
    for peer in (False, True):
        halo = SlabHalo(edges, h, N, rank, world, peer=peer)
        assert halo.peer == peer, "CUDA IPC peer access unavailable"
        g = torch.Generator(device="cuda")
        g.manual_seed(7 + rank)
        phi.copy_(torch.rand(N ** 3, generator=g, device="cuda", dtype=torch.float64))
        xav.copy_(torch.rand(N ** 3, generator=g, device="cuda", dtype=torch.float64))
        torch.cuda.synchronize()
        dist.barrier()
        halo.reduce_phi_(phi)
        torch.cuda.synchronize()
        dist.barrier()
        halo.gather_xh_(xav)
        torch.cuda.synchronize()
        o, c = halo.own_cells()
        first, count = halo.active_range()
        idx = torch.arange(first, first + count, device="cuda") % N
        out[peer] = (phi[o:o + c].clone(), xav.view(N, N * N)[idx].clone())
        halo.close()
        dist.barrier()
    assert torch.equal(out[False][0], out[True][0]) and torch.equal(out[False][1], out[True][1])
    p.device_close()
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_peer_halo_equals_nccl_halo():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_peer_worker, args=(2, port), nprocs=2, join=True)
