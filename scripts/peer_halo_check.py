"""Halo exchange of a slab-decomposed run, peer-memory kernels (CUDA IPC over NVLink) against NCCL send/recv: same results, time
per exchange.  usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/peer_halo_check.py [mesh] [R]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ASORA_QUIET"] = "1"
import numpy as np
import torch, torch.distributed as dist
import bench
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
with bench.quiet_stdout():
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dist.all_reduce(torch.zeros(1, device="cuda"))
import pyc2ray_b200 as p
from pyc2ray_b200.lib import _cabi
from pyc2ray_b200.parallel import SlabHalo, device_tensor
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
R = float(sys.argv[2]) if len(sys.argv) > 2 else 30.0
L, check = _cabi.L, _cabi.check
p.device_init(N, 8)
edges = [(r * N) // world for r in range(world + 1)]
h = int(R) + 1
phi = device_tensor(L.asora_device_buffer(_cabi.BUF_PHI_ION), N ** 3)
xav = device_tensor(L.asora_device_buffer(_cabi.BUF_XH_AV), N ** 3)
res = {}
for mode in (False, True):
    halo = SlabHalo(edges, h, N, rank, world, peer=mode)
    if mode and not halo.peer:
        print(f"rank {rank}: peer mode unavailable", flush=True)
    g = torch.Generator(device="cuda"); g.manual_seed(100 + rank)
    for rep in range(4):
        phi.copy_(torch.rand(N ** 3, generator=g, device="cuda", dtype=torch.float64))
        xav.copy_(torch.rand(N ** 3, generator=g, device="cuda", dtype=torch.float64))
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter(); halo.reduce_phi_(phi); torch.cuda.synchronize(); t1 = time.perf_counter()
        dist.barrier(); t2 = time.perf_counter(); halo.gather_xh_(xav); torch.cuda.synchronize(); t3 = time.perf_counter()
    o, c = halo.own_cells()
    res[mode] = (phi[o:o + c].clone(), xav.clone())
    print(f"rank {rank} {'peer' if halo.peer else 'nccl'}: reduce_phi {1e3*(t1-t0):.3f} ms, gather_xh {1e3*(t3-t2):.3f} ms "
          f"({h} planes, {h*N*N*8/1e6:.1f} MB per halo)", flush=True)
    halo.close(); dist.barrier()
# same random inputs in both modes (generator re-seeded) -> identical results
assert torch.equal(res[False][0], res[True][0]), "peer and NCCL halo reductions differ"
first, count = (edges[rank] - h) % N, (edges[rank + 1] - edges[rank]) + 2 * h
idx = (torch.arange(first, first + count, device="cuda") % N)
a = res[False][1].view(N, N * N)[idx]; b = res[True][1].view(N, N * N)[idx]
assert torch.equal(a, b), "peer and NCCL xh halo gathers differ"
print(f"rank {rank}: peer == nccl", flush=True)
p.device_close(); dist.destroy_process_group()
