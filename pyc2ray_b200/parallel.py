"""Source sharding over ranks and the phi_ion reduction (reference: pyc2ray/evolve.py:360-373,433-437).

One process per GPU.  The ray trace is embarrassingly parallel over sources; the only exchange on the
path is the sum of the per-rank rate grids, done as one in-place all-reduce (NCCL over NVLink on
GPUs; gloo on CPU tensors in the host-logic tests).  Every rank then runs the (cheap, deterministic)
chemistry pass on the full grid, which replaces the reference's rank-0 chemistry + four N^3 broadcasts
(evolve.py:439-497) with zero further traffic.
"""
import numpy as np

__all__ = ["shard_bounds", "allreduce_sum_", "reduce_scatter_sum_", "allgather_chunks_", "DeviceView", "device_tensor"]


def shard_bounds(NumSrc, rank, nprocs):
    """Contiguous block split, remainder to the last rank (evolve.py:362-367)."""
    perrank = NumSrc // nprocs
    i_start = int(rank * perrank)
    i_end = int((rank + 1) * perrank) if rank != nprocs - 1 else NumSrc
    return i_start, i_end


def allreduce_sum_(tensor, group=None):
    """In-place sum over ranks of a torch tensor (CUDA -> NCCL, CPU -> gloo)."""
    import torch.distributed as dist
    dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
    return tensor


def reduce_scatter_sum_(full, rank, nprocs, group=None):
    """Sum over ranks of ``full`` (n doubles, n divisible by nprocs), leaving rank r's chunk r summed in place; the
    other chunks of ``full`` are garbage afterwards.  NCCL: one in-place reduce-scatter (half the traffic of an
    all-reduce); gloo has no reduce-scatter, so the CPU host-logic tests use an all-reduce, which leaves the same
    chunk.  Returns the rank's chunk (a view)."""
    import torch.distributed as dist
    n = full.numel() // nprocs
    mine = full[rank * n:(rank + 1) * n]
    if full.is_cuda:
        dist.reduce_scatter_tensor(mine, full, op=dist.ReduceOp.SUM, group=group)
    else:
        dist.all_reduce(full, op=dist.ReduceOp.SUM, group=group)
    return mine


def allgather_chunks_(full, rank, nprocs, group=None):
    """Every rank contributes its chunk of ``full`` (in place); afterwards all ranks hold the whole tensor."""
    import torch.distributed as dist
    n = full.numel() // nprocs
    dist.all_gather_into_tensor(full, full[rank * n:(rank + 1) * n].clone() if not full.is_cuda else full[rank * n:(rank + 1) * n],
                                group=group)
    return full


class DeviceView:
    """Zero-copy ``__cuda_array_interface__`` view of a context-owned device buffer."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def device_tensor(ptr, n):
    """torch.float64 CUDA tensor aliasing ``n`` doubles at device address ``ptr``."""
    import torch
    return torch.as_tensor(DeviceView(ptr, n), device="cuda")


# ------------------------------------------------------------------------------------------------------------
# Slab decomposition: halo exchanges instead of N^3 collectives
# ------------------------------------------------------------------------------------------------------------
# With sources sharded in list order every rank's rates cover the whole box, so each iteration costs an
# all-reduce of N^3 doubles (1.07 GB at 512^3: ~3.6 ms on 8 B200s, against a 1.3 ms sweep of 10^5/8 sources at
# R = 10.76).  Sharding the sources by position instead -- contiguous ranges of the slowest grid axis with equal
# source counts -- confines a rank's rates to its own planes plus a halo of ceil(R) planes on either side, and the
# ionised fractions it needs to the same range.  Per iteration the ranks then exchange 4 halos of h*N^2 doubles
# with their two neighbours (23 MB each at 512^3, R = 10.76) and all-reduce three scalars.

def slab_edges(src_x0, N, nprocs, R):
    """Plane ranges [edges[r], edges[r+1]) with (nearly) equal source counts, or None when the slabs would be too thin
    (then the list-order sharding with an all-reduce is used): a halo must fit inside the neighbouring slab (width >= h)
    and a slab with both its halos must not meet itself around the box (width + 2h <= N; with two ranks, where both
    neighbours are the same rank, this is the familiar "two halos wide").  ``src_x0``: 0-indexed x of every source.
    Deterministic, so every rank computes the same edges."""
    h = int(np.floor(R)) + 1
    if nprocs < 2 or h * nprocs > N or (N + nprocs - 1) // nprocs + 2 * h > N:
        return None, h
    xs = np.sort(np.mod(np.asarray(src_x0, dtype=np.int64), N))
    edges = [0]
    for r in range(1, nprocs):
        e = int(xs[(r * xs.size) // nprocs]) if xs.size else (r * N) // nprocs
        edges.append(e)
    edges.append(N)
    # enforce the minimum width by pushing edges apart (left to right, then right to left), then the maximum width
    wmin, wmax = h, N - 2 * h
    for r in range(1, nprocs):
        edges[r] = max(edges[r], edges[r - 1] + wmin)
    for r in range(nprocs - 1, 0, -1):
        edges[r] = min(edges[r], edges[r + 1] - wmin)
    for r in range(1, nprocs):
        edges[r] = min(edges[r], edges[r - 1] + wmax)
    for r in range(nprocs - 1, 0, -1):
        edges[r] = max(edges[r], edges[r + 1] - wmax)
    widths = [edges[r + 1] - edges[r] for r in range(nprocs)]
    if min(widths) < wmin or max(widths) > wmax:
        return None, h
    return edges, h


class SlabHalo:
    """Halo exchanges of one rank in a slab-decomposed run; works on flat torch tensors of N^3 doubles (CUDA ->
    NCCL send/recv, CPU -> gloo).

    ``peer=True`` (CUDA ranks of one node, on the context's own PHI_ION / XH_AV buffers): the neighbours' buffers are
    mapped by CUDA IPC once (asora_ipc_export / asora_ipc_open) and the halo planes are read straight over NVLink by one
    kernel per neighbour (asora_peer_halo) instead of four NCCL send/recv pairs plus staging copies per exchange.  The
    cross-rank ordering comes from collectives on the same stream: a one-element all-reduce before the rates are read
    (every rank's sweep has finished), and the caller's own collective between the two exchanges of an iteration (the
    3-scalar all-reduce of the evolve loop: every rank has finished reading before anybody's next sweep overwrites)."""

    def __init__(self, edges, h, N, rank, nprocs, group=None, peer=False, stream_ordered=False):
        self.N, self.h, self.rank, self.nprocs, self.group = N, h, rank, nprocs, group
        # peer mode: the context's stream is torch's current stream (asora_set_stream), so kernels and collectives are
        # ordered by the stream and the host synchronisations between them can be left out
        self.stream_ordered = stream_ordered
        self.lo, self.hi = edges[rank], edges[rank + 1]
        self.left, self.right = (rank - 1) % nprocs, (rank + 1) % nprocs
        self.plane = N * N
        self._tmp = None
        self._peer = None
        if peer:
            self._open_peers()

    # -- peer-memory mode -----------------------------------------------------------------------------------------
    def _open_peers(self):
        import ctypes
        import torch
        import torch.distributed as dist
        from .lib import _cabi
        L, check = _cabi.L, _cabi.check
        bufs = (_cabi.BUF_PHI_ION, _cabi.BUF_XH_AV)
        mine = torch.zeros(len(bufs) * 64, dtype=torch.uint8)
        for i, b in enumerate(bufs):
            raw = ctypes.create_string_buffer(64)
            check(L.asora_ipc_export(b, raw))
            mine[64 * i:64 * (i + 1)] = torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8)
        mine = mine.cuda()
        allh = [torch.empty_like(mine) for _ in range(self.nprocs)]
        dist.all_gather(allh, mine, group=self.group)
        ptrs = {}
        try:
            for nb in {self.left, self.right}:
                hb = allh[nb].cpu().numpy().tobytes()
                ptrs[nb] = {}
                for i, b in enumerate(bufs):
                    p = ctypes.c_void_p()
                    check(L.asora_ipc_open(hb[64 * i:64 * (i + 1)], ctypes.byref(p)))
                    ptrs[nb][b] = p
        except RuntimeError as e:
            import warnings
            warnings.warn(f"SlabHalo: peer access to the neighbours' buffers failed ({e}); using NCCL send/recv")
            for d in ptrs.values():
                for p in d.values():
                    L.asora_ipc_close(p)
            ptrs = None
        # all ranks use the same mode: peer access must have worked everywhere
        ok = torch.tensor([1.0 if ptrs is not None else 0.0], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if ok.item() < 1.0:
            if ptrs is not None:
                for d in ptrs.values():
                    for p in d.values():
                        L.asora_ipc_close(p)
            ptrs = None
        self._peer = ptrs
        self._token = torch.zeros(1, device="cuda")

    @property
    def peer(self):
        return self._peer is not None

    def close(self):
        """Unmap the neighbours' buffers.  Not a collective; the owners free their buffers only at device_close, after the
        collectives that end a time step."""
        if self._peer is not None:
            from .lib import _cabi
            _cabi.L.asora_sync()
            for d in self._peer.values():
                for p in d.values():
                    _cabi.L.asora_ipc_close(p)
            self._peer = None

    def _rank_barrier(self):
        """Every rank has reached this point (and, having synchronised its sweep / chemistry before, finished that work)."""
        import torch
        import torch.distributed as dist
        dist.all_reduce(self._token, group=self.group)
        if not self.stream_ordered:
            torch.cuda.current_stream().synchronize()

    def _peer_done(self):
        from .lib import _cabi
        if not self.stream_ordered:
            _cabi.check(_cabi.L.asora_sync())   # my reads of the neighbours' memory are complete before I enter the next collective

    def _peer_halo(self, buf, nb, first_plane, add):
        from .lib import _cabi
        first_plane %= self.N
        assert first_plane + self.h <= self.N, "halo straddles the periodic boundary"
        _cabi.check(_cabi.L.asora_peer_halo(buf, self._peer[nb][buf], first_plane * self.plane, self.h * self.plane, int(add)))

    # -- geometry -------------------------------------------------------------------------------------------------
    def _planes(self, t, start, count):
        start %= self.N
        assert start + count <= self.N, "halo straddles the periodic boundary"
        return t[start * self.plane:(start + count) * self.plane]

    def active_range(self):
        """(first plane, number of planes) this rank's sweeps may touch."""
        return (self.lo - self.h) % self.N, (self.hi - self.lo) + 2 * self.h

    def own_cells(self):
        return self.lo * self.plane, (self.hi - self.lo) * self.plane

    def _exchange(self, send_left, send_right, recv_right, recv_left):
        import torch.distributed as dist
        ops = [dist.P2POp(dist.isend, send_left, self.left, self.group),
               dist.P2POp(dist.isend, send_right, self.right, self.group),
               dist.P2POp(dist.irecv, recv_right, self.right, self.group),
               dist.P2POp(dist.irecv, recv_left, self.left, self.group)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()

    def reduce_phi_(self, phi):
        """Add to this rank's own planes the rates its neighbours computed for them (in place).  Peer mode: ``phi`` must be
        the context's PHI_ION buffer."""
        import torch
        h, n = self.h, self.h * self.plane
        if self._peer is not None:
            from .lib import _cabi
            if not self.stream_ordered:
                _cabi.check(_cabi.L.asora_sync())                                  # my sweep has finished ...
            self._rank_barrier()                                                   # ... and so has everybody's
            self._peer_halo(_cabi.BUF_PHI_ION, self.right, self.hi - h, True)      # the right neighbour's left halo = my last h planes
            self._peer_halo(_cabi.BUF_PHI_ION, self.left, self.lo, True)           # the left neighbour's right halo = my first h planes
            self._peer_done()
            return phi
        if self._tmp is None or self._tmp.device != phi.device:
            self._tmp = torch.empty(2 * n, dtype=phi.dtype, device=phi.device)
        from_right, from_left = self._tmp[:n], self._tmp[n:]
        self._exchange(self._planes(phi, self.lo - h, h), self._planes(phi, self.hi, h), from_right, from_left)
        self._planes(phi, self.hi - h, h).add_(from_right)   # the right neighbour's left halo = my last h planes
        self._planes(phi, self.lo, h).add_(from_left)        # the left neighbour's right halo = my first h planes
        return phi

    def gather_xh_(self, xh_av, synced=False):
        """Refresh the halo planes of xh_av from the neighbours that own them (in place).  Peer mode: ``xh_av`` must be the
        context's XH_AV buffer; ``synced``: a collective on this stream already followed every rank's chemistry pass."""
        h = self.h
        if self._peer is not None:
            from .lib import _cabi
            import torch
            if synced:
                if not self.stream_ordered:
                    torch.cuda.current_stream().synchronize()   # the caller's collective has completed here
            else:
                if not self.stream_ordered:
                    _cabi.check(_cabi.L.asora_sync())
                self._rank_barrier()
            self._peer_halo(_cabi.BUF_XH_AV, self.right, self.hi, False)
            self._peer_halo(_cabi.BUF_XH_AV, self.left, self.lo - h, False)
            self._peer_done()
            return xh_av
        self._exchange(self._planes(xh_av, self.lo, h).contiguous(), self._planes(xh_av, self.hi - h, h).contiguous(),
                       self._planes(xh_av, self.hi, h), self._planes(xh_av, self.lo - h, h))
        return xh_av

    def assemble_(self, t):
        """Make every rank hold the whole grid: zero what this rank does not own, then sum over ranks."""
        o, c = self.own_cells()
        t[:o].zero_()
        t[o + c:].zero_()
        return allreduce_sum_(t, self.group)
