"""Device-side ghost exchange and CFL reduction for the slab decomposition (NVLink peer memory, or NCCL).

The step loop itself runs on the device (cuda_classic.DeviceStepper.evolve); what a multi-rank attempt adds is the ghost
cells at rank faces and the max of the CFL number over ranks, both on the solver's stream without a host round trip, and
identically ordered on every rank (no rank-local shortcut decides whether to exchange).  With the NCCL transports an
attempt is: two pack kernels (libclawcuda, claw_halo_pack) -> one grouped NCCL send/recv with each slab neighbour -> two
unpack kernels -> ghost fill + sweep -> one 8-byte ncclAllReduce(max) on the CFL slot of the control block -> the
control kernel, which decides accept/reject identically on every rank; pack / unpack find the current q buffer through
the control block, so the exchange of attempt n+1 is queued before the outcome of attempt n is known, and after a
rejected attempt the same rows are simply exchanged again.

Three transports (``CLAW_HALO`` or ``DeviceHalo(..., transport=...)``):

* ``peer`` (default): no exchange call at all.  The neighbours' buffers are mapped through CUDA IPC (``claw_peer_*``,
  include/clawcuda.h); the attempt itself writes the rows / planes next to a rank face into the neighbour's ghost cells
  over NVLink after the sweep and all-reduces the CFL number through per-rank mailboxes inside the control kernel, so
  ``claw_ctl_attempts`` (and its CUDA graph) queue N-rank attempts like 1-rank ones.  ``state.q`` must then be written on
  all ranks alike (a host upload is followed by a collective refresh of the neighbours' ghost cells).  Falls back to
  ``cabi`` -- on all ranks together -- where IPC mapping is not available.
* ``torch``: the process group of ``torch.distributed`` (pack / unpack kernels of libclawcuda around
  ``batch_isend_irecv``; ``all_reduce`` on the CFL slot).
* ``cabi``: the NCCL exchange behind the C ABI -- ``claw_halo_exchange`` / ``claw_allreduce_cfl`` on a communicator of the
  library's own (``claw_nccl_comm_init``; the 128-byte id travels through the torch process group once) -- which is what
  a non-Python host (an MPI / C driver) would call; one ctypes call per exchange instead of a dozen Python-level ops."""
import ctypes
import os

from .. import _lib
from . import util


_cabi_comm = {}          # (device index) -> ncclComm_t of libclawcuda, shared by all DeviceHalo objects of the process


def _cabi_communicator(dev, nranks, rank):
    """The library-side NCCL communicator of this process (created once: rank 0 draws the id, the torch process group
    broadcasts it)."""
    key = dev.device.index
    if key in _cabi_comm:
        return _cabi_comm[key]
    import torch
    dist = util._dist()
    L = dev.L
    ident = (ctypes.c_ubyte * 128)()
    if rank == 0:
        _lib.check(L.claw_nccl_unique_id(ctypes.cast(ident, ctypes.c_void_p)))
    t = torch.tensor(list(ident), dtype=torch.uint8, device=dev.device if dist.get_backend() == 'nccl' else 'cpu')
    dist.broadcast(t, src=0)
    ident = (ctypes.c_ubyte * 128)(*[int(v) for v in t.cpu().tolist()])
    comm = ctypes.c_void_p()
    _lib.check(L.claw_nccl_comm_init(ctypes.byref(comm), int(nranks), int(rank), ctypes.cast(ident, ctypes.c_void_p),
                                     int(dev.device.index)))
    _cabi_comm[key] = comm
    return comm


class DeviceHalo(object):
    def __init__(self, dev, patch, periodic, transport=None):
        import torch
        self.torch = torch
        self.dev = dev
        self.dist = util._dist()
        nranks, r = patch.num_ranks, patch.rank
        per_lo, per_up = periodic if isinstance(periodic, tuple) else (periodic, periodic)
        self.nb = util.slab_neighbours(nranks, r, per_lo, per_up)
        self.transport = transport or os.environ.get('CLAW_HALO', 'peer')
        if self.transport not in ('torch', 'cabi', 'peer'):
            raise ValueError("halo transport must be 'torch', 'cabi' or 'peer'")
        if self.transport == 'peer' and not self._peer_connect(dev, nranks, r):
            self.transport = 'cabi'                 # (decided by all ranks together) IPC mapping is not available here
        self.comm = _cabi_communicator(dev, nranks, r) if self.transport == 'cabi' else None
        self.nbi = [-1 if self.nb[k] is None else int(self.nb[k])
                    for k in ('send_up_to', 'recv_lo_from', 'send_lo_to', 'recv_up_from')]
        if self.comm is not None or self.transport == 'peer':
            return
        n = ctypes.c_size_t()
        _lib.check(dev.L.claw_halo_elems(dev.handle, ctypes.byref(n)), dev.handle)
        mk = lambda: torch.empty(n.value, dtype=torch.float64, device=dev.device)   # noqa: E731
        self.send_lo, self.send_up, self.recv_lo, self.recv_up = mk(), mk(), mk(), mk()

    def _peer_connect(self, dev, nranks, rank):
        """Peer-memory transport (include/clawcuda.h, claw_peer_*): every rank exports the IPC handles of its buffers and
        mailbox, the blocks are all-gathered through the process group, the neighbours' buffers are mapped.  Collective;
        returns False on every rank if any rank could not map its peers."""
        torch, dist = self.torch, self.dist
        on_gpu = dist.get_backend() == 'nccl'
        n = 320                                                # CLAW_PEER_INFO_BYTES
        info = (ctypes.c_ubyte * n)()
        ok = 0
        if getattr(dev, 'own_buffers', False):
            ok = 1 if dev.L.claw_peer_export(dev.handle, ctypes.cast(info, ctypes.c_void_p)) == 0 else 0
        mine = torch.tensor(list(info), dtype=torch.uint8, device=dev.device if on_gpu else 'cpu')
        parts = [torch.zeros_like(mine) for _ in range(nranks)]
        dist.all_gather(parts, mine)
        if ok:
            blob = bytes(torch.cat(parts).cpu().tolist())
            buf = (ctypes.c_ubyte * len(blob)).from_buffer_copy(blob)
            up = -1 if self.nb['send_up_to'] is None else int(self.nb['send_up_to'])
            lo = -1 if self.nb['send_lo_to'] is None else int(self.nb['send_lo_to'])
            ok = 1 if dev.L.claw_peer_connect(dev.handle, int(nranks), int(rank), ctypes.cast(buf, ctypes.c_void_p),
                                              up, lo) == 0 else 0
        flag = torch.tensor([ok], dtype=torch.int32, device=dev.device if on_gpu else 'cpu')
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        return bool(int(flag.item()))

    def _vp(self, t):
        return ctypes.c_void_p(t.data_ptr())

    def exchange_aux(self, dev):
        """One-off exchange of the static aux ghost rows/planes (petclaw/state.py:233-241)."""
        n = ctypes.c_size_t()
        _lib.check(dev.L.claw_halo_elems_aux(dev.handle, ctypes.byref(n)), dev.handle)
        if n.value == 0:
            return
        if self.transport == 'peer':
            _lib.check(dev.L.claw_peer_sync(dev.handle, 1), dev.handle)
            return
        if self.comm is not None:
            _lib.check(dev.L.claw_halo_exchange(dev.handle, self.comm, *self.nbi, 1), dev.handle)
            return
        torch = self.torch
        bufs = [torch.empty(n.value, dtype=torch.float64, device=dev.device) for _ in range(4)]
        self._exchange(dev, dev.L.claw_halo_pack_aux, dev.L.claw_halo_unpack_aux, *bufs)

    def exchange(self, dev):
        """Ghost rows / planes of the current q buffer from the slab neighbours (queued, not waited for)."""
        if self.transport == 'peer':              # refresh by hand + barrier (initial data, host uploads); collective
            _lib.check(dev.L.claw_peer_sync(dev.handle, 0), dev.handle)
            return
        if self.comm is not None:
            _lib.check(dev.L.claw_halo_exchange(dev.handle, self.comm, *self.nbi, 0), dev.handle)
            return
        self._exchange(dev, dev.L.claw_halo_pack, dev.L.claw_halo_unpack, self.send_lo, self.send_up,
                       self.recv_lo, self.recv_up)

    def _exchange(self, dev, pack, unpack, send_lo, send_up, recv_lo, recv_up):
        L, h, d, nb = dev.L, dev.handle, self.dist, self.nb
        if nb['send_up_to'] is not None:
            _lib.check(pack(h, 1, self._vp(send_up)), h)
        if nb['send_lo_to'] is not None:
            _lib.check(pack(h, 0, self._vp(send_lo)), h)
        ops = []
        if nb['send_up_to'] is not None:
            ops.append(d.P2POp(d.isend, send_up, nb['send_up_to']))
        if nb['recv_lo_from'] is not None:
            ops.append(d.P2POp(d.irecv, recv_lo, nb['recv_lo_from']))
        if nb['send_lo_to'] is not None:
            ops.append(d.P2POp(d.isend, send_lo, nb['send_lo_to']))
        if nb['recv_up_from'] is not None:
            ops.append(d.P2POp(d.irecv, recv_up, nb['recv_up_from']))
        if ops:
            for req in d.batch_isend_irecv(ops):
                req.wait()
        if nb['recv_lo_from'] is not None:
            _lib.check(unpack(h, 0, self._vp(recv_lo)), h)
        if nb['recv_up_from'] is not None:
            _lib.check(unpack(h, 1, self._vp(recv_up)), h)

    def all_reduce_cfl(self, dev):
        """max over ranks of the CFL slot written by the sweep kernels, in place on the device (petclaw/cfl.py:36)."""
        if self.transport == 'peer':
            cfl = ctypes.c_double(0.0)
            _lib.check(dev.L.claw_peer_allreduce_cfl(dev.handle, ctypes.byref(cfl)), dev.handle)
            return cfl.value
        if self.comm is not None:
            _lib.check(dev.L.claw_allreduce_cfl(dev.handle, self.comm), dev.handle)
            return
        self.dist.all_reduce(dev.cfl_tensor, op=self.dist.ReduceOp.MAX)
