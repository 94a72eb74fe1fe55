"""Device-side ghost exchange and CFL reduction for the slab decomposition (NCCL over NVLink).

Per step and rank: two pack kernels (libclawcuda, claw_halo_pack) -> one batched NCCL send/recv with
each slab neighbour -> two unpack kernels; after the sweep one 8-byte ncclAllReduce(max) on the
kernel's CFL slot, then the host reads the value it needs for accept/reject."""
import ctypes

from .. import _lib
from . import util


class DeviceHalo(object):
    def __init__(self, dev, patch, periodic):
        import torch
        self.torch = torch
        self.dev = dev
        self.dist = util._dist()
        nranks, r = patch.num_ranks, patch.rank
        self.lower_nb = r - 1 if r > 0 else (nranks - 1 if periodic else None)
        self.upper_nb = r + 1 if r < nranks - 1 else (0 if periodic else None)
        self.same_nb = (nranks == 2 and periodic)
        n = ctypes.c_size_t()
        _lib.check(dev.L.claw_halo_elems(dev.handle, ctypes.byref(n)), dev.handle)
        mk = lambda: torch.empty(n.value, dtype=torch.float64, device=dev.device)   # noqa: E731
        self.send_lo, self.send_up, self.recv_lo, self.recv_up = mk(), mk(), mk(), mk()

    def _vp(self, t):
        return ctypes.c_void_p(t.data_ptr())

    def exchange(self, dev):
        L, h, d = dev.L, dev.handle, self.dist
        ops = []
        if self.upper_nb is not None:
            _lib.check(L.claw_halo_pack(h, 1, self._vp(self.send_up)), h)
        if self.lower_nb is not None:
            _lib.check(L.claw_halo_pack(h, 0, self._vp(self.send_lo)), h)
        if self.same_nb:
            ops = [d.P2POp(d.isend, self.send_up, self.upper_nb), d.P2POp(d.irecv, self.recv_lo, self.lower_nb),
                   d.P2POp(d.isend, self.send_lo, self.lower_nb), d.P2POp(d.irecv, self.recv_up, self.upper_nb)]
        else:
            if self.upper_nb is not None:
                ops += [d.P2POp(d.isend, self.send_up, self.upper_nb), d.P2POp(d.irecv, self.recv_up, self.upper_nb)]
            if self.lower_nb is not None:
                ops += [d.P2POp(d.isend, self.send_lo, self.lower_nb), d.P2POp(d.irecv, self.recv_lo, self.lower_nb)]
        for req in d.batch_isend_irecv(ops):
            req.wait()
        if self.lower_nb is not None:
            _lib.check(L.claw_halo_unpack(h, 0, self._vp(self.recv_lo)), h)
        if self.upper_nb is not None:
            _lib.check(L.claw_halo_unpack(h, 1, self._vp(self.recv_up)), h)

    def reduce_cfl(self, dev):
        """max over ranks of the CFL slot written by the sweep kernel, on the device, then one read."""
        self.dist.all_reduce(dev.cfl_tensor, op=self.dist.ReduceOp.MAX)
        return float(dev.cfl_tensor.item())
