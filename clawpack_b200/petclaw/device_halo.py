"""Device-side ghost exchange and CFL reduction for the slab decomposition (NCCL over NVLink).

Per step and rank: two pack kernels (libclawcuda, claw_halo_pack) -> one batched NCCL send/recv with
each slab neighbour -> two unpack kernels; after the sweep one 8-byte ncclAllReduce(max) on the
kernel's CFL slot, then the host reads the value it needs for accept/reject.

The exchange of step n+1 is queued SPECULATIVELY right behind the sweep of step n, on the buffer that sweep wrote
(claw_halo_pack_next / claw_halo_unpack_next), before the host blocks on the CFL read-back: the host-side cost of
queueing pack + NCCL + unpack and the NCCL latency then hide behind the running sweep instead of leaving the GPU idle
between steps.  If the step is rejected the exchanged ghost rows belong to a buffer that is discarded, and the
current buffer still carries the ghost rows of the exchange before -- nothing has to be redone either way.
Measured on 8 B200 (profiles/r01_k_speculative_halo.jsonl): shallow water 16384^2 strong-scaled 1.57 -> 1.46 ms per
step; Euler 8192^2 per GPU unchanged (5.21 ms).  CLAW_SPECULATIVE_HALO=0 switches it off."""
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
        per_lo, per_up = periodic if isinstance(periodic, tuple) else (periodic, periodic)
        self.nb = util.slab_neighbours(nranks, r, per_lo, per_up)
        n = ctypes.c_size_t()
        _lib.check(dev.L.claw_halo_elems(dev.handle, ctypes.byref(n)), dev.handle)
        mk = lambda: torch.empty(n.value, dtype=torch.float64, device=dev.device)   # noqa: E731
        self.send_lo, self.send_up, self.recv_lo, self.recv_up = mk(), mk(), mk(), mk()
        import os
        self.speculative = os.environ.get('CLAW_SPECULATIVE_HALO', '1') != '0'

    def _vp(self, t):
        return ctypes.c_void_p(t.data_ptr())

    def exchange_aux(self, dev):
        """One-off exchange of the static aux ghost rows/planes (petclaw/state.py:233-241)."""
        n = ctypes.c_size_t()
        _lib.check(dev.L.claw_halo_elems_aux(dev.handle, ctypes.byref(n)), dev.handle)
        if n.value == 0:
            return
        torch = self.torch
        bufs = [torch.empty(n.value, dtype=torch.float64, device=dev.device) for _ in range(4)]
        self._exchange(dev, dev.L.claw_halo_pack_aux, dev.L.claw_halo_unpack_aux, *bufs)

    def exchange(self, dev):
        """Pre-step hook: ghost rows of the current buffer, unless the speculative exchange already delivered them."""
        if dev.ghosts_current:
            return
        self._exchange(dev, dev.L.claw_halo_pack, dev.L.claw_halo_unpack, self.send_lo, self.send_up,
                       self.recv_lo, self.recv_up)
        dev.ghosts_current = True

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

    def reduce_cfl(self, dev):
        """max over ranks of the CFL slot written by the sweep kernel, on the device, then one read."""
        self.dist.all_reduce(dev.cfl_tensor, op=self.dist.ReduceOp.MAX)
        if self.speculative:
            self._exchange(dev, dev.L.claw_halo_pack_next, dev.L.claw_halo_unpack_next, self.send_lo, self.send_up,
                           self.recv_lo, self.recv_up)
            dev.ghosts_current = True      # true for the new buffer if accepted, and still true for the old one if not
        return float(dev.cfl_tensor.item())
