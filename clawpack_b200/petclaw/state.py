"""Distributed State: local slab of q/aux; ghost rows of internal faces come from the neighbours
(replaces the DMDA-backed State of pyclaw/src/petclaw/state.py:50-59, 186-287)."""
import numpy as np

from ..pyclaw import state as base
from . import util


def exchange_ghost_rows_host(arr, num_ghost, patch, periodic):
    """Fill the ghost rows of the slowest dimension of ``arr`` (components, ..., n_last + 2g), host
    numpy, from the neighbouring ranks' interior rows (full extent of the other dimensions).
    ``periodic`` = (lower face periodic, upper face periodic)."""
    d = util._dist()
    nranks, r = patch.num_ranks, patch.rank
    if d is None or nranks == 1:
        return
    import torch
    g = num_ghost
    per_lo, per_up = periodic if isinstance(periodic, tuple) else (periodic, periodic)
    nb = util.slab_neighbours(nranks, r, per_lo, per_up)
    send_lo = torch.from_numpy(np.ascontiguousarray(arr[..., g:2 * g]))
    send_up = torch.from_numpy(np.ascontiguousarray(arr[..., -2 * g:-g]))
    recv_lo = torch.empty_like(send_lo)
    recv_up = torch.empty_like(send_up)
    ops = []
    if nb['send_up_to'] is not None:
        ops.append(d.P2POp(d.isend, send_up, nb['send_up_to'], tag=1))
    if nb['recv_lo_from'] is not None:
        ops.append(d.P2POp(d.irecv, recv_lo, nb['recv_lo_from'], tag=1))
    if nb['send_lo_to'] is not None:
        ops.append(d.P2POp(d.isend, send_lo, nb['send_lo_to'], tag=2))
    if nb['recv_up_from'] is not None:
        ops.append(d.P2POp(d.irecv, recv_up, nb['recv_up_from'], tag=2))
    if ops:
        for req in d.batch_isend_irecv(ops):
            req.wait()
    if nb['recv_lo_from'] is not None:
        arr[..., :g] = recv_lo.numpy()
    if nb['recv_up_from'] is not None:
        arr[..., -g:] = recv_up.numpy()


class State(base.State):
    """q / aux are the LOCAL slab; ``get_q_global`` gathers the whole array on every rank."""

    def get_qbc_from_q(self, num_ghost, qbc):
        qbc = super(State, self).get_qbc_from_q(num_ghost, qbc)
        exchange_ghost_rows_host(qbc, num_ghost, self.patch, getattr(self, '_periodic_last', (False, False)))
        return qbc

    def get_auxbc_from_aux(self, num_ghost, auxbc):
        auxbc = super(State, self).get_auxbc_from_aux(num_ghost, auxbc)
        exchange_ghost_rows_host(auxbc, num_ghost, self.patch, getattr(self, '_periodic_last_aux', (False, False)))
        return auxbc

    def _gather(self, local):
        d = util._dist()
        if d is None or self.patch.num_ranks == 1:
            return local.copy()
        import torch
        nranks = self.patch.num_ranks
        n_last = self.patch.dimensions[-1].num_cells
        use_cuda = d.get_backend() == 'nccl'
        parts = []
        for r in range(nranks):
            s0, s1 = util.slab_range(n_last, nranks, r)
            shape = list(local.shape[:-1]) + [s1 - s0]
            if r == self.patch.rank:
                t = torch.from_numpy(np.ascontiguousarray(local))
            else:
                t = torch.empty(shape, dtype=torch.float64)
            if use_cuda:
                t = t.cuda()
            d.broadcast(t, src=r)
            parts.append(t.cpu().numpy())
        return np.asfortranarray(np.concatenate(parts, axis=-1))

    def get_q_global(self):
        return self._gather(np.asarray(self.q))

    def get_aux_global(self):
        return self._gather(np.asarray(self.aux))
