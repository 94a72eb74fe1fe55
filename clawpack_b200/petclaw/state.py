"""Distributed State: local slab of q/aux; ghost rows of internal faces come from the neighbours
(replaces the DMDA-backed State of pyclaw/src/petclaw/state.py:50-59, 186-287)."""
import numpy as np

from ..pyclaw import state as base
from . import util


def exchange_ghost_rows_host(arr, num_ghost, patch, periodic):
    """Fill the ghost rows of the slowest dimension of ``arr`` (components, ..., n_last + 2g), host
    numpy, from the neighbouring ranks' interior rows (full extent of the other dimensions).  The
    blocking exchange is ordered by rank parity so that it cannot deadlock."""
    d = util._dist()
    nranks, r = patch.num_ranks, patch.rank
    if d is None or nranks == 1:
        return
    import torch
    g = num_ghost
    lower_nb = r - 1 if r > 0 else (nranks - 1 if periodic else None)
    upper_nb = r + 1 if r < nranks - 1 else (0 if periodic else None)
    send_lo = torch.from_numpy(np.ascontiguousarray(arr[..., g:2 * g]))
    send_up = torch.from_numpy(np.ascontiguousarray(arr[..., -2 * g:-g]))
    recv_lo = torch.empty_like(send_lo)
    recv_up = torch.empty_like(send_up)
    ops = []
    if upper_nb is not None:
        ops += [d.P2POp(d.isend, send_up, upper_nb), d.P2POp(d.irecv, recv_up, upper_nb)]
    if lower_nb is not None:
        ops += [d.P2POp(d.isend, send_lo, lower_nb), d.P2POp(d.irecv, recv_lo, lower_nb)]
    if nranks == 2 and periodic:
        # both neighbours are the same rank: tag the two messages by direction
        ops = [d.P2POp(d.isend, send_up, upper_nb, tag=1), d.P2POp(d.irecv, recv_up, upper_nb, tag=2),
               d.P2POp(d.isend, send_lo, lower_nb, tag=2), d.P2POp(d.irecv, recv_lo, lower_nb, tag=1)]
    for req in d.batch_isend_irecv(ops):
        req.wait()
    if lower_nb is not None:
        arr[..., :g] = recv_lo.numpy()
    if upper_nb is not None:
        arr[..., -g:] = recv_up.numpy()


class State(base.State):
    """q / aux are the LOCAL slab; ``get_q_global`` gathers the whole array on every rank."""

    def get_qbc_from_q(self, num_ghost, qbc):
        qbc = super(State, self).get_qbc_from_q(num_ghost, qbc)
        exchange_ghost_rows_host(qbc, num_ghost, self.patch, getattr(self, '_periodic_last', False))
        return qbc

    def get_auxbc_from_aux(self, num_ghost, auxbc):
        auxbc = super(State, self).get_auxbc_from_aux(num_ghost, auxbc)
        exchange_ghost_rows_host(auxbc, num_ghost, self.patch, getattr(self, '_periodic_last', False))
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
