"""Slab-decomposed Patch / Domain (replaces the DMDA of pyclaw/src/petclaw/geometry.py:10-73).

The local Grid spans this rank's rows of the slowest dimension; ``on_lower_boundary`` /
``on_upper_boundary`` say whether a face is a physical boundary (petclaw/geometry.py:30-37).
Cell centres of the local grid are evaluated with the GLOBAL formula lower + (i+0.5)*delta so that
they are bit-identical to the serial grid (the reference recomputes them per rank and may differ
in the last bit)."""
import numpy as np

from ..pyclaw import geometry as base
from . import util


class _SlabDimension(base.Dimension):
    """Rows [start, stop) of a global Dimension."""

    def __init__(self, gdim, start, stop):
        self._g_lower, self._g_delta, self._start = gdim.lower, gdim.delta, start
        super(_SlabDimension, self).__init__(gdim.lower + start * gdim.delta, gdim.lower + stop * gdim.delta,
                                             stop - start, name=gdim.name, units=gdim.units)
        self.on_lower_boundary = (start == 0)
        self.on_upper_boundary = (stop == gdim.num_cells)

    @property
    def delta(self):
        return self._g_delta

    @property
    def centers(self):
        return self._g_lower + (np.arange(self._start, self._start + self.num_cells) + 0.5) * self._g_delta

    @property
    def nodes(self):
        return self._g_lower + np.arange(self._start, self._start + self.num_cells + 1) * self._g_delta


class Patch(base.Patch):
    def __init__(self, dimensions):
        super(Patch, self).__init__(dimensions)
        dims = self.dimensions
        nranks, r = util.world_size(), util.rank()
        n_last = dims[-1].num_cells
        if n_last < 2 * nranks:
            raise Exception('slab decomposition needs at least 2 cells of the slowest dimension per rank '
                            '(%d cells, %d ranks)' % (n_last, nranks))
        self.slab = util.slab_range(n_last, nranks, r)
        self.num_ranks, self.rank = nranks, r
        local = []
        for d in dims[:-1]:
            c = base.Dimension(d.lower, d.upper, d.num_cells, name=d.name, units=d.units)
            c.on_lower_boundary = True
            c.on_upper_boundary = True
            local.append(c)
        local.append(_SlabDimension(dims[-1], *self.slab))
        self.grid = base.Grid(local)

    def __deepcopy__(self, memo={}):
        import copy
        result = self.__class__(copy.deepcopy(self.dimensions))
        result.grid.mapc2p = self.grid._mapc2p
        return result


class Domain(base.Domain):
    @staticmethod
    def _patch_class():
        return Patch

    def __init__(self, *arg):
        if len(arg) == 1:
            geom = arg[0] if isinstance(arg[0], (list, tuple)) else [arg[0]]
            if isinstance(geom[0], base.Dimension):
                self.patches = [Patch(list(geom))]
                return
        super(Domain, self).__init__(*arg)
