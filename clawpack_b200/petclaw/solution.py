"""Solution built from the distributed State (pyclaw/src/petclaw/solution.py).

Frame output replaces PetClaw's parallel petsc/hdf5 writers (petclaw/fileio/) by a gather: every rank contributes
its slab (``get_q_global`` -- the replacement of ``petclaw/state.py:267-287``), rank 0 writes ONE set of files in
the serial ascii layout for the global grid, so a file written by an N-GPU run is byte-identical to the serial one."""
import numpy as np

from ..pyclaw import geometry as _geom
from ..pyclaw import state as _state
from ..pyclaw.solution import Solution as _Solution
from . import util


class Solution(_Solution):
    def __init__(self, *arg, **kargs):
        kargs.setdefault('claw_package', 'clawpack_b200.petclaw')
        super(Solution, self).__init__(*arg, **kargs)

    def write(self, frame, path='./', file_format='ascii', file_prefix=None, write_aux=False, options={},
              write_p=False):
        if util.world_size() == 1:
            return super(Solution, self).write(frame, path, file_format, file_prefix, write_aux, options, write_p)
        st = self.state
        q = st.get_q_global()                                  # collective: every rank takes part
        aux = st.get_aux_global() if (st.num_aux > 0 and write_aux) else None
        if write_p and st.p is None:
            raise Exception('write_p is set but state.p has not been computed (set state.mp and a compute_p function)')
        pglob = st._gather(np.asarray(st.p)) if write_p else None      # collective as well
        if util.rank() == 0:
            gdims = [_geom.Dimension(d.lower, d.upper, d.num_cells, name=d.name, units=d.units)
                     for d in st.patch.dimensions]             # the patch keeps the GLOBAL dimensions
            gdom = _geom.Domain(gdims)
            gst = _state.State(gdom, st.num_eqn, st.num_aux)
            gst.q = q
            if aux is not None:
                gst.aux = aux
            if pglob is not None:
                gst.p = pglob
            gst.t = st.t
            gst.problem_data = st.problem_data
            gst.grid.mapc2p = st.grid._mapc2p
            from ..pyclaw import solution as _serial
            g = _serial.Solution(gst, gdom, claw_package='clawpack_b200.pyclaw')
            g.write(frame, path, file_format, file_prefix, write_aux and aux is not None, options, write_p)
        util.barrier()
