"""State: q / aux on one patch plus problem_data, t, index_capa
(pyclaw/src/pyclaw/state.py:119-152, 192-289, 316-333).

q and aux are Fortran-ordered ``(num_eqn, mx, my[, mz])`` numpy arrays as in the reference.
When a CUDA solver owns the state (kernel_language='CUDA', device-resident path) the
authoritative copy lives in HBM; reading ``state.q`` downloads it on demand and any host access
marks the host copy as possibly modified, so that it is uploaded again before the next step."""
import numpy as np

from . import geometry


class State(object):
    _grid_attrs = ('num_dim', 'p_centers', 'p_edges', 'c_centers', 'c_edges', 'num_cells', 'lower', 'upper',
                   'delta', 'centers', 'edges', 'gauges')

    def __init__(self, geom, num_eqn, num_aux=0):
        if isinstance(geom, geometry.Patch):
            self.patch = geom
        elif isinstance(geom, geometry.Domain):
            self.patch = geom.patches[0]
        else:
            raise Exception('A PyClaw State object must be initialized with a PyClaw Patch object.')
        self.p = None
        self.F = None
        self.problem_data = {}
        self._device = None          # set by a CUDA solver (cuda_classic.DeviceStepper): owner of the device copy
        self._t = 0.
        self.index_capa = -1
        self.keep_gauges = False
        self.gauge_data = []
        self._q = self.new_array(num_eqn)
        self._aux = self.new_array(num_aux)

    def __getattr__(self, key):
        if key in State._grid_attrs:
            return getattr(self.grid, key)
        raise AttributeError("'State' object has no attribute '" + key + "'")

    # ---- t: advanced on the device while step attempts are queued (cuda_classic.DeviceStepper.evolve) ----------
    @property
    def t(self):
        dev = self.__dict__.get('_device')
        if dev is not None and dev.pending:
            dev.drain()
        return self._t

    @t.setter
    def t(self, value):
        dev = self.__dict__.get('_device')
        if dev is not None:
            if dev.pending:
                dev.drain()
            dev.host_ctl_dirty = True
        self._t = value

    # ---- aux: static on the device path; assigning a new array re-uploads it before the next step ------------
    @property
    def aux(self):
        return self._aux

    @aux.setter
    def aux(self, value):
        self._aux = value
        dev = self.__dict__.get('_device')
        if dev is not None:
            dev.aux_dirty = True

    # ---- q with lazy device synchronisation ------------------------------------------------
    @property
    def q(self):
        dev = self.__dict__.get('_device')
        if dev is not None:
            dev.sync_host(self)
        return self._q

    @q.setter
    def q(self, value):
        self._q = value
        dev = self.__dict__.get('_device')
        if dev is not None:
            dev.host_modified()

    @property
    def num_eqn(self):
        if self._q is None:
            raise Exception('state.num_eqn has not been set.')
        return self._q.shape[0]

    @property
    def num_aux(self):
        return self.aux.shape[0] if self.aux is not None else 0

    @property
    def grid(self):
        return self.patch.grid

    @property
    def mp(self):
        return self.p.shape[0] if self.p is not None else 0

    @mp.setter
    def mp(self, mp):
        if self.p is not None:
            raise Exception('Cannot change state.mp after aux is initialized.')
        self.p = self.new_array(mp)

    @property
    def mF(self):
        return self.F.shape[0] if self.F is not None else 0

    @mF.setter
    def mF(self, mF):
        if self.F is not None:
            raise Exception('Cannot change state.mF after aux is initialized.')
        self.F = self.new_array(mF)

    def new_array(self, dof):
        if dof == 0:
            return None
        return np.empty([dof] + list(self.grid.num_cells), order='F')

    def is_valid(self):
        ok = self._q.flags['F_CONTIGUOUS']
        if self.aux is not None:
            ok = ok and self.aux.flags['F_CONTIGUOUS']
        return bool(ok)

    def set_cparam(self, fortran_module):
        """problem_data -> the solver's cparam members (state.py:192-217); missing names raise."""
        if hasattr(fortran_module, 'cparam'):
            names = [p for p in fortran_module.cparam.__dir__() if '__' not in p]
            if not set(names) <= set(self.problem_data.keys()):
                raise Exception('Some required value(s) in the cparam common block in the Riemann solver '
                                'have not been set in problem_data.')
            for k, v in self.problem_data.items():
                if k in names:
                    setattr(fortran_module.cparam, k, v)

    def set_num_ghost(self, num_ghost):
        pass

    def _interior(self, num_ghost):
        g = num_ghost
        return (slice(None),) + (slice(g, -g),) * self.patch.num_dim

    def set_q_from_qbc(self, num_ghost, qbc):
        self.q = qbc[self._interior(num_ghost)]

    def set_aux_from_auxbc(self, num_ghost, auxbc):
        self.aux = auxbc[self._interior(num_ghost)]

    def get_qbc_from_q(self, num_ghost, qbc):
        qbc[self._interior(num_ghost)] = self.q
        return qbc

    def get_auxbc_from_aux(self, num_ghost, auxbc):
        auxbc[self._interior(num_ghost)] = self.aux
        return auxbc

    def get_q_global(self):
        return self.q.copy()

    def get_aux_global(self):
        return self.aux.copy()

    def sum_F(self, i):
        return np.sum(np.abs(self.F[i, ...]))

    def __deepcopy__(self, memo={}):
        import copy
        result = self.__class__(copy.deepcopy(self.patch), self.num_eqn, self.num_aux)
        result.t = copy.deepcopy(self.t)
        result._q = np.array(self.q, order='F', copy=True)
        if self.aux is not None:
            result.aux = np.array(self.aux, order='F', copy=True)
        result.problem_data = copy.deepcopy(self.problem_data)
        result.index_capa = self.index_capa
        return result

    def __str__(self):
        s = 'PyClaw State object\nPatch dimensions: %s\nTime  t=%s\n' % (self.patch.num_cells_global, self.t)
        s += 'Number of conserved quantities: %s\n' % self.num_eqn
        if self.aux is not None:
            s += 'Number of auxiliary fields: %s\n' % self.num_aux
        if self.problem_data:
            s += 'problem_data: ' + str(self.problem_data)
        return s
