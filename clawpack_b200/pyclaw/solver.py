"""Base solver: boundary conditions, time-step control and the accept/reject loop.

Behavioural twin of pyclaw/src/pyclaw/solver.py (BC enum :13-19, attribute freeze :139-144,
``_apply_bcs`` :351-437, ``_bc_lower/_bc_upper`` :439-519, ``get_dt`` :531-552,
``evolve_to_time`` :554-666).  The one structural change: keeping / restoring the pre-step state
goes through three hooks (``_save_for_retry``, ``_commit_step``, ``_retry_from_saved``) so that a
device-resident solver can replace the reference's full-array ``q_backup`` copy with a buffer
swap (see classic/solver.py and cuda_classic.py).
"""
import logging

import numpy as np


class CFLError(Exception):
    """Error raised when cfl_max is exceeded."""


class BC():
    """Boundary-condition ids."""
    custom = 0
    extrap = 1
    periodic = 2
    wall = 3


def default_compute_gauge_values(q, aux):
    return q


class Solver(object):
    """Abstract solver.  ``Solver(riemann_solver)`` looks up num_eqn / num_waves / fwave from
    ``riemann.static`` by the solver object's ``__name__`` tail (solver.py:230-238)."""

    def __setattr__(self, key, value):
        # attribute set is frozen once __init__ has finished (solver.py:139-144)
        if not hasattr(self, '_isinitialized'):
            self.__dict__['_isinitialized'] = False
        if self._isinitialized and not hasattr(self, key):
            raise TypeError('%s has no attribute %s' % (self.__class__, key))
        object.__setattr__(self, key, value)

    @property
    def all_bcs(self):
        return self.bc_lower, self.bc_upper

    @all_bcs.setter
    def all_bcs(self, all_bcs):
        for i in range(self.num_dim):
            self.bc_lower[i] = all_bcs
            self.bc_upper[i] = all_bcs

    def __init__(self, riemann_solver=None, claw_package=None):
        self.logger = logging.getLogger('pyclaw.solver')
        self.dt_initial = 0.1
        self.dt_max = 1e99
        self.max_steps = 10000
        self.dt_variable = True
        self.num_waves = None
        self.qbc = None
        self.auxbc = None
        self.rp = None
        self.fmod = None
        self._is_set_up = False
        self._use_old_bc_sig = False
        self.accept_step = True
        self.before_step = None

        import sys
        if claw_package is not None and claw_package in sys.modules:
            self.claw_package = sys.modules[claw_package]
        else:
            # e.g. 'clawpack_b200.pyclaw.classic.solver' -> 'clawpack_b200.pyclaw'
            name = self.__module__.rpartition('.')[0].rpartition('.')[0]
            if name not in sys.modules:
                raise NotImplementedError('Unable to determine solver package, please provide one')
            self.claw_package = sys.modules[name]

        self.dt = self.dt_initial
        self.cfl = self.claw_package.CFL(self.cfl_desired)
        self.status = {'cflmax': -np.inf, 'dtmin': np.inf, 'dtmax': -np.inf, 'numsteps': 0}

        self.bc_lower = [None] * self.num_dim
        self.bc_upper = [None] * self.num_dim
        self.aux_bc_lower = [None] * self.num_dim
        self.aux_bc_upper = [None] * self.num_dim
        self.user_bc_lower = None
        self.user_bc_upper = None
        self.user_aux_bc_lower = None
        self.user_aux_bc_upper = None

        self.num_eqn = None
        self.num_waves = None
        self.fwave = None
        self.compute_gauge_values = default_compute_gauge_values

        if riemann_solver is not None:
            self.rp = riemann_solver
            rp_name = riemann_solver.__name__.split('.')[-1]
            if 'ptwise' in rp_name:
                rp_name = rp_name.replace('_ptwise', '')
            from .. import riemann
            self.num_eqn = riemann.static.num_eqn.get(rp_name, None)
            self.num_waves = riemann.static.num_waves.get(rp_name, None)
            self.fwave = riemann.static.fwave.get(rp_name, None)

        self._isinitialized = True

    # ---- validation / setup ---------------------------------------------------------------
    def is_valid(self):
        reason = None
        if any(b == BC.custom for b in self.bc_lower) and self.user_bc_lower is None:
            reason = 'Lower custom BC function has not been set.'
        if any(b == BC.custom for b in self.bc_upper) and self.user_bc_upper is None:
            reason = 'Upper custom BC function has not been set.'
        if self.num_waves is None:
            reason = 'solver.num_waves has not been set.'
        if self.num_eqn is None:
            reason = 'solver.num_eqn has not been set.'
        if self.fwave is None:
            reason = 'solver.fwave has not been set.'
        if (None in self.bc_lower) or (None in self.bc_upper):
            reason = 'One of the boundary conditions has not been set.'
        if reason is not None:
            self.logger.debug(reason)
        return reason is None, reason

    def setup(self, solution):
        self._is_set_up = True

    def __del__(self):
        self.__dict__['_is_set_up'] = False

    def __str__(self):
        out = 'Solver Status:\n'
        for k, v in self.status.items():
            out = '\n'.join((out, '%s = %s' % (k.rjust(25), v)))
        return out

    # ---- boundary conditions (host arrays) ---------------------------------------------------
    def _allocate_bc_arrays(self, state):
        import inspect
        for fun in (self.user_bc_lower, self.user_bc_upper, self.user_aux_bc_lower, self.user_aux_bc_upper):
            if fun is not None and len(inspect.getfullargspec(fun)[0]) == 5:
                self.logger.warning('The custom boundary condition function signature (state,dim,t,qbc,num_ghost) '
                                    'is deprecated; use (state,dim,t,qbc,auxbc,num_ghost).')
                self._use_old_bc_sig = True
        shape = [n + 2 * self.num_ghost for n in state.grid.num_cells]
        self.qbc = np.zeros([state.num_eqn] + shape, order='F')
        self.auxbc = np.empty([state.num_aux] + shape, order='F')
        self._apply_bcs(state)

    def _apply_bcs(self, state):
        """Copy q (aux) into qbc (auxbc) and fill ghost cells: per dimension ascending, lower face
        then upper face, aux before q (solver.py:377-437)."""
        self.qbc = state.get_qbc_from_q(self.num_ghost, self.qbc)
        if state.num_aux > 0:
            self.auxbc = state.get_auxbc_from_aux(self.num_ghost, self.auxbc)
        grid = state.grid
        for idim, dim in enumerate(grid.dimensions):
            for upper in (False, True):
                on_this = grid.on_upper_boundary[idim] if upper else grid.on_lower_boundary[idim]
                on_other = grid.on_lower_boundary[idim] if upper else grid.on_upper_boundary[idim]
                if not on_this:
                    continue
                jobs = []
                if state.num_aux > 0:
                    jobs.append((self.auxbc, self.aux_bc_upper if upper else self.aux_bc_lower,
                                 self.user_aux_bc_upper if upper else self.user_aux_bc_lower, 'aux'))
                jobs.append((self.qbc, self.bc_upper if upper else self.bc_lower,
                             self.user_bc_upper if upper else self.user_bc_lower, 'q'))
                for array, types, custom_fun, variable in jobs:
                    if types[idim] == BC.custom:
                        if not self._use_old_bc_sig:
                            custom_fun(state, dim, state.t, self.qbc, self.auxbc, self.num_ghost)
                        else:
                            custom_fun(state, dim, state.t, array, self.num_ghost)
                    elif types[idim] == BC.periodic and not on_other:
                        pass   # periodic across ranks: the ghost exchange already did it
                    else:
                        view = np.rollaxis(array, idim + 1, 1)
                        if upper:
                            self._bc_upper(types[idim], state, dim, state.t, view, idim, variable)
                        else:
                            self._bc_lower(types[idim], state, dim, state.t, view, idim, variable)

    def _bc_lower(self, bc_type, state, dim, t, array, idim, name):
        g = self.num_ghost
        if bc_type == BC.extrap:
            for i in range(g):
                array[:, i, ...] = array[:, g, ...]
        elif bc_type == BC.periodic:
            array[:, :g, ...] = array[:, -2 * g:-g, ...]
        elif bc_type == BC.wall:
            for i in range(g):
                array[:, i, ...] = array[:, 2 * g - 1 - i, ...]
                if name == 'q':
                    r = self.reflect_index[idim]
                    array[r, i, ...] = -array[r, 2 * g - 1 - i, ...]
        elif bc_type is None:
            raise Exception('Lower boundary condition not specified for either q or aux.')
        else:
            raise NotImplementedError('Boundary condition %s not implemented' % bc_type)

    def _bc_upper(self, bc_type, state, dim, t, array, idim, name):
        g = self.num_ghost
        if bc_type == BC.extrap:
            for i in range(g):
                array[:, -i - 1, ...] = array[:, -g - 1, ...]
        elif bc_type == BC.periodic:
            array[:, -g:, ...] = array[:, g:2 * g, ...]
        elif bc_type == BC.wall:
            for i in range(g):
                array[:, -i - 1, ...] = array[:, -2 * g + i, ...]
                if name == 'q':
                    r = self.reflect_index[idim]
                    array[r, -i - 1, ...] = -array[r, -2 * g + i, ...]
        elif bc_type is None:
            raise Exception('Upper boundary condition not specified for either q or aux.')
        else:
            raise NotImplementedError('Boundary condition %s not implemented' % bc_type)

    # ---- time stepping --------------------------------------------------------------------------
    def accept_reject_step(self, state):
        return not (self.cfl.get_cached_max() > self.cfl_max)

    def get_dt_new(self):
        cfl = self.cfl.get_cached_max()
        self.dt = min(self.dt_max, self.dt * self.cfl_desired / cfl)

    def get_dt(self, t, tstart, tend, take_one_step):
        cfl = self.cfl.get_cached_max()
        if self.dt_variable and self.dt_old is not None:
            if cfl > 0.0:
                self.get_dt_new()
                self.status['dtmin'] = min(self.dt, self.status['dtmin'])
                self.status['dtmax'] = max(self.dt, self.status['dtmax'])
            else:
                self.dt = self.dt_max
        else:
            self.dt_old = self.dt
        if not take_one_step:
            if t + self.dt > tend and tstart < tend:
                self.dt = tend - t
            if tend - t - self.dt < 1.e-14 * t:
                self.dt = tend - t

    # hooks around one attempted step (host default = the reference's q_backup copy, solver.py:611,645)
    def _save_for_retry(self, state):
        return state.q.copy('F')

    def _commit_step(self, state):
        pass

    def _retry_from_saved(self, state, saved):
        state.q = saved

    def evolve_to_time(self, solution, tend=None):
        """Advance ``solution`` to ``tend`` (one step if None); returns ``self.status``."""
        if not self._is_set_up:
            self.setup(solution)
        take_one_step = tend is None
        tstart = solution.t
        num_steps = 0

        if not self.dt_variable:
            if take_one_step:
                self.max_steps = 1
            else:
                self.max_steps = int((tend - tstart + 1e-10) / self.dt)
                if abs(self.max_steps * self.dt - (tend - tstart)) > 1e-5 * (tend - tstart):
                    raise Exception('dt does not divide (tend-tstart) and dt is fixed!')
        if self.dt_variable == 1 and self.cfl_desired > self.cfl_max:
            raise Exception('Variable time-stepping and desired CFL > maximum CFL')
        if not take_one_step and tend <= tstart:
            self.logger.info('Already at or beyond end time: no evolution required.')
            self.max_steps = 0

        for n in range(self.max_steps):
            state = solution.state
            if self.dt_variable:
                saved = self._save_for_retry(state)
                told = solution.t
            if self.before_step is not None:
                self.before_step(self, solution.states[0])

            self.step(solution, take_one_step, tstart, tend)

            cfl = self.cfl.get_cached_max()
            self.accept_step = self.accept_reject_step(state)
            self.step_log.append((self.dt, cfl, self.accept_step))
            if self.accept_step:
                self._commit_step(state)
                self.status['cflmax'] = max(cfl, self.status['cflmax'])
                if self.dt_variable:
                    solution.t += self.dt
                else:
                    solution.t = tstart + (n + 1) * self.dt
                self.logger.debug('Step %i  CFL = %f   dt = %f   t = %f' % (n, cfl, self.dt, solution.t))
                self.write_gauge_values(solution)
                num_steps += 1
                self.status['numsteps'] += 1
            else:
                self.logger.debug('Rejecting time step, CFL number too large')
                if self.dt_variable:
                    self._retry_from_saved(state, saved)
                    solution.t = told
                else:
                    self.status['cflmax'] = max(cfl, self.status['cflmax'])
                    raise Exception('CFL too large, giving up!')
            if take_one_step or solution.t >= tend:
                break

        if not take_one_step:
            if self.dt_variable and solution.t < tend and num_steps == self.max_steps:
                raise Exception('Maximum number of timesteps have been taken')
        return self.status

    def step(self, solution):
        raise NotImplementedError('No stepping routine has been defined!')

    def write_gauge_values(self, solution):
        """Record (t, compute_gauge_values(q, aux)) at every gauge cell (solver.py:680-715): appended to
        ``state.gauge_data`` when ``state.keep_gauges`` and written to the gauge files when they are open.  On the
        device-resident path only the gauge cells are read back (claw_get_cells), not the whole array."""
        state = solution.state
        gauges = state.grid.gauges
        if len(gauges) == 0:
            return
        dev = getattr(self, '_dev', None)
        qcells = dev.get_cells(state, gauges) if (dev is not None and not dev.host_valid) else None
        for i, gauge in enumerate(gauges):
            ix = tuple(gauge)
            aux = state.aux[(slice(None),) + ix] if state.num_aux > 0 else None
            q = qcells[i] if qcells is not None else state.q[(slice(None),) + ix]
            p = self.compute_gauge_values(q, aux)
            if not hasattr(p, '__iter__'):
                p = [p]
            t = solution.t
            if state.keep_gauges:
                gauge_data = state.gauge_data
                if len(gauge_data) == len(gauges):
                    gauge_data[i] = np.vstack((gauge_data[i], np.append(t, p)))
                else:
                    gauge_data.append(np.append(t, p))
            if len(state.grid.gauge_files) > i:
                state.grid.gauge_files[i].write(str(t) + ' ' + ' '.join(str(j) for j in p) + '\n')
