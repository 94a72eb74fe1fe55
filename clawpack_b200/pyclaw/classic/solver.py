"""Classic (wave-propagation) solvers ClawSolver2D / ClawSolver3D with ``kernel_language='CUDA'``.

Attribute names, defaults and the step sequence follow pyclaw/src/pyclaw/classic/solver.py
(:71-93 attributes, :96-146 ``step``, :162-193 ``_set_mthlim/_set_method``, :195-234 ``setup``,
:393-549 2-D, :552-706 3-D).  Three kernel back ends:

``kernel_language='CUDA'`` (new, default here)
    device-resident path (cuda_classic.DeviceStepper): q stays in HBM between steps, one ghost
    fill + one fused sweep kernel per step, accept/reject by buffer swap.  Used whenever no
    host callback needs the state every step (no ``before_step``, ``step_source`` or BC.custom).
    Otherwise the same kernels are driven through the f2py-shaped plug-in object
    ``cuda_classic.fmod(num_dim)`` on host arrays, exactly like the Fortran module.
``kernel_language='Fortran'``
    the reference's f2py modules ``clawpack.pyclaw.classic.classic2/3`` if they are importable,
    or any user-supplied ``solver.fmod`` (reference precedent: examples/shallow_sphere/Rossby_wave.py:345).
    There is no built-in CPU implementation in this package.
"""
import numpy as np

from ..limiters import tvd
from ..solver import Solver


class ClawSolver(Solver):
    """Generic classic solver; see ClawSolver2D / ClawSolver3D."""
    # On the device-resident path these live in the device's control block while step attempts are queued
    # (cuda_classic.DeviceStepper.evolve): reading one waits for the queue, writing one hands authority back to the host.
    from ..cuda_classic import DeviceMirrored as _M
    dt = _M('dt')
    dt_max = _M('dt_max')
    dt_variable = _M('dt_variable')
    dt_old = _M('dt_old')
    cfl_max = _M('cfl_max')
    cfl_desired = _M('cfl_desired')
    cfl = _M('cfl')
    status = _M('status')
    step_log = _M('step_log')
    del _M

    def __init__(self, riemann_solver=None, claw_package=None):
        self.num_ghost = 2
        self.limiters = tvd.minmod
        self.order = 2
        self.source_split = 1
        self.fwave = False
        self.step_source = None
        self.kernel_language = 'CUDA'
        self.verbosity = 0
        self.cfl_max = 1.0
        self.cfl_desired = 0.9
        self._mthlim = self.limiters
        self._method = None
        self.dt_old = None
        # CUDA-specific knobs (must exist before the attribute set is frozen, solver.py:139-144)
        self.cuda_precision = 'strict'      # 'strict': no FMA contraction, bit-exact vs the reference arithmetic
        self.cuda_device = None             # CUDA device ordinal (default: current torch device)
        self.cuda_seg_rows = 0              # rows per march segment (0 = auto)
        self._dev = None                    # cuda_classic.DeviceStepper when the device-resident path is active
        self.step_log = []                  # (dt, cfl, accepted) per attempted step
        super(ClawSolver, self).__init__(riemann_solver, claw_package)

    # ---- one time step: classic/solver.py:96-146 --------------------------------------------
    def step(self, solution, take_one_step, tstart, tend):
        self.get_dt(solution.t, tstart, tend, take_one_step)
        self.cfl.set_global_max(0.)
        if self.source_split == 2 and self.step_source is not None:
            self.step_source(self, solution.states[0], self.dt / 2.0)
        self.step_hyperbolic(solution)
        if self.cfl.get_cached_max() >= self.cfl_max:
            return False
        if self.step_source is not None:
            if self.source_split == 2:
                self.step_source(self, solution.states[0], self.dt / 2.0)
            if self.source_split == 1:
                self.step_source(self, solution.states[0], self.dt)
        return True

    def _check_cfl_settings(self):
        pass

    def _allocate_workspace(self, solution):
        pass

    def step_hyperbolic(self, solution):
        raise Exception('Dummy routine, please override!')

    def _set_mthlim(self):
        self._mthlim = self.limiters
        if not isinstance(self.limiters, list):
            self._mthlim = [self._mthlim]
        if len(self._mthlim) == 1:
            self._mthlim = self._mthlim * self.num_waves
        if len(self._mthlim) != self.num_waves:
            raise Exception('Length of solver.limiters is not equal to 1 or to solver.num_waves')

    def _set_method(self, state):
        m = np.empty(7, dtype=np.intc, order='F')
        m[0] = self.dt_variable
        m[1] = self.order
        if self.num_dim == 1:
            m[2] = 0
        elif self.dimensional_split:
            m[2] = -1
        else:
            m[2] = self.transverse_waves
        m[3] = self.verbosity
        m[4] = 0
        m[5] = state.index_capa + 1
        m[6] = state.num_aux
        self._method = m

    def _device_path_eligible(self, dev_is_mine=False):
        from ..solver import BC
        if (self.fmod is not None and not dev_is_mine) or self.before_step is not None:
            return False
        # a source term keeps the run on the device only as a device functor (pyclaw.device_callbacks) with Godunov
        # splitting: the half step of Strang splitting before the sweep would modify the buffer a rejected step keeps
        if self.step_source is not None and not (hasattr(self.step_source, 'device_source') and self.source_split == 1):
            return False
        # BC.custom on q needs the host every step unless the fill is a device functor; BC.custom on aux alone does
        # not: aux is static on the device path, so the user's aux ghost fill is evaluated once at setup
        # (cuda_classic.DeviceStepper)
        for lst, fun in ((self.bc_lower, self.user_bc_lower), (self.bc_upper, self.user_bc_upper)):
            if any(b == BC.custom for b in lst) and not hasattr(fun, 'device_bc'):
                return False
        return True

    def setup(self, solution):
        solution.state.set_num_ghost(self.num_ghost)
        self._check_cfl_settings()
        self._set_mthlim()
        if self.kernel_language == 'CUDA':
            from .. import cuda_classic
            self._set_method(solution.state)
            solution.state.set_cparam(self.rp)
            if self.num_dim > 1 and self._device_path_eligible():
                self._dev = cuda_classic.DeviceStepper(self, solution.state)
            else:
                if self.fmod is None:        # a user-supplied plug-in (solver.fmod = ...) is kept, as in the reference
                    self.fmod = cuda_classic.fmod(self.num_dim, precision=self.cuda_precision)
                self._allocate_workspace(solution)
                self._allocate_bc_arrays(solution.states[0])
        elif self.kernel_language == 'Fortran':
            if self.fmod is None:
                so_name = 'clawpack.pyclaw.classic.classic' + str(self.num_dim)
                try:
                    self.fmod = __import__(so_name, fromlist=['clawpack.pyclaw.classic'])
                except ImportError as e:
                    raise ImportError("kernel_language='Fortran' needs the reference's f2py module %s (not "
                                      "available: %s); use kernel_language='CUDA' or set solver.fmod" % (so_name, e))
            self._set_fortran_parameters(solution)
            self._allocate_workspace(solution)
            self._allocate_bc_arrays(solution.states[0])
        else:
            raise Exception('Only CUDA and Fortran kernels are supported in multi-D.')
        super(ClawSolver, self).setup(solution)

    def _set_fortran_parameters(self, solution):
        self._set_method(solution.state)
        solution.state.set_cparam(self.fmod)
        solution.state.set_cparam(self.rp)

    def teardown(self):
        if self._dev is not None:
            self._dev.close()            # downloads the final state first, then detaches from the State
            self._dev = None

    def evolve_to_time(self, solution, tend=None):
        """solver.py:554-666.  On the device-resident path the loop (get_dt, the step, accept/reject, the retry with a
        smaller dt) runs on the device: attempts are queued without reading the CFL number back, and ``solver.dt``,
        ``solver.cfl``, ``solver.status``, ``solver.step_log`` and ``solution.t`` catch up when they are read."""
        if not self._is_set_up:
            self.setup(solution)
        if self._dev is not None and not self._device_path_eligible(dev_is_mine=True):
            self._leave_device_path(solution)        # a host callback was set after setup(): plug-in mode from now on
        if self._dev is not None:
            return self._dev.evolve(self, solution, tend)
        return super(ClawSolver, self).evolve_to_time(solution, tend)

    def _leave_device_path(self, solution):
        from .. import cuda_classic
        self._dev.close()
        self._dev = None
        self.fmod = cuda_classic.fmod(self.num_dim, precision=self.cuda_precision)
        self._allocate_workspace(solution)
        self._allocate_bc_arrays(solution.states[0])

    # ---- retry hooks: buffer swap instead of q_backup when device-resident -------------------
    def _save_for_retry(self, state):
        if self._dev is not None:
            return None
        return super(ClawSolver, self)._save_for_retry(state)

    def _commit_step(self, state):
        if self._dev is not None:
            self._dev.accept()

    def _retry_from_saved(self, state, saved):
        if self._dev is not None:
            self._dev.reject()
        else:
            super(ClawSolver, self)._retry_from_saved(state, saved)

    def _rp_args(self, names):
        """(rpn, rpt[, rptt]) arguments for fmod: f2py C pointers when the Riemann solver is an
        f2py module (classic/solver.py:516-521), else the solver object itself."""
        out = []
        for nm in names:
            f = getattr(self.rp, nm, None)
            out.append(f._cpointer if f is not None and hasattr(f, '_cpointer') else self.rp)
        return out


class ClawSolver1D(ClawSolver):
    """1-D classic solver (classic/solver.py:249-387, Fortran branch).  ``kernel_language='CUDA'`` always runs in
    plug-in mode: every step calls ``fmod.step1`` (claw_step1_host) on the host arrays -- 1-D problems are small and
    the device-resident machinery would buy nothing.  The reference's pure-Python 1-D kernel is not part of this
    build."""

    def __init__(self, riemann_solver=None, claw_package=None):
        self.num_dim = 1
        self.reflect_index = [1]
        super(ClawSolver1D, self).__init__(riemann_solver, claw_package)

    def step_hyperbolic(self, solution):
        state = solution.states[0]
        grid = state.grid
        self._apply_bcs(state)
        mx = grid.num_cells[0]
        dx, dt = grid.delta[0], self.dt
        rp1, = self._rp_args(('rp1',))
        self.qbc, cfl = self.fmod.step1(self.num_ghost, mx, self.qbc, self.auxbc, dx, dt, self._method, self._mthlim,
                                        self.fwave, rp1)
        self.cfl.update_global_max(cfl)
        state.set_q_from_qbc(self.num_ghost, self.qbc)
        if state.num_aux > 0:
            state.set_aux_from_auxbc(self.num_ghost, self.auxbc)


class ClawSolver2D(ClawSolver):
    """2-D classic solver.  Extra options (classic/solver.py:435-452): ``dimensional_split``
    (default True) and ``transverse_waves`` in {no_trans=0, trans_inc=1, trans_cor=2}."""
    no_trans = 0
    trans_inc = 1
    trans_cor = 2

    def __init__(self, riemann_solver=None, claw_package=None):
        self.dimensional_split = True
        self.transverse_waves = self.trans_inc
        self.num_dim = 2
        self.reflect_index = [1, 2]
        self.aux1 = None
        self.aux2 = None
        self.aux3 = None
        self.work = None
        super(ClawSolver2D, self).__init__(riemann_solver, claw_package)

    def _check_cfl_settings(self):
        recommended = 0.5 if (not self.dimensional_split) and (self.transverse_waves == 0) else 1.0
        if self.cfl_max > recommended:
            import warnings
            warnings.warn('cfl_max is set higher than the recommended value of %s' % recommended)

    def _allocate_workspace(self, solution):
        state = solution.state
        num_aux = state.num_aux if state.aux is not None else 1
        maxm = max(state.grid.num_cells)
        g = self.num_ghost
        self.aux1 = np.empty((num_aux, maxm + 2 * g), order='F')
        self.aux2 = np.empty((num_aux, maxm + 2 * g), order='F')
        self.aux3 = np.empty((num_aux, maxm + 2 * g), order='F')
        mwork = (maxm + 2 * g) * (5 * state.num_eqn + self.num_waves + state.num_eqn * self.num_waves)
        self.work = np.empty((mwork), order='F')

    def step_hyperbolic(self, solution):
        state = solution.states[0]
        if self._dev is not None:
            cfl = self._dev.step(state, self.dt)
            self.cfl.update_global_max(cfl)
            return
        grid = state.grid
        dx, dy = grid.delta
        mx, my = grid.num_cells
        maxm = max(mx, my)
        self._apply_bcs(state)
        qold = self.qbc.copy('F')
        rpn2, rpt2 = self._rp_args(('rpn2', 'rpt2'))
        if self.dimensional_split or self.transverse_waves == 0:
            rpt2 = rpn2
        if self.dimensional_split:
            self.qbc, cfl_x = self.fmod.step2ds(maxm, self.num_ghost, mx, my, qold, self.qbc, self.auxbc, dx, dy,
                                                self.dt, self._method, self._mthlim, self.aux1, self.aux2,
                                                self.aux3, self.work, 1, self.fwave, rpn2, rpt2)
            self.qbc, cfl_y = self.fmod.step2ds(maxm, self.num_ghost, mx, my, self.qbc, self.qbc, self.auxbc, dx, dy,
                                                self.dt, self._method, self._mthlim, self.aux1, self.aux2,
                                                self.aux3, self.work, 2, self.fwave, rpn2, rpt2)
            cfl = max(cfl_x, cfl_y)
        else:
            self.qbc, cfl = self.fmod.step2(maxm, self.num_ghost, mx, my, qold, self.qbc, self.auxbc, dx, dy,
                                            self.dt, self._method, self._mthlim, self.aux1, self.aux2, self.aux3,
                                            self.work, self.fwave, rpn2, rpt2)
        self.cfl.update_global_max(cfl)
        state.set_q_from_qbc(self.num_ghost, self.qbc)
        if state.num_aux > 0:
            state.set_aux_from_auxbc(self.num_ghost, self.auxbc)


class ClawSolver3D(ClawSolver):
    """3-D classic solver (classic/solver.py:552-706).  ``transverse_waves`` is the two-digit code
    no_trans=0, trans_inc=11, trans_cor=22 (default)."""
    no_trans = 0
    trans_inc = 11
    trans_cor = 22

    def __init__(self, riemann_solver=None, claw_package=None):
        self.dimensional_split = True
        self.transverse_waves = self.trans_cor
        self.num_dim = 3
        self.reflect_index = [1, 2, 3]
        self.aux1 = None
        self.aux2 = None
        self.aux3 = None
        self.work = None
        super(ClawSolver3D, self).__init__(riemann_solver, claw_package)

    def _allocate_workspace(self, solution):
        state = solution.states[0]
        num_aux = state.num_aux if state.aux is not None else 1
        maxm = max(state.grid.num_cells)
        g = self.num_ghost
        self.aux1 = np.empty((num_aux, maxm + 2 * g, 3), order='F')
        self.aux2 = np.empty((num_aux, maxm + 2 * g, 3), order='F')
        self.aux3 = np.empty((num_aux, maxm + 2 * g, 3), order='F')
        mwork = (maxm + 2 * g) * (31 * state.num_eqn + self.num_waves + state.num_eqn * self.num_waves)
        self.work = np.empty((mwork), order='F')

    def step_hyperbolic(self, solution):
        state = solution.states[0]
        if self._dev is not None:
            cfl = self._dev.step(state, self.dt)
            self.cfl.update_global_max(cfl)
            return
        grid = state.grid
        dx, dy, dz = grid.delta
        mx, my, mz = grid.num_cells
        maxm = max(mx, my, mz)
        self._apply_bcs(state)
        qnew = self.qbc
        qold = qnew.copy('F')
        rpn3, rpt3, rptt3 = self._rp_args(('rpn3', 'rpt3', 'rptt3'))
        if self.dimensional_split or self.transverse_waves == 0:
            rpt3 = rpn3
            rptt3 = rpn3
        if self.dimensional_split:
            q, cfl_x = self.fmod.step3ds(maxm, self.num_ghost, mx, my, mz, qold, qnew, self.auxbc, dx, dy, dz, self.dt,
                                         self._method, self._mthlim, self.aux1, self.aux2, self.aux3, self.work, 1,
                                         self.fwave, rpn3, rpt3, rptt3)
            q, cfl_y = self.fmod.step3ds(maxm, self.num_ghost, mx, my, mz, q, q, self.auxbc, dx, dy, dz, self.dt,
                                         self._method, self._mthlim, self.aux1, self.aux2, self.aux3, self.work, 2,
                                         self.fwave, rpn3, rpt3, rptt3)
            q, cfl_z = self.fmod.step3ds(maxm, self.num_ghost, mx, my, mz, q, q, self.auxbc, dx, dy, dz, self.dt,
                                         self._method, self._mthlim, self.aux1, self.aux2, self.aux3, self.work, 3,
                                         self.fwave, rpn3, rpt3, rptt3)
            cfl = max(cfl_x, cfl_y, cfl_z)
        else:
            q, cfl = self.fmod.step3(maxm, self.num_ghost, mx, my, mz, qold, qnew, self.auxbc, dx, dy, dz, self.dt,
                                     self._method, self._mthlim, self.aux1, self.aux2, self.aux3, self.work,
                                     self.fwave, rpn3, rpt3, rptt3)
        self.qbc = q
        self.cfl.update_global_max(cfl)
        state.set_q_from_qbc(self.num_ghost, self.qbc)
        if state.num_aux > 0:
            state.set_aux_from_auxbc(self.num_ghost, self.auxbc)
