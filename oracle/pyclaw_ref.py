"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.

numpy restatement of PyClaw's host-side logic around the classic step, driving the C++
restatement of the Fortran kernels in ``libclawref.so`` (oracle/clawref2.hpp, clawref3.hpp):

* ghost-cell fill           pyclaw/src/pyclaw/solver.py:351-519   (`_apply_bcs`, `_bc_lower`, `_bc_upper`)
* one hyperbolic step       pyclaw/src/pyclaw/classic/solver.py:498-549 (2-D), :648-706 (3-D)
* dt control / accept-reject pyclaw/src/pyclaw/solver.py:524-666, classic/solver.py:96-146
* output-time loop          pyclaw/src/pyclaw/controller.py:336-359

Pinned against the reference's stored golden vectors in tests/test_oracle.py (tests/golden/).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

RP_IDS = {
    'advection_2D': 1,
    'acoustics_2D': 2,
    'euler_4wave_2D': 3,
    'shallow_roe_with_efix_2D': 4,
    'vc_acoustics_3D': 5,
    'euler_3D': 6,
    'vc_acoustics_2D': 7,
    'vc_advection_2D': 8,
    'euler_5wave_2D': 9,
    'psystem_2D': 10,
    'euler_hlle_2D': 11,
    'shallow_hlle_2D': 12,
    'acoustics_mapped_2D': 13,
    'shallow_bathymetry_fwave_2D': 14,
    'burgers_3D': 15,
    'burgers_2D': 16,
    'vc_advection_3D': 17,
    'kpp_2D': 18,
    'vc_elasticity_2D': 19,
    'vc_advection_vector_2D': 20,
    'advection_1D': 21,
    'acoustics_1D': 22,
    'burgers_1D': 23,
    'shallow_roe_with_efix_1D': 24,
    'shallow_hlle_1D': 25,
    'traffic_1D': 26,
    'advection_color_1D': 27,
    'acoustics_variable_1D': 28,
    'shallow_bathymetry_fwave_1D': 29,
    'euler_with_efix_1D': 30,
    'euler_hlle_1D': 31,
    'nonlinear_elasticity_fwave_1D': 32,
    'cubic_1D': 33,
    'traffic_vc_1D': 34,
    'traffic_vc_fwave_1D': 35,
    'traffic_vc_tracer_1D': 36,
}
# COMMON /cparam/ member order of each solver (rpn2_advection.f90, rpn2_acoustics.f90:52,
# rpn2_euler_4wave.f90:71, rpn2_shallow_roe_with_efix.f90:59; vc_acoustics_3D has none)
CPARAM_NAMES = {
    'advection_2D': ('u', 'v'),
    'acoustics_2D': ('rho', 'bulk', 'cc', 'zz'),
    'euler_4wave_2D': ('gamma',),
    'shallow_roe_with_efix_2D': ('grav',),
    'vc_acoustics_3D': (),
    'euler_3D': ('gamma',),           # rpn3_euler.f90: COMMON /cparam/ gamma
    'vc_acoustics_2D': (),            # rpn2_vc_acoustics.f90: coefficients live in aux
    'vc_advection_2D': (),            # rpn2_vc_advection.f90: edge velocities live in aux
    'euler_5wave_2D': ('gamma',),     # rpn2_euler_5wave.f90:64
    'psystem_2D': (),                 # rpn2_psystem.f90: coefficients live in aux
    'euler_hlle_2D': ('gamma',),      # rpn2_euler_hlle.f90:32
    'shallow_hlle_2D': ('grav',),     # rpn2_shallow_hlle.f90:31
    'acoustics_mapped_2D': (),        # rpn2_acoustics_mapped.f90: everything lives in aux
    'shallow_bathymetry_fwave_2D': ('grav', 'dry_tolerance', 'sea_level'),   # rpn2_shallow_bathymetry_fwave.f90:44-45
    'burgers_2D': (),
    'kpp_2D': (),                     # rpn2_kpp.f90: no parameters
    'vc_elasticity_2D': (),           # rpn2_vc_elasticity.f90: material in aux
    'vc_advection_vector_2D': (),     # examples/advection_reaction_2d/rpn2_vc_advection_vector.f90
    'vc_advection_3D': (),
    'burgers_3D': ('coeff',),         # rpn3_burgers.f90:45: common /cparam/ coeff(3)
    'advection_1D': ('u',),           # rp1_advection.f90
    'acoustics_1D': ('rho', 'bulk', 'cc', 'zz'),   # rp1_acoustics.f90
    'burgers_1D': (),
    'shallow_roe_with_efix_1D': ('grav',),
    'shallow_hlle_1D': ('grav',),
    'traffic_1D': ('umax',),   # rp1_traffic.f90:39-40
    'advection_color_1D': (),   # rp1_advection_color.f90: velocity in aux
    'acoustics_variable_1D': (),   # rp1_acoustics_variable.f90: Z, c in aux
    'shallow_bathymetry_fwave_1D': ('grav', 'dry_tolerance', 'sea_level'),   # rp1_shallow_bathymetry_fwave.f90:42-43
    'euler_with_efix_1D': ('gamma',),   # rp1_euler_with_efix.f90:66
    'euler_hlle_1D': ('gamma',),   # rp1_euler_hlle.f90:47
    'nonlinear_elasticity_fwave_1D': (),   # rp1_nonlinear_elasticity_fwave.f90: rho, coefA, coefB in aux
    'cubic_1D': (),   # rp1_cubic.f90
    'traffic_vc_1D': (), 'traffic_vc_fwave_1D': (), 'traffic_vc_tracer_1D': (),   # rp1_traffic_vc*.f90: speed limit in aux
}
NUM_EQN = {'advection_2D': 1, 'acoustics_2D': 3, 'euler_4wave_2D': 4, 'shallow_roe_with_efix_2D': 3,
           'vc_acoustics_3D': 4, 'euler_3D': 5, 'vc_acoustics_2D': 3, 'vc_advection_2D': 1, 'euler_5wave_2D': 5,
           'psystem_2D': 3, 'euler_hlle_2D': 4, 'shallow_hlle_2D': 3,
           'acoustics_mapped_2D': 3, 'shallow_bathymetry_fwave_2D': 3, 'burgers_3D': 1, 'burgers_2D': 1, 'vc_advection_3D': 1, 'advection_1D': 1, 'acoustics_1D': 2, 'burgers_1D': 1, 'shallow_roe_with_efix_1D': 2, 'shallow_hlle_1D': 2}
NUM_WAVES = {'advection_2D': 1, 'acoustics_2D': 2, 'euler_4wave_2D': 4, 'shallow_roe_with_efix_2D': 3,
             'vc_acoustics_3D': 2, 'euler_3D': 3, 'vc_acoustics_2D': 2, 'vc_advection_2D': 1, 'euler_5wave_2D': 5,
             'psystem_2D': 2, 'euler_hlle_2D': 2, 'shallow_hlle_2D': 2,
             'acoustics_mapped_2D': 2, 'shallow_bathymetry_fwave_2D': 3, 'burgers_3D': 1, 'burgers_2D': 1, 'vc_advection_3D': 1, 'advection_1D': 1, 'acoustics_1D': 2, 'burgers_1D': 1, 'shallow_roe_with_efix_1D': 2, 'shallow_hlle_1D': 2}

NUM_EQN.update({'vc_advection_vector_2D': 2, 'traffic_vc_1D': 1, 'traffic_vc_fwave_1D': 1, 'traffic_vc_tracer_1D': 2, 'kpp_2D': 1, 'vc_elasticity_2D': 5, 'traffic_1D': 1, 'advection_color_1D': 1, 'acoustics_variable_1D': 2, 'shallow_bathymetry_fwave_1D': 2, 'euler_with_efix_1D': 3, 'euler_hlle_1D': 3, 'nonlinear_elasticity_fwave_1D': 2, 'cubic_1D': 1})
NUM_WAVES.update({'vc_advection_vector_2D': 1, 'traffic_vc_1D': 1, 'traffic_vc_fwave_1D': 1, 'traffic_vc_tracer_1D': 2, 'kpp_2D': 1, 'vc_elasticity_2D': 4, 'traffic_1D': 1, 'advection_color_1D': 1, 'acoustics_variable_1D': 2, 'shallow_bathymetry_fwave_1D': 2, 'euler_with_efix_1D': 3, 'euler_hlle_1D': 2, 'nonlinear_elasticity_fwave_1D': 2, 'cubic_1D': 1})

BC_CUSTOM, BC_EXTRAP, BC_PERIODIC, BC_WALL = 0, 1, 2, 3

_lib = None


def build():
    """Compile libclawref.so with the committed recipe (oracle/Makefile)."""
    import subprocess
    subprocess.run(['make', '-C', _HERE], check=True, stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, 'libclawref.so')
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int)
        i, d = ctypes.c_int, ctypes.c_double
        L.clawref_step2.argtypes = [i] * 7 + [dp, dp, dp, d, d, d, ip, ip, dp, i, i, dp, i]
        L.clawref_step2ds.argtypes = [i] * 7 + [dp, dp, dp, d, d, d, ip, ip, dp, i, i, i, dp, i]
        L.clawref_count_ops2.argtypes = [i] * 7 + [dp, dp, d, d, d, ip, ip, i, i, i, dp,
                                                   ctypes.POINTER(ctypes.c_uint64)]
        L.clawref_step1.argtypes = [i] * 5 + [dp, dp, d, d, ip, ip, dp, i, i, dp]
        if hasattr(L, 'clawref_step3'):
            L.clawref_step3.argtypes = [i] * 8 + [dp, dp, dp, d, d, d, d, ip, ip, dp, i, i, dp, i]
            L.clawref_step3ds.argtypes = [i] * 8 + [dp, dp, dp, d, d, d, d, ip, ip, dp, i, i, i, dp, i]
            L.clawref_count_ops3.argtypes = [i] * 8 + [dp, dp, d, d, d, d, ip, ip, i, i, i, dp,
                                                       ctypes.POINTER(ctypes.c_uint64)]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def max_threads():
    return int(lib().clawref_max_threads())


# ----------------------------------------------------------------------------------------
#  Ghost-cell fill -- solver.py:351-519
# ----------------------------------------------------------------------------------------
def fill_ghosts(arr, bc_lower, bc_upper, num_ghost, reflect_index=None, is_q=True,
                on_lower=None, on_upper=None, custom_lower=None, custom_upper=None):
    """arr: (n, mx+2g, my+2g[, mz+2g]) F-ordered with the interior already copied in.
    For each dimension ascending: lower then upper (solver.py:377-437); strips span the full
    extent of the other dimensions including their ghost cells."""
    g = num_ghost
    ndim = arr.ndim - 1
    for idim in range(ndim):
        a = np.moveaxis(arr, idim + 1, 1)   # view, like np.rollaxis(array, idim+1, 1)  (solver.py:406)
        lo_on = True if on_lower is None else on_lower[idim]
        up_on = True if on_upper is None else on_upper[idim]
        if lo_on:
            t = bc_lower[idim]
            if t == BC_EXTRAP:                                  # solver.py:459-461
                for i in range(g):
                    a[:, i, ...] = a[:, g, ...]
            elif t == BC_PERIODIC:
                if up_on:                                       # :400-402 (skipped unless rank owns both ends)
                    a[:, :g, ...] = a[:, -2 * g:-g, ...]          # :463
            elif t == BC_WALL:                                  # :464-473
                for i in range(g):
                    a[:, i, ...] = a[:, 2 * g - 1 - i, ...]
                    if is_q:
                        r = reflect_index[idim]
                        a[r, i, ...] = -a[r, 2 * g - 1 - i, ...]
            elif t == BC_CUSTOM and custom_lower is not None:   # :393-398 (un-rolled array handed over)
                custom_lower(idim, arr, g)
            else:
                raise NotImplementedError('oracle: lower BC %r' % (t,))
        if up_on:
            t = bc_upper[idim]
            if t == BC_EXTRAP:                                  # :500-502
                for i in range(g):
                    a[:, -i - 1, ...] = a[:, -g - 1, ...]
            elif t == BC_PERIODIC:
                if lo_on:
                    a[:, -g:, ...] = a[:, g:2 * g, ...]           # :504
            elif t == BC_WALL:                                  # :505-514
                for i in range(g):
                    a[:, -i - 1, ...] = a[:, -2 * g + i, ...]
                    if is_q:
                        r = reflect_index[idim]
                        a[r, -i - 1, ...] = -a[r, -2 * g + i, ...]
            elif t == BC_CUSTOM and custom_upper is not None:
                custom_upper(idim, arr, g)
            else:
                raise NotImplementedError('oracle: upper BC %r' % (t,))
    return arr


class RefSolver(object):
    """Oracle twin of ClawSolver2D/3D (kernel_language='Fortran').  Attribute names follow
    classic/solver.py:71-93, 435-452, 597-615."""

    def __init__(self, rp_name, num_dim, problem_data, num_aux=0):
        self.rp_name = rp_name
        self.rp_id = RP_IDS[rp_name]
        self.num_dim = num_dim
        self.num_eqn = NUM_EQN[rp_name]
        self.num_waves = NUM_WAVES[rp_name]
        self.num_aux = num_aux
        self.num_ghost = 2
        self.limiters = 1                   # tvd.minmod  (classic/solver.py:79)
        self.order = 2
        self.fwave = False
        self.cfl_max = 1.0
        self.cfl_desired = 0.9
        self.dimensional_split = True
        self.transverse_waves = {1: 0, 2: 1, 3: 22}[num_dim]  # trans_inc (2-D :446), trans_cor (3-D :605)
        self.dt_initial = 0.1
        self.dt_max = 1e99
        self.max_steps = 10000
        self.dt_variable = True
        self.index_capa = -1
        self.bc_lower = [None] * num_dim
        self.bc_upper = [None] * num_dim
        self.aux_bc_lower = [None] * num_dim
        self.aux_bc_upper = [None] * num_dim
        self.reflect_index = list(range(1, num_dim + 1))
        self.user_bc_lower = None           # f(idim, qbc, num_ghost): BC.custom fill of q (solver.py:393-398)
        self.user_bc_upper = None
        self.step_source = None             # f(q, aux, dt): in-place source step on the interior q (classic/solver.py:126-144)
        self.source_split = 1
        self.before_step = None             # f(q, aux): solver.py:609-610
        self.gauges = []                    # interior cell index tuples (geometry.py:392-405)
        self.compute_gauge_values = lambda q, aux: q      # solver.py:22-26
        self.gauge_data = []                # per gauge: rows (t, p...) (solver.py:680-708)
        self.user_aux_bc_lower = None       # f(idim, auxbc, num_ghost): BC.custom fill of aux
        self.user_aux_bc_upper = None
        self.nthreads = 1
        self.cparam = np.array([float(v) for k in CPARAM_NAMES[rp_name] for v in np.atleast_1d(problem_data[k])] + [0.0],
                               dtype=np.float64)
        self.dt = self.dt_initial
        self.dt_old = None
        self.cfl = self.cfl_desired          # Solver.__init__: CFL(self.cfl_desired)  (solver.py:199)
        self.status = {'cflmax': -np.inf, 'dtmin': np.inf, 'dtmax': -np.inf, 'numsteps': 0}
        self.step_log = []                   # (dt, cfl, accepted) per attempted step
        self.qbc = None
        self.auxbc = None
        self.t = 0.0

    # -- setup: classic/solver.py:195-219, 162-193 -----------------------------------------
    def setup(self, q, aux, delta):
        self.delta = [float(x) for x in delta]
        self.q = np.asfortranarray(q, dtype=np.float64).copy(order='F')
        self.aux = None if aux is None else np.asfortranarray(aux, dtype=np.float64).copy(order='F')
        lim = self.limiters
        if not isinstance(lim, list):
            lim = [lim]
        if len(lim) == 1:
            lim = lim * self.num_waves
        if len(lim) != self.num_waves:
            raise Exception('Length of solver.limiters is not equal to 1 or to solver.num_waves')
        self._mthlim = np.array(lim, dtype=np.intc)
        m = np.zeros(7, dtype=np.intc)
        m[0] = int(self.dt_variable)
        m[1] = self.order
        m[2] = -1 if self.dimensional_split else self.transverse_waves
        if self.num_dim == 1:
            m[2] = 0                                                         # classic/solver.py:181-182
        m[3] = 0
        m[4] = 0
        m[5] = self.index_capa + 1
        m[6] = self.num_aux
        self._method = m
        g = self.num_ghost
        shp = [n + 2 * g for n in self.q.shape[1:]]
        self.qbc = np.zeros([self.num_eqn] + shp, order='F')                # solver.py:341-343
        if self.num_aux > 0:
            self.auxbc = np.zeros([self.num_aux] + shp, order='F')          # np.empty in the reference (:347)
        else:
            self.auxbc = np.zeros([1] + shp, order='F')
        self.apply_bcs()                                                     # solver.py:349
        self.dt = self.dt_initial

    def _interior(self, a):
        g = self.num_ghost
        return a[(slice(None),) + (slice(g, -g),) * self.num_dim]

    def apply_bcs(self):
        g = self.num_ghost
        self._interior(self.qbc)[...] = self.q                              # state.py:257-272
        if self.num_aux > 0:
            self._interior(self.auxbc)[...] = self.aux
            fill_ghosts(self.auxbc, self.aux_bc_lower, self.aux_bc_upper, g, is_q=False,
                        custom_lower=self.user_aux_bc_lower, custom_upper=self.user_aux_bc_upper)
        fill_ghosts(self.qbc, self.bc_lower, self.bc_upper, g, self.reflect_index, is_q=True,
                    custom_lower=self.user_bc_lower, custom_upper=self.user_bc_upper)

    # -- step_hyperbolic: classic/solver.py:498-549 / 648-706 --------------------------------
    def step_hyperbolic(self):
        L = lib()
        g = self.num_ghost
        self.apply_bcs()
        qold = self.qbc.copy('F')                                            # :514
        cfl = ctypes.c_double(0.0)
        n = self.q.shape[1:]
        maxm = max(n)
        common = (maxm, self.num_eqn, self.num_waves, self.num_aux, g) + tuple(n)
        tail = (int(self.fwave), self.rp_id, _dp(self.cparam), self.nthreads)
        if self.num_dim == 1:                                                # classic/solver.py:279-303: in place on qbc
            rc = L.clawref_step1(self.num_eqn, self.num_waves, g, self.num_aux, n[0], _dp(self.qbc), _dp(self.auxbc),
                                 self.delta[0], self.dt, _ip(self._method), _ip(self._mthlim), ctypes.byref(cfl),
                                 int(self.fwave), self.rp_id, _dp(self.cparam))
            assert rc == 0, rc
            c = cfl.value
        elif self.num_dim == 2:
            dx, dy = self.delta
            if self.dimensional_split:
                rc = L.clawref_step2ds(*common, _dp(qold), _dp(self.qbc), _dp(self.auxbc), dx, dy, self.dt,
                                       _ip(self._method), _ip(self._mthlim), ctypes.byref(cfl), 1, *tail)
                assert rc == 0, rc
                cfl_x = cfl.value
                rc = L.clawref_step2ds(*common, _dp(self.qbc), _dp(self.qbc), _dp(self.auxbc), dx, dy, self.dt,
                                       _ip(self._method), _ip(self._mthlim), ctypes.byref(cfl), 2, *tail)
                assert rc == 0, rc
                c = max(cfl_x, cfl.value)                                    # :535
            else:
                rc = L.clawref_step2(*common, _dp(qold), _dp(self.qbc), _dp(self.auxbc), dx, dy, self.dt,
                                     _ip(self._method), _ip(self._mthlim), ctypes.byref(cfl), *tail)
                assert rc == 0, rc
                c = cfl.value
        else:
            dx, dy, dz = self.delta
            if self.dimensional_split:
                cs = []
                src = qold
                for idir in (1, 2, 3):                                       # :680-690
                    rc = L.clawref_step3ds(*common, _dp(src), _dp(self.qbc), _dp(self.auxbc), dx, dy, dz, self.dt,
                                           _ip(self._method), _ip(self._mthlim), ctypes.byref(cfl), idir, *tail)
                    assert rc == 0, rc
                    cs.append(cfl.value)
                    src = self.qbc
                c = max(cs)
            else:
                rc = L.clawref_step3(*common, _dp(qold), _dp(self.qbc), _dp(self.auxbc), dx, dy, dz, self.dt,
                                     _ip(self._method), _ip(self._mthlim), ctypes.byref(cfl), *tail)
                assert rc == 0, rc
                c = cfl.value
        self.cfl = max(self.cfl, c)                                          # cfl.update_global_max (:543)
        self.q = self._interior(self.qbc)                                    # a VIEW (state.py:225-239)

    # -- dt selection: solver.py:531-552 -----------------------------------------------------
    def get_dt(self, t, tstart, tend, take_one_step):
        cfl = self.cfl
        if self.dt_variable and self.dt_old is not None:
            if cfl > 0.0:
                self.dt = min(self.dt_max, self.dt * self.cfl_desired / cfl)
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

    # -- evolve_to_time: solver.py:554-666 + ClawSolver.step (classic/solver.py:96-146) ------
    def evolve_to_time(self, tend=None):
        take_one_step = tend is None
        tstart = self.t
        num_steps = 0
        max_steps = self.max_steps
        if not self.dt_variable:
            if take_one_step:
                max_steps = 1
            else:
                max_steps = int((tend - tstart + 1e-10) / self.dt)
                if abs(max_steps * self.dt - (tend - tstart)) > 1e-5 * (tend - tstart):
                    raise Exception('dt does not divide (tend-tstart) and dt is fixed!')
        if self.dt_variable == 1 and self.cfl_desired > self.cfl_max:
            raise Exception('Variable time-stepping and desired CFL > maximum CFL')
        if not take_one_step and tend <= tstart:
            max_steps = 0
        for n in range(max_steps):
            if self.dt_variable:
                q_backup = self.q.copy('F')
                told = self.t
            if self.before_step is not None:                                 # solver.py:609-610
                self.before_step(self.q, self.aux)
            self.get_dt(self.t, tstart, tend, take_one_step)                 # classic/solver.py:122
            self.cfl = 0.0                                                   # cfl.set_global_max(0.) (:124)
            if self.source_split == 2 and self.step_source is not None:      # :126-127
                self.step_source(self.q, self.aux, self.dt / 2.0)
            self.step_hyperbolic()
            cfl = self.cfl
            if not (cfl >= self.cfl_max) and self.step_source is not None:   # :134-144
                if self.source_split == 2:
                    self.step_source(self.q, self.aux, self.dt / 2.0)
                if self.source_split == 1:
                    self.step_source(self.q, self.aux, self.dt)
            accept = not (cfl > self.cfl_max)                                # solver.py:524-529
            self.step_log.append((self.dt, cfl, accept))
            if accept:
                self.status['cflmax'] = max(cfl, self.status['cflmax'])
                if self.dt_variable:
                    self.t += self.dt
                else:
                    self.t = tstart + (n + 1) * self.dt
                self.write_gauge_values()                                    # solver.py:636
                num_steps += 1
                self.status['numsteps'] += 1
            else:
                if self.dt_variable:
                    self.q = q_backup
                    self.t = told
                else:
                    self.status['cflmax'] = max(cfl, self.status['cflmax'])
                    raise Exception('CFL too large, giving up!')
            if take_one_step:
                break
            elif self.t >= tend:
                break
        if not take_one_step:
            if self.dt_variable and self.t < tend and num_steps == max_steps:
                raise Exception('Maximum number of timesteps have been taken')
        return self.status

    # -- gauges: solver.py:680-708 (keep_gauges branch) ------------------------------------------
    def write_gauge_values(self):
        for i, g in enumerate(self.gauges):
            ix = (slice(None),) + tuple(g)
            p = self.compute_gauge_values(self.q[ix], self.aux[ix] if self.aux is not None else None)
            if not hasattr(p, '__iter__'):
                p = [p]
            row = np.append(self.t, p)
            if len(self.gauge_data) == len(self.gauges):
                self.gauge_data[i] = np.vstack((self.gauge_data[i], row))
            else:
                self.gauge_data.append(row)

    # -- Controller.run, output_style 1: controller.py:336-359 ---------------------------------
    def run(self, tfinal, num_output_times=10, keep_frames=False):
        self.write_gauge_values()                                            # controller.py:308-309
        frames = [self.q.copy('F')] if keep_frames else None
        for t in np.linspace(self.t, tfinal, num_output_times + 1)[1:]:
            self.evolve_to_time(t)
            if keep_frames:
                frames.append(np.array(self.q, order='F', copy=True))
        return frames


def count_ops(solver, dt):
    """Exact operation counts of one hyperbolic step from the solver's current state."""
    L = lib()
    solver.apply_bcs()
    g = solver.num_ghost
    n = solver.q.shape[1:]
    counts = (ctypes.c_uint64 * 12)()
    common = (max(n), solver.num_eqn, solver.num_waves, solver.num_aux, g) + tuple(n)
    if solver.num_dim == 2:
        rc = L.clawref_count_ops2(*common, _dp(solver.qbc), _dp(solver.auxbc), solver.delta[0], solver.delta[1], dt,
                                  _ip(solver._method), _ip(solver._mthlim), int(solver.dimensional_split),
                                  int(solver.fwave), solver.rp_id, _dp(solver.cparam), counts)
    else:
        rc = L.clawref_count_ops3(*common, _dp(solver.qbc), _dp(solver.auxbc), solver.delta[0], solver.delta[1],
                                  solver.delta[2], dt, _ip(solver._method), _ip(solver._mthlim),
                                  int(solver.dimensional_split), int(solver.fwave), solver.rp_id,
                                  _dp(solver.cparam), counts)
    assert rc == 0, rc
    names = ('add', 'mul', 'div', 'sqrt', 'cmp', 'abs')
    ncell = int(np.prod(n))
    out = {k: counts[i] / ncell for i, k in enumerate(names)}
    out['repeated'] = {k: counts[6 + i] / ncell for i, k in enumerate(names)}   # redone verbatim by the reference
    return out
