"""User-supplied Riemann solvers (clawpack_b200/csrc/rp_plugin.cu, riemann.compile_cuda_solver): the B200 form of the
reference's `solver.rp = my_f2py_module` (pyclaw/src/pyclaw/classic/solver.py:516-521).  Two functors written the way
a user would port a Fortran solver (tests/plugins/*.cuh) are compiled into plug-ins with nvcc (works without a GPU),
registered with libclawcuda, and -- on a GPU -- run through the PyClaw API and the f2py-shaped host calls against the
oracle's solvers of the same equations: bit-exact in strict arithmetic, 1e-12 in fast."""
import ctypes
import os

import numpy as np
import pytest

from clawpack_b200 import _lib, riemann, scenarios as S
from refutil import ROOT, ref_solver

PLUG = os.path.join(ROOT, 'tests', 'plugins')


def _acoustics():
    return riemann.compile_cuda_solver(os.path.join(PLUG, 'user_acoustics.cuh'), 'UserAcoustics2D',
                                       cparam=('rho', 'bulk', 'cc', 'zz'), name='user_acoustics_2D')


def _vc_acoustics():
    return riemann.compile_cuda_solver(os.path.join(PLUG, 'user_vc_acoustics.cuh'), 'UserVcAcoustics2D',
                                       name='user_vc_acoustics_2D')


def _acoustics3d():
    return riemann.compile_cuda_solver(os.path.join(PLUG, 'user_acoustics3d.cuh'), 'UserAcoustics3D',
                                       name='user_acoustics_3D', num_dim=3)


def test_3d_plugin_builds_without_a_gpu_and_registers():
    rp = _acoustics3d()
    assert rp.rp_id >= 1000 and (rp.num_dim, rp.num_eqn, rp.num_waves, rp.num_aux_read) == (3, 4, 2, 2)
    P = ctypes.CDLL(rp.plugin_path)
    assert hasattr(P, 'claw_rp_plugin_launch3') and not hasattr(P, 'claw_rp_plugin_launch2')
    L = _lib.lib()
    h = ctypes.c_void_p()
    n = (ctypes.c_int * 3)(8, 8, 8)
    rc = L.claw_create(ctypes.byref(h), 2, rp.rp_id, 4, 2, 2, n, 2, 0, None)        # a 3-D solver on a 2-D grid
    assert rc != 0 and b'expects ndim=3' in L.claw_last_error(None)
    from clawpack_b200 import pyclaw
    solver = pyclaw.ClawSolver3D(rp)
    assert solver.num_eqn == 4 and solver.num_waves == 2


def test_plugins_build_without_a_gpu_and_register():
    rp = _acoustics()
    assert rp.rp_id >= 1000 and (rp.num_eqn, rp.num_waves, rp.num_aux_read, rp.num_dim) == (3, 2, 0, 2)
    assert rp.cparam_names == ('rho', 'bulk', 'cc', 'zz') and not rp.no_transverse
    rpv = _vc_acoustics()
    assert rpv.rp_id >= 1000 and rpv.rp_id != rp.rp_id and (rpv.num_eqn, rpv.num_waves, rpv.num_aux_read) == (3, 2, 2)
    assert riemann.load_cuda_solver(rp.plugin_path).rp_id == rp.rp_id            # same path, same id
    # the three exported symbols, straight from the shared object
    P = ctypes.CDLL(rp.plugin_path)
    info = (ctypes.c_int * 6)()
    P.claw_rp_plugin_info(info)
    assert list(info)[:5] == [2, 3, 2, 0, 0] and info[5] > 100 and P.claw_rp_plugin_abi() >= 1
    assert hasattr(P, 'claw_rp_plugin_launch2')
    # Solver.__init__ resolves num_eqn / num_waves from the object like for a built-in solver
    from clawpack_b200 import pyclaw
    solver = pyclaw.ClawSolver2D(rp)
    assert solver.num_eqn == 3 and solver.num_waves == 2


def test_plugin_errors_are_reported_not_fatal(tmp_path):
    L = _lib.lib()
    rid = ctypes.c_int(0)
    assert L.claw_rp_plugin_load(str(tmp_path / 'nope.so').encode(), ctypes.byref(rid)) != 0
    assert b'claw_rp_plugin_load' in L.claw_last_error(None)
    # a shared object that is not a plug-in
    assert L.claw_rp_plugin_load(_lib.LIB_PATH.encode(), ctypes.byref(rid)) != 0
    assert b'does not export' in L.claw_last_error(None)
    info = (ctypes.c_int * 5)()
    assert L.claw_rp_plugin_info(4711, info) != 0
    # a registered plug-in checks the sizes the solver is created with (before any device is touched)
    rp = _acoustics()
    h = ctypes.c_void_p()
    n = (ctypes.c_int * 3)(8, 8, 1)
    rc = L.claw_create(ctypes.byref(h), 2, rp.rp_id, 4, 2, 0, n, 2, 0, None)
    assert rc != 0 and b'expects' in L.claw_last_error(None)
    rc = L.claw_create(ctypes.byref(h), 2, 1000 + 500, 3, 2, 0, n, 2, 0, None)
    assert rc != 0 and b'not a registered' in L.claw_last_error(None)
    with pytest.raises(RuntimeError):
        riemann.build_cuda_solver(os.path.join(PLUG, 'user_acoustics.cuh'), 'NoSuchFunctor', out_dir=str(tmp_path))


# ---- GPU: parity against the oracle's solver of the same equations ------------------------------------------------
def _run(sc, nsteps, precision, rp_object=None):
    sc = dict(sc)
    if rp_object is not None:
        sc['rp_object'] = rp_object
    claw = S.make_claw(sc, kernel_language='CUDA', cuda_precision=precision)
    sol, solver = claw.solution, claw.solver
    solver.setup(sol)
    solver.dt = solver.dt_initial
    for _ in range(nsteps):
        solver.evolve_to_time(sol)
    return np.array(sol.state.q, copy=True), list(solver.step_log)


def _ref(sc, nsteps):
    r = ref_solver(sc)
    for _ in range(nsteps):
        r.evolve_to_time()
    return np.array(r.q, copy=True), list(r.step_log)


PLUGIN_CASES = {
    'acoustics_split': (lambda: S.acoustics_2d(mx=100, my=90), _acoustics),
    'acoustics_unsplit_tw2': (lambda: dict(S.acoustics_2d(mx=75, my=64), dimensional_split=False, transverse_waves=2),
                              _acoustics),
    'acoustics_unsplit_tw1_superbee': (lambda: dict(S.acoustics_2d(mx=40, my=61), dimensional_split=False,
                                                    transverse_waves=1, limiters=[S.SUPERBEE, S.VANLEER]), _acoustics),
    'vc_acoustics_unsplit': (lambda: S.acoustics_2d_variable(mx=64, my=37, transverse_waves=2), _vc_acoustics),
    'vc_acoustics_split': (lambda: S.acoustics_2d_variable(mx=57, my=40, dimensional_split=True), _vc_acoustics),
}


@pytest.mark.gpu
@pytest.mark.parametrize('name', sorted(PLUGIN_CASES))
def test_plugin_strict_bit_exact_vs_oracle(name):
    mk, plug = PLUGIN_CASES[name]
    sc = mk()
    q, log = _run(sc, 8, 'strict', plug())
    qr, logr = _ref(sc, 8)
    assert log == logr
    assert np.array_equal(q, qr), float(np.max(np.abs(q - qr)))


@pytest.mark.gpu
@pytest.mark.parametrize('name', sorted(PLUGIN_CASES))
def test_plugin_fast_within_tolerance_and_equal_to_builtin_strict(name):
    mk, plug = PLUGIN_CASES[name]
    sc = mk()
    q, log = _run(sc, 8, 'fast', plug())
    qr, logr = _ref(sc, 8)
    assert [a for _, _, a in log] == [a for _, _, a in logr]
    assert np.max(np.abs(q - qr)) <= 1e-12 * max(np.max(np.abs(qr)), 1e-300)
    # and the plug-in's strict build computes what the built-in functor's strict build computes
    qs, _ = _run(sc, 8, 'strict', plug())
    qb, _ = _run(sc, 8, 'strict')
    assert np.array_equal(qs, qb)


PLUGIN_CASES_3D = {
    'unsplit_22': lambda: S.acoustics_3d_general(mx=20, my=17, mz=15),
    'unsplit_11_superbee': lambda: dict(S.acoustics_3d_general(mx=14, my=30, mz=12), transverse_waves=11, limiters=S.SUPERBEE),
    'split': lambda: S.acoustics_3d_general(mx=13, my=12, mz=33, dimensional_split=True),
    'reference_example': lambda: S.acoustics_3d(mx=24, my=24, mz=24),
}


@pytest.mark.gpu
@pytest.mark.parametrize('name', sorted(PLUGIN_CASES_3D))
def test_3d_plugin_vs_oracle(name):
    sc = PLUGIN_CASES_3D[name]()
    q, log = _run(sc, 5, 'strict', _acoustics3d())
    qr, logr = _ref(sc, 5)
    assert log == logr
    assert np.array_equal(q, qr), float(np.max(np.abs(q - qr)))
    qf, logf = _run(sc, 5, 'fast', _acoustics3d())
    assert [a for _, _, a in logf] == [a for _, _, a in logr]
    assert np.max(np.abs(qf - qr)) <= 1e-12 * max(np.max(np.abs(qr)), 1e-300)


@pytest.mark.gpu
def test_plugin_through_the_f2py_shaped_host_call():
    """claw_step2_host / claw_step2ds_host accept the plug-in's rp_id like a built-in one (kernel_language='CUDA'
    with a host callback, which keeps q on the host between steps)."""
    sc = dict(S.acoustics_2d(mx=50, my=44), rp_object=_acoustics(), before_step=lambda q, aux: None)
    claw = S.make_claw(sc, kernel_language='CUDA', cuda_precision='strict')
    sol, solver = claw.solution, claw.solver
    solver.setup(sol)
    assert solver._dev is None                 # a host callback: every step goes through claw_step2ds_host
    solver.dt = solver.dt_initial
    for _ in range(5):
        solver.evolve_to_time(sol)
    qr, logr = _ref(S.acoustics_2d(mx=50, my=44), 5)
    assert list(solver.step_log) == logr and np.array_equal(sol.state.q, qr)
