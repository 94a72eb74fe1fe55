"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on identical data.

Bars (BASELINE.json north_star): relative L-inf <= 1e-12 after one step, <= 1e-10 after 100 steps,
identical accept/reject sequences.  In 'strict' arithmetic (no FMA contraction) the kernels are
op-for-op the reference, so we additionally demand BIT-EXACT equality with the oracle."""
import os

import numpy as np
import pytest

from clawpack_b200 import scenarios as S
from refutil import goldens, ref_solver

pytestmark = pytest.mark.gpu


def _cuda_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


CASES = {
    'advection_split': lambda: S.advection_2d(mx=50, my=50, dimensional_split=True),
    'advection_unsplit_notrans': lambda: S.advection_2d(dimensional_split=False, transverse_waves=0),
    'advection_unsplit_inc': lambda: S.advection_2d(dimensional_split=False, transverse_waves=1),
    'advection_unsplit_cor': lambda: S.advection_2d(dimensional_split=False, transverse_waves=2),
    'acoustics_split': lambda: S.acoustics_2d(mx=100, my=100),
    'acoustics_unsplit_cor': lambda: dict(S.acoustics_2d(mx=60, my=44), dimensional_split=False, transverse_waves=2,
                                          cfl_max=1.0, cfl_desired=0.9),
    'euler_split': lambda: S.quadrants(mx=100, my=100),
    'euler_unsplit_cor': lambda: dict(S.quadrants(mx=70, my=90), dimensional_split=False, transverse_waves=2,
                                      limiters=S.MC),
    'euler_unsplit_inc_superbee': lambda: dict(S.quadrants(mx=130, my=37), dimensional_split=False,
                                               transverse_waves=1, limiters=[S.SUPERBEE, S.MINMOD, S.VANLEER, S.MC]),
    'euler_shockbubble': lambda: S.shock_bubble(mx=160, my=120),
    'shallow_split': lambda: S.radial_dam_break(mx=150, my=150),
    'shallow_unsplit_cor': lambda: dict(S.radial_dam_break(mx=64, my=50), dimensional_split=False, transverse_waves=2),
    'euler_first_order': lambda: dict(S.quadrants(mx=48, my=48), order=1, dimensional_split=False, transverse_waves=1),
    'euler_hlle_split': lambda: S.quadrants(mx=70, my=50, riemann_solver='hlle'),
    'euler_hlle_unsplit_notrans': lambda: dict(S.quadrants(mx=40, my=66, riemann_solver='hlle'), dimensional_split=False),
    'shallow_hlle_split': lambda: S.radial_dam_break(mx=60, my=45, riemann_solver='hlle'),
    'burgers2_unsplit_cor': lambda: S.burgers_2d(),
    'burgers2_unsplit_inc_minmod': lambda: dict(S.burgers_2d(mx=31, my=77, transverse_waves=1), limiters=S.MINMOD),
    'burgers2_split': lambda: S.burgers_2d(mx=90, my=33, dimensional_split=True),
    'euler_tiny': lambda: dict(S.quadrants(mx=3, my=2), dimensional_split=False, transverse_waves=2),
    'euler_wide_strip_edge': lambda: dict(S.quadrants(mx=249, my=9), dimensional_split=False, transverse_waves=2),
}


def _run_cuda(sc, nsteps, precision):
    claw = S.make_claw(sc, kernel_language='CUDA', cuda_precision=precision)
    sol, solver = claw.solution, claw.solver
    solver.setup(sol)
    solver.dt = solver.dt_initial
    for _ in range(nsteps):
        solver.evolve_to_time(sol)
    return np.array(sol.state.q, copy=True), list(solver.step_log), solver


def _run_ref(sc, nsteps):
    r = ref_solver(sc)
    for _ in range(nsteps):
        r.evolve_to_time()
    return np.array(r.q, copy=True), list(r.step_log)


def _rel_linf(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


@pytest.mark.parametrize('name', sorted(CASES))
def test_strict_bit_exact_vs_oracle(name):
    sc = CASES[name]()
    q, log, _ = _run_cuda(sc, 6, 'strict')
    qr, logr = _run_ref(sc, 6)
    assert [s[2] for s in log] == [s[2] for s in logr], 'accept/reject sequence differs'
    assert log == logr, 'dt / CFL sequence differs: %r vs %r' % (log[:3], logr[:3])
    assert np.array_equal(q, qr), 'max abs diff %g' % np.max(np.abs(q - qr))


@pytest.mark.parametrize('name', sorted(CASES))
def test_fast_within_tolerance(name):
    sc = CASES[name]()
    q, log, _ = _run_cuda(sc, 2, 'fast')       # dt_initial is a reject for most cases: 1 reject + 1 accept
    qr, logr = _run_ref(sc, 2)
    assert [s[2] for s in log] == [s[2] for s in logr]
    assert _rel_linf(q, qr) <= 1e-12           # north_star: 1e-12 relative L-inf per step


@pytest.mark.parametrize('name', ['euler_unsplit_cor', 'shallow_split', 'acoustics_split', 'euler_shockbubble'])
def test_100_steps(name):
    sc = CASES[name]()
    qr, logr = _run_ref(sc, 101)
    for precision, tol in (('strict', 0.0), ('fast', 1e-10)):
        q, log, _ = _run_cuda(sc, 101, precision)
        assert [s[2] for s in log] == [s[2] for s in logr], 'accept/reject sequence differs (%s)' % precision
        assert _rel_linf(q, qr) <= tol, (precision, _rel_linf(q, qr))


GOLDEN_RUNS = [
    ('advection_2d__classic_dim_split', lambda: S.advection_2d(dimensional_split=True), 1e-6),
    ('advection_2d__classic_unsplit_no_trans', lambda: S.advection_2d(dimensional_split=False, transverse_waves=0), 1e-6),
    ('advection_2d__classic_unsplit_trans_inc', lambda: S.advection_2d(dimensional_split=False, transverse_waves=1), 1e-6),
    ('advection_2d__classic_unsplit_trans_cor', lambda: S.advection_2d(dimensional_split=False, transverse_waves=2), 1e-6),
    ('euler_2d__quadrants_roe', lambda: S.quadrants(), 1e-6),
    ('shallow_2d__radialdambreak_classic_roe', lambda: S.radial_dam_break(), 1e-5),
]


@pytest.mark.parametrize('key,make,tol', GOLDEN_RUNS)
@pytest.mark.parametrize('precision', ['strict', 'fast'])
def test_full_run_vs_reference_golden(key, make, tol, precision):
    """Whole Controller.run() on the GPU against the vector the reference's own Fortran path stored
    (same L1 metric and tolerance as the reference tests, e.g. examples/euler_2d/test_euler2d.py:10-26)."""
    sc = make()
    claw = S.make_claw(sc, kernel_language='CUDA', cuda_precision=precision)
    claw.run()
    q0 = claw.frames[claw.num_output_times].state.get_q_global()[0].reshape(-1)
    dx, dy = claw.solution.domain.grid.delta
    expected = goldens()[key].reshape(-1)
    assert dx * dy * np.sum(np.abs(q0 - expected)) < tol
    # and far tighter than the reference tolerance: we sit on the oracle, which sits on the golden
    assert np.max(np.abs(q0 - expected)) < 1e-11


def test_acoustics_golden():
    sc = S.acoustics_2d()
    claw = S.make_claw(sc, kernel_language='CUDA')
    claw.run()
    p = claw.frames[-1].q[0]
    expected = goldens()['acoustics_2d__verify_classic']
    # reference tolerance: relative L2 < 1e-3 (examples/acoustics_2d_homogeneous/test_2d_acoustics.py:14)
    assert np.linalg.norm(p - expected) / np.linalg.norm(expected) < 1e-3


def test_plugin_fmod_matches_oracle_fmod():
    """The f2py-shaped plug-in (host buffers) against the oracle, unsplit + both split sweeps."""
    from clawpack_b200 import riemann
    from clawpack_b200.pyclaw import cuda_classic
    from oracle_fmod import OracleFmod
    sc = S.quadrants(mx=90, my=70)
    r = ref_solver(sc)
    r.apply_bcs()
    rp = riemann.euler_4wave_2D
    rp.cparam.gamma = 1.4
    dx, dy = r.delta
    dt = 0.002
    mthlim = [4, 4, 4, 4]
    cu, orc = cuda_classic.fmod(2), OracleFmod()
    for method2 in (2, -1):
        method = np.array([1, 2, method2, 0, 0, 0, 0], dtype=np.intc)
        if method2 >= 0:
            qa, qb = r.qbc.copy('F'), r.qbc.copy('F')
            qa, ca = cu.step2(90, 2, 90, 70, r.qbc.copy('F'), qa, None, dx, dy, dt, method, mthlim, 0, 0, 0, 0, False, rp, rp)
            qb, cb = orc.step2(90, 2, 90, 70, r.qbc.copy('F'), qb, None, dx, dy, dt, method, mthlim, 0, 0, 0, 0, False, rp, rp)
            assert ca == cb
            assert np.array_equal(qa[:, 2:-2, 2:-2], qb[:, 2:-2, 2:-2])
        else:
            for ids in (1, 2):
                qa, qb = r.qbc.copy('F'), r.qbc.copy('F')
                qa, ca = cu.step2ds(90, 2, 90, 70, r.qbc.copy('F'), qa, None, dx, dy, dt, method, mthlim, 0, 0, 0, 0, ids, False, rp, rp)
                qb, cb = orc.step2ds(90, 2, 90, 70, r.qbc.copy('F'), qb, None, dx, dy, dt, method, mthlim, 0, 0, 0, 0, ids, False, rp, rp)
                assert ca == cb, ids
                assert np.array_equal(qa, qb), ids     # whole array incl. the updated ghost rows / columns


def test_plugin_pipelined_large_grid_matches_oracle():
    """claw_step2_host switches to the chunked transfer/compute pipeline above 2^20 cells: same result, bit for
    bit, as the oracle plug-in (threaded oracle = serial oracle, tests/test_oracle.py)."""
    from clawpack_b200 import riemann
    from clawpack_b200.pyclaw import cuda_classic
    from oracle_fmod import OracleFmod
    mx, my = 1200, 1100
    sc = S.shock_bubble(mx=mx, my=my)
    r = ref_solver(sc)
    r.apply_bcs()
    rp = riemann.euler_4wave_2D
    rp.cparam.gamma = 1.4
    dx, dy = r.delta
    dt = 0.9 * dx / 3.5
    method = np.array([1, 2, 2, 0, 0, 0, 0], dtype=np.intc)
    mthlim = [4, 4, 4, 4]
    cu, orc = cuda_classic.fmod(2), OracleFmod(nthreads=8)
    qa, qb = r.qbc.copy('F'), r.qbc.copy('F')
    for _ in range(2):      # twice: the second call reuses the pipeline's streams / staging
        qa, ca = cu.step2(mx, 2, mx, my, qa.copy('F'), qa, None, dx, dy, dt, method, mthlim, 0, 0, 0, 0, False, rp, rp)
        qb, cb = orc.step2(mx, 2, mx, my, qb.copy('F'), qb, None, dx, dy, dt, method, mthlim, 0, 0, 0, 0, False, rp, rp)
        assert ca == cb
        assert np.array_equal(qa[:, 2:-2, 2:-2], qb[:, 2:-2, 2:-2])
        # ghost cells of qnew are left as passed (include/clawcuda.h)
        qa[:, :2, :], qa[:, -2:, :], qa[:, :, :2], qa[:, :, -2:] = qb[:, :2, :], qb[:, -2:, :], qb[:, :, :2], qb[:, :, -2:]
    cu.L.claw_host_release()


def test_no_cpu_fallback_symbols_loaded():
    """The CUDA path must run in libclawcuda.so, not in the oracle."""
    import ctypes
    from clawpack_b200 import _lib
    L = _lib.lib()
    assert L.claw_device_count() >= 1
    assert b'clawcuda' in L.claw_version()


# ---- 1-D: claw_step1_host (csrc/sweep1.cuh, rp1.cuh) behind ClawSolver1D ------------------------------
CASES_1D = {
    'advection_1d': lambda: S.advection_1d(nx=300, tfinal=0.3),
    'acoustics_1d': lambda: S.acoustics_1d(num_cells=257, tfinal=0.3),
    'burgers_1d': lambda: S.burgers_1d(mx=500),
    'burgers_1d_first_order': lambda: dict(S.burgers_1d(mx=123), order=1),
    'dam_break_1d_roe': lambda: S.dam_break_1d(mx=500),
    'dam_break_1d_hlle': lambda: S.dam_break_1d(mx=250, riemann_solver='hlle'),
    'dam_break_1d_tiny': lambda: S.dam_break_1d(mx=3),
    'acoustics_1d_tile_edge': lambda: dict(S.acoustics_1d(num_cells=124 * 2 - 4 + 1), limiters=[S.SUPERBEE, S.MINMOD]),
    'acoustics_1d_wall': lambda: dict(S.acoustics_1d(num_cells=90), bc_lower=[S.WALL], bc_upper=[S.EXTRAP]),
}


@pytest.mark.parametrize('name', sorted(CASES_1D))
def test_1d_strict_bit_exact_vs_oracle(name):
    sc = CASES_1D[name]()
    q, log, solver = _run_cuda(sc, 40, 'strict')
    qr, logr = _run_ref(sc, 40)
    assert solver._dev is None and solver.fmod is not None          # 1-D always runs through the plug-in call
    assert log == logr, 'dt / CFL sequence differs: %r vs %r' % (log[:3], logr[:3])
    assert np.array_equal(q, qr), 'max abs diff %g' % np.max(np.abs(q - qr))


@pytest.mark.parametrize('name', sorted(CASES_1D))
def test_1d_fast_within_tolerance(name):
    sc = CASES_1D[name]()
    qr, logr = _run_ref(sc, 101)
    q, log, _ = _run_cuda(sc, 101, 'fast')
    assert [s[2] for s in log] == [s[2] for s in logr]
    assert _rel_linf(q, qr) <= 1e-10


@pytest.mark.parametrize('precision', ['strict', 'fast'])
def test_1d_reference_goldens(precision):
    """burgers_1d/test_burgers_1d.py:34-35, shallow_1d/test_shallow1d.py:25-34 (L1 < 1e-6 against the stored vectors)
    and the error values of advection_1d/test_advection.py:30-31, acoustics_1d_homogeneous/test_acoustics.py:28-29,
    through Controller.run() with kernel_language='CUDA'."""
    g2 = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_goldens2.npz'))
    for sc, key in ((S.burgers_1d(), 'burgers_1d__fortran_classic'),
                    (S.dam_break_1d(riemann_solver='roe'), 'shallow_1d__dambreak_fortran_classic_roe'),
                    (S.dam_break_1d(riemann_solver='hlle'), 'shallow_1d__dambreak_fortran_classic_hlle')):
        claw = S.make_claw(sc, kernel_language='CUDA', cuda_precision=precision)
        claw.run()
        q = claw.frames[claw.num_output_times].state.get_q_global()[0].reshape(-1)
        dx = claw.solution.domain.grid.delta[0]
        assert dx * np.sum(np.abs(q - g2[key])) < 1e-6, key
        assert np.max(np.abs(q - g2[key])) < 1e-10, key
    for sc, val, tol in ((S.advection_1d(), 3.203924e-04, 1e-4), (S.acoustics_1d(), 0.001981, 1e-5)):
        claw = S.make_claw(sc, kernel_language='CUDA', cuda_precision=precision)
        claw.run()
        q0 = claw.frames[0].state.get_q_global().reshape(-1)
        qf = claw.frames[claw.num_output_times].state.get_q_global().reshape(-1)
        err = claw.solution.domain.grid.delta[0] * np.sum(np.abs(qf - q0))
        assert abs(err - val) < tol, (sc['rp'], err)
