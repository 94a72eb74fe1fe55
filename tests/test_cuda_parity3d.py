"""GPU parity tests of the 3-D path (step3 / step3ds / flux3 + vc_acoustics_3D, euler_3D) through the C ABI
against the oracle, which is itself pinned to round-off by the reference's stored 3-D vectors
(tests/test_oracle.py).  Same bars as the 2-D tests: strict arithmetic bit-exact, fast arithmetic within
1e-12 relative L-inf per step, identical accept/reject sequences."""
import numpy as np
import pytest

from clawpack_b200 import scenarios as S
from refutil import goldens, ref_solver
from test_cuda_parity import _rel_linf, _run_cuda, _run_ref

pytestmark = pytest.mark.gpu


def _ac(**kw):
    base = dict(mx=24, my=20, mz=18)
    base.update(kw)
    return S.acoustics_3d(**base)


_ac_general = S.acoustics_3d_general


def _euler_blob(mx=20, my=14, mz=12, **kw):
    """euler_3D with smooth 3-D variation and non-zero velocities in all directions (extrap BCs)."""
    sc = S.euler_3d_shocktube(mx=mx, my=my, mz=mz)
    n = (mx, my, mz)
    cs = [S.centers(-1., 1., k) for k in n]
    X, Y, Z = np.meshgrid(*cs, indexing='ij')
    gamma = 1.4
    rho = 1.0 + 0.5 * np.exp(-4. * ((X - 0.1) ** 2 + (Y + 0.2) ** 2 + (Z - 0.3) ** 2))
    u, v, w = 0.3 * np.sin(2 * X + Y), -0.2 * np.cos(Y - Z), 0.25 * np.sin(Z + 0.5 * X)
    p = 1.0 + 2.0 * np.exp(-6. * (X ** 2 + Y ** 2 + Z ** 2))
    q = np.empty((5,) + n, order='F')
    q[0] = rho
    q[1], q[2], q[3] = rho * u, rho * v, rho * w
    q[4] = p / (gamma - 1.) + 0.5 * rho * (u * u + v * v + w * w)
    sc = dict(sc, q=q, dimensional_split=False, transverse_waves=22, cfl_desired=0.9, dt_initial=0.1)
    sc.update(kw)
    return sc


def _euler_capa(**kw):
    sc = _euler_blob(mx=12, my=10, mz=14, **kw)
    X, Y, Z = np.meshgrid(*[S.centers(-1., 1., k) for k in (12, 10, 14)], indexing='ij')
    aux = np.empty((1, 12, 10, 14), order='F')
    aux[0] = 1.0 + 0.25 * np.sin(2 * X + Y) * np.cos(Z)
    return dict(sc, aux=aux, index_capa=0, aux_bc_lower=sc['bc_lower'], aux_bc_upper=sc['bc_upper'])


CASES3 = {
    'ac_unsplit_22': lambda: _ac(),
    'ac_unsplit_11': lambda: _ac(transverse_waves=11, mx=17, my=9, mz=11),
    'ac_unsplit_10': lambda: dict(_ac(transverse_waves=10, mx=13, my=15, mz=8), cfl_max=0.5, cfl_desired=0.45),
    'ac_unsplit_20': lambda: dict(_ac(transverse_waves=20, mx=12, my=8, mz=13), cfl_max=0.5, cfl_desired=0.45),
    'ac_unsplit_21': lambda: _ac(transverse_waves=21, mx=9, my=12, mz=10),
    'ac_unsplit_0': lambda: dict(_ac(transverse_waves=0, mx=10, my=11, mz=12), cfl_max=0.5, cfl_desired=0.45),
    'ac_split': lambda: _ac(dimensional_split=True),
    'ac_split_homog': lambda: S.acoustics_3d(mx=40, my=4, mz=4, heterogeneous=False),
    'ac_periodic_22': lambda: _ac(periodic=True, mx=16, my=16, mz=16),
    'ac_first_order': lambda: dict(_ac(transverse_waves=11, mx=10, my=9, mz=8), order=1),
    'ac_limiters': lambda: dict(_ac(mx=14, my=10, mz=9), limiters=[S.SUPERBEE, S.VANLEER]),
    'ac_general_22': lambda: _ac_general(),
    'ac_general_21_superbee': lambda: _ac_general(mx=11, my=17, mz=9, transverse_waves=21, limiters=S.SUPERBEE),
    'ac_general_11_o1': lambda: _ac_general(mx=8, my=9, mz=19, transverse_waves=11, order=1),
    'ac_general_split': lambda: _ac_general(mx=13, my=8, mz=11, dimensional_split=True),
    # burgers_3D: scalar, coefficient per direction, entropy fix; no reference example (parity unpinned)
    'burgers3_unsplit_22': lambda: S.burgers_3d(),
    'burgers3_unsplit_11': lambda: S.burgers_3d(mx=30, my=9, mz=11, transverse_waves=11),
    'burgers3_unsplit_21_minmod': lambda: dict(S.burgers_3d(mx=13, my=17, mz=10, transverse_waves=21), limiters=S.MINMOD),
    'burgers3_split': lambda: S.burgers_3d(mx=16, my=12, mz=20, dimensional_split=True),
    # vc_advection_3D: edge velocities in aux, every transverse mode (parity unpinned by the reference)
    'vcadv3_unsplit_22': lambda: S.advection_3d_variable(),
    'vcadv3_unsplit_21': lambda: S.advection_3d_variable(mx=13, my=30, mz=9, transverse_waves=21),
    'vcadv3_unsplit_11_superbee': lambda: dict(S.advection_3d_variable(mx=11, my=10, mz=31, transverse_waves=11), limiters=S.SUPERBEE),
    'vcadv3_split': lambda: S.advection_3d_variable(mx=16, my=12, mz=14, dimensional_split=True),
    # capacity function in 3-D (step3.f90:229-273, step3ds.f90:200-205; pinned by the oracle's identities only, no
    # reference run uses one): a third aux plane the Riemann solver does not read; an Euler run whose ONLY aux plane is it
    'ac_capa_unsplit_22': lambda: S.acoustics_3d_capa(),
    'ac_capa_unsplit_21_superbee': lambda: S.acoustics_3d_capa(mx=11, my=17, mz=9, transverse_waves=21, limiters=S.SUPERBEE),
    'ac_capa_unsplit_10_o1': lambda: S.acoustics_3d_capa(mx=31, my=9, mz=12, transverse_waves=10, order=1),
    'ac_capa_split': lambda: S.acoustics_3d_capa(mx=13, my=8, mz=11, dimensional_split=True),
    'euler_capa_unsplit_22': lambda: _euler_capa(),
    'ac_tiny': lambda: _ac(mx=3, my=2, mz=4),
    'ac_tile_edges': lambda: _ac(mx=57, my=7, mz=13),          # 28-lane strips, 6-row tiles: ragged edges
    'ac_tile_edges2': lambda: _ac(mx=31, my=13, mz=6),
    'euler_split': lambda: S.euler_3d_shocktube(mx=6, my=5, mz=40),
    'euler_unsplit_22': lambda: _euler_blob(),
    'euler_unsplit_11_mc': lambda: _euler_blob(mx=9, my=31, mz=8, transverse_waves=11, limiters=S.MC),
}


@pytest.mark.parametrize('name', sorted(CASES3))
def test_strict_bit_exact_vs_oracle_3d(name):
    sc = CASES3[name]()
    q, log, _ = _run_cuda(sc, 5, 'strict')
    qr, logr = _run_ref(sc, 5)
    assert [s[2] for s in log] == [s[2] for s in logr], 'accept/reject sequence differs'
    assert log == logr, 'dt / CFL sequence differs: %r vs %r' % (log[:3], logr[:3])
    assert np.array_equal(q, qr), 'max abs diff %g' % np.max(np.abs(q - qr))


@pytest.mark.parametrize('name', sorted(CASES3))
def test_fast_within_tolerance_3d(name):
    sc = CASES3[name]()
    q, log, _ = _run_cuda(sc, 2, 'fast')
    qr, logr = _run_ref(sc, 2)
    assert [s[2] for s in log] == [s[2] for s in logr]
    assert _rel_linf(q, qr) <= 1e-12


@pytest.mark.parametrize('name', ['ac_general_22', 'ac_split', 'euler_unsplit_22'])
def test_100_steps_3d(name):
    sc = CASES3[name]()
    qr, logr = _run_ref(sc, 101)
    for precision, tol in (('strict', 0.0), ('fast', 1e-10)):
        q, log, _ = _run_cuda(sc, 101, precision)
        assert [s[2] for s in log] == [s[2] for s in logr], 'accept/reject sequence differs (%s)' % precision
        assert _rel_linf(q, qr) <= tol, (precision, _rel_linf(q, qr))


@pytest.mark.parametrize('precision', ['strict', 'fast'])
def test_acoustics3d_reference_golden(precision):
    """Controller.run() of the reference's 3-D acoustics example (15^3, unsplit (2,22), aux, wall/periodic) on the
    GPU against the reference's stored vector verify_classic_heterogeneous.txt."""
    sc = S.acoustics_3d(mx=15, my=15, mz=15)
    claw = S.make_claw(sc, kernel_language='CUDA', cuda_precision=precision)
    claw.run()
    p = claw.frames[-1].q[0].reshape(-1)
    assert np.max(np.abs(p - goldens()['acoustics_3d__verify_classic_heterogeneous'])) < 1e-12


@pytest.mark.parametrize('precision', ['strict', 'fast'])
def test_euler3d_reference_goldens(precision):
    """examples/euler_3d/test_euler3d.py: shocktube (device-resident, split) and shock_bubble (BC.custom -> the
    f2py-shaped plug-in step3 on host arrays), L1 < 1e-6 as in the reference test."""
    G = goldens()
    sc = S.euler_3d_shocktube()
    claw = S.make_claw(sc, kernel_language='CUDA', cuda_precision=precision)
    claw.run()
    rho = claw.frames[claw.num_output_times].state.get_q_global()[0].reshape(-1)
    assert np.prod(claw.solution.domain.grid.delta) * np.sum(np.abs(rho - G['euler_3d__shocktube'])) < 1e-6
    assert np.max(np.abs(rho - G['euler_3d__shocktube'])) < 1e-11
    sc = S.euler_3d_shock_bubble(G['euler_3d__shock_bubble_q0'], G['euler_3d__shock_bubble_consts'])
    claw = S.make_claw(sc, kernel_language='CUDA', cuda_precision=precision)
    claw.run()
    assert claw.solver._dev is None and claw.solver.fmod is not None      # BC.custom => plug-in mode
    rho = claw.frames[claw.num_output_times].state.get_q_global()[0].reshape(-1)
    assert np.prod(claw.solution.domain.grid.delta) * np.sum(np.abs(rho - G['euler_3d__shock_bubble'])) < 1e-6
    assert np.max(np.abs(rho - G['euler_3d__shock_bubble'])) < 1e-11


def test_plugin_fmod3_matches_oracle_fmod():
    """step3 / step3ds on host arrays (claw_step3_host, claw_step3ds_host) against the oracle plug-in."""
    from clawpack_b200 import riemann
    from clawpack_b200.pyclaw import cuda_classic
    from oracle_fmod import OracleFmod
    sc = _ac(mx=19, my=11, mz=14)
    r = ref_solver(sc)
    r.apply_bcs()
    rp = riemann.vc_acoustics_3D
    dx, dy, dz = r.delta
    dt = 0.02
    mthlim = [4, 4]
    n = (19, 11, 14)
    cu, orc = cuda_classic.fmod(3), OracleFmod()
    method = np.array([1, 2, 22, 0, 0, 0, 2], dtype=np.intc)
    qa, ca = cu.step3(19, 2, *n, r.qbc.copy('F'), r.qbc.copy('F'), r.auxbc, dx, dy, dz, dt, method, mthlim, 0, 0, 0, 0,
                      False, rp, rp, rp)
    qb, cb = orc.step3(19, 2, *n, r.qbc.copy('F'), r.qbc.copy('F'), r.auxbc, dx, dy, dz, dt, method, mthlim, 0, 0, 0,
                       0, False, rp, rp, rp)
    assert ca == cb
    assert np.array_equal(qa[:, 2:-2, 2:-2, 2:-2], qb[:, 2:-2, 2:-2, 2:-2])
    method = np.array([1, 2, -1, 0, 0, 0, 2], dtype=np.intc)
    qa, qb = r.qbc.copy('F'), r.qbc.copy('F')
    for idir in (1, 2, 3):
        qa, ca = cu.step3ds(19, 2, *n, qa.copy('F'), qa, r.auxbc, dx, dy, dz, dt, method, mthlim, 0, 0, 0, 0, idir,
                            False, rp, rp, rp)
        qb, cb = orc.step3ds(19, 2, *n, qb, qb, r.auxbc, dx, dy, dz, dt, method, mthlim, 0, 0, 0, 0, idir, False, rp,
                             rp, rp)
        assert ca == cb, idir
        assert np.array_equal(qa, qb), idir             # whole array incl. the swept ghost layer
