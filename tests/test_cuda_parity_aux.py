"""GPU parity tests for the 2-D solvers that read aux arrays, capacity functions and f-waves
(csrc/sweep2_aux.cuh, rp2_aux.cuh) against the oracle, through the C ABI.  Same bars as test_cuda_parity.py:
bit-exact in 'strict' arithmetic, 1e-12 / 1e-10 relative L-inf in 'fast', identical step logs."""
import os

import numpy as np
import pytest

from clawpack_b200 import scenarios as S
from refutil import ROOT, ref_solver
from test_cuda_parity import _rel_linf, _run_cuda, _run_ref

pytestmark = pytest.mark.gpu


def goldens2():
    return np.load(os.path.join(ROOT, 'tests', 'golden', 'reference_goldens2.npz'))


CASES = {
    'vcac_unsplit_inc': lambda: S.acoustics_2d_variable(),
    'vcac_unsplit_cor': lambda: S.acoustics_2d_variable(mx=64, my=37, transverse_waves=2),
    'vcac_unsplit_notrans': lambda: dict(S.acoustics_2d_variable(mx=31, my=58, transverse_waves=0), cfl_max=0.5,
                                         cfl_desired=0.45),
    'vcac_split': lambda: S.acoustics_2d_variable(mx=57, my=40, dimensional_split=True),
    'vcac_first_order': lambda: dict(S.acoustics_2d_variable(mx=40, my=40), order=1),
    'vcac_superbee_periodic': lambda: dict(S.acoustics_2d_variable(mx=45, my=30, transverse_waves=2),
                                           limiters=[S.SUPERBEE, S.VANLEER], bc_lower=[S.PERIODIC, S.WALL],
                                           bc_upper=[S.PERIODIC, S.EXTRAP], aux_bc_lower=[S.PERIODIC, S.WALL],
                                           aux_bc_upper=[S.PERIODIC, S.EXTRAP]),
    'vcac_tiny': lambda: S.acoustics_2d_variable(mx=3, my=2, transverse_waves=2),
    'vcac_strip_edges': lambda: S.acoustics_2d_variable(mx=57, my=5, transverse_waves=2),   # 2 warps + 1 column
}
CASES.update({
    # capacity function + aux edge velocities + custom aux ghost cells (mapped annulus grid)
    'annulus_unsplit_cor': lambda: S.advection_annulus(),
    'annulus_unsplit_inc': lambda: S.advection_annulus(mr=33, mtheta=61, transverse_waves=1),
    'annulus_split': lambda: S.advection_annulus(mr=30, mtheta=58, dimensional_split=True),
    'annulus_first_order': lambda: S.advection_annulus(mr=29, mtheta=40, order=1, transverse_waves=1),
    # euler_5wave_2D: device-resident (no host callbacks) ...
    'euler5_split': lambda: S.quadrants_tracer(),
    'euler5_unsplit_cor': lambda: S.quadrants_tracer(mx=45, my=70, dimensional_split=False, transverse_waves=2),
    'euler5_unsplit_inc': lambda: S.quadrants_tracer(mx=33, my=20, dimensional_split=False, transverse_waves=1),
    # ... and the reference example: host source term + BC.custom inflow -> f2py-shaped plug-in mode (with an aux array)
    'sbi_split': lambda: S.shock_bubble_interaction(goldens2()['euler_2d__shock_bubble_interaction_q0']),
    'sbi_unsplit_cor': lambda: S.shock_bubble_interaction(goldens2()['euler_2d__shock_bubble_interaction_q0'][:, :80, :],
                                                          dimensional_split=False, transverse_waves=2),
})


CASES.update({
    # psystem_2D (f-wave, 4 aux planes, before_step -> plug-in mode); relations 1 (linear) and 3 (cubic) are
    # exp-free and therefore bit-exact; the example's exponential relation is compared with a tolerance below
    'psystem_linear': lambda: S.psystem_2d(cells_per_layer=3, stress_rel=1, size=10.0),
    'psystem_cubic_inc': lambda: S.psystem_2d(cells_per_layer=4, stress_rel=3, size=7.0, transverse_waves=1),
    'psystem_linear_split': lambda: S.psystem_2d(cells_per_layer=3, stress_rel=1, size=9.0, dimensional_split=True),
    # ... and device-resident (no before_step: the strain plane of aux keeps its initial zeros)
    'psystem_linear_resident': lambda: dict(S.psystem_2d(cells_per_layer=3, stress_rel=1, size=10.0), before_step=None),
})


CASES.update({
    # acoustics_mapped_2D: 9 aux planes, capacity function, mapped grid; BC.custom inflow -> plug-in mode ...
    'mapped_acoustics_cor': lambda: S.acoustics_2d_mapped(num_cells=60, transverse_waves=2),
    'mapped_acoustics_inc': lambda: S.acoustics_2d_mapped(num_cells=45),
    # ... and device-resident with plain extrapolation at x-lower (a pulse is put into the initial data)
    'mapped_acoustics_resident': lambda: _mapped_resident(),
})


CASES.update({
    # shallow_bathymetry_fwave_2D (3 f-waves, bathymetry in aux): wave-maker BC.custom -> plug-in mode, and
    # device-resident with walls all round
    'sill_split': lambda: S.sill(my=12, tfinal=2.0, num_output_times=2),
    # (unsplit only with complete ghost fills: the wave maker leaves ghost cells stale, and the unsplit Fortran
    # step leaves partial sums in them that claw_step2_host documents it does not reproduce)
    'sill_unsplit_notrans': lambda: dict(S.sill(my=10), dimensional_split=False, transverse_waves=0, cfl_max=0.5,
                                         cfl_desired=0.45, custom_bc_lower_full=None, bc_lower=[S.EXTRAP, S.WALL]),
    'sill_resident': lambda: dict(S.sill(my=16), custom_bc_lower_full=None, bc_lower=[S.WALL, S.WALL]),
})


def _mapped_resident():
    sc = S.acoustics_2d_mapped(num_cells=50, transverse_waves=2)
    sc['custom_bc_lower_full'] = None
    sc['bc_lower'] = [S.EXTRAP, S.EXTRAP]
    x = (np.arange(50) + 0.5) / 50
    sc['q'][0] = np.exp(-200 * (x[:, None] - 0.2) ** 2) * np.ones((1, 50))
    return sc


@pytest.mark.parametrize('name', sorted(CASES))
def test_strict_bit_exact_vs_oracle(name):
    sc = CASES[name]()
    q, log, _ = _run_cuda(sc, 6, 'strict')
    qr, logr = _run_ref(sc, 6)
    assert log == logr, 'dt / CFL sequence differs: %r vs %r' % (log[:3], logr[:3])
    assert np.array_equal(q, qr), 'max abs diff %g' % np.max(np.abs(q - qr))


@pytest.mark.parametrize('name', sorted(CASES))
def test_fast_within_tolerance(name):
    sc = CASES[name]()
    q, log, _ = _run_cuda(sc, 2, 'fast')
    qr, logr = _run_ref(sc, 2)
    assert [s[2] for s in log] == [s[2] for s in logr]
    assert _rel_linf(q, qr) <= 1e-12


@pytest.mark.parametrize('name', ['vcac_unsplit_inc', 'vcac_split', 'annulus_unsplit_cor'])
def test_100_steps(name):
    sc = CASES[name]()
    qr, logr = _run_ref(sc, 101)
    for precision, tol in (('strict', 0.0), ('fast', 1e-10)):
        q, log, _ = _run_cuda(sc, 101, precision)
        assert [s[2] for s in log] == [s[2] for s in logr], 'accept/reject sequence differs (%s)' % precision
        assert _rel_linf(q, qr) <= tol, (precision, _rel_linf(q, qr))


@pytest.mark.parametrize('precision', ['strict', 'fast'])
def test_acoustics_2d_variable_golden(precision):
    """pyclaw/examples/acoustics_2d_variable/test_acoustics_2d_variable.py:8-23 (classic, 50x50): max-abs < 1e-1 there;
    the CUDA path reproduces the stored vector to round-off."""
    sc = S.acoustics_2d_variable()
    claw = S.make_claw(sc, kernel_language='CUDA', cuda_precision=precision)
    claw.run()
    p = claw.frames[claw.num_output_times].state.get_q_global()[0]
    expected = goldens2()['acoustics_2d_variable__pressure_classic']
    assert np.max(np.abs(p.reshape(-1) - expected.reshape(-1))) < 1e-12


@pytest.mark.parametrize('precision', ['strict', 'fast'])
def test_advection_annulus_golden(precision):
    """pyclaw/examples/advection_2d_annulus/test_2d_advection_annulus.py:9-30: area-weighted L1 < 1e-6."""
    sc = S.advection_annulus()
    claw = S.make_claw(sc, kernel_language='CUDA', cuda_precision=precision)
    claw.run()
    q = claw.frames[claw.num_output_times].state.get_q_global()[0].reshape(-1)
    dx, dy = claw.solution.domain.grid.delta
    area = (sc['aux'][2] * dx * dy).reshape(-1)
    expected = goldens2()['advection_annulus__classic']
    assert np.sum(np.abs(area * (q - expected))) < 1e-6
    assert np.max(np.abs(q - expected)) < 1e-12


@pytest.mark.parametrize('precision', ['strict', 'fast'])
def test_shock_bubble_interaction_golden(precision):
    """pyclaw/examples/euler_2d/test_euler2d.py:35-37 (160x40, tfinal 0.2, one output time): L1 < 1e-6.  The host
    source term and the BC.custom inflow put the solver in plug-in mode: every step goes through claw_step2ds_host."""
    sc = S.shock_bubble_interaction(goldens2()['euler_2d__shock_bubble_interaction_q0'])
    claw = S.make_claw(sc, kernel_language='CUDA', cuda_precision=precision)
    claw.run()
    assert claw.solver._dev is None and claw.solver.fmod is not None        # plug-in mode
    rho = claw.frames[claw.num_output_times].state.get_q_global()[0].reshape(-1)
    dx, dy = claw.solution.domain.grid.delta
    expected = goldens2()['euler_2d__shock_bubble_interaction']
    assert dx * dy * np.sum(np.abs(rho - expected)) < 1e-6
    assert np.max(np.abs(rho - expected)) < (1e-11 if precision == 'strict' else 1e-8)


@pytest.mark.parametrize('precision', ['strict', 'fast'])
def test_psystem_exponential_within_tolerance(precision):
    """The example's exponential stress-strain relation goes through exp(): CUDA's and glibc's results may differ in
    the last bit, so both arithmetic modes are held to the north-star tolerance (1e-12 per step), with identical
    accept/reject sequences."""
    sc = S.psystem_2d(cells_per_layer=4, size=10.0)
    q, log, _ = _run_cuda(sc, 3, precision)
    qr, logr = _run_ref(sc, 3)
    assert [s[2] for s in log] == [s[2] for s in logr]
    assert _rel_linf(q, qr) <= 3e-12


@pytest.mark.parametrize('resident', [False, True])
def test_psystem_2d_gauges_golden(resident, tmp_path):
    """pyclaw/examples/psystem_2d/test_2d_psystem.py:9-25: the four gauge records, in memory and in the gauge files,
    against the reference's stored files (np.allclose there; 12 stored digits) -- plug-in mode, because of the
    before_step callback.  `resident` drops the callback to put the same run on the device-resident path, where the
    gauge cells are read through claw_get_cells: the last record must equal the downloaded final state."""
    sc = S.psystem_2d()
    if resident:
        sc['before_step'] = None
    claw = S.make_claw(sc, kernel_language='CUDA', cuda_precision='strict')
    claw.outdir = str(tmp_path)
    claw.output_format = None
    claw.run()
    state = claw.solution.state
    assert (claw.solver._dev is not None) == resident
    if resident:
        # last record of every gauge == the downloaded final state at the gauge cell
        for i, g in enumerate(state.grid.gauges):
            p = sc['compute_gauge_values'](state.q[:, g[0], g[1]], state.aux[:, g[0], g[1]])
            assert np.array_equal(state.gauge_data[i][-1, 1:], np.asarray(p))
            assert state.gauge_data[i][-1, 0] == claw.solution.t
        return
    g2 = goldens2()
    for i, name in enumerate(('0.25_0.25', '17.85_1.25', '3.25_18.75', '11.75_11.75')):
        gold = g2['psystem_2d__gauge' + name]
        assert state.gauge_data[i].shape == gold.shape
        assert np.allclose(gold, state.gauge_data[i], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize('precision', ['strict', 'fast'])
def test_acoustics_2d_mapped_golden(precision):
    """pyclaw/examples/acoustics_2d_mapped/test_acoustics_2d_mapped.py:6-21: max-abs < 1e-7 against pressure.npz."""
    sc = S.acoustics_2d_mapped()
    claw = S.make_claw(sc, kernel_language='CUDA', cuda_precision=precision)
    claw.run()
    p = claw.frames[-1].q[0, :, :]
    assert np.max(np.abs(p - goldens2()['acoustics_2d_mapped__pressure'])) < (1e-12 if precision == 'strict' else 1e-9)


@pytest.mark.parametrize('precision', ['strict', 'fast'])
def test_sill_golden(precision):
    """pyclaw/examples/shallow_2d/test_shallow2d.py:36-37: L1 < 1e-6 against expected_sols['sill']."""
    sc = S.sill()
    claw = S.make_claw(sc, kernel_language='CUDA', cuda_precision=precision)
    claw.run()
    h = claw.frames[claw.num_output_times].state.get_q_global()[0].reshape(-1)
    dx, dy = claw.solution.domain.grid.delta
    g = goldens2()['shallow_2d__sill']
    assert dx * dy * np.sum(np.abs(h - g)) < 1e-6
    assert np.max(np.abs(h - g)) < (1e-12 if precision == 'strict' else 1e-9)


def test_missing_transverse_solver_is_refused():
    """rpt2_dummy solvers and shallow_bathymetry_fwave_2D (whose rpt2 reads an unassigned value) cannot take transverse
    corrections: the library says so instead of producing garbage like the reference would."""
    from clawpack_b200._lib import ClawCudaError
    for sc in (dict(S.sill(my=8), dimensional_split=False, transverse_waves=2, custom_bc_lower_full=None,
                    bc_lower=[S.WALL, S.WALL]),
               dict(S.quadrants(mx=12, my=12, riemann_solver='hlle'), dimensional_split=False, transverse_waves=1)):
        claw = S.make_claw(sc, kernel_language='CUDA')
        with pytest.raises(ClawCudaError):
            claw.run()


# ---- device-side source term and BC.custom (pyclaw.device_callbacks) --------------------------------------------------
def test_shock_bubble_interaction_device_resident_bit_exact():
    """The reference's shock-bubble example (euler_5wave_2D + step_Euler_radial + incoming_shock) with the source term and
    the inflow fill as DEVICE functors: the solver stays on the device-resident path (no plug-in object, no host round
    trip per step) and reproduces the oracle run with the host callbacks bit for bit in strict arithmetic, step log
    included; and the reference's stored vector to its tolerance (test_euler2d.py:35-37: L1 < 1e-6)."""
    q0 = goldens2()['euler_2d__shock_bubble_interaction_q0']
    sc = S.shock_bubble_interaction(q0, device_callbacks=True)
    claw = S.make_claw(sc, kernel_language='CUDA', cuda_precision='strict')
    claw.run()
    assert claw.solver._dev is not None and claw.solver.fmod is None          # device-resident, not plug-in mode
    q = claw.frames[claw.num_output_times].state.get_q_global()
    r = ref_solver(S.shock_bubble_interaction(q0))
    r.run(sc['tfinal'], sc['num_output_times'])
    assert claw.solver.step_log == r.step_log
    assert np.array_equal(q, r.q), np.max(np.abs(q - r.q))
    e = goldens2()['euler_2d__shock_bubble_interaction']
    dx, dy = claw.solution.domain.grid.delta
    assert dx * dy * np.sum(np.abs(q[0].reshape(-1) - e)) < 1e-6


def test_shock_bubble_interaction_device_resident_fast_and_host_call():
    """Fast arithmetic within tolerance; and the device functor objects are ordinary host callbacks as well (plug-in
    mode when kernel_language='CUDA' is given a before_step hook that forces the host path)."""
    q0 = goldens2()['euler_2d__shock_bubble_interaction_q0']
    sc = S.shock_bubble_interaction(q0, tfinal=0.05, device_callbacks=True)
    r = ref_solver(S.shock_bubble_interaction(q0, tfinal=0.05))
    r.run(sc['tfinal'], sc['num_output_times'])
    claw = S.make_claw(sc, kernel_language='CUDA', cuda_precision='fast')
    claw.run()
    assert claw.solver._dev is not None
    q = claw.frames[claw.num_output_times].state.get_q_global()
    assert [s[2] for s in claw.solver.step_log] == [s[2] for s in r.step_log]
    assert np.max(np.abs(q - r.q)) / np.max(np.abs(r.q)) < 1e-10
    claw = S.make_claw(sc, kernel_language='CUDA', cuda_precision='strict')
    claw.solver.before_step = lambda solver, state: None                      # host hook: plug-in mode
    claw.run()
    assert claw.solver._dev is None and claw.solver.fmod is not None
    assert np.array_equal(claw.frames[claw.num_output_times].state.get_q_global(), r.q)
