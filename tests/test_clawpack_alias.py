"""The ``clawpack`` import namespace (clawpack/__init__.py of this repo): reference scripts import
``from clawpack import pyclaw, riemann`` / ``import clawpack.petclaw`` (reference clawpack/__init__.py:3-24).

* CPU: the alias resolves to the very same module objects as ``clawpack_b200``; and -- in this container, where
  /root/reference exists -- the reference's own example files are executed UNMODIFIED through the alias with the oracle
  plug-in as ``solver.fmod`` (the reference's plug-in point, classic/solver.py:208-211) and land on the vectors the
  reference stored for them.  This pins the host-side API surface the examples touch (Controller, Solution, State,
  Domain, BC, limiters, custom BCs, source terms, problem_data -> cparam) without a GPU.
* GPU: the same flow, written the way a reference script is written, runs on the CUDA path through the alias."""
import importlib.util
import os
import sys

import numpy as np
import pytest

from refutil import goldens

EXAMPLES = '/root/reference/pyclaw/examples'
needs_reference = pytest.mark.skipif(not os.path.isdir(EXAMPLES), reason='the reference tree is only present in the build container')


def test_alias_is_the_same_package():
    import clawpack
    import clawpack_b200
    from clawpack import pyclaw, riemann
    import clawpack.petclaw as petclaw
    from clawpack.pyclaw.classic.solver import ClawSolver2D
    from clawpack.riemann.euler_4wave_2D_constants import num_eqn
    assert pyclaw is clawpack_b200.pyclaw and riemann is clawpack_b200.riemann
    assert petclaw is sys.modules['clawpack_b200.petclaw']
    assert ClawSolver2D is clawpack_b200.pyclaw.ClawSolver2D and num_eqn == 4
    assert pyclaw.ClawSolver2D(riemann.euler_4wave_2D).kernel_language == 'CUDA'
    for missing in ('sharpclaw', 'visclaw', 'amrclaw', 'geoclaw'):
        with pytest.raises(ImportError):
            importlib.import_module('clawpack.' + missing)
    assert clawpack.__version__.startswith('5.')


def _load_example(rel):
    path = os.path.join(EXAMPLES, rel)
    spec = importlib.util.spec_from_file_location('refex_' + os.path.basename(rel)[:-3], path)
    mod = importlib.util.module_from_spec(spec)
    cwd = os.getcwd()
    os.chdir(os.path.dirname(path))          # some examples read files next to themselves at import time
    try:
        spec.loader.exec_module(mod)
    finally:
        os.chdir(cwd)
    return mod


def _run_with_oracle(claw):
    from oracle_fmod import OracleFmod
    claw.solver.fmod = OracleFmod()           # the reference's plug-in point; no GPU involved
    claw.output_format = None
    claw.keep_copy = True
    claw.run()
    return claw


@needs_reference
def test_reference_acoustics_2d_example_unmodified():
    """examples/acoustics_2d_homogeneous/acoustics_2d.py + test_2d_acoustics.py:14 (relative L2 < 1e-3)."""
    ex = _load_example('acoustics_2d_homogeneous/acoustics_2d.py')
    claw = _run_with_oracle(ex.setup(outdir=None, disable_output=True))
    p = claw.frames[-1].q[0]
    e = goldens()['acoustics_2d__verify_classic']
    assert np.linalg.norm(p - e) / np.linalg.norm(e) < 1e-3


@needs_reference
def test_reference_radial_dam_break_example_unmodified():
    """examples/shallow_2d/radial_dam_break.py + test_shallow2d.py (L1 < 1e-5 against expected_sols.npy)."""
    ex = _load_example('shallow_2d/radial_dam_break.py')
    claw = _run_with_oracle(ex.setup(outdir=None))
    h = claw.frames[-1].q[0]
    e = goldens()['shallow_2d__radialdambreak_classic_roe'].reshape(h.shape)
    dx, dy = claw.solution.domain.grid.delta
    assert dx * dy * np.sum(np.abs(h - e)) < 1e-5


@needs_reference
def test_reference_advection_2d_example_unmodified():
    """examples/advection_2d/advection_2d.py + test_2d_advection.py (dimension-split classic run)."""
    ex = _load_example('advection_2d/advection_2d.py')
    claw = _run_with_oracle(ex.setup(outdir=None))
    q = claw.frames[-1].q[0]
    e = goldens()['advection_2d__classic_dim_split'].reshape(q.shape)
    dx, dy = claw.solution.domain.grid.delta
    assert dx * dy * np.sum(np.abs(q - e)) < 1e-6


@needs_reference
def test_reference_shock_bubble_example_unmodified():
    """examples/euler_2d/shock_bubble_interaction.py (euler_5wave_2D, radial source term, custom inflow BC, scipy
    sub-cell quadrature of the bubble) + test_euler2d.py:10-26."""
    ex = _load_example('euler_2d/shock_bubble_interaction.py')
    claw = ex.setup(mx=160, my=40, tfinal=0.2, num_output_times=1, disable_output=True)      # test_euler2d.py:37
    _run_with_oracle(claw)
    rho = claw.frames[claw.num_output_times].state.get_q_global()[0].reshape([-1])
    e = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'reference_goldens2.npz'))['euler_2d__shock_bubble_interaction']
    dx, dy = claw.solution.domain.grid.delta
    assert dx * dy * np.sum(np.abs(rho - e)) < 1e-6


@pytest.mark.gpu
def test_reference_style_script_runs_on_cuda_through_the_alias():
    """A script written like examples/acoustics_2d_homogeneous/acoustics_2d.py:31-99, line for line in the reference's
    API and import style, on the CUDA path: must reproduce the reference's stored vector."""
    from clawpack import riemann
    from clawpack import pyclaw
    solver = pyclaw.ClawSolver2D(riemann.acoustics_2D)
    solver.dimensional_split = True
    solver.cfl_max = 0.5
    solver.cfl_desired = 0.45
    solver.limiters = pyclaw.limiters.tvd.MC
    solver.bc_lower[0] = pyclaw.BC.extrap
    solver.bc_upper[0] = pyclaw.BC.extrap
    solver.bc_lower[1] = pyclaw.BC.extrap
    solver.bc_upper[1] = pyclaw.BC.extrap
    mx, my = 100, 100
    x = pyclaw.Dimension(-1.0, 1.0, mx, name='x')
    y = pyclaw.Dimension(-1.0, 1.0, my, name='y')
    domain = pyclaw.Domain([x, y])
    state = pyclaw.State(domain, solver.num_eqn)
    rho, bulk = 1.0, 4.0
    cc = np.sqrt(bulk / rho)
    zz = rho * cc
    state.problem_data['rho'] = rho
    state.problem_data['bulk'] = bulk
    state.problem_data['zz'] = zz
    state.problem_data['cc'] = cc
    solver.dt_initial = np.min(domain.grid.delta) / state.problem_data['cc'] * solver.cfl_desired
    X, Y = state.p_centers
    r = np.sqrt(X ** 2 + Y ** 2)
    width, rad = 0.2, 0.5
    state.q[0, :, :] = (np.abs(r - rad) <= width) * (1. + np.cos(np.pi * (r - rad) / width))
    state.q[1, :, :] = 0.
    state.q[2, :, :] = 0.
    claw = pyclaw.Controller()
    claw.keep_copy = True
    claw.output_format = None
    claw.solution = pyclaw.Solution(state, domain)
    claw.solver = solver
    claw.num_output_times = 10
    claw.tfinal = 0.12
    claw.run()
    assert solver._dev is None or solver.kernel_language == 'CUDA'
    p = claw.frames[-1].q[0]
    e = goldens()['acoustics_2d__verify_classic']
    assert np.linalg.norm(p - e) / np.linalg.norm(e) < 1e-3
