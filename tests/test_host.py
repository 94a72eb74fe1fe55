"""CPU tests of the PyClaw-shaped host layer and of the C ABI surface (no GPU compute)."""
import ctypes
import os
import re

import numpy as np
import pytest

from clawpack_b200 import pyclaw, riemann, scenarios as S
from oracle_fmod import OracleFmod
from refutil import goldens, ref_solver

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_loop_with_plugin_matches_oracle_loop(oracle_lib):
    """kernel_language='Fortran' + an injected fmod (here the test-only oracle plug-in) drives exactly the
    reference's step sequence: identical q and identical (dt, cfl, accepted) logs."""
    for sc in (S.advection_2d(dimensional_split=False, transverse_waves=2),
               S.quadrants(mx=40, my=40, tfinal=0.2), S.radial_dam_break(mx=40, my=40, tfinal=0.5)):
        claw = S.make_claw(sc, kernel_language='Fortran', fmod=OracleFmod())
        claw.run()
        r = ref_solver(sc)
        r.run(sc['tfinal'], sc['num_output_times'])
        assert np.array_equal(claw.frames[-1].q, r.q)
        assert claw.solver.step_log == r.step_log
        assert any(not s[2] for s in r.step_log) or sc['rp'] == 'shallow_roe_with_efix_2D'


def test_host_loop_reference_golden(oracle_lib):
    sc = S.quadrants()
    claw = S.make_claw(sc, kernel_language='Fortran', fmod=OracleFmod())
    claw.run()
    rho = claw.frames[claw.num_output_times].state.get_q_global()[0].reshape(-1)
    dx, dy = claw.solution.domain.grid.delta
    assert dx * dy * np.sum(np.abs(rho - goldens()['euler_2d__quadrants_roe'])) < 1e-6


def test_host_loop_3d_with_plugin_matches_oracle_loop(oracle_lib):
    """ClawSolver3D (classic/solver.py:648-706) with the oracle plug-in: unsplit (step3) with aux arrays and
    wall/periodic BCs, dimension-split (3 aliased step3ds calls), and BC.custom + wall on euler_3D."""
    G = goldens()
    cases = [S.acoustics_3d(mx=12, my=10, mz=9, tfinal=0.3, num_output_times=2),
             S.acoustics_3d(mx=16, my=4, mz=4, heterogeneous=False, tfinal=0.3, num_output_times=2),
             S.euler_3d_shocktube(mx=4, my=5, mz=24, tfinal=0.05, num_output_times=1),
             S.euler_3d_shock_bubble(G['euler_3d__shock_bubble_q0'][:, :20, :6, :6].copy(order='F'),
                                     G['euler_3d__shock_bubble_consts'], tfinal=0.02, num_output_times=1)]
    for sc in cases:
        claw = S.make_claw(sc, kernel_language='Fortran', fmod=OracleFmod())
        claw.run()
        r = ref_solver(sc)
        r.run(sc['tfinal'], sc['num_output_times'])
        assert np.array_equal(claw.frames[-1].q, r.q)
        assert claw.solver.step_log == r.step_log


def test_host_loop_3d_reference_golden(oracle_lib):
    sc = S.acoustics_3d(mx=15, my=15, mz=15)
    claw = S.make_claw(sc, kernel_language='Fortran', fmod=OracleFmod())
    claw.run()
    p = claw.frames[-1].q[0].reshape(-1)
    assert np.max(np.abs(p - goldens()['acoustics_3d__verify_classic_heterogeneous'])) < 1e-13


def test_api_surface():
    solver = pyclaw.ClawSolver2D(riemann.euler_4wave_2D)
    assert (solver.num_eqn, solver.num_waves, solver.num_ghost) == (4, 4, 2)
    assert solver.kernel_language == 'CUDA' and solver.dimensional_split is True
    assert solver.transverse_waves == pyclaw.ClawSolver2D.trans_inc == 1
    assert solver.limiters == pyclaw.limiters.tvd.minmod == 1 and pyclaw.limiters.tvd.MC == 4
    assert (pyclaw.BC.custom, pyclaw.BC.extrap, pyclaw.BC.periodic, pyclaw.BC.wall) == (0, 1, 2, 3)
    with pytest.raises(TypeError):          # attribute set is frozen (solver.py:139-144)
        solver.no_such_option = 1
    solver.all_bcs = pyclaw.BC.extrap
    assert solver.bc_lower == [1, 1] and solver.bc_upper == [1, 1]
    s3 = pyclaw.ClawSolver3D(riemann.vc_acoustics_3D)
    assert s3.transverse_waves == 22 and s3.num_eqn == 4 and s3.num_waves == 2
    dom = pyclaw.Domain([0., 0.], [1., 1.], [4, 3])
    sol = pyclaw.Solution(4, dom)
    assert sol.q.shape == (4, 4, 3) and sol.q.flags['F_CONTIGUOUS']
    assert np.allclose(dom.grid.p_centers[0][:, 0], [0.125, 0.375, 0.625, 0.875])
    st = pyclaw.State(dom, 3)
    st.problem_data['grav'] = 1.0
    st.set_cparam(riemann.shallow_roe_with_efix_2D)
    assert riemann.shallow_roe_with_efix_2D.cparam.grav == 1.0
    with pytest.raises(Exception):          # missing cparam member raises (state.py:212-215)
        pyclaw.State(dom, 4).set_cparam(riemann.euler_4wave_2D)


def test_cuda_path_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    from clawpack_b200._lib import ClawCudaError
    claw = S.make_claw(S.advection_2d(), kernel_language='CUDA')
    with pytest.raises(ClawCudaError):
        claw.run()


def test_fortran_path_needs_reference_module():
    claw = S.make_claw(S.advection_2d(), kernel_language='Fortran')
    with pytest.raises(ImportError):
        claw.run()


def test_abi_exports_every_declared_symbol():
    """libclawcuda.so loads on a CPU-only box and exports every function include/clawcuda.h declares."""
    from clawpack_b200 import _lib
    L = _lib.lib()
    header = open(os.path.join(ROOT, 'include', 'clawcuda.h')).read()
    declared = set(re.findall(r'\b(claw_[a-z0-9_]+)\s*\(', header))
    declared.discard('claw_solver')
    assert len(declared) >= 35
    for name in sorted(declared):
        assert hasattr(L, name), 'missing export ' + name
    assert set(_lib.EXPORTS) == declared, set(_lib.EXPORTS) ^ declared
    assert b'clawcuda' in L.claw_version()
    assert L.claw_device_count() >= 0
    # error paths never abort: bad arguments give a code + message
    h = ctypes.c_void_p()
    n = (ctypes.c_int * 3)(8, 8, 1)
    rc = L.claw_create(ctypes.byref(h), 4, 3, 4, 4, 0, n, 2, 0, None)
    assert rc != 0 and b'ndim' in L.claw_last_error(None)
    rc = L.claw_create(ctypes.byref(h), 2, 3, 3, 4, 0, n, 2, 0, None)
    assert rc != 0 and b'expects' in L.claw_last_error(None)


def test_slab_partition():
    from clawpack_b200.petclaw.util import slab_range
    for n, p in ((8192, 8), (100, 3), (17, 4)):
        parts = [slab_range(n, p, r) for r in range(p)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        sizes = [b - a for a, b in parts]
        assert max(sizes) - min(sizes) <= 1
