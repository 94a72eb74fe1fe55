"""The sweep KERNEL source executed on the CPU (tests/host_functors/cuda_emul.h: every CUDA thread a fiber, resumed
round-robin from synchronisation point to synchronisation point) against the oracle, bit for bit in strict arithmetic.

tests/test_host_functors.py checks the Riemann-solver functors; this file checks the kernels AROUND them -- thread/cell
mapping, shared-memory and shuffle exchange, tile edges, aux addressing, the f-wave branch, the block-wide CFL
reduction -- on small grids, so that a kernel change can be validated before GPU time is spent.  The `-m gpu` tests
remain the proof on the hardware."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from clawpack_b200 import scenarios as S
from refutil import ref_solver

HERE = os.path.dirname(os.path.abspath(__file__))
CXX = '/usr/bin/g++' if os.path.exists('/usr/bin/g++') else 'g++'


def _compile(tmp_path_factory, name):
    out = str(tmp_path_factory.mktemp(name) / ('lib%s.so' % name))
    subprocess.run([CXX, '-O1', '-ffp-contract=off', '-std=c++17', '-fPIC', '-shared', '-Wno-unknown-pragmas', '-o', out,
                    os.path.join(HERE, 'host_functors', name + '.cpp')], check=True)
    return ctypes.CDLL(out)


def _cparam(rp):
    cp = getattr(rp, 'cparam', None)
    vals = cp.values() if cp is not None else []
    return np.array(list(vals) + [0.0], dtype=np.float64)


class EmulatedFmod1D(object):
    """``fmod.step1`` of ClawSolver1D backed by the emulated sweep1 kernel (argument list of claw_step1_host)."""

    def __init__(self, lib):
        self.L = lib
        i, d, vp = ctypes.c_int, ctypes.c_double, ctypes.c_void_p
        ip, dp = ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double)
        lib.emul_step1_host.argtypes = [i] * 5 + [vp, vp, d, d, ip, ip, i, i, dp, i, dp]
        lib.emul_step1_host.restype = i

    def step1(self, num_ghost, mx, q, aux, dx, dt, method, mthlim, use_fwave, rp1):
        q = np.asfortranarray(q, dtype=np.float64)
        aux = None if aux is None or aux.size == 0 else np.asfortranarray(aux, dtype=np.float64)
        m = np.ascontiguousarray(method, dtype=np.intc)
        ml = np.array(mthlim, dtype=np.intc)
        cp = _cparam(rp1)
        cfl = ctypes.c_double(0.0)
        rc = self.L.emul_step1_host(q.shape[0], len(ml), 0 if aux is None else aux.shape[0], num_ghost, mx,
                                    q.ctypes.data_as(ctypes.c_void_p),
                                    ctypes.c_void_p(0) if aux is None else aux.ctypes.data_as(ctypes.c_void_p),
                                    dx, dt, m.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
                                    ml.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), int(bool(use_fwave)), int(rp1.rp_id),
                                    cp.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), len(cp) - 1, ctypes.byref(cfl))
        assert rc == 0, rc
        return q, cfl.value


@pytest.fixture(scope='module')
def emul1(tmp_path_factory):
    return EmulatedFmod1D(_compile(tmp_path_factory, 'sweep1_emul'))


CASES_1D = {
    # the solvers of the first batch (already proven on the GPU): they validate the emulation itself
    'advection': lambda: S.advection_1d(nx=300, tfinal=0.3),
    'acoustics_tile_edge': lambda: dict(S.acoustics_1d(num_cells=124 * 2 - 4 + 1), limiters=[S.SUPERBEE, S.MINMOD]),
    'burgers_first_order': lambda: dict(S.burgers_1d(mx=123), order=1),
    'dam_break_roe': lambda: S.dam_break_1d(mx=250),
    'dam_break_hlle_tiny': lambda: S.dam_break_1d(mx=3, riemann_solver='hlle'),
    # rp_id 26..36: not yet run on a GPU
    'traffic': lambda: S.traffic_1d(mx=260),
    'advection_color_aux_periodic': lambda: S.advection_1d_variable(mx=124 * 2 - 4 + 1),
    'acoustics_variable_aux': lambda: S.acoustics_1d_variable(mx=200),
    'sill_fwave_aux': lambda: S.sill_1d(mx=300),
    'sill_tiny': lambda: S.sill_1d(mx=5),
    'shocktube_hlle': lambda: S.shocktube_1d(mx=200),
    'blast_roe_walls': lambda: S.blast_1d(mx=200),
    'stegoton_linear_custom_bc': lambda: S.stegoton_1d(xupper=20.0, exponential=False),
    'stegoton_exponential': lambda: S.stegoton_1d(xupper=20.0),
    'cubic': lambda: S.cubic_1d(mx=150),
    'traffic_vc': lambda: S.traffic_vc_1d(mx=150),
    'traffic_vc_fwave': lambda: S.traffic_vc_1d(mx=150, kind='fwave'),
    'traffic_vc_tracer': lambda: S.traffic_vc_1d(mx=131, kind='tracer'),
}


@pytest.mark.parametrize('name', sorted(CASES_1D))
def test_emulated_sweep1_kernel_bit_exact_vs_oracle(emul1, name):
    sc = CASES_1D[name]()
    claw = S.make_claw(sc, kernel_language='Fortran', fmod=emul1)
    sol, solver = claw.solution, claw.solver
    solver.setup(sol)
    solver.dt = solver.dt_initial
    r = ref_solver(sc)
    for _ in range(25):
        solver.evolve_to_time(sol)
        r.evolve_to_time()
    assert solver.step_log == r.step_log, (solver.step_log[:3], r.step_log[:3])
    assert np.array_equal(sol.state.q, r.q), float(np.max(np.abs(sol.state.q - r.q)))
