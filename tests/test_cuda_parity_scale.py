"""GPU parity at BASELINE scale, in the BENCHMARKED arithmetic mode (fast), against the threaded oracle.

The small-grid tests (test_cuda_parity*.py) pin every branch; these pin what bench.py actually runs: the launch
geometry of large grids (hundreds of strips x tens of row segments, 256-row march segments, the cp.async row ring over
thousands of march steps) and the accumulated rounding of the fast mode over 100 steps.  Bars (BASELINE.json north_star,
tighter than pyclaw/examples/euler_2d/test_euler2d.py:10-26): relative L-inf <= 1e-12 after one step, <= 1e-10 after
100 steps, identical accept/reject sequences -- and strict arithmetic stays bit-exact at these sizes too.

Oracle cost (threads = all host cores): Euler 2048^2 ~0.2 s per step, 8192^2 ~3 s; shallow water 4096^2 ~0.3 s;
3-D acoustics 128^3 (serial step3) ~2 s per step."""
import os

import numpy as np
import pytest

from clawpack_b200 import scenarios as S
from refutil import ref_solver

pytestmark = pytest.mark.gpu


def _threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _cuda(sc, nsteps, precision):
    claw = S.make_claw(sc, kernel_language='CUDA', cuda_precision=precision)
    sol, solver = claw.solution, claw.solver
    solver.setup(sol)
    solver.dt = solver.dt_initial
    for _ in range(nsteps):
        solver.evolve_to_time(sol)
    q = np.array(sol.state.q, copy=True)
    log = list(solver.step_log)
    solver.teardown()
    return q, log


def _ref(sc, nsteps, nthreads=None):
    r = ref_solver(sc, nthreads=nthreads or _threads())
    for _ in range(nsteps):
        r.evolve_to_time()
    return r.q, list(r.step_log)


def _rel(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def _check(sc, nsteps, tol, nthreads=None, strict_too=True):
    qr, logr = _ref(sc, nsteps, nthreads)
    q, log = _cuda(sc, nsteps, 'fast')
    assert [s[2] for s in log] == [s[2] for s in logr], 'accept/reject sequence differs'
    assert np.allclose([s[0] for s in log], [s[0] for s in logr], rtol=1e-11, atol=0), 'dt sequence differs'
    err = _rel(q, qr)
    assert err <= tol, 'relative L-inf %.3g > %.1g' % (err, tol)
    if strict_too:
        q, log = _cuda(sc, nsteps, 'strict')
        assert log == logr
        assert np.array_equal(q, qr), 'strict mode: max abs diff %g' % np.max(np.abs(q - qr))


def test_c2_euler_2048_first_accepted_step():
    """BASELINE config 2 scheme at 2048^2: the deliberate reject (dt_initial = 0.1) + one accepted step."""
    _check(S.shock_bubble(mx=2048, my=2048), 2, 1e-12)


def test_c2_euler_2048_100_steps():
    """100 accepted steps at 2048^2 (74 strips x 8 segments of 256 rows per launch), fast and strict."""
    _check(S.shock_bubble(mx=2048, my=2048), 101, 1e-10)


def test_c2_euler_8192_single_step():
    """The bench.py grid itself: 8192^2, reject + one accepted step, fast mode (the strict run is left to 2048^2)."""
    _check(S.shock_bubble(mx=8192, my=8192), 2, 1e-12, strict_too=False)


def test_c2_euler_ragged_large():
    """Sizes that are not multiples of the strip width (28) or the segment length: 3001 x 1777."""
    _check(S.shock_bubble(mx=3001, my=1777), 4, 1e-12)


def test_c3_shallow_4096():
    """BASELINE config 3 scheme (dimension-split, MC) at 4096^2: reject + 3 accepted steps; and 100 steps at 1024^2."""
    _check(S.radial_dam_break(mx=4096, my=4096), 4, 1e-12)
    _check(S.radial_dam_break(mx=1024, my=1024), 101, 1e-10, strict_too=False)


def test_c4_acoustics3d_128():
    """BASELINE config 4 scheme (unsplit, transverse_waves=22, MC; the kernels specialised for this method set) at
    128^3 against the serial step3 oracle: reject + 2 accepted steps."""
    _check(S.acoustics_3d(mx=128, my=128, mz=128), 3, 1e-12, nthreads=1)


def test_c1_acoustics_200_through_graph_batches():
    """BASELINE config 1 (200^2) run to an end time, so that the attempts go through the CUDA-graph batches of the
    device-resident step control: same step log and bits as the oracle's host loop."""
    sc = S.acoustics_2d(mx=200, my=200)
    r = ref_solver(sc)
    r.dt = r.dt_initial
    for tend in (0.03, 0.0601, 0.12):
        r.evolve_to_time(tend)
    claw = S.make_claw(sc, kernel_language='CUDA', cuda_precision='strict')
    claw.solver.setup(claw.solution)
    claw.solver.dt = claw.solver.dt_initial
    for tend in (0.03, 0.0601, 0.12):
        claw.solver.evolve_to_time(claw.solution, tend)
    assert claw.solver.step_log == r.step_log
    assert claw.solution.t == r.t
    assert np.array_equal(claw.solution.state.q, r.q)
    assert claw.solver.status['numsteps'] == sum(1 for s in r.step_log if s[2])


def test_tile_ordering_forced_on_mid_size_grids():
    """The tile ordering of the uniform-region kernels (sweep2_ws.cuh, HotTiles) only switches on for launches of four
    waves of CTAs or more (8192^2-class grids).  CLAW_TUNE=32 forces it on any launch, so that 100-step runs against the
    oracle exercise it with the lists of every step feeding the next: Euler 1024 x 768 (unsplit) and shallow water 768^2
    (dimension-split), fast mode, the tolerance of the north star.  The switch is read once per process: subprocess."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "from test_cuda_parity_scale import _check, S\n"
            "_check(S.shock_bubble(mx=1024, my=768), 101, 1e-10, strict_too=False)\n"
            "_check(S.radial_dam_break(mx=768, my=768), 101, 1e-10, strict_too=False)\n"
            "print('tile ordering parity ok')\n") % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                     os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CLAW_TUNE='32')
    p = subprocess.run([sys.executable, '-c', code], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                       timeout=900)
    assert p.returncode == 0 and 'tile ordering parity ok' in p.stdout, p.stdout[-3000:]


def test_dense_data_switch_on_mid_size_grids():
    """The dense-data switch (clawcuda.cu: dense_update): on data without uniform regions the fast-path kernels report
    (almost) every tile heavy and the stepping moves to the kernels without the fast path; on the BASELINE data it stays
    where it is.  Both ways the run stays within the north-star tolerance of the oracle, through the switch and the graph
    re-capture.  CLAW_TUNE=32 lets launches of mid-size grids report their tiles (read once per process: subprocess)."""
    import subprocess
    import sys
    code = '''
import sys, ctypes
sys.path.insert(0, %r); sys.path.insert(0, %r)
import numpy as np
from test_cuda_parity_scale import _ref, _rel, S
from clawpack_b200 import _lib

def run(sc, nsteps):
    claw = S.make_claw(sc, kernel_language='CUDA', cuda_precision='fast')
    sol, solver = claw.solution, claw.solver
    solver.setup(sol)
    solver.dt = solver.dt_initial
    states = []
    st = (ctypes.c_int * 3)()
    for _ in range(nsteps):
        solver.evolve_to_time(sol)
        _ = sol.t                                  # (read-back: the host does not run ahead of the GPU in this test)
        _lib.check(solver._dev.L.claw_dense_state(solver._dev.handle, st), solver._dev.handle)
        states.append(st[0])
    q = np.array(sol.state.q, copy=True)
    log = list(solver.step_log)
    solver.teardown()
    return q, log, states, st[2]

def dense(sc):
    q = sc['q']
    idx = np.indices(q.shape[1:], sparse=True)
    pert = 1.0 + 1e-3 * np.sin(0.37 * idx[0] + 0.11 * idx[1] + 0.5)
    q[0] *= pert
    q[-1] *= pert if q.shape[0] == 4 else 1.0
    return sc

for make, n in ((lambda: S.shock_bubble(mx=1024, my=768), 40), (lambda: S.radial_dam_break(mx=768, my=768), 40)):
    for is_dense in (False, True):
        sc = dense(make()) if is_dense else make()
        q, log, states, enabled = run(sc, n)
        assert enabled == 1
        qr, logr = _ref(sc, n)
        assert [s[2] for s in log] == [s[2] for s in logr]
        assert np.allclose([s[0] for s in log], [s[0] for s in logr], rtol=1e-11, atol=0)
        assert _rel(q, qr) <= 1e-10, _rel(q, qr)
        if is_dense:
            assert states[0] == 0 and 1 in states[:12] and sum(states[12:]) >= len(states[12:]) // 2, states   # switched within a few
            # batches and stays there but for the probes (every 4th batch in this test)
        else:
            assert not any(states), states
# the way back: dense data first, then the state is replaced by the sparse BASELINE data (a host upload); the next probe
# (every 4th batch here, CLAW_DENSE_PROBE) finds few heavy tiles and the stepping returns to the fast-path kernels
sc = dense(S.shock_bubble(mx=1024, my=768))
claw = S.make_claw(sc, kernel_language='CUDA', cuda_precision='fast')
sol, solver = claw.solution, claw.solver
solver.setup(sol)
solver.dt = solver.dt_initial
st = (ctypes.c_int * 3)()
seen = []
for _ in range(14):
    solver.evolve_to_time(sol)
    _ = sol.t
    _lib.check(solver._dev.L.claw_dense_state(solver._dev.handle, st), solver._dev.handle)
    seen.append(st[0])
assert seen[-1] == 1 or seen[-2] == 1, seen              # dense (a probe batch may be the last one)
sol.state.q = S.shock_bubble(mx=1024, my=768)['q']
after = []
for _ in range(16):
    solver.evolve_to_time(sol)
    _ = sol.t
    _lib.check(solver._dev.L.claw_dense_state(solver._dev.handle, st), solver._dev.handle)
    after.append(st[0])
assert after[-1] == 0 and after[-2] == 0 and after[-3] == 0, (seen, after)
assert np.isfinite(sol.state.q).all()
solver.teardown()
print('dense switch ok')
''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CLAW_TUNE='32', CLAW_DENSE_PROBE='4')
    p = subprocess.run([sys.executable, '-c', code], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                       timeout=900)
    assert p.returncode == 0 and 'dense switch ok' in p.stdout, p.stdout[-3000:]
