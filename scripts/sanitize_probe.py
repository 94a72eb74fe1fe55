"""A handful of device-resident runs for compute-sanitizer (memcheck / racecheck / initcheck): the headline Euler kernel
with the uniform-region path and the tile ordering forced on (CLAW_TUNE=32), the shallow-water split kernel, the 3-D
kernels, the device-functor example.  No oracle here -- the parity tests do that; this only has to touch the code paths.
    CLAW_TUNE=32 compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/sanitize_probe.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from clawpack_b200 import scenarios as S  # noqa: E402


def run(sc, nsteps, precision='fast'):
    claw = S.make_claw(sc, kernel_language='CUDA', cuda_precision=precision)
    sol, solver = claw.solution, claw.solver
    solver.setup(sol)
    solver.dt = solver.dt_initial
    for _ in range(nsteps):
        solver.evolve_to_time(sol)
    q = np.array(sol.state.q, copy=True)
    solver.teardown()
    assert np.all(np.isfinite(q))
    return q


def main():
    run(S.shock_bubble(mx=601, my=389), 8)                       # ragged: last strip partly outside, short last segment
    run(S.shock_bubble(mx=1024, my=768), 6)
    run(S.radial_dam_break(mx=500, my=333), 8)
    run(S.acoustics_2d(mx=200, my=200), 8)
    run(S.acoustics_3d(mx=40, my=36, mz=32), 3)
    run(S.shock_bubble(mx=300, my=200), 4, 'strict')
    g2 = np.load(os.path.join(ROOT, 'tests', 'golden', 'reference_goldens2.npz'))
    sc = S.shock_bubble_interaction(g2['euler_2d__shock_bubble_interaction_q0'], tfinal=0.02, device_callbacks=True)
    claw = S.make_claw(sc, kernel_language='CUDA', cuda_precision='fast')
    claw.run()
    print('sanitize probe done')


if __name__ == '__main__':
    main()
