"""Load balance of a slab-decomposed strong-scaling run, measured on ONE GPU: the sweep kernel of config 3 (or 4) on the
slab that rank r of `world` ranks would own, run as a standalone problem (the neighbours' ghost rows are replaced by the
slab's own boundary conditions, which does not change the cost), timed on the initial data and after a few steps.  With
the uniform-region fast path the cost of a slab depends on how much of the fronts it holds; the step time of the N-rank
run is the slowest rank's.
    python scripts/slab_probe.py --config c3 --world 8 --ranks 0 3 --steps 20"""
import argparse
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--config', default='c3')
    ap.add_argument('--n', type=int, default=0)
    ap.add_argument('--world', type=int, default=8)
    ap.add_argument('--ranks', type=int, nargs='*', default=[0, 3])
    ap.add_argument('--steps', type=int, default=20)
    args = ap.parse_args()
    import bench
    from clawpack_b200 import _lib, scenarios as S
    n = args.n or bench.default_n(args.config)
    rows = n // args.world
    for r in args.ranks:
        sc = bench.build_scenario(args.config, n, 1, slab=(r * rows, (r + 1) * rows))
        last = len(sc['num_cells']) - 1
        lo, up = sc['lower'][last], sc['upper'][last]
        d = (up - lo) / n
        sc['num_cells'][last] = rows
        sc['lower'][last] = lo + r * rows * d
        sc['upper'][last] = lo + (r + 1) * rows * d
        sc['slab'] = None
        claw = S.make_claw(sc, kernel_language='CUDA', cuda_precision='fast')
        solver, solution = claw.solver, claw.solution
        solver.setup(solution)
        solver.dt = solver.dt_initial
        dev = solver._dev
        out = {'config': args.config, 'n': n, 'world': args.world, 'rank': r, 'cells': int(np.prod(sc['num_cells']))}
        for tag, nsteps in (('initial', 1), ('after_%d_steps' % args.steps, args.steps)):
            for _ in range(nsteps):
                solver.evolve_to_time(solution)
            ms = ctypes.c_double()
            _lib.check(dev.L.claw_time_sweep(dev.handle, _lib.dp(dev.delta), float(solver.dt), 5, ctypes.byref(ms)), dev.handle)
            out['kernel_ms_' + tag] = ms.value
        print(json.dumps(out), flush=True)
        solver.teardown()
        del claw, solver, solution, dev


if __name__ == '__main__':
    main()
