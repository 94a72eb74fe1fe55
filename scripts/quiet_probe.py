"""Where does the time of the headline kernel go on piecewise-constant data?  Times the sweep kernel of BASELINE config 2
(or 3) alone (claw_time_sweep) on variants of the initial data: 'actual' (shock + bubble), 'uniform' (one state
everywhere: the pure uniform-region path plus the warm-up rows of every march segment), 'shock' (the planar shock only:
one or two heavy strips through every segment, i.e. the tail of the launch).  Kernel switches through the environment
(CLAW_VARIANT, CLAW_TUNE) are read once per process.
    python scripts/quiet_probe.py --config c2 --n 8192 --seg 0 32 64 128"""
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
    ap.add_argument('--config', default='c2')
    ap.add_argument('--n', type=int, default=0)
    ap.add_argument('--seg', type=int, nargs='*', default=[0])
    ap.add_argument('--data', nargs='*', default=['actual', 'uniform', 'shock'])
    ap.add_argument('--reps', type=int, default=5)
    args = ap.parse_args()
    import bench
    from clawpack_b200 import _lib, scenarios as S
    n = args.n or bench.default_n(args.config)
    sc = bench.build_scenario(args.config, n, 1)
    q0 = sc['q']
    for data in args.data:
        q = q0.copy(order='F')
        if data == 'uniform':
            q[...] = q0[:, -1:, -1:]
        elif data == 'shock':
            q[...] = q0[:, :, -1:]
        elif data == 'dense':
            idx = np.indices(q.shape[1:], sparse=True)
            pert = 1.0 + 1e-3 * np.sin(0.37 * idx[0] + 0.11 * idx[1] + 0.5)
            q[0] *= pert
            if args.config == 'c2':
                q[3] *= pert
        sc['q'] = q
        for seg in args.seg:
            claw = S.make_claw(sc, kernel_language='CUDA', cuda_precision='fast', cuda_seg_rows=seg)
            solver, solution = claw.solver, claw.solution
            solver.setup(solution)
            solver.dt = 1e-5 if args.config == 'c2' else 1e-6
            dev = solver._dev
            dev._upload_if_needed(solution.state)
            _lib.check(dev.L.claw_apply_bcs(dev.handle), dev.handle)
            ms = ctypes.c_double()
            _lib.check(dev.L.claw_time_sweep(dev.handle, _lib.dp(dev.delta), float(solver.dt), 2, ctypes.byref(ms)),
                       dev.handle)                                       # warm-up
            _lib.check(dev.L.claw_time_sweep(dev.handle, _lib.dp(dev.delta), float(solver.dt), args.reps, ctypes.byref(ms)),
                       dev.handle)
            linfo = (ctypes.c_int * 5)()
            _lib.check(dev.L.claw_last_launch_info(dev.handle, linfo), dev.handle)
            cells = float(np.prod(q.shape[1:]))
            print(json.dumps({'config': args.config, 'n': n, 'data': data, 'seg_rows': seg, 'grid': [linfo[0], linfo[1]],
                              'tune': os.environ.get('CLAW_TUNE', '0'), 'variant': os.environ.get('CLAW_VARIANT', '0'),
                              'kernel_ms': ms.value,
                              'GBps': bench.BYTES_PER_CELL[args.config] * cells / (ms.value * 1e-3) / 1e9}), flush=True)
            solver.teardown()
            del claw, solver, solution, dev
        del q


if __name__ == '__main__':
    main()
