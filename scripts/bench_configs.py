#!/usr/bin/env python
"""Throughput + roofline of the OTHER BASELINE.json configurations (bench.py measures configs[1]/[4]):

    c1  acoustics_2D, dimension-split, MC, 200x200            (configs[0]; tiny: launch-latency bound)
    c3  shallow_roe_with_efix_2D, dimension-split, MC, 16384^2 (configs[2], 1-GPU leg)
    c4  vc_acoustics_3D, unsplit transverse_waves=22, MC, 512^3 (configs[3], 1-GPU leg; --n3 to shrink)
    c4g the same scheme on dense data (aux and state varying everywhere: no limiter short-cuts)
    c4s c4 dimension-split

One JSON line per config on stdout: cell-updates/s through the PyClaw API (device-resident), the sweep
kernels alone (claw_time_sweep), exact per-cell operation counts from the counting oracle and the
resulting FP64-issue / HBM roofline fractions.  Not part of the driver contract; results are
copied into profiles/ and DESIGN.md.

    python scripts/bench_configs.py --configs c1,c3,c4 --steps 10
"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

W_DIV, W_SQRT = 8.0, 9.0      # FP64-pipe issues of the fast-path division / square root (see bench.py)


def ops_of(sc_small, nsteps=4):
    from oracle import pyclaw_ref
    from refutil import ref_solver
    r = ref_solver(sc_small)
    for _ in range(nsteps):
        r.evolve_to_time()
    ops = pyclaw_ref.count_ops(r, r.dt)
    rep = ops.pop('repeated')
    ops['reference_repeats_flops'] = rep['add'] + rep['mul'] + rep['div'] + rep['sqrt']
    ops['flops'] = ops['add'] + ops['mul'] + ops['div'] + ops['sqrt']
    ops['fp64_issues'] = ops['add'] + ops['mul'] + ops['cmp'] + W_DIV * ops['div'] + W_SQRT * ops['sqrt']
    return ops


def run(name, sc, sc_small, bytes_per_cell, steps, warmup, precision, peak_gdfma, hbm_peak):
    import torch
    from clawpack_b200 import _lib, scenarios as S
    t0 = time.time()
    claw = S.make_claw(sc, kernel_language='CUDA', cuda_precision=precision)
    solver, solution = claw.solver, claw.solution
    solver.setup(solution)
    solver.dt = solver.dt_initial
    dev = solver._dev
    for _ in range(warmup):
        solver.evolve_to_time(solution)
    torch.cuda.synchronize()
    l0 = dev.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        solver.evolve_to_time(solution)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / steps
    launches = dev.launch_count() - l0
    avg = ctypes.c_double()
    _lib.check(dev.L.claw_time_sweep(dev.handle, _lib.dp(dev.delta), float(solver.dt), 3, ctypes.byref(avg)), dev.handle)
    cells = int(np.prod(sc['num_cells']))
    ops = ops_of(sc_small)
    log = solver.step_log[-steps:]
    issues = ops['fp64_issues'] * cells / (avg.value * 1e-3)
    gbs = bytes_per_cell * cells / (avg.value * 1e-3) / 1e9
    t_f, t_h = ops['fp64_issues'] * cells / (peak_gdfma * 1e9), bytes_per_cell * cells / (hbm_peak * 1e9)
    out = {'config': name, 'workload': '%s %s, %s, limiters=%s, fp64, arithmetic=%s' % (
               sc['rp'], 'x'.join(map(str, sc['num_cells'])),
               'dimension-split' if sc['dimensional_split'] else 'unsplit transverse_waves=%s' % sc['transverse_waves'],
               sc['limiters'], precision),
           'cell_updates_per_s': cells / (ms * 1e-3), 'ms_per_step': ms, 'sweep_kernels_ms': avg.value,
           'gpu_launches_per_step': launches / steps,
           'accepted': sum(1 for s in log if s[2]), 'rejected': sum(1 for s in log if not s[2]),
           'algorithmic_per_cell': {k: round(v, 2) for k, v in ops.items()},
           'roofline': {'bound': 'fp64' if t_f > t_h else 'hbm',
                        'fp64': {'achieved_T_issues_s': issues / 1e12, 'peak': peak_gdfma / 1e3,
                                 'frac': issues / 1e9 / peak_gdfma},
                        'hbm': {'achieved_GBs': gbs, 'peak': hbm_peak, 'frac': gbs / hbm_peak,
                                'bytes_per_cell': bytes_per_cell}},
           'setup_s': round(time.time() - t0, 1)}
    solver.teardown()
    del claw, solver, solution, dev
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--configs', default='c1,c3,c4,c4s')
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--precision', default='fast')
    ap.add_argument('--n2', type=int, default=16384)
    ap.add_argument('--n3', type=int, default=512)
    args = ap.parse_args()
    from clawpack_b200 import _lib, scenarios as S
    L = _lib.lib()
    peak = ctypes.c_double()
    _lib.check(L.claw_measure_fp64_peak(0, ctypes.byref(peak)))
    hbm_peak = 6650.0
    pf = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(pf):
        hbm_peak = float(json.load(open(pf))['hbm_gbs'])
    n2, n3 = args.n2, args.n3
    table = {
        'c1': lambda: (S.acoustics_2d(mx=200, my=200), S.acoustics_2d(mx=200, my=200), 48),
        'c3': lambda: (S.radial_dam_break(mx=n2, my=n2), S.radial_dam_break(mx=256, my=256), 48),
        'c4': lambda: (S.acoustics_3d(mx=n3, my=n3, mz=n3), S.acoustics_3d(mx=48, my=48, mz=48), 80),
        'c4g': lambda: (S.acoustics_3d_general(mx=n3, my=n3, mz=n3), S.acoustics_3d_general(mx=48, my=48, mz=48), 80),
        'c4s': lambda: (S.acoustics_3d(mx=n3, my=n3, mz=n3, dimensional_split=True),
                        S.acoustics_3d(mx=48, my=48, mz=48, dimensional_split=True), 80),
    }
    for name in args.configs.split(','):
        sc, small, bpc = table[name]()
        out = run(name, sc, small, bpc, args.steps, args.warmup, args.precision, peak.value, hbm_peak)
        print(json.dumps(out), flush=True)


if __name__ == '__main__':
    main()
