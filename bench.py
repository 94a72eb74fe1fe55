#!/usr/bin/env python
"""bench.py -- headline benchmark: fp64 cell-updates/s of the classic wave-propagation step.

Workload (BASELINE.json configs[1] / configs[4]): 2-D Euler (euler_4wave_2D), unsplit with
transverse corrections (transverse_waves=2), MC limiter, 8192 x 8192 cells PER GPU, fp64,
synthetic shock/bubble initial data (clawpack_b200/scenarios.py:shock_bubble).  N>1 is the weak-scaling
sweep: the domain is tiled in y, one 8192^2 slab per rank, 2-row ghost exchange over NCCL
send/recv and a CFL max all-reduce every step.

A "step" is one hyperbolic step attempt through the PyClaw API (solver.evolve_to_time(solution)):
dt selection, ghost fill, fused sweep kernel, CFL read-back, accept/reject.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference      # the reference's CPU algorithm (oracle restatement), all host threads

Prints ONE JSON line (see the task's measurement contract).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

METRIC = 'fp64 cell-updates/s, Euler 2D 8192^2 per GPU (euler_4wave_2D, unsplit, transverse_waves=2, MC)'
UNIT = 'cell-updates/s'
# FP64-pipe issue weights of the op classes (SASS of sm_100a fast paths: div = MUFU.RCP64H + 8 DFMA/DMUL,
# sqrt = MUFU.RSQ64H + 9, compare/min/max = 1 DSETP; see DESIGN.md "Roofline")
W_DIV, W_SQRT = 8.0, 9.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='cuda', choices=['cuda', 'reference'])
    ap.add_argument('--n', type=int, default=8192, help='cells per side per GPU')
    ap.add_argument('--precision', default='fast', choices=['fast', 'strict'])
    ap.add_argument('--seg-rows', type=int, default=0)
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-n', type=int, default=1024, help='grid side of the bounded CPU-baseline sample')
    return ap.parse_args()


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower() == 'active':
                    reasons.add(nm)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': smax, 'reasons': sorted(reasons),
                'samples': len(sm)}


def algorithmic_ops(n_sample=192):
    """Exact per-cell operation counts of the reference algorithm on this workload (counting oracle),
    taken a few steps into the run so that the wave structure is representative."""
    from clawpack_b200 import scenarios as S
    from oracle import pyclaw_ref
    from refutil import ref_solver
    sc = S.shock_bubble(mx=n_sample, my=n_sample)
    r = ref_solver(sc)
    for _ in range(4):
        r.evolve_to_time()
    ops = pyclaw_ref.count_ops(r, r.dt)
    rep = ops.pop('repeated')      # Roe averages the reference rebuilds inside rpt2: not algorithmic work
    ops['reference_repeats_flops'] = rep['add'] + rep['mul'] + rep['div'] + rep['sqrt']
    ops['flops'] = ops['add'] + ops['mul'] + ops['div'] + ops['sqrt']
    ops['fp64_issues'] = ops['add'] + ops['mul'] + ops['cmp'] + W_DIV * ops['div'] + W_SQRT * ops['sqrt']
    return ops


def cpu_baseline(n, nthreads, min_seconds=10.0, max_steps=400):
    """The oracle (C++ restatement of the reference's f2py Fortran step) on a bounded sample of the
    same workload, threaded over pencils."""
    from clawpack_b200 import scenarios as S
    from refutil import ref_solver
    sc = S.shock_bubble(mx=n, my=n)
    r = ref_solver(sc, nthreads=nthreads)
    for _ in range(6):            # untimed: the deliberate reject (dt_initial = 0.1) + OpenMP pool spin-up
        r.evolve_to_time()
    t0 = time.perf_counter()
    steps = 0
    while steps < max_steps and (time.perf_counter() - t0 < min_seconds or steps < 3):
        r.evolve_to_time()
        steps += 1
    dt = time.perf_counter() - t0
    return {'value': n * n * steps / dt, 'unit': UNIT, 'cores': nthreads, 'kind': 'port',
            'sample': '%dx%d shock-bubble grid, %d steps, %.1f s, oracle/libclawref.so (C++ restatement of '
                      'step2/flux2/limiter + rpn2_euler_4wave/rpt2_euler, g++ -O2 -ffp-contract=off, '
                      'OpenMP over pencils)' % (n, n, steps, dt)}, dt / steps


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from oracle import pyclaw_ref
    pyclaw_ref.build()
    nthreads = pyclaw_ref.max_threads()
    n = args.cpu_n
    from clawpack_b200 import scenarios as S
    from refutil import ref_solver
    sc = S.shock_bubble(mx=n, my=n)
    r = ref_solver(sc, nthreads=nthreads)
    for _ in range(max(args.warmup, 6)):   # incl. the deliberate reject and the OpenMP pool spin-up
        r.evolve_to_time()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r.evolve_to_time()
    dt = time.perf_counter() - t0
    value = n * n * args.steps / dt
    sample = ('%dx%d sample of the 8192^2 workload per step (the reference itself stops at pencils > 1800 cells, '
              'rpn2_euler_4wave.f90:77-80), %d threads' % (n, n, nthreads))
    out = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
           'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True,
           'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
           'config': {'workload': 'euler_4wave_2D shock-bubble, unsplit, transverse_waves=2, MC limiter, fp64; '
                                  'bounded %dx%d sample per step' % (n, n)},
           'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': nthreads, 'kind': 'port', 'sample': sample},
           'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
           'gpu_launches': 0}
    print(json.dumps(out))


def main():
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from clawpack_b200 import _lib, scenarios as S

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (the CUDA path has no CPU fallback)')
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    n = args.n

    # ---- build the solver through the public PyClaw-shaped API --------------------------------
    if world > 1:
        from clawpack_b200 import petclaw as pyclaw
        sc = S.shock_bubble(mx=n, my=n * world, ly=float(world), ly_total=float(world))   # global definition
        claw = pyclaw.make_claw_distributed(sc, cuda_precision=args.precision, cuda_seg_rows=args.seg_rows)
    else:
        sc = S.shock_bubble(mx=n, my=n)
        claw = S.make_claw(sc, kernel_language='CUDA', cuda_precision=args.precision, cuda_seg_rows=args.seg_rows)
    solver, solution = claw.solver, claw.solution
    solver.setup(solution)
    solver.dt = solver.dt_initial
    dev = solver._dev
    cells_local = n * n
    cells_total = cells_local * world

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (includes the deliberate first reject) ------------------------------------------
    for _ in range(max(args.warmup, 3)):
        solver.evolve_to_time(solution)
    barrier()
    launches0 = dev.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        solver.evolve_to_time(solution)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        tms = torch.tensor([ms], dtype=torch.float64, device='cuda')
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    clocks = sampler.stop() if rank == 0 else None
    launches = dev.launch_count() - launches0
    log = solver.step_log[-args.steps:]
    value = cells_total * args.steps / (ms * 1e-3)

    # ---- dominant kernel alone, for the roofline --------------------------------------------------
    avg_ms = ctypes.c_double()
    _lib.check(dev.L.claw_time_sweep(dev.handle, _lib.dp(dev.delta), float(solver.dt), 5, ctypes.byref(avg_ms)),
               dev.handle)
    sweep_ms = avg_ms.value
    linfo = (ctypes.c_int * 5)()
    _lib.check(dev.L.claw_last_launch_info(dev.handle, linfo), dev.handle)

    out = None
    if rank == 0:
        peak = ctypes.c_double()
        _lib.check(dev.L.claw_measure_fp64_peak(local_rank, ctypes.byref(peak)))
        ops = algorithmic_ops()
        peaks_file = os.path.join(ROOT, 'MEASURED_PEAKS.json')
        hbm_peak, hbm_src = 6650.0, 'fallback (B200_PROFILING.md)'
        if os.path.exists(peaks_file):
            hbm_peak, hbm_src = float(json.load(open(peaks_file))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
        # Roofline of the dominant kernel.  The stencil is FP64-pipe bound (arithmetic intensity ~19 flop/B vs a
        # machine balance of ~5.5), so: achieved = ALGORITHMIC flops of the reference algorithm per launch
        # (add, mul, div, sqrt each counted as ONE flop; exact counts from the counting oracle, Roe averages once per
        # interface) / measured kernel time; peak = 2 flop x the DFMA issue rate measured in this run.  A division
        # or square root costs ~9-10 FP64-pipe issues for its one flop, so 1.0 is not attainable by this algorithm;
        # `fp64_pipe` carries the hardware view (ncu: cycles the FP64 pipe was busy) of the last committed capture.
        flops_per_s = ops['flops'] * cells_local / (sweep_ms * 1e-3)
        hbm_gbs = 64.0 * cells_local / (sweep_ms * 1e-3) / 1e9
        ncu = {}
        ncu_file = os.path.join(ROOT, 'profiles', 'latest_ncu.json')
        if os.path.exists(ncu_file):
            ncu = json.load(open(ncu_file)).get('sweep2_unsplit_euler_%s' % args.precision, {})
        roofline = {
            'bound': 'fp64',
            'achieved': flops_per_s / 1e12, 'peak': 2.0 * peak.value / 1e3, 'unit': 'TFLOP/s',
            'frac': flops_per_s / 1e9 / (2.0 * peak.value),
            'traffic': ncu.get('dram_bytes_per_launch'),
            'kernel': ('sweep2_unsplit_ws<RpEuler4Wave2D,1 warp/CTA,%d CTAs/SM,MC|spec|cp.async row prefetch> (%s)'
                       % (linfo[4], args.precision))
                      if linfo[2] == 32 else 'sweep2_unsplit<RpEuler4Wave2D,%d> (%s)' % (linfo[2], args.precision),
            'kernel_ms': sweep_ms,
            'launch': {'grid': [linfo[0], linfo[1]], 'block': linfo[2], 'smem_bytes': linfo[3],
                       'ctas_per_sm': linfo[4]},
            'peak_source': '2 x DFMA issue rate measured in this run (claw_measure_fp64_peak: %.2f T DFMA/s; nominal '
                           '18.6)' % (peak.value / 1e3),
            'algorithmic_per_cell': {k: round(v, 2) for k, v in ops.items() if k != 'fp64_issues'},
            'algorithmic_bytes_per_launch': 64 * cells_local,
            'fp64_pipe': ncu.get('fp64_pipe'),
            'hbm': {'achieved': hbm_gbs, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': hbm_gbs / hbm_peak,
                    'bytes_per_cell': 64, 'peak_source': hbm_src},
        }
        out = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': 'euler_4wave_2D shock-bubble %dx%d per GPU, unsplit, transverse_waves=2, MC limiter, '
                                   'fp64, arithmetic=%s' % (n, n, args.precision),
                       'cells_per_gpu': cells_local, 'parallelism': 'y-slabs x%d' % world,
                       'l2_policy': 'inputs larger than L2 (2 x %.2f GB state per GPU vs 126 MB L2)'
                                    % (4 * 8 * (n + 4) ** 2 / 1e9),
                       'accepted_steps': sum(1 for s in log if s[2]), 'rejected_steps': sum(1 for s in log if not s[2]),
                       'last_cfl': log[-1][1] if log else None},
            'roofline': roofline, 'clocks': clocks, 'gpu_launches': launches,
        }

    # ---- end-to-end through the f2py-shaped plug-in on pinned HOST buffers (rank 0, N=1 workload) -----
    if rank == 0 and not args.no_e2e:
        out['e2e'] = e2e_plugin(args, solver, n)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import pyclaw_ref
        pyclaw_ref.build()
        out['cpu_baseline'], _ = cpu_baseline(args.cpu_n, pyclaw_ref.max_threads())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))


def e2e_plugin(args, solver, n):
    """Same metric through the reference-facing plug-in call claw_step2_host (include/clawcuda.h): qold/qnew
    with ghost cells in pinned host memory, H2D + sweep + D2H inside the timed region."""
    import torch
    from clawpack_b200 import riemann
    from clawpack_b200.pyclaw import cuda_classic
    fm = cuda_classic.fmod(2, precision=args.precision)
    shape = (n + 4, n + 4, 4)                       # C-order view of the F-ordered (4, n+4, n+4) array
    qold_t = torch.empty(shape, dtype=torch.float64).pin_memory()
    qnew_t = torch.empty(shape, dtype=torch.float64).pin_memory()
    qold = qold_t.numpy().transpose(2, 1, 0)        # F-contiguous (4, n+4, n+4)
    qnew = qnew_t.numpy().transpose(2, 1, 0)
    state_q = solver._dev and None
    from clawpack_b200 import scenarios as S
    sc = S.shock_bubble(mx=n, my=n)
    qold[:, 2:-2, 2:-2] = sc['q']
    for g in range(2):                              # extrapolation ghost cells
        qold[:, g, :] = qold[:, 2, :]
        qold[:, -1 - g, :] = qold[:, -3, :]
    for g in range(2):
        qold[:, :, g] = qold[:, :, 2]
        qold[:, :, -1 - g] = qold[:, :, -3]
    qnew[...] = qold
    rp = riemann.euler_4wave_2D
    rp.cparam.gamma = 1.4
    dx = 1.0 / n
    dt = 0.9 * dx / 3.5
    method = np.array([1, 2, 2, 0, 0, 0, 0], dtype=np.intc)
    steps = max(3, min(args.steps, 5))
    fm.step2(n, 2, n, n, qold, qnew, None, dx, dx, dt, method, [4, 4, 4, 4], None, None, None, None, False, rp, rp)
    t0 = time.perf_counter()
    for _ in range(steps):
        _, cfl = fm.step2(n, 2, n, n, qold, qnew, None, dx, dx, dt, method, [4, 4, 4, 4], None, None, None, None,
                          False, rp, rp)
    torch.cuda.synchronize()
    t = (time.perf_counter() - t0) / steps
    fm.L.claw_host_release()
    return {'value': n * n / t, 'unit': UNIT, 'h2d_bytes_per_step': int(qold.nbytes),
            'd2h_bytes_per_step': int(4 * 8 * n * n + 8), 'ms_per_step': 1e3 * t,
            'api': 'claw_step2_host (f2py-shaped step2 plug-in), pinned host qold/qnew', 'steps': steps}


if __name__ == '__main__':
    main()
