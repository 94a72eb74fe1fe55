#!/usr/bin/env python
"""bench.py -- fp64 cell-updates/s of the classic wave-propagation step on B200.

Headline workload (BASELINE.json configs[1] / configs[4]; --config c2): 2-D Euler (euler_4wave_2D), unsplit with
transverse corrections (transverse_waves=2), MC limiter, 8192 x 8192 cells PER GPU, fp64, synthetic shock/bubble
initial data (clawpack_b200/scenarios.py:shock_bubble).  N > 1 is the weak-scaling sweep: the domain is tiled in y, one
8192^2 slab per rank, 2-row ghost exchange over NCCL send/recv and a CFL max all-reduce every step.

The other BASELINE configurations ride along in the same JSON line (`other_configs`) and can be made the primary
line with --config:
    c1  acoustics_2D 200x200, dimension-split, MC             (configs[0]; 1 GPU only: 40 k cells do not shard)
    c3  shallow_roe_with_efix_2D 16384^2, dimension-split, MC  (configs[2]; STRONG scaling over y-slabs)
    c4  vc_acoustics_3D 512^3, unsplit transverse_waves=22, MC (configs[3]; STRONG scaling over z-slabs)

A "step" is one hyperbolic step attempt through the PyClaw API (solver.evolve_to_time): dt selection, ghost fill
(+ ghost exchange), fused sweep kernel(s), CFL reduction, accept/reject -- all queued on the device without a host
round trip per step (cuda_classic.DeviceStepper.evolve).

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference      # the reference's CPU algorithm (oracle restatement), all host threads

Prints ONE JSON line (see the task's measurement contract).
"""
import argparse
import ctypes
import hashlib
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

UNIT = 'cell-updates/s'
METRICS = {
    'c2': 'fp64 cell-updates/s, Euler 2D 8192^2 per GPU (euler_4wave_2D, unsplit, transverse_waves=2, MC)',
    'c1': 'fp64 cell-updates/s, acoustics 2D 200^2 (acoustics_2D, dimension-split, MC)',
    'c3': 'fp64 cell-updates/s, shallow water 2D 16384^2 (shallow_roe_with_efix_2D, dimension-split, MC), strong scaling',
    'c4': 'fp64 cell-updates/s, acoustics 3D 512^3 (vc_acoustics_3D, unsplit transverse_waves=22, MC), strong scaling',
}
BYTES_PER_CELL = {'c1': 48, 'c2': 64, 'c3': 48, 'c4': 80}       # 8 * (2 * meqn + maux): one read + one write of q, one read of aux


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='cuda', choices=['cuda', 'reference'])
    ap.add_argument('--config', default='c2', choices=['c1', 'c2', 'c3', 'c4', 'c5'])
    ap.add_argument('--n', type=int, default=0, help='cells per side (default: the BASELINE size of the config)')
    ap.add_argument('--precision', default='fast', choices=['fast', 'strict'])
    ap.add_argument('--seg-rows', type=int, default=0)
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-other-configs', action='store_true')
    ap.add_argument('--no-parity', action='store_true')
    ap.add_argument('--no-dense', action='store_true')
    ap.add_argument('--cpu-n', type=int, default=1024, help='grid side of the bounded CPU-baseline sample (cuda arm)')
    args = ap.parse_args()
    if args.config == 'c5':
        args.config = 'c2'
    return args


def default_n(config):
    return {'c1': 200, 'c2': 8192, 'c3': 16384, 'c4': 512}[config]


# ---------------------------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------------------------
def global_cells(config, n, world):
    if config == 'c2':
        return [n, n * world]                 # weak scaling: one n x n slab per rank
    if config == 'c4':
        return [n, n, n]
    return [n, n]


def build_scenario(config, n, world=1, slab=None):
    """The GLOBAL problem of `config` on `world` ranks (slab: only this rank's rows / planes of the initial data)."""
    from clawpack_b200 import scenarios as S
    if config == 'c2':
        return S.shock_bubble(mx=n, my=n * world, ly=float(world), ly_total=float(world), slab=slab)
    if config == 'c1':
        return S.acoustics_2d(mx=n, my=n)
    if config == 'c3':
        return S.radial_dam_break(mx=n, my=n, slab=slab)
    if config == 'c4':
        return S.acoustics_3d(mx=n, my=n, mz=n, slab=slab)
    raise ValueError(config)


def small_scenario(config):
    """Small instance of the same scheme for the operation count (counting oracle)."""
    from clawpack_b200 import scenarios as S
    return {'c2': lambda: S.shock_bubble(mx=192, my=192), 'c1': lambda: S.acoustics_2d(mx=200, my=200),
            'c3': lambda: S.radial_dam_break(mx=256, my=256), 'c4': lambda: S.acoustics_3d(mx=48, my=48, mz=48)}[config]()


def workload_config(config, n, world):
    """`config` of the JSON line: identical for the CUDA arm and the reference arm."""
    cells = global_cells(config, n, world)
    names = {'c2': 'euler_4wave_2D shock-bubble %dx%d per GPU, unsplit, transverse_waves=2, MC limiter, fp64' % (n, n),
             'c1': 'acoustics_2D %dx%d, dimension-split, MC limiter, fp64' % (n, n),
             'c3': 'shallow_roe_with_efix_2D radial dam break %dx%d, dimension-split, MC limiter, fp64' % (n, n),
             'c4': 'vc_acoustics_3D %dx%dx%d, unsplit transverse_waves=22, MC limiter, fp64' % (n, n, n)}
    state_gb = BYTES_PER_CELL[config] * float(np.prod(cells)) / world / 1e9
    return {'workload': names[config], 'baseline_config': config, 'global_cells': cells,
            'cells_per_gpu': int(np.prod(cells) // world),
            'parallelism': ('y-slabs x%d' if config != 'c4' else 'z-slabs x%d') % world,
            'l2_policy': 'inputs larger than L2 (%.2f GB of state per GPU per step vs 126 MB L2)' % state_gb
                         if state_gb > 0.3 else 'flush: a 256 MB buffer is overwritten between timed batches'}


def algorithmic_ops(config):
    """Exact per-cell operation counts of the reference algorithm on this workload (counting oracle), taken a few
    steps into the run so that the wave structure is representative."""
    from oracle import pyclaw_ref
    from refutil import ref_solver
    r = ref_solver(small_scenario(config))
    for _ in range(4):
        r.evolve_to_time()
    ops = pyclaw_ref.count_ops(r, r.dt)
    rep = ops.pop('repeated')      # Roe averages the reference rebuilds inside rpt2: not algorithmic work
    ops['reference_repeats_flops'] = rep['add'] + rep['mul'] + rep['div'] + rep['sqrt']
    ops['flops'] = ops['add'] + ops['mul'] + ops['div'] + ops['sqrt']
    if config in ('c2', 'c3'):
        # the count on the DENSE variant of the data (same perturbation as the dense timing below): every wave is
        # non-zero there, so every limiter ratio is formed (limiter.f90:84 skips waves of zero norm)
        q = r.q
        idx = np.indices(q.shape[1:], sparse=True)
        pert = 1.0 + 1e-3 * np.sin(0.37 * idx[0] + 0.11 * idx[1] + 0.5)
        q[0] *= pert
        if config == 'c2':
            q[3] *= pert
        d = pyclaw_ref.count_ops(r, r.dt * 0.5)
        ops['dense_flops'] = d['add'] + d['mul'] + d['div'] + d['sqrt']
    return ops


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


# ---------------------------------------------------------------------------------------------------------------
# the reference arm: the reference's CPU algorithm (oracle restatement), all host cores
# ---------------------------------------------------------------------------------------------------------------
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_sample_scenario(config, n_full, reduced):
    """The CPU legs run the same scheme; `reduced` bounds the grid so that the leg finishes in ~10-30 s."""
    from clawpack_b200 import scenarios as S
    if config == 'c2':
        n = reduced or n_full
        return S.shock_bubble(mx=n, my=n), [n, n]
    if config == 'c1':
        return S.acoustics_2d(mx=n_full, my=n_full), [n_full, n_full]
    if config == 'c3':
        n = reduced or 4096
        return S.radial_dam_break(mx=n, my=n), [n, n]
    n = reduced or 96
    return S.acoustics_3d(mx=n, my=n, mz=n), [n, n, n]


def run_reference(args):
    """`--impl reference`: the reference's own algorithm for the path -- the C++ restatement of the f2py Fortran step
    (oracle/libclawref.so; the Fortran itself cannot be built in this image) -- on ALL host cores (torchrun exports
    OMP_NUM_THREADS=1 to its workers; the thread count is therefore taken from the CPU affinity mask and passed
    explicitly).  Headline config at N=1: the full 8192^2 grid, the driver's own step count (same config as the CUDA arm).
    For N > 1 (rank 0 alone) and for the larger configs each step is a bounded sample of the workload."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from oracle import pyclaw_ref
    pyclaw_ref.build()
    nthreads = host_cores()
    from refutil import ref_solver
    config = args.config
    n_full = args.n or default_n(config)
    world = args.gpus
    reduced = 0 if (config in ('c1', 'c2')) else None
    sc, dims = cpu_sample_scenario(config, n_full, reduced)
    if config == 'c4':
        nthreads_used = 1            # the unsplit 3-D oracle (step3) is serial
    else:
        nthreads_used = nthreads
    r = ref_solver(sc, nthreads=nthreads_used)
    cells = int(np.prod(dims))
    for _ in range(max(args.warmup, 2)):       # includes the deliberate reject (dt_initial) and the OpenMP pool spin-up
        r.evolve_to_time()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r.evolve_to_time()
    dt = time.perf_counter() - t0
    value = cells * args.steps / dt
    full = [n_full] * len(dims)
    if dims == full and world == 1:
        sample = 'the full %s grid, %d steps, %d threads' % ('x'.join(map(str, dims)), args.steps, nthreads_used)
    else:
        sample = ('%s grid per step (bounded sample of the %s workload on %d GPU(s): CPU cell-updates/s do not depend '
                  'on the grid size), %d steps, %d threads'
                  % ('x'.join(map(str, dims)), 'x'.join(map(str, global_cells(config, n_full, world))), world,
                     args.steps, nthreads_used))
    out = {'impl': 'reference', 'metric': METRICS[config], 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
           'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True,
           'scaling': 'weak' if config == 'c2' else 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
           'config': workload_config(config, n_full, world),
           'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': nthreads_used, 'kind': 'port',
                            'sample': sample + '; oracle/libclawref.so = C++ restatement of the f2py Fortran step '
                                               '(g++ -O2 -ffp-contract=off, OpenMP over pencils)',
                            'host_cores': nthreads},
           'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
           'gpu_launches': 0}
    print(json.dumps(out))


def cpu_baseline(config, n_full, cpu_n, min_seconds=10.0, max_steps=400):
    """`cpu_baseline` of the CUDA arm's line: the oracle on a bounded sample of the same workload."""
    from oracle import pyclaw_ref
    pyclaw_ref.build()
    from refutil import ref_solver
    nthreads = host_cores()
    sc, dims = cpu_sample_scenario(config, n_full, cpu_n if config == 'c2' else None)
    used = 1 if config == 'c4' else nthreads
    r = ref_solver(sc, nthreads=used)
    for _ in range(6):            # untimed: the deliberate reject (dt_initial = 0.1) + OpenMP pool spin-up
        r.evolve_to_time()
    t0 = time.perf_counter()
    steps = 0
    while steps < max_steps and (time.perf_counter() - t0 < min_seconds or steps < 3):
        r.evolve_to_time()
        steps += 1
    dt = time.perf_counter() - t0
    return {'value': int(np.prod(dims)) * steps / dt, 'unit': UNIT, 'cores': used, 'kind': 'port', 'host_cores': nthreads,
            'sample': '%s grid of the same scheme, %d steps, %.1f s, oracle/libclawref.so (C++ restatement of the f2py '
                      'Fortran step, g++ -O2 -ffp-contract=off, OpenMP over pencils)'
                      % ('x'.join(map(str, dims)), steps, dt)}


# ---------------------------------------------------------------------------------------------------------------
# the CUDA arm
# ---------------------------------------------------------------------------------------------------------------
class Ctx(object):
    pass


def make_claw(config, n, ctx, precision, seg_rows=0):
    from clawpack_b200 import scenarios as S
    if ctx.world > 1:
        from clawpack_b200 import petclaw
        from clawpack_b200.petclaw import util
        n_last = global_cells(config, n, ctx.world)[-1]
        slab = util.slab_range(n_last, ctx.world, ctx.rank)
        sc = build_scenario(config, n, ctx.world, slab=slab)
        claw = petclaw.make_claw_distributed(sc, cuda_precision=precision, cuda_seg_rows=seg_rows)
    else:
        sc = build_scenario(config, n, 1)
        claw = S.make_claw(sc, kernel_language='CUDA', cuda_precision=precision, cuda_seg_rows=seg_rows)
    sc.pop('q', None)
    sc.pop('aux', None)
    return claw


def sass_identity(kernel_pattern, obj):
    """(mangled name, sha1 of the SASS text) of the one kernel in `obj` whose mangled name contains the pattern."""
    try:
        sys.path.insert(0, os.path.join(ROOT, 'scripts'))
        import sass_mix
        ks = sass_mix.kernels(obj)
        names = [k for k in ks if kernel_pattern in k]
        if len(names) != 1:
            return None, None
        text = '\n'.join('%x %s' % (a, t) for a, t in ks[names[0]])
        return names[0], hashlib.sha1(text.encode()).hexdigest()
    except Exception:
        return None, None


NCU_KEYS = {'c2': ('sweep2_unsplit_wsINS0_15RpEuler4Wave2DTILi3ELi0EEELi1ELi8ELi71ELi1ELi8ELi0E', 'sweep_inst_%s.o'),
            'c3': ('sweep2_split_wsINS0_18RpShallowRoeEfix2DELi1ELi16ELi1ELi1ELi4ELi4E', 'sweep_inst_%s.o'),
            'c1': ('sweep2_split_wsINS0_13RpAcoustics2DELi1ELi16ELi1ELi1ELi4ELi4E', 'sweep_inst_%s.o'),
            'c4': ('sweep3_marchINS0_15RpVcAcoustics3DELi2ELi32ELi8ELi1ELi41E', 'sweep3_inst_%s.o')}


def ncu_evidence(config, precision):
    """DRAM traffic and FP64-pipe figures of the last committed `ncu --set full` capture of this config's dominant
    kernel (profiles/latest_ncu.json) -- only if that capture was taken from the kernel this build runs: the mangled
    name and the SHA-1 of its SASS must match, otherwise the figures are withheld and the entry is flagged stale."""
    f = os.path.join(ROOT, 'profiles', 'latest_ncu.json')
    if not os.path.exists(f):
        return {'traffic': None, 'fp64_pipe': None, 'ncu': 'no capture committed'}
    rec = json.load(open(f)).get('%s_%s' % (config, precision))
    if not rec:
        return {'traffic': None, 'fp64_pipe': None, 'ncu': 'no capture of this config / arithmetic mode'}
    pat, objname = NCU_KEYS[config]
    name, sha = sass_identity(pat, os.path.join(ROOT, 'clawpack_b200', 'csrc', 'build', objname % precision))
    if name is None:
        return {'traffic': None, 'fp64_pipe': None, 'ncu': 'stale: cannot identify the kernel of this build', 'stale': True}
    if rec.get('kernel_symbol') != name or rec.get('sass_sha1') != sha:
        return {'traffic': None, 'fp64_pipe': None, 'stale': True,
                'ncu': 'stale: %s was captured from another build of the kernel (sass sha1 %s, this build %s)'
                       % (rec.get('source'), str(rec.get('sass_sha1'))[:12], sha[:12])}
    return {'traffic': rec.get('dram_bytes_per_launch'), 'fp64_pipe': rec.get('fp64_pipe'), 'stale': False,
            'ncu': rec.get('source'), 'kernel_symbol': name, 'sass_sha1': sha}


def flush_l2(ctx):
    if ctx.flush is None:
        ctx.flush = ctx.torch.empty(32 * 1024 * 1024, dtype=ctx.torch.float64, device='cuda')   # 256 MB > 126 MB L2
    ctx.flush.fill_(1.0)


def time_steps(ctx, solver, solution, steps, warmup, sampler=None):
    """`steps` attempts of evolve_to_time(solution) (take-one-step mode), CUDA events, max over ranks."""
    torch, dist = ctx.torch, ctx.dist
    for _ in range(warmup):
        solver.evolve_to_time(solution)
    ctx.barrier()
    dev = solver._dev
    l0 = dev.launch_count()
    if sampler is not None:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    ev0.record()
    for _ in range(steps):
        solver.evolve_to_time(solution)
    ev1.record()
    ctx.barrier()
    ms = ev0.elapsed_time(ev1)
    if ctx.world > 1:
        tms = torch.tensor([ms], dtype=torch.float64, device='cuda')
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    return ms, dev.launch_count() - l0


def time_to_tend(ctx, solver, solution, nsteps_target):
    """Small grids (c1): evolve_to_time(solution, tend) over ~nsteps_target steps, so that the attempts go through the
    CUDA graph in batches; returns (ms, attempts, launches).  The L2 is flushed before the timed batch."""
    torch = ctx.torch
    for _ in range(5):
        solver.evolve_to_time(solution)
    dt = solver.dt
    solver.evolve_to_time(solution, solution.t + 8 * dt)          # graph capture happens here
    dev = solver._dev
    flush_l2(ctx)
    ctx.barrier()
    n0, l0 = len(solver.step_log), dev.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    solver.evolve_to_time(solution, solution.t + nsteps_target * dt)
    ev1.record()
    ctx.barrier()
    return ev0.elapsed_time(ev1), len(solver.step_log) - n0, dev.launch_count() - l0


def run_config(ctx, args, config, n, steps, warmup, primary):
    """One configuration through the PyClaw API; returns its dict (rank 0) or None."""
    from clawpack_b200 import _lib
    torch = ctx.torch
    t_setup = time.time()
    claw = make_claw(config, n, ctx, args.precision, args.seg_rows if config == 'c2' else 0)
    solver, solution = claw.solver, claw.solution
    solver.setup(solution)
    solver.dt = solver.dt_initial
    dev = solver._dev
    cells_global = int(np.prod(global_cells(config, n, ctx.world)))
    sampler = ClockSampler(ctx.local_rank) if (primary and ctx.rank == 0) else None
    if config == 'c1':
        if sampler is not None:
            sampler.start()
        ms, attempts, launches = time_to_tend(ctx, solver, solution, max(steps, 20) * 10)
    else:
        ms, launches = time_steps(ctx, solver, solution, steps, max(warmup, 3), sampler)
        attempts = steps
    clocks = sampler.stop() if sampler is not None else None
    log = solver.step_log[-attempts:]
    value = cells_global * attempts / (ms * 1e-3)
    # the sweep kernel(s) of one step alone, back to back, for the roofline
    avg_ms = ctypes.c_double()
    _lib.check(dev.L.claw_time_sweep(dev.handle, _lib.dp(dev.delta), float(solver.dt), 5 if config != 'c1' else 50,
                                     ctypes.byref(avg_ms)), dev.handle)
    linfo = (ctypes.c_int * 5)()
    _lib.check(dev.L.claw_last_launch_info(dev.handle, linfo), dev.handle)
    # The same kernel on DENSE data: the BASELINE initial data are piecewise constant, and the kernels of c1-c3 skip the
    # arithmetic of exactly uniform regions (results unchanged), which makes them HBM-bound there.  For the FP64-bound
    # regime every cell of the current state gets a smooth 1e-3 perturbation (no two neighbours carry the same bits) and
    # the sweep kernel is timed again on it.
    dense_ms = None
    dense_plain = False
    if config in ('c2', 'c3') and ctx.world == 1 and not args.no_dense:
        state = solution.state
        q = state.q
        idx = np.indices(q.shape[1:], sparse=True)
        pert = 1.0 + 1e-3 * np.sin(0.37 * idx[0] + 0.11 * idx[1] + 0.5)
        q[0] *= pert
        if config == 'c2':
            q[3] *= pert
        state.q = q
        dev._upload_if_needed(state)
        dm = ctypes.c_double()
        _lib.check(dev.L.claw_time_sweep(dev.handle, _lib.dp(dev.delta), float(solver.dt) * 0.5, 5, ctypes.byref(dm)),
                   dev.handle)
        dense_ms = dm.value
        dstate = (ctypes.c_int * 3)()
        _lib.check(dev.L.claw_dense_state(dev.handle, dstate), dev.handle)
        dense_plain = bool(dstate[1])
        del q, pert
    out = None
    if ctx.rank == 0:
        ops = algorithmic_ops(config)
        cells_local = cells_global // ctx.world
        flops_per_s = ops['flops'] * cells_local / (avg_ms.value * 1e-3)
        hbm_gbs = BYTES_PER_CELL[config] * cells_local / (avg_ms.value * 1e-3) / 1e9
        t_f = ops['flops'] / (2.0 * ctx.peak_gdfma * 1e9)                     # s per cell if FP64-bound
        t_h = BYTES_PER_CELL[config] / (ctx.hbm_peak * 1e9)                   # s per cell if HBM-bound
        bound = 'fp64' if t_f > t_h else 'hbm'
        # c1-c3 on the BASELINE data: uniform regions cost no arithmetic (see above), the measured kernel time is far
        # below the FP64 time of the algorithmic operation count -- the binding roofline of THIS workload is HBM
        if flops_per_s / 1e9 / (2.0 * ctx.peak_gdfma) > 1.0 or (dense_ms is not None and dense_ms > 1.5 * avg_ms.value):
            bound = 'hbm'
        ev = ncu_evidence(config, args.precision)
        kernels = {'c2': 'sweep2_unsplit_ws<RpEuler4Wave2DT<3>, 1 warp/CTA, MC|spec|cp.async row ring>',
                   'c1': 'sweep2_split_ws<RpAcoustics2D, 1 warp/CTA, LIM=MC> (x- and y-sweep fused in one march)',
                   'c3': 'sweep2_split_ws<RpShallowRoeEfix2D, 1 warp/CTA, LIM=MC> (x- and y-sweep fused in one march)',
                   'c4': 'sweep3_x<RpVcAcoustics3D,32x6,SPEC> + sweep3_march<..,y> + sweep3_march<..,z> (three launches)'}
        roofline = {
            'bound': bound,
            'achieved': flops_per_s / 1e12 if bound == 'fp64' else hbm_gbs,
            'peak': 2.0 * ctx.peak_gdfma / 1e3 if bound == 'fp64' else ctx.hbm_peak,
            'unit': 'TFLOP/s' if bound == 'fp64' else 'GB/s',
            'frac': flops_per_s / 1e9 / (2.0 * ctx.peak_gdfma) if bound == 'fp64' else hbm_gbs / ctx.hbm_peak,
            'traffic': ev['traffic'],
            'kernel': '%s (%s)' % (kernels[config], args.precision),
            'kernel_ms': avg_ms.value,
            'launch': {'grid': [linfo[0], linfo[1]], 'block': linfo[2], 'smem_bytes': linfo[3], 'ctas_per_sm': linfo[4]},
            'peak_source': ('2 x DFMA issue rate measured in this run (claw_measure_fp64_peak: %.2f T DFMA/s; nominal 18.6)'
                            % (ctx.peak_gdfma / 1e3)) if bound == 'fp64' else ctx.hbm_src,
            'algorithmic_flops_per_cell': round(ops['flops'], 2),
            'algorithmic_per_cell': {k: round(v, 2) for k, v in ops.items()},
            'algorithmic_bytes_per_cell': BYTES_PER_CELL[config],
            'algorithmic_bytes_per_launch': BYTES_PER_CELL[config] * cells_local,
            'fp64_pipe': ev['fp64_pipe'], 'ncu': ev['ncu'], 'ncu_stale': ev.get('stale'),
            'hbm': {'achieved': hbm_gbs, 'peak': ctx.hbm_peak, 'unit': 'GB/s', 'frac': hbm_gbs / ctx.hbm_peak,
                    'peak_source': ctx.hbm_src},
            'fp64': {'achieved': flops_per_s / 1e12, 'peak': 2.0 * ctx.peak_gdfma / 1e3, 'unit': 'TFLOP/s',
                     'frac': flops_per_s / 1e9 / (2.0 * ctx.peak_gdfma),
                     'note': 'algorithmic flops of the reference on this data / kernel time; > 1 means the kernel skipped '
                             'the arithmetic of uniform regions (exact zeros) and is not FP64-bound on this data'},
        }
        if dense_ms is not None:
            dflops = ops.get('dense_flops', ops['flops']) * cells_local / (dense_ms * 1e-3)
            roofline['dense'] = {'data': 'the same state with a smooth 1e-3 perturbation in every cell (no uniform regions)',
                                 'bound': 'fp64', 'kernel_ms': dense_ms, 'achieved': dflops / 1e12,
                                 'algorithmic_flops_per_cell': round(ops.get('dense_flops', ops['flops']), 2),
                                 'peak': 2.0 * ctx.peak_gdfma / 1e3, 'unit': 'TFLOP/s',
                                 'frac': dflops / 1e9 / (2.0 * ctx.peak_gdfma),
                                 'cell_updates_per_s': cells_local / (dense_ms * 1e-3),
                                 'kernel': ('the same kernel WITHOUT the uniform-region fast path: the dense-data switch takes it '
                                            'when most tiles report the full path (DESIGN.md 3.1a)') if dense_plain else
                                           'the kernel with the uniform-region fast path'}
        out = {'metric': METRICS[config], 'value': value, 'unit': UNIT, 'n_gpus': ctx.world, 'steps': attempts,
               'warmup': max(warmup, 3), 'ms_per_step': ms / attempts, 'higher_is_better': True,
               'scaling': 'weak' if config == 'c2' else 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
               'config': workload_config(config, n, ctx.world),          # identical in both arms
               'run': {'arithmetic': args.precision, 'accepted_steps': sum(1 for s in log if s[2]),
                       'rejected_steps': sum(1 for s in log if not s[2]), 'last_cfl': log[-1][1] if log else None,
                       'stepping': 'evolve_to_time(solution, tend): attempts queued in CUDA-graph batches'
                       if config == 'c1' else 'evolve_to_time(solution) per step, queued without a CFL read-back'},
               'roofline': roofline, 'clocks': clocks, 'gpu_launches': launches, 'setup_s': round(time.time() - t_setup, 1)}
        if ctx.world > 1 and dev.halo is not None:
            out['run']['halo_transport'] = {
                'peer': 'peer memory over NVLink (CUDA IPC): ghost rows written into the neighbour\'s buffers after the '
                        'sweep, CFL max through mailboxes inside the control kernel; no collective per step',
                'torch': 'NCCL send/recv + all-reduce through torch.distributed',
                'cabi': 'NCCL send/recv + all-reduce behind the C ABI (claw_halo_exchange / claw_allreduce_cfl)'}[dev.halo.transport]
    solver.teardown()
    del claw, solver, solution, dev
    torch.cuda.empty_cache()
    return out


def parity_check(ctx, precision):
    """A small slab case on the N ranks of this run against the SERIAL oracle before anything is timed: 2-D Euler
    unsplit (the benchmarked kernel, strict arithmetic: bit for bit; and the benchmarked arithmetic mode: within the
    north-star tolerance), 6 attempts including the deliberate reject."""
    from clawpack_b200 import scenarios as S
    mx, my = 192, 64 * ctx.world
    res = {'nranks': ctx.world, 'case': 'euler_4wave_2D shock-bubble %dx%d, unsplit tw=2, MC, 6 step attempts vs the serial '
                                        'oracle' % (mx, my)}
    ref_q = ref_log = None
    if ctx.rank == 0:
        from refutil import ref_solver
        r = ref_solver(S.shock_bubble(mx=mx, my=my, ly=float(ctx.world), ly_total=float(ctx.world)))
        for _ in range(6):
            r.evolve_to_time()
        ref_q, ref_log = r.q, r.step_log
    for mode in ('strict', precision):
        if ctx.world > 1:
            from clawpack_b200 import petclaw
            sc = S.shock_bubble(mx=mx, my=my, ly=float(ctx.world), ly_total=float(ctx.world))
            claw = petclaw.make_claw_distributed(sc, cuda_precision=mode)
        else:
            claw = S.make_claw(S.shock_bubble(mx=mx, my=my), kernel_language='CUDA', cuda_precision=mode)
        claw.solver.setup(claw.solution)
        claw.solver.dt = claw.solver.dt_initial
        for _ in range(6):
            claw.solver.evolve_to_time(claw.solution)
        q = claw.solution.state.get_q_global()
        log = list(claw.solver.step_log)
        if ctx.rank == 0:
            if mode == 'strict':
                res['bit_identical'] = bool(np.array_equal(q, ref_q) and log == ref_log)
            if mode == precision:
                res['arithmetic'] = mode
                res['rel_linf'] = float(np.max(np.abs(q - ref_q)) / np.max(np.abs(ref_q)))
                res['same_accept_reject'] = [s[2] for s in log] == [s[2] for s in ref_log]
                res['within_tolerance'] = bool(res['rel_linf'] <= 1e-11 and res['same_accept_reject'])
        claw.solver.teardown()
        if mode == precision == 'strict':
            break
    return res if ctx.rank == 0 else None


def e2e_plugin(ctx, args, n):
    """Same metric through the reference-facing plug-in call claw_step2_host (include/clawcuda.h): qold/qnew with ghost
    cells in pinned host memory, H2D + sweep + D2H inside the timed region.  Every rank runs its own slab (the ghost rows
    between slabs are the caller's business at this boundary, as they are for PetClaw's f2py call): the value is
    N x cells / (max over ranks of the time per step)."""
    import torch
    from clawpack_b200 import riemann, scenarios as S
    from clawpack_b200.pyclaw import cuda_classic
    fm = cuda_classic.fmod(2, precision=args.precision)
    shape = (n + 4, n + 4, 4)                       # C-order view of the F-ordered (4, n+4, n+4) array
    qold_t = torch.empty(shape, dtype=torch.float64).pin_memory()
    qnew_t = torch.empty(shape, dtype=torch.float64).pin_memory()
    qold = qold_t.numpy().transpose(2, 1, 0)        # F-contiguous (4, n+4, n+4)
    qnew = qnew_t.numpy().transpose(2, 1, 0)
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
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        _, cfl = fm.step2(n, 2, n, n, qold, qnew, None, dx, dx, dt, method, [4, 4, 4, 4], None, None, None, None,
                          False, rp, rp)
    torch.cuda.synchronize()
    t = (time.perf_counter() - t0) / steps
    if ctx.world > 1:
        tt = torch.tensor([t], dtype=torch.float64, device='cuda')
        ctx.dist.all_reduce(tt, op=ctx.dist.ReduceOp.MAX)
        t = float(tt.item())
    fm.L.claw_host_release()
    return {'value': ctx.world * n * n / t, 'unit': UNIT, 'h2d_bytes_per_step': int(qold.nbytes) * ctx.world,
            'd2h_bytes_per_step': int(4 * 8 * n * n + 8) * ctx.world, 'ms_per_step': 1e3 * t,
            'api': 'claw_step2_host (f2py-shaped step2 plug-in), pinned host qold/qnew, one %dx%d slab per rank'
                   % (n, n), 'n_ranks': ctx.world, 'steps': steps}


def main():
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from clawpack_b200 import _lib

    ctx = Ctx()
    ctx.torch, ctx.dist = torch, dist
    ctx.world = int(os.environ.get('WORLD_SIZE', '1'))
    ctx.rank = int(os.environ.get('RANK', '0'))
    ctx.local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    ctx.flush = None
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (the CUDA path has no CPU fallback)')
    torch.cuda.set_device(ctx.local_rank)
    if ctx.world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', ctx.local_rank))

    def barrier():
        if ctx.world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    ctx.barrier = barrier

    peak = ctypes.c_double()
    _lib.check(_lib.lib().claw_measure_fp64_peak(ctx.local_rank, ctypes.byref(peak)))
    ctx.peak_gdfma = peak.value
    ctx.hbm_peak, ctx.hbm_src = 6650.0, 'fallback (B200_PROFILING.md)'
    pf = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(pf):
        ctx.hbm_peak, ctx.hbm_src = float(json.load(open(pf))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'

    config = args.config
    n = args.n or default_n(config)
    if config == 'c1' and ctx.world > 1:
        raise SystemExit('c1 (200x200) is a single-GPU configuration')

    parity = None
    if not args.no_parity:
        parity = parity_check(ctx, args.precision)
    out = run_config(ctx, args, config, n, args.steps, args.warmup, primary=True)
    if ctx.rank == 0:
        out['parity'] = parity

    if config == 'c2' and not args.no_e2e:
        e2e = e2e_plugin(ctx, args, n)
        if ctx.rank == 0:
            out['e2e'] = e2e
    if ctx.rank == 0 and ctx.world == 1 and not args.no_cpu_baseline:
        out['cpu_baseline'] = cpu_baseline(config, n, args.cpu_n)

    if config == 'c2' and not args.no_other_configs:
        others = []
        for oc in ('c1', 'c3', 'c4'):
            if oc == 'c1' and ctx.world > 1:
                continue
            r = run_config(ctx, args, oc, default_n(oc), 10, 3, primary=False)
            if ctx.rank == 0:
                others.append({'config': oc, 'workload': r['config']['workload'], 'metric': r['metric'],
                               'value': r['value'], 'unit': UNIT, 'ms_per_step': r['ms_per_step'], 'steps': r['steps'],
                               'n_gpus': ctx.world, 'scaling': r['scaling'],
                               'accepted_steps': r['run']['accepted_steps'],
                               'rejected_steps': r['run']['rejected_steps'], 'gpu_launches': r['gpu_launches'],
                               'roofline': {k: r['roofline'][k] for k in ('bound', 'achieved', 'peak', 'unit', 'frac',
                                                                         'kernel', 'kernel_ms', 'traffic',
                                                                         'algorithmic_flops_per_cell',
                                                                         'algorithmic_bytes_per_cell', 'ncu_stale')
                                            if k in r['roofline']}})
                if 'dense' in r['roofline']:
                    others[-1]['roofline']['dense'] = r['roofline']['dense']
                others[-1]['roofline']['fp64_frac'] = r['roofline']['fp64']['frac']
                others[-1]['roofline']['hbm_frac'] = r['roofline']['hbm']['frac']
        if ctx.rank == 0:
            out['other_configs'] = others
    if ctx.world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if ctx.rank == 0:
        print(json.dumps(out))


if __name__ == '__main__':
    main()
