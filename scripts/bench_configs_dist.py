#!/usr/bin/env python
"""Multi-GPU legs of BASELINE configs 3 and 4 (bench.py measures configs[1]/[4]):

    c3  shallow_roe_with_efix_2D 16384^2, dimension-split, MC -- STRONG scaling: the global grid is cut into y-slabs
    c4  vc_acoustics_3D 512^3, unsplit transverse_waves=22, MC -- z-slabs with a 2-plane halo exchange per step

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/bench_configs_dist.py --configs c3,c4 --steps 10

One JSON line per config from rank 0: whole-job cell-updates/s (CUDA events, max over ranks), ms/step."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--configs', default='c3,c4')
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--precision', default='fast')
    ap.add_argument('--n2', type=int, default=16384)
    ap.add_argument('--n3', type=int, default=512)
    args = ap.parse_args()
    import numpy as np
    import torch
    import torch.distributed as dist
    from clawpack_b200 import petclaw, scenarios as S
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    table = {
        'c3': lambda: S.radial_dam_break(mx=args.n2, my=args.n2),
        'c4': lambda: S.acoustics_3d(mx=args.n3, my=args.n3, mz=args.n3, periodic=True),
    }
    for name in args.configs.split(','):
        sc = table[name]()
        claw = petclaw.make_claw_distributed(sc, cuda_precision=args.precision)
        del sc['q']
        solver, solution = claw.solver, claw.solution
        solver.setup(solution)
        solver.dt = solver.dt_initial
        for _ in range(max(args.warmup, 3)):
            solver.evolve_to_time(solution)
        dist.barrier()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(args.steps):
            solver.evolve_to_time(solution)
        ev1.record()
        dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item()) / args.steps
        cells = int(np.prod(sc['num_cells']))
        log = solver.step_log[-args.steps:]
        if rank == 0:
            print(json.dumps({'config': name, 'n_gpus': world, 'workload': '%s %s' % (sc['rp'], 'x'.join(map(str, sc['num_cells']))),
                              'scaling': 'strong', 'cell_updates_per_s': cells / (ms * 1e-3), 'ms_per_step': ms,
                              'accepted': sum(1 for s in log if s[2]), 'rejected': sum(1 for s in log if not s[2]),
                              'arithmetic': args.precision}), flush=True)
        solver.teardown()
        del claw, solver, solution
        torch.cuda.empty_cache()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
