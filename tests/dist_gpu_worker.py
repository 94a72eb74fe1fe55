"""Worker of the multi-GPU parity test (NCCL): slab-decomposed CUDA run vs the serial oracle.
Launch: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tests/dist_gpu_worker.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import numpy as np
import torch
import torch.distributed as dist

from clawpack_b200 import petclaw, scenarios as S
from refutil import ref_solver

CASES = {
    'euler_unsplit_extrap': lambda: dict(S.shock_bubble(mx=300, my=512), num_output_times=1, tfinal=0.01),
    'advection_unsplit_periodic': lambda: dict(S.advection_2d(mx=64, my=96, dimensional_split=False, transverse_waves=2),
                                               tfinal=0.2, num_output_times=2),
    'shallow_split_wall': lambda: dict(S.radial_dam_break(mx=150, my=120, tfinal=0.3), num_output_times=2),
    # 3-D z-slabs: unsplit (2,22) with aux (aux ghost planes exchanged once), wall below / periodic above;
    # dimension-split (both received ghost planes are swept at rank-to-rank faces => equals the SERIAL run)
    'acoustics3d_wall_periodic': lambda: S.acoustics_3d(mx=40, my=36, mz=32, tfinal=0.3, num_output_times=2),
    'acoustics3d_split_periodic': lambda: S.acoustics_3d(mx=30, my=20, mz=24, periodic=True, dimensional_split=True,
                                                        tfinal=0.3, num_output_times=2),
    'euler3d_split_extrap': lambda: S.euler_3d_shocktube(mx=10, my=10, mz=64, tfinal=0.1, num_output_times=2),
}


def main():
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()
    ok = True
    # second pass: the same cases with the exchange behind the C ABI (claw_halo_exchange / claw_allreduce_cfl on the
    # library's own NCCL communicator) -- the path a non-Python host would use
    # third pass: the peer-memory transport (ghost cells written into the neighbour's buffers over NVLink inside the
    # attempt, CFL through mailboxes; claw_peer_*)
    transports = os.environ.get('CLAW_TEST_TRANSPORTS', 'torch,cabi,peer').split(',')
    for transport, case in [(t, c) for t in transports for c in sorted(CASES)]:
        os.environ['CLAW_HALO'] = transport
        sc = CASES[case]()
        claw = petclaw.make_claw_distributed(sc, kernel_language='CUDA', cuda_precision='strict')
        claw.run()
        used = claw.solver._dev.halo.transport if claw.solver._dev is not None and claw.solver._dev.halo is not None else '-'
        q = claw.solution.state.get_q_global()
        log = claw.solver.step_log
        if rank == 0:
            r = ref_solver(sc)
            r.run(sc['tfinal'], sc['num_output_times'])
            same_q, same_log = np.array_equal(q, r.q), (log == r.step_log)
            print('case %s (%s -> %s) world %d: q bit-identical to serial oracle: %s, step log identical: %s (%d steps)'
                  % (case, transport, used, world, same_q, same_log, len(log)), flush=True)
            ok = ok and same_q and same_log
        claw.solver.teardown()
        dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
