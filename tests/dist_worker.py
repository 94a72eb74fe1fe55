"""Worker of the world_size>1 CPU tests (gloo): runs a scenario with the slab-decomposed host logic
(clawpack_b200.petclaw) driving the TEST-ONLY oracle plug-in, gathers the result and compares it on
rank 0 with the serial oracle run.  Usage: python dist_worker.py <case> <port>  (RANK/WORLD_SIZE in env)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import numpy as np
import torch.distributed as dist

from clawpack_b200 import petclaw, scenarios as S
from oracle_fmod import OracleFmod
from refutil import ref_solver

CASES = {
    'advection_unsplit_periodic': lambda: dict(S.advection_2d(mx=40, my=36, dimensional_split=False, transverse_waves=2),
                                               tfinal=0.3, num_output_times=2),
    'euler_unsplit_extrap': lambda: dict(S.quadrants(mx=30, my=41, tfinal=0.1), dimensional_split=False,
                                         transverse_waves=2, limiters=S.MC, num_output_times=2),
    'shallow_split_wall': lambda: dict(S.radial_dam_break(mx=36, my=30, tfinal=0.4), num_output_times=2),
    # 3-D, unsplit (2,22) with aux arrays: the reference example's lower wall / upper periodic faces (z-slabs),
    # and the all-periodic variant
    'acoustics3d_wall_periodic': lambda: S.acoustics_3d(mx=10, my=9, mz=13, tfinal=0.4, num_output_times=2),
    'acoustics3d_periodic': lambda: S.acoustics_3d(mx=8, my=7, mz=12, periodic=True, tfinal=0.4, num_output_times=2),
    # 2-D with aux planes (ghost exchange of the static aux array across y-slabs) and a scalar solver without rpt2
    'elasticity_unsplit_aux': lambda: S.elasticity_2d(mx=26, my=31, tfinal=0.1, num_output_times=2),
    'kpp_split': lambda: S.kpp_2d(mx=24, my=27, tfinal=0.3, num_output_times=2),
    'euler3d_unsplit_extrap': lambda: dict(S.euler_3d_shocktube(mx=5, my=6, mz=20, tfinal=0.05, num_output_times=1),
                                           dimensional_split=False),
}


def main():
    case, port = sys.argv[1], sys.argv[2]
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = port
    dist.init_process_group('gloo')
    rank, world = dist.get_rank(), dist.get_world_size()
    sc = CASES[case]()
    claw = petclaw.make_claw_distributed(sc, kernel_language='Fortran', fmod=OracleFmod())
    outdir = os.environ.get('CLAW_TEST_OUTDIR')
    if outdir:                                     # frame files: gathered on rank 0, serial ascii layout
        claw.output_format = os.environ.get('CLAW_TEST_FORMAT', 'ascii')
        claw.outdir = outdir
    claw.run()
    q = claw.solution.state.get_q_global()
    log = claw.solver.step_log
    ok = True
    if rank == 0:
        r = ref_solver(sc)
        r.run(sc['tfinal'], sc['num_output_times'])
        same_q = np.array_equal(q, r.q)
        same_log = (log == r.step_log)
        print('case %s world %d: q bit-identical to serial: %s, step log identical: %s (%d steps, %d rejected)'
              % (case, world, same_q, same_log, len(log), sum(1 for s in log if not s[2])))
        ok = same_q and same_log
        if outdir:
            from clawpack_b200 import pyclaw
            back = pyclaw.Solution(sc['num_output_times'], path=outdir)
            tol = 0.0 if claw.output_format == 'binary' else 5e-9 * max(1.0, np.max(np.abs(r.q)))     # binary is lossless
            same_file = (back.state.q.shape == r.q.shape and np.max(np.abs(back.state.q - r.q)) <= tol)
            print('frame file of the gathered solution matches the serial result: %s' % same_file)
            ok = ok and same_file
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
