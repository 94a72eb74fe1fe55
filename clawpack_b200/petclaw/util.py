"""Rank helpers and the scenario -> distributed Controller factory."""
import numpy as np


def _dist():
    import torch.distributed as dist
    return dist if (dist.is_available() and dist.is_initialized()) else None


def rank():
    d = _dist()
    return d.get_rank() if d else 0


def world_size():
    d = _dist()
    return d.get_world_size() if d else 1


def barrier():
    d = _dist()
    if d is not None and world_size() > 1:
        d.barrier()


def slab_range(n, nranks, r):
    """Cells [start, stop) of the slowest dimension owned by rank r (balanced block partition, the
    same rule as PETSc's default DMDA split: the first n % nranks ranks get one extra cell)."""
    base, extra = divmod(n, nranks)
    start = r * base + min(r, extra)
    return start, start + base + (1 if r < extra else 0)


def slab_neighbours(nranks, r, per_lo, per_up):
    """Who fills my ghost rows and who needs my interior rows (slowest dimension).  ``per_lo`` / ``per_up``:
    the lower / upper physical face is periodic (the reference allows them to differ, e.g.
    examples/acoustics_3d_variable: lower wall, upper periodic -- the DMDA is always periodic and a
    non-periodic BC simply overwrites the wrapped ghost values, petclaw/state.py:192-195).
    Channel A = upper interior -> the upper neighbour's lower ghosts; channel B = lower interior ->
    the lower neighbour's upper ghosts.  Every rank posts (send A, recv A, send B, recv B) in that
    order, so two messages between the same pair are matched in order."""
    nb = {}
    nb['send_up_to'] = r + 1 if r < nranks - 1 else (0 if per_lo else None)            # A
    nb['recv_lo_from'] = r - 1 if r > 0 else (nranks - 1 if per_lo else None)          # A
    nb['send_lo_to'] = r - 1 if r > 0 else (nranks - 1 if per_up else None)            # B
    nb['recv_up_from'] = r + 1 if r < nranks - 1 else (0 if per_up else None)          # B
    return nb


def make_claw_distributed(sc, kernel_language='CUDA', **solver_kw):
    """Build the Controller for a GLOBAL scenario dict; every rank keeps only its slab.  Initial data
    come from the global arrays, so 1-rank and N-rank runs start from identical bits
    (SURVEY.md appendix B)."""
    from .. import petclaw as pyclaw
    from .. import riemann
    ndim = len(sc['num_cells'])
    Solver = pyclaw.ClawSolver2D if ndim == 2 else pyclaw.ClawSolver3D
    solver = Solver(getattr(riemann, sc['rp']))
    solver.kernel_language = kernel_language
    solver.limiters = sc['limiters']
    solver.order = sc['order']
    solver.dimensional_split = sc['dimensional_split']
    if sc['transverse_waves'] is not None:
        solver.transverse_waves = sc['transverse_waves']
    solver.cfl_max = sc['cfl_max']
    solver.cfl_desired = sc['cfl_desired']
    solver.dt_initial = sc['dt_initial']
    for i in range(ndim):
        solver.bc_lower[i] = sc['bc_lower'][i]
        solver.bc_upper[i] = sc['bc_upper'][i]
        if sc['aux'] is not None:
            solver.aux_bc_lower[i] = sc['aux_bc_lower'][i]
            solver.aux_bc_upper[i] = sc['aux_bc_upper'][i]
    for k, v in solver_kw.items():
        setattr(solver, k, v)
    names = ['x', 'y', 'z'][:ndim]
    dims = [pyclaw.Dimension(lo, up, n, name=nm)
            for lo, up, n, nm in zip(sc['lower'], sc['upper'], sc['num_cells'], names)]
    domain = pyclaw.Domain(dims)
    num_aux = 0 if sc['aux'] is None else sc['aux'].shape[0]
    state = pyclaw.State(domain, sc['q'].shape[0], num_aux)
    s0, s1 = state.patch.slab
    if sc.get('slab') is not None:               # the scenario was built for this rank's slab only (global centres)
        if tuple(sc['slab']) != (s0, s1):
            raise Exception('scenario slab %r does not match the decomposition %r' % (tuple(sc['slab']), (s0, s1)))
        state.q[...] = np.asarray(sc['q'])
        if num_aux:
            state.aux[...] = np.asarray(sc['aux'])
    else:
        state.q[...] = np.asarray(sc['q'])[..., s0:s1]
        if num_aux:
            state.aux[...] = np.asarray(sc['aux'])[..., s0:s1]
    state.problem_data.update(sc['problem_data'])
    claw = pyclaw.Controller()
    claw.tfinal = sc['tfinal']
    claw.num_output_times = sc['num_output_times']
    claw.solution = pyclaw.Solution(state, domain)
    claw.solver = solver
    claw.output_format = None
    claw.keep_copy = False
    return claw
