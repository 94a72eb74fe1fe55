"""ClawSolver2D / ClawSolver3D on a slab-decomposed patch: the serial classes plus the per-step ghost
exchange and the CFL all-reduce (pyclaw/src/petclaw/classic/solver.py just re-exports the serial
classes on top of the DMDA State; here the device-resident path needs its own hooks)."""
from ...pyclaw.classic import solver as base
from ...pyclaw.solver import BC
from .. import util


class _DistributedMixin(object):
    def setup(self, solution):
        state = solution.state
        last = self.num_dim - 1
        periodic = (self.bc_lower[last] == BC.periodic and self.bc_upper[last] == BC.periodic)
        if (self.bc_lower[last] == BC.periodic) != (self.bc_upper[last] == BC.periodic):
            raise Exception('periodic boundary conditions must be set on both faces of a dimension')
        state._periodic_last = periodic
        if periodic and state.patch.num_ranks > 1:
            # the wrap-around is an ordinary ghost exchange: no rank owns a physical face in that dimension
            state.grid.dimensions[last].on_lower_boundary = False
            state.grid.dimensions[last].on_upper_boundary = False
        super(_DistributedMixin, self).setup(solution)
        if self._dev is not None and state.patch.num_ranks > 1:
            from ..device_halo import DeviceHalo
            halo = DeviceHalo(self._dev, state.patch, periodic)
            self._dev.pre_step_hook = halo.exchange
            self._dev.post_step_hook = halo.reduce_cfl

    def step_hyperbolic(self, solution):
        if self._dev is not None and self._dev.post_step_hook is not None:
            cfl = self._dev.step(solution.states[0], self.dt)      # already the global max
            self.cfl.update_global_max(cfl, already_global=True)
            return
        super(_DistributedMixin, self).step_hyperbolic(solution)


class ClawSolver2D(_DistributedMixin, base.ClawSolver2D):
    def __init__(self, riemann_solver=None, claw_package=None):
        super(ClawSolver2D, self).__init__(riemann_solver, claw_package or 'clawpack_b200.petclaw')


class ClawSolver3D(_DistributedMixin, base.ClawSolver3D):
    def __init__(self, riemann_solver=None, claw_package=None):
        super(ClawSolver3D, self).__init__(riemann_solver, claw_package or 'clawpack_b200.petclaw')
