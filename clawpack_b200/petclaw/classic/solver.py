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
        # each face of the slab dimension wraps around independently (the reference's 3-D acoustics example
        # has a wall below and a periodic face above); the ghost fill skips a periodic face on a rank that
        # does not own both ends (solver.py:400-402, 430-432), the exchange provides the wrapped values
        periodic = (self.bc_lower[last] == BC.periodic, self.bc_upper[last] == BC.periodic)
        state._periodic_last = periodic
        periodic_aux = (self.aux_bc_lower[last] == BC.periodic, self.aux_bc_upper[last] == BC.periodic)
        state._periodic_last_aux = periodic_aux
        super(_DistributedMixin, self).setup(solution)
        if self._dev is not None and state.patch.num_ranks > 1:
            from ..device_halo import DeviceHalo
            halo = DeviceHalo(self._dev, state.patch, periodic)
            if state.num_aux > 0:
                self._dev.halo_aux = DeviceHalo(self._dev, state.patch, periodic_aux)
                self._dev.halo_aux.exchange_aux(self._dev)
            self._dev.halo = halo

    def step_hyperbolic(self, solution):
        if self._dev is not None and self._dev.halo is not None:
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
