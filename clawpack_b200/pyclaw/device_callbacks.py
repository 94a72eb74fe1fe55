"""Callbacks that also exist as DEVICE functors.

The reference lets the user hook Python functions into the step: ``solver.step_source`` (classic/solver.py:126-144) and
``BC.custom`` fills (solver.py:393-398).  A Python function needs the state on the host, so a solver with such a hook runs
in plug-in mode (a host round trip per step).  The objects below are callables with the reference's signatures -- they
work as ordinary host callbacks anywhere -- and additionally name a kernel of libclawcuda (include/clawcuda.h:
``claw_set_source``, ``CLAW_BC_FIXED_STATE``), so that ``kernel_language='CUDA'`` keeps the run device-resident:

    solver.step_source = device_callbacks.EulerRadialSource(gamma)        # instead of step_Euler_radial
    solver.bc_lower[0] = pyclaw.BC.custom
    solver.user_bc_lower = device_callbacks.FixedStateBC([rinf, rinf * vinf, 0., einf, 0.])   # instead of incoming_shock
"""
import numpy as np

SRC_EULER_RADIAL_2D = 1        # include/clawcuda.h: CLAW_SRC_EULER_RADIAL_2D
BC_FIXED_STATE = 4             # include/clawcuda.h: CLAW_BC_FIXED_STATE


class EulerRadialSource(object):
    """Geometric source terms of the 2-D Euler equations with cylindrical symmetry, two-stage Runge-Kutta: the
    reference's ``step_Euler_radial`` (pyclaw/examples/euler_2d/shock_bubble_interaction.py:112-143).  ``aux[0]`` is the
    radial coordinate of the cell centres; components 0..3 of q are (rho, rho u, rho v, E), further components (the
    tracer of euler_5wave_2D) are left alone."""

    def __init__(self, gamma):
        self.gamma = float(gamma)
        self.device_source = (SRC_EULER_RADIAL_2D, [self.gamma])

    def __call__(self, solver, state, dt):
        gamma = self.gamma
        dt2 = dt / 2.
        q = state.q
        rad = state.aux[0, :, :]
        rho = q[0, :, :]
        u = q[1, :, :] / rho
        v = q[2, :, :] / rho
        press = (gamma - 1.) * (q[3, :, :] - 0.5 * rho * (u ** 2 + v ** 2))
        qstar = np.empty(q.shape)
        qstar[0, :, :] = q[0, :, :] - dt2 / rad * q[2, :, :]
        qstar[1, :, :] = q[1, :, :] - dt2 / rad * rho * u * v
        qstar[2, :, :] = q[2, :, :] - dt2 / rad * rho * v * v
        qstar[3, :, :] = q[3, :, :] - dt2 / rad * v * (q[3, :, :] + press)
        rho = qstar[0, :, :]
        u = qstar[1, :, :] / rho
        v = qstar[2, :, :] / rho
        press = (gamma - 1.) * (qstar[3, :, :] - 0.5 * rho * (u ** 2 + v ** 2))
        q[0, :, :] = q[0, :, :] - dt / rad * qstar[2, :, :]
        q[1, :, :] = q[1, :, :] - dt / rad * rho * u * v
        q[2, :, :] = q[2, :, :] - dt / rad * rho * v * v
        q[3, :, :] = q[3, :, :] - dt / rad * v * (qstar[3, :, :] + press)


class FixedStateBC(object):
    """BC.custom fill that writes one fixed state into every ghost cell of a face: the reference's ``incoming_shock``
    (shock_bubble_interaction.py:93-109).  ``upper=True`` for a ``user_bc_upper`` hook.  The same object may serve
    several dimensions (the host call gets ``dim``; on the device every BC.custom face of that side takes the state)."""

    def __init__(self, values, upper=False):
        self.values = [float(v) for v in values]
        self.upper = bool(upper)
        self.device_bc = (BC_FIXED_STATE, self.values)

    def __call__(self, state, dim, t, qbc, auxbc, num_ghost):
        idim = state.grid.dimensions.index(dim) if dim in state.grid.dimensions else [d.name for d in state.grid.dimensions].index(dim.name)
        view = np.rollaxis(qbc, idim + 1, 1)
        for m, v in enumerate(self.values):
            if self.upper:
                view[m, -num_ghost:, ...] = v
            else:
                view[m, :num_ghost, ...] = v
