"""PyClaw-compatible front end of the B200 classic solvers.

    from clawpack_b200 import pyclaw, riemann
    solver = pyclaw.ClawSolver2D(riemann.euler_4wave_2D)      # kernel_language = 'CUDA'

Same names as ``clawpack.pyclaw`` for everything on the classic 2-D/3-D hot path
(pyclaw/src/pyclaw/__init__.py:20-40): Controller, Solution, State, Domain/Patch/Dimension, BC,
CFL, ClawSolver2D/3D, limiters.  SharpClaw, the 1-D solvers, file I/O, gauges and plotting are
outside this build's scope."""
import logging as _logging

from . import limiters  # noqa: F401
from .cfl import CFL  # noqa: F401
from .geometry import Dimension, Domain, Grid, Patch  # noqa: F401
from .state import State  # noqa: F401
from .solution import Solution  # noqa: F401
from .solver import BC, CFLError, Solver  # noqa: F401
from .classic.solver import ClawSolver, ClawSolver1D, ClawSolver2D, ClawSolver3D  # noqa: F401
from .controller import Controller  # noqa: F401

_logging.getLogger('pyclaw').addHandler(_logging.NullHandler())

__all__ = ['Controller', 'Solution', 'State', 'Dimension', 'Domain', 'Grid', 'Patch', 'BC', 'CFL', 'Solver',
           'ClawSolver', 'ClawSolver2D', 'ClawSolver3D', 'limiters']
