"""Multi-GPU variant of the PyClaw front end: ``from clawpack_b200 import petclaw as pyclaw``.

Replaces PetClaw's PETSc DMDA + MPI (pyclaw/src/petclaw/) by one process per GPU with
torch.distributed: the patch is cut into slabs along the SLOWEST dimension (y in 2-D, z in 3-D), so
the two ghost rows/planes a rank needs from each neighbour are contiguous; per step there is one
ghost exchange (NCCL send/recv; replaces DMDA.globalToLocal, petclaw/state.py:223-231) and one
scalar max all-reduce of the CFL number (replaces Vec.max, petclaw/cfl.py:22,36).  With the 'gloo'
backend the same host logic runs on CPU tensors (used by the world_size-2 tests)."""
from ..pyclaw import limiters  # noqa: F401
from ..pyclaw.solver import BC, CFLError  # noqa: F401
from ..pyclaw.geometry import Dimension, Grid  # noqa: F401
from .cfl import CFL  # noqa: F401
from .geometry import Domain, Patch  # noqa: F401
from .state import State  # noqa: F401
from .solution import Solution  # noqa: F401
from .classic.solver import ClawSolver2D, ClawSolver3D  # noqa: F401
from .controller import Controller  # noqa: F401
from .util import make_claw_distributed, rank, world_size  # noqa: F401
