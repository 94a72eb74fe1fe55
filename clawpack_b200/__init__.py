"""clawpack_b200 -- B200-native classic wave-propagation step behind PyClaw's API.

``from clawpack_b200 import pyclaw, riemann`` mirrors ``from clawpack import pyclaw, riemann``
for the classic 2-D/3-D solvers with a new ``kernel_language='CUDA'``; ``petclaw`` is the
multi-GPU (one process per GPU, torch.distributed/NCCL) variant of the same classes."""
__version__ = '0.1.0'

from . import riemann  # noqa: F401,E402
from . import pyclaw  # noqa: F401,E402
