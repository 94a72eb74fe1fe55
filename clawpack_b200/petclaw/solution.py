"""Solution built from the distributed State (pyclaw/src/petclaw/solution.py)."""
from ..pyclaw.solution import Solution as _Solution


class Solution(_Solution):
    def __init__(self, *arg, **kargs):
        kargs.setdefault('claw_package', 'clawpack_b200.petclaw')
        super(Solution, self).__init__(*arg, **kargs)
