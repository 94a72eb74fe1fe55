"""CFL number with a max all-reduce over ranks (pyclaw/src/petclaw/cfl.py:4-36)."""
from . import util


class CFL(object):
    def __init__(self, global_max):
        self._local_max = global_max
        self._global_max = global_max

    def _reduce(self, value):
        d = util._dist()
        if d is None or d.get_world_size() == 1:
            return value
        import torch
        dev = 'cuda' if d.get_backend() == 'nccl' else 'cpu'
        t = torch.tensor([value], dtype=torch.float64, device=dev)
        d.all_reduce(t, op=d.ReduceOp.MAX)
        return float(t.item())

    def get_global_max(self):
        self._global_max = self._reduce(self._local_max)
        return self._global_max

    def get_cached_max(self):
        return self._global_max

    def set_local_max(self, new_local_max):
        self._local_max = new_local_max

    def set_global_max(self, new_global_max):
        self._global_max = new_global_max

    def update_global_max(self, new_local_max, already_global=False):
        """``already_global``: the device path all-reduced the value on the GPU before reading it."""
        v = new_local_max if already_global else self._reduce(new_local_max)
        self._global_max = max(self._global_max, v)
