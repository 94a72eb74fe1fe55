"""Running maximum of the CFL number over one step (pyclaw/src/pyclaw/cfl.py:6-29)."""


class CFL(object):
    def __init__(self, global_max):
        self._global_max = global_max

    def get_global_max(self):
        return self._global_max

    def get_cached_max(self):
        return self._global_max

    def set_local_max(self, new_local_max):
        self._global_max = new_local_max

    def set_global_max(self, new_global_max):
        self._global_max = new_global_max

    def update_global_max(self, new_local_max):
        self._global_max = max(self._global_max, new_local_max)
