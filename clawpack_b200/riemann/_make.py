"""Factory for the solver "module objects" of clawpack.riemann.

In the reference each solver is an f2py extension module exposing ``rpn2``/``rpt2`` (C pointers
handed to the Fortran stepper, classic/solver.py:516-521) and a ``cparam`` common block
(state.set_cparam, state.py:192-217).  Here a solver is a tiny module-like object carrying the
id understood by libclawcuda (include/clawcuda.h) and the ordered cparam member names.
"""
import types


class _CParam(object):
    """Stand-in for the f2py view of COMMON /cparam/: attributes are the member names."""

    def __init__(self, names):
        object.__setattr__(self, '_names', tuple(names))
        for n in names:
            object.__setattr__(self, n, 0.0)

    def __dir__(self):
        return list(self._names)

    def values(self):
        import numpy as np
        return [float(v) for n in self._names for v in np.atleast_1d(getattr(self, n))]      # arrays (coeff(3)) flatten


def make(name, rp_id, cparam_names, num_dim):
    mod = types.ModuleType('clawpack_b200.riemann.' + name)
    mod.rp_id = rp_id
    mod.num_dim = num_dim
    mod.cparam_names = tuple(cparam_names)
    if cparam_names:
        mod.cparam = _CParam(cparam_names)
    mod.__doc__ = 'CUDA device functor %s (libclawcuda rp_id=%d)' % (name, rp_id)
    return mod
