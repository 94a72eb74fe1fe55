"""TEST-ONLY plug-in object with the f2py signatures of the reference's classic2/classic3
modules, backed by the ORACLE (oracle/libclawref.so).  Injected as ``solver.fmod`` with
``kernel_language='Fortran'`` to exercise the PyClaw-shaped host logic on CPU."""
import ctypes

import numpy as np

from oracle import pyclaw_ref as R


def _cparam(rp):
    cp = getattr(rp, 'cparam', None)
    vals = cp.values() if cp is not None else []
    return np.array(list(vals) + [0.0], dtype=np.float64)


class OracleFmod(object):
    def __init__(self, nthreads=1):
        self.L = R.lib()
        self.nthreads = nthreads

    def _aux(self, aux, shape):
        if aux is None or aux.size == 0:
            return np.zeros((1,) + tuple(shape), order='F')
        return np.asfortranarray(aux)

    def step1(self, num_ghost, mx, q, aux, dx, dt, method, mthlim, use_fwave, rp1):
        m = np.ascontiguousarray(method, dtype=np.intc)
        ml = np.array(mthlim, dtype=np.intc)
        aux = self._aux(aux, q.shape[1:])
        cp = _cparam(rp1)
        cfl = ctypes.c_double()
        rc = self.L.clawref_step1(q.shape[0], len(ml), num_ghost, int(m[6]), mx, R._dp(q), R._dp(aux), dx, dt,
                                  R._ip(m), R._ip(ml), ctypes.byref(cfl), int(use_fwave), int(rp1.rp_id), R._dp(cp))
        assert rc == 0
        return q, cfl.value

    def step2(self, maxm, num_ghost, mx, my, qold, qnew, aux, dx, dy, dt, method, mthlim, aux1, aux2, aux3, work,
              use_fwave, rpn2, rpt2):
        m = np.ascontiguousarray(method, dtype=np.intc)
        ml = np.array(mthlim, dtype=np.intc)
        aux = self._aux(aux, qold.shape[1:])
        cp = _cparam(rpn2)
        cfl = ctypes.c_double()
        rc = self.L.clawref_step2(maxm, qold.shape[0], len(ml), int(m[6]), num_ghost, mx, my, R._dp(qold), R._dp(qnew),
                                  R._dp(aux), dx, dy, dt, R._ip(m), R._ip(ml), ctypes.byref(cfl), int(use_fwave),
                                  int(rpn2.rp_id), R._dp(cp), self.nthreads)
        assert rc == 0
        return qnew, cfl.value

    def step2ds(self, maxm, num_ghost, mx, my, qold, qnew, aux, dx, dy, dt, method, mthlim, aux1, aux2, aux3, work,
                ids, use_fwave, rpn2, rpt2):
        m = np.ascontiguousarray(method, dtype=np.intc)
        ml = np.array(mthlim, dtype=np.intc)
        aux = self._aux(aux, qold.shape[1:])
        cp = _cparam(rpn2)
        cfl = ctypes.c_double()
        rc = self.L.clawref_step2ds(maxm, qold.shape[0], len(ml), int(m[6]), num_ghost, mx, my, R._dp(qold),
                                    R._dp(qnew), R._dp(aux), dx, dy, dt, R._ip(m), R._ip(ml), ctypes.byref(cfl),
                                    int(ids), int(use_fwave), int(rpn2.rp_id), R._dp(cp), self.nthreads)
        assert rc == 0
        return qnew, cfl.value

    def step3(self, maxm, num_ghost, mx, my, mz, qold, qnew, aux, dx, dy, dz, dt, method, mthlim, aux1, aux2, aux3,
              work, use_fwave, rpn3, rpt3, rptt3):
        m = np.ascontiguousarray(method, dtype=np.intc)
        ml = np.array(mthlim, dtype=np.intc)
        aux = self._aux(aux, qold.shape[1:])
        cp = _cparam(rpn3)
        cfl = ctypes.c_double()
        rc = self.L.clawref_step3(maxm, qold.shape[0], len(ml), int(m[6]), num_ghost, mx, my, mz, R._dp(qold),
                                  R._dp(qnew), R._dp(aux), dx, dy, dz, dt, R._ip(m), R._ip(ml), ctypes.byref(cfl),
                                  int(use_fwave), int(rpn3.rp_id), R._dp(cp), self.nthreads)
        assert rc == 0
        return qnew, cfl.value

    def step3ds(self, maxm, num_ghost, mx, my, mz, qold, qnew, aux, dx, dy, dz, dt, method, mthlim, aux1, aux2,
                aux3, work, idir, use_fwave, rpn3, rpt3, rptt3):
        m = np.ascontiguousarray(method, dtype=np.intc)
        ml = np.array(mthlim, dtype=np.intc)
        aux = self._aux(aux, qold.shape[1:])
        cp = _cparam(rpn3)
        cfl = ctypes.c_double()
        rc = self.L.clawref_step3ds(maxm, qold.shape[0], len(ml), int(m[6]), num_ghost, mx, my, mz, R._dp(qold),
                                    R._dp(qnew), R._dp(aux), dx, dy, dz, dt, R._ip(m), R._ip(ml), ctypes.byref(cfl),
                                    int(idir), int(use_fwave), int(rpn3.rp_id), R._dp(cp), self.nthreads)
        assert rc == 0
        return qnew, cfl.value
