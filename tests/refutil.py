"""Test helpers: build the ORACLE twin (oracle/pyclaw_ref.RefSolver) of a scenario dict."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import pyclaw_ref  # noqa: E402


def ref_solver(sc, nthreads=1):
    ndim = len(sc['num_cells'])
    num_aux = 0 if sc['aux'] is None else sc['aux'].shape[0]
    s = pyclaw_ref.RefSolver(sc['rp'], ndim, sc['problem_data'], num_aux=num_aux)
    s.limiters = sc['limiters']
    s.order = sc['order']
    if ndim > 1:
        s.dimensional_split = sc['dimensional_split']
        if sc['transverse_waves'] is not None:
            s.transverse_waves = sc['transverse_waves']
    else:
        s.dimensional_split = False
    s.cfl_max = sc['cfl_max']
    s.cfl_desired = sc['cfl_desired']
    s.dt_initial = sc['dt_initial']
    s.bc_lower = list(sc['bc_lower'])
    s.bc_upper = list(sc['bc_upper'])
    if num_aux:
        s.aux_bc_lower = list(sc['aux_bc_lower'])
        s.aux_bc_upper = list(sc['aux_bc_upper'])
    if sc.get('custom_bc_lower') is not None:
        f = sc['custom_bc_lower']
        s.user_bc_lower = lambda idim, qbc, g: f(qbc, g)
    if sc.get('custom_bc_lower_full') is not None:       # BC.custom fill reading auxbc and t (solver.py:393-398)
        ff = sc['custom_bc_lower_full']
        s.user_bc_lower = lambda idim, qbc, g: ff(qbc, s.auxbc, s.t, g)
    if sc.get('custom_aux_bc_lower') is not None:
        fl = sc['custom_aux_bc_lower']
        s.user_aux_bc_lower = lambda idim, auxbc, g: fl(auxbc, g)
    if sc.get('custom_aux_bc_upper') is not None:
        fu = sc['custom_aux_bc_upper']
        s.user_aux_bc_upper = lambda idim, auxbc, g: fu(auxbc, g)
    s.index_capa = sc.get('index_capa', -1)
    if sc.get('step_source') is not None:
        s.step_source = sc['step_source']
        s.source_split = sc.get('source_split', 1)
    if sc.get('before_step') is not None:
        s.before_step = sc['before_step']
    if sc.get('before_step_full') is not None:
        fbf = sc['before_step_full']
        s.before_step = lambda q, aux: fbf(q, aux, s.t, s.dt)
    if sc.get('fwave'):
        s.fwave = True
    if sc.get('gauges'):
        s.gauges = [[int((g[n] - sc['lower'][n]) // ((sc['upper'][n] - sc['lower'][n]) / float(sc['num_cells'][n])))
                     for n in range(ndim)] for g in sc['gauges']
                    if all(sc['lower'][n] <= g[n] < sc['upper'][n] for n in range(ndim))]   # geometry.py:398
        if sc.get('compute_gauge_values') is not None:
            s.compute_gauge_values = sc['compute_gauge_values']
    s.nthreads = nthreads
    delta = [(u - l) / float(n) for l, u, n in zip(sc['lower'], sc['upper'], sc['num_cells'])]
    s.setup(sc['q'], sc['aux'], delta)
    return s


def goldens():
    return np.load(os.path.join(ROOT, 'tests', 'golden', 'reference_goldens.npz'))


def goldens4():
    """1-D vectors of tests/golden/make_golden4.py (traffic, advection_1d_variable, stegoton_1d, shallow_1d sill,
    euler_1d shock-tube density)."""
    return np.load(os.path.join(ROOT, 'tests', 'golden', 'reference_goldens4.npz'))
