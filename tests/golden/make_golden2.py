"""Second batch of reference regression vectors (2-D solvers with aux arrays / capacity functions / f-waves and
the 5-equation Euler example) -> tests/golden/reference_goldens2.npz.

Run in the build container only (reads /root/reference):   python tests/golden/make_golden2.py
The vectors are DATA written by the reference's own Fortran path; no source is copied."""
import os
import numpy as np

REF = '/root/reference/pyclaw/examples'
out = {}
out['acoustics_2d_variable__pressure_classic'] = np.loadtxt(os.path.join(REF, 'acoustics_2d_variable/pressure_classic.txt'))
d = np.load(os.path.join(REF, 'advection_2d_annulus/expected_sols.npy'), allow_pickle=True).item()
out['advection_annulus__classic'] = np.asarray(d['classic'], dtype=np.float64)
d = np.load(os.path.join(REF, 'euler_2d/expected_sols.npy'), allow_pickle=True).item()
out['euler_2d__shock_bubble_interaction'] = np.asarray(d['shock_bubble_interaction'], dtype=np.float64)
out['euler_2d__quadrants_hlle'] = np.asarray(d['quadrants_hlle'], dtype=np.float64)
d = np.load(os.path.join(REF, 'shallow_2d/expected_sols.npy'), allow_pickle=True).item()
out['shallow_2d__sill'] = np.asarray(d['sill'], dtype=np.float64)
out['shallow_2d__radialdambreak_classic_hlle'] = np.asarray(d['radialdambreak_classic_hlle'], dtype=np.float64)
d = np.load(os.path.join(REF, 'burgers_1d/expected_sols.npy'), allow_pickle=True).item()
out['burgers_1d__fortran_classic'] = np.asarray(d['fortran_classic'], dtype=np.float64)
d = np.load(os.path.join(REF, 'shallow_1d/expected_sols.npy'), allow_pickle=True).item()
out['shallow_1d__dambreak_fortran_classic_roe'] = np.asarray(d['dambreak_fortran_classic_roe'], dtype=np.float64)
out['shallow_1d__dambreak_fortran_classic_hlle'] = np.asarray(d['dambreak_fortran_classic_hlle'], dtype=np.float64)
out['acoustics_2d_mapped__pressure'] = np.asarray(np.load(os.path.join(REF, 'acoustics_2d_mapped/pressure.npz'))['arr_0'],
                                                   dtype=np.float64)
for g in ('0.25_0.25', '11.75_11.75', '17.85_1.25', '3.25_18.75'):
    out['psystem_2d__gauge' + g] = np.loadtxt(os.path.join(REF, 'psystem_2d/verify_gauge%s.txt' % g))


def shock_bubble_interaction_q0(mx=160, my=40):
    """Initial state of pyclaw/examples/euler_2d/shock_bubble_interaction.py (sub-cell bubble fractions by
    scipy.integrate.quad), produced by running the reference example's own qinit on a State of this package
    (same attribute names); `clawpack` is aliased to this package only so that the example module imports."""
    import sys
    ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    sys.path.insert(0, ROOT)
    import clawpack_b200
    from clawpack_b200 import pyclaw, riemann
    sys.modules.setdefault('clawpack', clawpack_b200)
    sys.modules.setdefault('clawpack.riemann', riemann)
    sys.modules.setdefault('clawpack.riemann.euler_5wave_2D_constants', riemann.euler_5wave_2D_constants)
    sys.modules.setdefault('clawpack.pyclaw', pyclaw)
    clawpack_b200.riemann = riemann
    sys.path.insert(0, os.path.join(REF, 'euler_2d'))
    import shock_bubble_interaction as sbi
    x = pyclaw.Dimension(0.0, 2.0, mx, name='x')
    y = pyclaw.Dimension(0.0, 0.5, my, name='y')
    domain = pyclaw.Domain([x, y])
    state = pyclaw.State(domain, 5, 1)
    sbi.qinit(state)
    return np.array(state.q, order='F', copy=True)


out['euler_2d__shock_bubble_interaction_q0'] = shock_bubble_interaction_q0()

np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'reference_goldens2.npz'), **out)
for k, v in out.items():
    print(k, v.shape, v.dtype)
