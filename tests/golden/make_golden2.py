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
for g in ('0.25_0.25', '11.75_11.75', '17.85_1.25', '3.25_18.75'):
    out['psystem_2d__gauge' + g] = np.loadtxt(os.path.join(REF, 'psystem_2d/verify_gauge%s.txt' % g))
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'reference_goldens2.npz'), **out)
for k, v in out.items():
    print(k, v.shape, v.dtype)
