"""Extract the reference's stored regression vectors that pin the classic 2-D/3-D step
 into one small fixture, tests/golden/reference_goldens.npz.

Run in the build container only (reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
The vectors are DATA written by the reference's own Fortran path (gfortran); no source is copied.
"""
import os
import numpy as np

REF = '/root/reference/pyclaw/examples'
out = {}

d = np.load(os.path.join(REF, 'advection_2d/expected_sols.npy'), allow_pickle=True).item()
for k in ('classic_dim_split', 'classic_unsplit_no_trans', 'classic_unsplit_trans_inc', 'classic_unsplit_trans_cor'):
    out['advection_2d__' + k] = np.asarray(d[k], dtype=np.float64)

out['acoustics_2d__verify_classic'] = np.loadtxt(os.path.join(REF, 'acoustics_2d_homogeneous/verify_classic.txt'))

d = np.load(os.path.join(REF, 'euler_2d/expected_sols.npy'), allow_pickle=True).item()
out['euler_2d__quadrants_roe'] = np.asarray(d['quadrants_roe'], dtype=np.float64)

d = np.load(os.path.join(REF, 'shallow_2d/expected_sols.npy'), allow_pickle=True).item()
out['shallow_2d__radialdambreak_classic_roe'] = np.asarray(d['radialdambreak_classic_roe'], dtype=np.float64)

p = os.path.join(REF, 'acoustics_3d_variable/verify_classic_heterogeneous.txt')
if os.path.exists(p):
    out['acoustics_3d__verify_classic_heterogeneous'] = np.loadtxt(p)

np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'reference_goldens.npz'), **out)
for k, v in out.items():
    print(k, v.shape, v.dtype)
