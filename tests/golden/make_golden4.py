"""Fourth batch of reference regression vectors (1-D classic solvers added after the first batch, advection_reaction_2d)
-> tests/golden/reference_goldens4.npz.

Run in the build container only (reads /root/reference):   python tests/golden/make_golden4.py
The vectors are DATA written by the reference's own Fortran path; no source is copied."""
import os
import numpy as np

REF = '/root/reference/pyclaw/examples'
out = {}
d = np.load(os.path.join(REF, 'traffic/expected_sols.npy'), allow_pickle=True).item()
out['traffic__classic'] = np.asarray(d['classic'], dtype=np.float64)                         # traffic/test_traffic.py:21-22
d = np.load(os.path.join(REF, 'advection_1d_variable/expected_sols.npy'), allow_pickle=True).item()
out['advection_1d_variable__fortran_classic'] = np.asarray(d['fortran_classic'], dtype=np.float64)   # test_variable_coefficient_advection.py:35-36
d = np.load(os.path.join(REF, 'stegoton_1d/expected_sols.npy'), allow_pickle=True).item()
out['stegoton_1d__fortran_classic'] = np.asarray(d['fortran_classic'], dtype=np.float64)     # stegoton_1d/test_stego.py:21-22
d = np.load(os.path.join(REF, 'shallow_1d/expected_sols.npy'), allow_pickle=True).item()
out['shallow_1d__sill_Fortran'] = np.asarray(d['sill_Fortran'], dtype=np.float64)            # shallow_1d/test_shallow1d.py:41-44
# euler_1d/test_shocktube.py:17-26: the Fortran HLLE run must stay within 0.04 (max norm) of the Python HLLC run, whose
# density is this stored vector (to 1e-4 in the 2-norm)
out['euler_1d__shocktube_regression_density'] = np.loadtxt(os.path.join(REF, 'euler_1d/shocktube_regression_density.txt'))
# advection_reaction_2d/test_advection_reaction.py:13-17: check_diff(reltol=1e-5) of q[0] against this 100x100 array
out['advection_reaction_2d__density'] = np.loadtxt(os.path.join(REF, 'advection_reaction_2d/advection_reaction.txt'))
HERE = os.path.dirname(os.path.abspath(__file__))
np.savez_compressed(os.path.join(HERE, 'reference_goldens4.npz'), **out)
for k, v in out.items():
    print(k, v.shape)
