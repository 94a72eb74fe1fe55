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

d = np.load(os.path.join(REF, 'euler_3d/expected_sols.npy'), allow_pickle=True).item()
out['euler_3d__shocktube'] = np.asarray(d['shocktube'], dtype=np.float64)
out['euler_3d__shock_bubble'] = np.asarray(d['shock_bubble'], dtype=np.float64)


def shock_bubble_3d_initial(num_cells=(64, 16, 16)):
    """Initial state of pyclaw/examples/euler_3d/shock_bubble.py:105-141 (sub-cell bubble fractions by
    scipy dblquad), produced by IMPORTING the reference module for its constants and its `bubble`
    integrand; the grid formulas are Dimension.centers (geometry.py:546-553)."""
    import sys
    sys.path.insert(0, os.path.join(REF, 'euler_3d'))
    import shock_bubble as sb          # the reference example (imports only numpy + scipy at module level)
    from scipy import integrate
    mx, my, mz = num_cells
    lower, upper = (0., 0., 0.), (2., 0.5, 0.5)
    delta = [(u - l) / float(n) for l, u, n in zip(lower, upper, num_cells)]
    cs = [np.array([l + (i + 0.5) * dl for i in range(n)]) for l, dl, n in zip(lower, delta, num_cells)]
    X, Y, Z = np.meshgrid(*cs, indexing='ij')
    r0 = np.sqrt((X - sb.x0) ** 2 + (Y - sb.y0) ** 2 + (Z - sb.z0) ** 2)
    q = np.zeros((5, mx, my, mz), order='F')
    q[0] = sb.rho_shock * (X < sb.xshock) + sb.rhoout * (X >= sb.xshock)
    q[1] = sb.rho_shock * sb.v_shock * (X < sb.xshock)
    q[4] = sb.e_shock * (X < sb.xshock) + sb.pout / sb.gamma1 * (X >= sb.xshock)
    dx, dy, dz = delta
    dx2, dy2, dz2 = [dl / 2. for dl in delta]
    dmax = max(delta)
    for i in range(mx):
        for j in range(my):
            for k in range(mz):
                if r0[i, j, k] - dmax > sb.r_bubble:
                    continue
                xdown = X[i, j, k] - dx2
                xup = X[i, j, k] + dx2
                ydown = lambda x: Y[i, j, k] - dy2
                yup = lambda x: Y[i, j, k] + dy2
                zdown = Z[i, j, k] - dz2
                zup = Z[i, j, k] + dz2
                infrac, abserr = integrate.dblquad(sb.bubble, xdown, xup, ydown, yup, args=(zdown, zup, 0),
                                                   epsabs=1.e-3, epsrel=1.e-2)
                infrac = infrac / (dx * dy * dz)
                q[0, i, j, k] = sb.rhoout * (1. - infrac) + sb.rhoin * infrac
                q[4, i, j, k] = (sb.pout * (1. - infrac) + sb.pin * infrac) / sb.gamma1
    consts = np.array([sb.gamma, sb.rho_shock, sb.v_shock, sb.e_shock])
    return q, consts


q0, consts = shock_bubble_3d_initial()
out['euler_3d__shock_bubble_q0'] = q0
out['euler_3d__shock_bubble_consts'] = consts      # gamma, rho_shock, v_shock, e_shock (inflow BC state)

np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'reference_goldens.npz'), **out)
for k, v in out.items():
    print(k, v.shape, v.dtype)
