import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from clawpack_b200 import scenarios as S, riemann
from clawpack_b200.pyclaw import cuda_classic
from oracle_fmod import OracleFmod
from refutil import ref_solver
sc = S.radial_dam_break()
r = ref_solver(sc)
r.evolve_to_time()   # rejected
r.get_dt(0, 0, None, True)
r.apply_bcs()
dt = r.dt
print('dt', dt)
rp = riemann.shallow_roe_with_efix_2D; rp.cparam.grav = 1.0
dx, dy = r.delta
method = np.array([1, 2, -1, 0, 0, 0, 0], dtype=np.intc); mth = [4, 4, 4]
cu, orc = cuda_classic.fmod(2), OracleFmod()
mx, my = 150, 150
q0 = r.qbc.copy('F')
qa, ca = cu.step2ds(150, 2, mx, my, q0.copy('F'), q0.copy('F'), None, dx, dy, dt, method, mth, 0, 0, 0, 0, 1, False, rp, rp)
qb, cb = orc.step2ds(150, 2, mx, my, q0.copy('F'), q0.copy('F'), None, dx, dy, dt, method, mth, 0, 0, 0, 0, 1, False, rp, rp)
print('x-sweep equal', np.array_equal(qa, qb), ca == cb, np.argwhere(qa != qb)[:10].tolist())
print('signed zero diffs', np.argwhere(np.signbit(qa) != np.signbit(qb))[:10].tolist())
qa2, ca2 = cu.step2ds(150, 2, mx, my, qb.copy('F'), qb.copy('F'), None, dx, dy, dt, method, mth, 0, 0, 0, 0, 2, False, rp, rp)
qb2, cb2 = orc.step2ds(150, 2, mx, my, qb.copy('F'), qb.copy('F'), None, dx, dy, dt, method, mth, 0, 0, 0, 0, 2, False, rp, rp)
bad = np.argwhere(qa2 != qb2)
print('y-sweep equal', np.array_equal(qa2, qb2), ca2 == cb2, bad[:10].tolist())
np.set_printoptions(precision=17, linewidth=200)
for b in bad[:2]:
    m, i, j = b
    print('cell', b, 'gpu', repr(qa2[m, i, j]), 'ref', repr(qb2[m, i, j]))
    print('column stencil q(:, i, j-2..j+2):')
    print(repr(qb[:, i, j-2:j+3]))
np.save('gpurun_out/dbg_qx.npy', qb); np.save('gpurun_out/dbg_gpu.npy', qa2); np.save('gpurun_out/dbg_ref.npy', qb2)
