"""world_size-2 and -3 runs of the slab-decomposed host logic on CPU (gloo backend): partition, ghost
exchange, boundary ownership, periodic wrap-around and the CFL max all-reduce must reproduce the
serial run bit for bit (the unsplit and 2-D split algorithms are decomposition-invariant,
SURVEY.md appendix A.15)."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _launch(case, world, port, outdir=None, fmt=None):
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), OMP_NUM_THREADS='1')
        if outdir:
            env['CLAW_TEST_OUTDIR'] = outdir
        if fmt:
            env['CLAW_TEST_FORMAT'] = fmt
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, 'dist_worker.py'), case, str(port)],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
    return [p.returncode for p in procs], outs


@pytest.mark.parametrize('case,world,port', [
    ('advection_unsplit_periodic', 2, 29611),
    ('euler_unsplit_extrap', 2, 29612),
    ('shallow_split_wall', 2, 29613),
    ('euler_unsplit_extrap', 3, 29614),
    ('advection_unsplit_periodic', 3, 29615),
    ('acoustics3d_wall_periodic', 2, 29616),
    ('acoustics3d_periodic', 2, 29617),
    ('acoustics3d_wall_periodic', 3, 29618),
    ('euler3d_unsplit_extrap', 2, 29619),
    ('elasticity_unsplit_aux', 2, 29622),
    ('elasticity_unsplit_aux', 3, 29623),
    ('kpp_split', 2, 29624),
])
def test_slab_decomposition_matches_serial(case, world, port, oracle_lib):
    codes, outs = _launch(case, world, port)
    assert codes == [0] * world, '\n'.join(outs)
    assert 'q bit-identical to serial: True, step log identical: True' in outs[0], outs[0]


def test_distributed_frame_output(oracle_lib, tmp_path):
    """Solution.write on a slab-decomposed run (the replacement of petclaw/fileio + get_q_global,
    petclaw/state.py:267-287): slabs gathered, rank 0 writes the serial ascii layout of the GLOBAL grid."""
    out = str(tmp_path / 'frames')
    codes, outs = _launch('shallow_split_wall', 2, 29621, outdir=out)
    assert codes == [0, 0], '\n'.join(outs)
    assert 'frame file of the gathered solution matches the serial result: True' in outs[0], outs[0]
    assert sorted(f for f in os.listdir(out) if f.startswith('fort.q')) == ['fort.q0000', 'fort.q0001', 'fort.q0002']


def test_distributed_binary_frame_output(oracle_lib, tmp_path):
    """The same with output_format='binary': the gathered frame is the serial result bit for bit."""
    out = str(tmp_path / 'frames')
    codes, outs = _launch('elasticity_unsplit_aux', 2, 29625, outdir=out, fmt='binary')
    assert codes == [0, 0], '\n'.join(outs)
    assert 'frame file of the gathered solution matches the serial result: True' in outs[0], outs[0]
    assert sorted(f for f in os.listdir(out) if f.startswith('fort.b')) == ['fort.b0000', 'fort.b0001', 'fort.b0002']
