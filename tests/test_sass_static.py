"""Static guards on the compiled hot kernels (no GPU needed: cuobjdump on the in-tree objects, scripts/sass_mix.py).
They pin the properties the measured numbers in profiles/ depend on, so that a source change which silently adds
spills, drops the cp.async prefetch ring or bloats the march loop is caught on the CPU."""
import os
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'scripts'))
import sass_mix  # noqa: E402

OBJ = os.path.join(ROOT, 'clawpack_b200', 'csrc', 'build', 'sweep_inst_fast.o')
pytestmark = pytest.mark.skipif(shutil.which('cuobjdump') is None or not os.path.exists(OBJ),
                                reason='needs cuobjdump and the built objects (python -c "import __graft_entry__ as g; g.build()")')


@pytest.fixture(scope='module')
def fast_kernels():
    return sass_mix.kernels(OBJ)


def _count(c, pred):
    return sum(v for k, v in c.items() if pred(k))


def test_headline_euler_kernel_shape(fast_kernels):
    """sweep2_unsplit_ws<RpEuler4Wave2DT<3>,1,8,MC|spec|cp.async ring,QD=8> (fast): the kernel bench.py times."""
    names = [k for k in fast_kernels if 'sweep2_unsplit_wsINS0_15RpEuler4Wave2DTILi3ELi0EEELi1ELi8ELi71ELi1ELi8ELi0E' in k]
    assert len(names) == 1, names
    c = sass_mix.mix(fast_kernels[names[0]])
    n = sum(c.values())
    fp64 = _count(c, lambda k: k.split('.')[0] in ('DFMA', 'DMUL', 'DADD'))
    assert _count(c, lambda k: k.split('.')[0] in ('LDL', 'STL')) == 0          # spill-free at 255 registers
    assert _count(c, lambda k: k.startswith('LDGSTS')) >= 4                      # cp.async row prefetch ring
    assert _count(c, lambda k: k.startswith('SHFL')) >= 90                       # neighbour exchange by shuffle
    assert _count(c, lambda k: k.startswith('BAR')) <= 2                         # no barrier in the march loop
    # measured build (round 2, with the uniform-region fast path, its tight inner loop for runs of quiet rows and the
    # 8-stage ring, tile ordering, wide tight loop as a local function): ~4500 instructions, ~940 of them FP64 arithmetic; the full march step is ~1420
    # instructions with ~800 FP64, the quiet step ~160 with 10, a row of the tight loop 75
    assert n <= 4800 and 800 <= fp64 <= 1000, (n, fp64)
    assert _count(c, lambda k: k.startswith('VOTE')) >= 2                        # the warp votes of the fast path
    assert _count(c, lambda k: k.startswith('MUFU.RCP64H') or k.startswith('MUFU.RSQ64H')) <= 32   # shared reciprocals


def test_no_fast_sweep_kernel_of_the_baseline_configs_spills(fast_kernels):
    """Default kernels of BASELINE configs 2/5 (Euler unsplit), 3 (shallow split) and 1 (acoustics split)."""
    for pat in ('sweep2_unsplit_wsINS0_15RpEuler4Wave2DTILi3ELi0EEELi1ELi8ELi71ELi1ELi8ELi0E',
                'sweep2_split_wsINS0_18RpShallowRoeEfix2DELi1ELi16ELi1ELi1ELi4E',
                'sweep2_split_wsINS0_13RpAcoustics2D'):
        names = [k for k in fast_kernels if pat in k]
        assert names, pat
        for k in names:
            c = sass_mix.mix(fast_kernels[k])
            assert _count(c, lambda x: x.split('.')[0] in ('LDL', 'STL')) == 0, k


def test_specialised_split_kernel_is_smaller_than_the_generic_one(fast_kernels):
    """The default kernel of BASELINE config 3 (limiter as a template argument) against the generic shallow-water split kernel."""
    # (both without the uniform-region fast path, QD = 0: the default adds ~900 instructions of quiet-row handling)
    gen = [k for k in fast_kernels if 'sweep2_split_wsINS0_18RpShallowRoeEfix2DELi1ELi16ELi1ELi1ELi0ELi0E' in k]
    spec = [k for k in fast_kernels if 'sweep2_split_wsINS0_18RpShallowRoeEfix2DELi1ELi16ELi1ELi1ELi4ELi0E' in k]
    assert len(gen) == 1 and len(spec) == 1
    assert len(fast_kernels[spec[0]]) < 0.75 * len(fast_kernels[gen[0]])


def test_specialised_3d_kernels_are_smaller_and_spill_free():
    """The default kernels of BASELINE config 4 (method set as template arguments) against the generic vc_acoustics_3D kernels."""
    obj = os.path.join(ROOT, 'clawpack_b200', 'csrc', 'build', 'sweep3_inst_fast.o')
    if not os.path.exists(obj):
        pytest.skip('sweep3_inst_fast.o not built')
    k = sass_mix.kernels(obj)
    for gen, spec in (('sweep3_xINS0_15RpVcAcoustics3DELi8ELi1ELi0E', 'sweep3_xINS0_15RpVcAcoustics3DELi8ELi1ELi41E'),
                      ('sweep3_marchINS0_15RpVcAcoustics3DELi1ELi32ELi8ELi1ELi0E',
                       'sweep3_marchINS0_15RpVcAcoustics3DELi1ELi32ELi8ELi1ELi41E'),
                      ('sweep3_marchINS0_15RpVcAcoustics3DELi2ELi32ELi8ELi1ELi0E',
                       'sweep3_marchINS0_15RpVcAcoustics3DELi2ELi32ELi8ELi1ELi41E')):
        g = [n for n in k if gen in n]
        s = [n for n in k if spec in n]
        assert len(g) == 1 and len(s) == 1, (gen, g, s)
        assert len(k[s[0]]) < 0.82 * len(k[g[0]]), (gen, len(k[g[0]]), len(k[s[0]]))
        c = sass_mix.mix(k[s[0]])
        assert _count(c, lambda x: x.split('.')[0] in ('LDL', 'STL')) == 0, spec


def test_tma_ring_variant_uses_bulk_copies_and_mbarriers(fast_kernels):
    """CLAW_VARIANT=91 (sweep2_unsplit_ws<..., TMA=1>): the row ring is fed by cp.async.bulk (UBLKCP) with mbarrier
    completion (SYNCS); the default kernel has neither (measured slower on B200)."""
    tma = [k for k in fast_kernels if 'RpEuler4Wave2DTILi3ELi0EEELi1ELi8ELi71ELi1ELi8ELi1E' in k]
    dflt = [k for k in fast_kernels if 'RpEuler4Wave2DTILi3ELi0EEELi1ELi8ELi71ELi1ELi8ELi0E' in k]
    assert len(tma) == 1 and len(dflt) == 1
    c = sass_mix.mix(fast_kernels[tma[0]])
    assert _count(c, lambda k: k.startswith('UBLKCP')) >= 4 and _count(c, lambda k: k.startswith('SYNCS')) >= 2
    c = sass_mix.mix(fast_kernels[dflt[0]])
    assert _count(c, lambda k: k.startswith('UBLKCP')) == 0 and _count(c, lambda k: k.startswith('LDGSTS')) >= 4
