"""Build libclawcuda.so in-tree for sm_100a with nvcc (cross-compiles without a GPU).

The kernel translation unit is compiled twice:
  strict: -fmad=false  -> namespace clawcuda::strict (op-for-op the reference's unfused arithmetic)
  fast:   -fmad=true   -> namespace clawcuda::fast   (FMA contraction)
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), 'libclawcuda.so')
EXTRA = os.environ.get('CLAW_EXTRA_NVCC', '').split()     # tuning experiments only
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
HOSTCXX = '/usr/bin/g++' if os.path.exists('/usr/bin/g++') else 'g++'
COMMON = ['nvcc', '-ccbin', HOSTCXX] + ARCH + ['-lineinfo', '-O3', '-std=c++17', '-Xcompiler', '-fPIC', '--threads', '4']

SOURCES = ['clawcuda.cu', 'sweep_inst.cu', 'sweep3_inst.cu', 'sweep2.cuh', 'sweep3.cuh', 'rp2.cuh', 'rp3.cuh', 'launch.h',
           '../../include/clawcuda.h']
KERNEL_UNITS = ['sweep_inst.cu', 'sweep3_inst.cu', 'sweep1_inst.cu']


def _stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    for f in os.listdir(HERE):
        if f.endswith(('.cu', '.cuh', '.h', '.py')) and os.path.getmtime(os.path.join(HERE, f)) > t:
            return True
    return os.path.getmtime(os.path.join(HERE, '../../include/clawcuda.h')) > t


def build(force=False, verbose=False):
    if not force and not _stale():
        return OUT
    bdir = os.path.join(HERE, 'build')
    os.makedirs(bdir, exist_ok=True)
    objs = []
    jobs = []
    extra = ['-Xptxas', '-v'] if verbose else []
    for unit in KERNEL_UNITS:
        if not os.path.exists(os.path.join(HERE, unit)):
            continue
        stem = unit[:-3]
        for ns, flags in (('strict', ['-DCLAW_NS=strict', '-DCLAW_STRICT=1', '-fmad=false']),
                          ('fast', ['-DCLAW_NS=fast', '-fmad=true'])):
            o = os.path.join(bdir, '%s_%s.o' % (stem, ns))
            jobs.append(subprocess.Popen(COMMON + extra + EXTRA + flags + ['-c', unit, '-o', o], cwd=HERE))
            objs.append(o)
    o = os.path.join(bdir, 'clawcuda.o')
    have3 = os.path.exists(os.path.join(HERE, 'sweep3_inst.cu'))
    jobs.append(subprocess.Popen(COMMON + (['-DCLAW_HAVE_3D=1'] if have3 else []) + ['-c', 'clawcuda.cu', '-o', o], cwd=HERE))
    objs.append(o)
    for j in jobs:
        if j.wait() != 0:
            raise RuntimeError('nvcc failed')
    subprocess.check_call(['nvcc', '-ccbin', HOSTCXX] + ARCH + ['-shared', '-o', OUT] + objs + ['-lcudart', '-ldl'], cwd=HERE)
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
