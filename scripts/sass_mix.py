"""Static instruction mix of a kernel in a cubin / object file (no GPU needed):
    python scripts/sass_mix.py <file.cubin|.o> <substring of the mangled kernel name> [--loop]
Counts SASS opcodes (FP64 pipe: DFMA/DMUL/DADD/DSETP/MUFU.*64H; register moves; shuffles; shared / global accesses)
for the whole kernel and, with --loop, for the largest backward-branch loop (the march loop of the sweep kernels)."""
import collections
import re
import subprocess
import sys


def kernels(path):
    out = subprocess.run(['cuobjdump', '-sass', path], capture_output=True, text=True).stdout
    cur, body = None, {}
    for line in out.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = m.group(1)
            body[cur] = []
            continue
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', line)
        if m and cur:
            body[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return body


def mix(instrs):
    c = collections.Counter()
    for _, text in instrs:
        text = re.sub(r'^@!?U?P\w+\s+', '', text)
        c[text.split()[0]] += 1
    return c


def summarise(c):
    n = sum(c.values())
    grp = collections.OrderedDict()
    grp['fp64 (DFMA/DMUL/DADD)'] = sum(v for k, v in c.items() if k.split('.')[0] in ('DFMA', 'DMUL', 'DADD'))
    grp['fp64 compare (DSETP)'] = sum(v for k, v in c.items() if k.startswith('DSETP'))
    grp['MUFU'] = sum(v for k, v in c.items() if k.startswith('MUFU'))
    grp['register moves (MOV / IMAD.MOV / CS2R)'] = sum(v for k, v in c.items() if k.startswith(('MOV', 'IMAD.MOV', 'CS2R')))
    grp['selects (FSEL/SEL)'] = sum(v for k, v in c.items() if k.split('.')[0] in ('FSEL', 'SEL'))
    grp['SHFL'] = sum(v for k, v in c.items() if k.startswith('SHFL'))
    grp['shared (LDS/STS/LDGSTS)'] = sum(v for k, v in c.items() if k.split('.')[0] in ('LDS', 'STS', 'LDGSTS'))
    grp['global (LDG/STG)'] = sum(v for k, v in c.items() if k.split('.')[0] in ('LDG', 'STG'))
    grp['local / spill (LDL/STL)'] = sum(v for k, v in c.items() if k.split('.')[0] in ('LDL', 'STL'))
    grp['branch / sync'] = sum(v for k, v in c.items() if k.split('.')[0] in ('BRA', 'BSSY', 'BSYNC', 'BAR', 'EXIT', 'WARPSYNC'))
    other = n - sum(grp.values())
    print('  total %d' % n)
    for k, v in grp.items():
        print('  %-42s %5d  %5.1f %%' % (k, v, 100.0 * v / max(n, 1)))
    print('  %-42s %5d  %5.1f %%' % ('other integer / address', other, 100.0 * other / max(n, 1)))


if __name__ == '__main__':
    path, pat = sys.argv[1], sys.argv[2]
    for name, ins in kernels(path).items():
        if pat not in name:
            continue
        print(name)
        summarise(mix(ins))
        if '--loop' in sys.argv:
            best = None
            for addr, text in ins:
                m = re.search(r'BRA\s+(?:\w+,\s*)?(0x[0-9a-f]+)', text)
                if m and int(m.group(1), 16) < addr:
                    span = (int(m.group(1), 16), addr)
                    if best is None or span[1] - span[0] > best[1] - best[0]:
                        best = span
            if best:
                loop = [(a, t) for a, t in ins if best[0] <= a <= best[1]]
                print(' largest loop 0x%x..0x%x' % best)
                summarise(mix(loop))
