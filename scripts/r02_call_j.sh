#!/bin/bash
# round 2, call j: uniform-region fast path as the default (Euler unsplit, shallow / acoustics split): whole GPU suite + bench
out=gpurun_out; tag=r02_j
python -m pytest tests -q -m gpu -x > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest_gpu.log
tail -8 $out/${tag}_pytest_gpu.log
python bench.py --steps 20 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; tail -3 $out/${tag}_bench.err
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r02_j_bench.json').read().splitlines() if l.startswith('{')][-1])
r=d['roofline']
print('c2', d['value'], d['ms_per_step'], r['bound'], r['frac'], r['kernel_ms'], 'dense', r.get('dense'), 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])
for o in d['other_configs']:
    print(o['config'], o['value'], o['ms_per_step'], o['roofline']['bound'], o['roofline']['frac'], o['roofline']['kernel_ms'], o['roofline'].get('dense'))
P
