#!/bin/bash
# round 2, call k: tight quiet loops -- parity subset + bench
out=gpurun_out; tag=r02_k
python -m pytest tests/test_cuda_parity_scale.py tests/test_cuda_parity.py tests/test_clawpack_alias.py -q -m gpu -x > $out/${tag}_pytest.log 2>&1; tail -4 $out/${tag}_pytest.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $out/${tag}_bench.json 2> $out/${tag}_bench.err; tail -3 $out/${tag}_bench.err
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r02_k_bench.json').read().splitlines() if l.startswith('{')][-1])
r=d['roofline']
print('c2', d['value'], d['ms_per_step'], r['bound'], r['frac'], r['kernel_ms'], 'dense', r.get('dense',{}).get('kernel_ms'), r.get('dense',{}).get('frac'), 'launches', d['gpu_launches'])
for o in d['other_configs']:
    print(o['config'], o['value'], o['ms_per_step'], o['roofline']['bound'], o['roofline']['frac'], o['roofline']['kernel_ms'], o['roofline'].get('dense',{}).get('kernel_ms'))
P
for sr in 96 128; do python bench.py --steps 20 --warmup 3 --seg-rows $sr --no-e2e --no-cpu-baseline --no-other-configs --no-parity --no-dense 2>>$out/${tag}.err | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); r=d['roofline']
print('seg $sr', r['kernel_ms'], d['ms_per_step'], r['hbm']['frac'])"; done
