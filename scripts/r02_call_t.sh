#!/bin/bash
out=gpurun_out; tag=r02_t
for t in 0 16; do CLAW_TUNE=$t python bench.py --config c1 --steps 200 --warmup 3 --no-e2e --no-cpu-baseline --no-other-configs --no-parity --no-dense 2>>$out/${tag}.err | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); r=d['roofline']
print('tune $t', r['kernel_ms'], d['ms_per_step'], d['gpu_launches'], r['launch'])"; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/${tag}_launches_c1.csv python bench.py --config c1 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-other-configs --no-parity --no-dense > /dev/null 2>>$out/${tag}.err
tail -12 $out/${tag}_launches_c1.csv | cut -c1-200
python bench.py --steps 20 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
python -c "
import json
d=json.loads([l for l in open('$out/${tag}_bench.json').read().splitlines() if l.startswith('{')][-1])
r=d['roofline']; print('c2', d['value'], d['ms_per_step'], r['bound'], r['frac'], r['kernel_ms'], r.get('dense',{}).get('kernel_ms'), d['e2e'])
for o in d['other_configs']: print(o['config'], o['value'], o['ms_per_step'], o['roofline']['bound'], o['roofline']['frac'], o['roofline'].get('dense',{}).get('kernel_ms'))
"
