#!/bin/bash
out=gpurun_out; tag=r02_m
for c in c2 c3; do python bench.py --config $c --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-other-configs --no-parity --no-dense 2>>$out/${tag}.err | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); r=d['roofline']
print('$c', r['kernel_ms'], d['ms_per_step'], d['value'], r['hbm']['frac'], r['launch'])"; done
