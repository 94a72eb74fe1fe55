#!/bin/bash
out=gpurun_out; tag=r02_l
for v in 0 73; do CLAW_VARIANT=$v python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-other-configs --no-parity --no-dense 2>>$out/${tag}.err | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); r=d['roofline']
print('variant $v', r['kernel_ms'], d['ms_per_step'], r['hbm']['frac'], r['launch'])"; done
ncu --set full --clock-control none --import-source on -k regex:sweep2_unsplit_ws -s 6 -c 1 -o $out/${tag}_euler_quiet -f \
   python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-other-configs --no-parity --no-dense > $out/${tag}_ncu.log 2>&1
tail -2 $out/${tag}_ncu.log
