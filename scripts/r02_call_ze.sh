#!/bin/bash
# 3-D kernels A/B: double exchange (one barrier per step) on the x-sweep only / nowhere; segment choice
out=gpurun_out; tag=r02_ze
B="--config c4 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-other-configs --no-parity"
run() { name=$1; shift; env "$@" python bench.py $B > $out/${tag}_$name.json 2>> $out/${tag}.err; }
run dbl1_auto CLAW_DBL3=1
run dbl0_auto CLAW_DBL3=0
run dbl0_seg512 CLAW_DBL3=0 CLAW_SEG3=512
run dbl1_seg128 CLAW_DBL3=1 CLAW_SEG3=128
run dbl1_seg64 CLAW_DBL3=1 CLAW_SEG3=64
python - <<PY
import json
for f in ('dbl1_auto','dbl0_auto','dbl0_seg512','dbl1_seg128','dbl1_seg64'):
    try:
        d=json.loads([l for l in open('$out/${tag}_%s.json'%f).read().splitlines() if l.startswith('{')][-1])
        print(f, d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['roofline']['launch'])
    except Exception as e: print(f, 'ERR', e)
PY
tail -3 $out/${tag}.err
CLAW_DBL3=1 python -m pytest tests/test_cuda_parity3d.py -x -q -m gpu 2>&1 | tail -2
