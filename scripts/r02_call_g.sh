#!/bin/bash
# round 2, call g: uniform-region fast path of the Euler kernel -- timing (ring depth 2/4/8) and parity with it switched on
out=gpurun_out; tag=r02_g
: > $out/${tag}_variants.jsonl
for v in 0 70 71 72; do
  CLAW_VARIANT=$v python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-other-configs --no-parity 2>>$out/${tag}.err | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); r=d['roofline']
print(json.dumps({'variant':$v,'launch':r['launch'],'kernel_ms':r['kernel_ms'],'ms_per_step':d['ms_per_step'],'value':d['value'],'fp64_frac':r['frac'],'hbm':r['hbm'],'cfl':d['run']['last_cfl']}))" >> $out/${tag}_variants.jsonl
done
cat $out/${tag}_variants.jsonl
CLAW_VARIANT=72 python -m pytest tests/test_cuda_parity_scale.py tests/test_cuda_parity.py -q -m gpu -k "euler or c2 or shock" > $out/${tag}_pytest_quiet.log 2>&1; tail -5 $out/${tag}_pytest_quiet.log
