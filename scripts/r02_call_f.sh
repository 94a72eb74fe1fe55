#!/bin/bash
# round 2, call f: BASELINE-scale parity tests, alias test, Euler kernel experiments (warps per CTA / per-step barrier)
out=gpurun_out; tag=r02_f
python -m pytest tests/test_cuda_parity_scale.py tests/test_clawpack_alias.py -q -m gpu -x > $out/${tag}_pytest_scale.log 2>&1; tail -12 $out/${tag}_pytest_scale.log
: > $out/${tag}_variants.jsonl
for v in 0 60 61 62 63 64 65; do
  CLAW_VARIANT=$v python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-other-configs --no-parity 2>>$out/${tag}.err | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); r=d['roofline']
print(json.dumps({'variant':$v,'launch':r['launch'],'kernel_ms':r['kernel_ms'],'ms_per_step':d['ms_per_step'],'frac':r['frac'],'cfl':d['run']['last_cfl']}))" >> $out/${tag}_variants.jsonl
done
cat $out/${tag}_variants.jsonl
