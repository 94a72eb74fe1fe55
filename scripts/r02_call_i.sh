#!/bin/bash
out=gpurun_out; tag=r02_i
: > $out/${tag}_segrows.jsonl
for sr in 0 128 64 32 16; do
  CLAW_VARIANT=72 python bench.py --steps 20 --warmup 3 --seg-rows $sr --no-e2e --no-cpu-baseline --no-other-configs --no-parity 2>>$out/${tag}.err | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); r=d['roofline']
print(json.dumps({'seg_rows':$sr,'launch':r['launch'],'kernel_ms':r['kernel_ms'],'ms_per_step':d['ms_per_step'],'value':d['value'],'hbm_frac':r['hbm']['frac']}))" >> $out/${tag}_segrows.jsonl
done
cat $out/${tag}_segrows.jsonl
