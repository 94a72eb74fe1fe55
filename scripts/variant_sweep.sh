#!/bin/bash
# Times the Euler 8192^2 sweep kernel for each tuning variant (CLAW_VARIANT, see csrc/sweep_inst.cu).
# usage: scripts/variant_sweep.sh "0 1 2 ..." out.jsonl
out=${2:-gpurun_out/variants.jsonl}
: > $out
for v in $1; do
  CLAW_VARIANT=$v python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>>$out.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print(json.dumps({'variant':$v,'kernel':r['kernel'],'launch':r['launch'],'kernel_ms':r['kernel_ms'],'ms_per_step':d['ms_per_step'],'value':d['value'],'frac':r['frac'],'cfl':d['config']['last_cfl']}))" >> $out
done
cat $out
