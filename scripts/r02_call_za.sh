#!/bin/bash
# TMA row ring (cp.async.bulk + mbarrier, variants 91 / 92 = 8- / 16-stage) against the per-lane cp.async ring (default, 73 = 16-stage)
out=gpurun_out; tag=r02_za
for v in 0 91 92 73; do
CLAW_VARIANT=$v timeout 300 python scripts/quiet_probe.py --config c2 --seg 64 --data uniform actual >> $out/${tag}_c2.jsonl 2>>$out/${tag}.err
done
for v in 0 91; do
CLAW_VARIANT=$v timeout 300 python scripts/quiet_probe.py --config c2 --seg 64 --data dense --reps 3 >> $out/${tag}_c2_dense.jsonl 2>>$out/${tag}.err
done
CLAW_VARIANT=91 timeout 300 python -m pytest tests/test_cuda_parity_scale.py -m gpu -x -q -k "c2_euler" > $out/${tag}_parity.log 2>&1
tail -3 $out/${tag}_parity.log
cat $out/${tag}_*.jsonl | cut -c1-260
tail -5 $out/${tag}.err
