#!/bin/bash
# in-flight LDGSTS per SM: ring depth x warps per SM, both kernels; dense-path cost of fewer warps (c3)
out=gpurun_out; tag=r02_s
for v in 81 82 83 84 85; do
CLAW_TUNE=4 CLAW_VARIANT=$v python scripts/quiet_probe.py --config c3 --seg 32 64 --data uniform actual >> $out/${tag}_c3.jsonl 2>>$out/${tag}.err
done
for v in 81 82; do
CLAW_TUNE=0 CLAW_VARIANT=$v python scripts/quiet_probe.py --config c3 --seg 32 --data uniform >> $out/${tag}_c3_wide.jsonl 2>>$out/${tag}.err
done
for v in 0 82 83; do
CLAW_TUNE=4 CLAW_VARIANT=$v python scripts/quiet_probe.py --config c3 --seg 64 --data dense --reps 3 >> $out/${tag}_c3_dense.jsonl 2>>$out/${tag}.err
done
for v in 0 71; do
CLAW_TUNE=8 CLAW_VARIANT=$v python scripts/quiet_probe.py --config c2 --seg 32 64 --data uniform actual >> $out/${tag}_c2.jsonl 2>>$out/${tag}.err
done
CLAW_TUNE=0 CLAW_VARIANT=71 python scripts/quiet_probe.py --config c2 --seg 64 --data uniform actual >> $out/${tag}_c2_wide.jsonl 2>>$out/${tag}.err
cat $out/${tag}_*.jsonl | cut -c1-220
tail -5 $out/${tag}.err
