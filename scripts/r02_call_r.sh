#!/bin/bash
# what limits the uniform-region path of the shallow-water split kernel at 4.3 TB/s?  ring depth / warps per SM (narrow loop: CLAW_TUNE=4)
out=gpurun_out; tag=r02_r
for v in 0 81 82 83; do
CLAW_TUNE=4 CLAW_VARIANT=$v python scripts/quiet_probe.py --config c3 --seg 64 --data uniform >> $out/${tag}_c3.jsonl 2>>$out/${tag}.err
done
CLAW_TUNE=4 CLAW_VARIANT=82 python scripts/quiet_probe.py --config c3 --seg 32 128 --data uniform >> $out/${tag}_c3.jsonl 2>>$out/${tag}.err
cat $out/${tag}_*.jsonl
tail -5 $out/${tag}.err
