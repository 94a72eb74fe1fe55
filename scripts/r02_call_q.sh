#!/bin/bash
# wide tight loop (16-byte loads/stores, register pipeline): A/B through CLAW_TUNE (4: off in the split kernel, 8: off in the unsplit kernel)
out=gpurun_out; tag=r02_q
CLAW_TUNE=0 python scripts/quiet_probe.py --config c3 --seg 32 64 128 --data actual uniform > $out/${tag}_c3_tune0.jsonl 2>$out/${tag}.err
CLAW_TUNE=4 python scripts/quiet_probe.py --config c3 --seg 64 --data actual uniform > $out/${tag}_c3_tune4.jsonl 2>>$out/${tag}.err
CLAW_TUNE=0 python scripts/quiet_probe.py --config c2 --seg 32 64 128 --data actual uniform shock > $out/${tag}_tune0.jsonl 2>>$out/${tag}.err
CLAW_TUNE=8 python scripts/quiet_probe.py --config c2 --seg 64 --data actual uniform > $out/${tag}_tune8.jsonl 2>>$out/${tag}.err
cat $out/${tag}_*.jsonl
tail -5 $out/${tag}.err
python -m pytest tests -x -q -m gpu > $out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest.log
tail -5 $out/${tag}_pytest.log
