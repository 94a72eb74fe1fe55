#!/bin/bash
# Riemann-solver plug-ins on the GPU; dense-data cost of the uniform-region fast path (variant 2 = same kernel without it)
out=gpurun_out; tag=r02_zc
python -m pytest tests/test_rp_plugin.py tests/test_cuda_parity.py -x -q -m gpu > $out/${tag}_pytest_plugin.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest_plugin.log
tail -4 $out/${tag}_pytest_plugin.log
for v in 0 2; do
CLAW_VARIANT=$v timeout 300 python scripts/quiet_probe.py --config c2 --seg 0 --data dense actual --reps 3 >> $out/${tag}_c2_dense.jsonl 2>>$out/${tag}.err
CLAW_VARIANT=$v timeout 300 python scripts/quiet_probe.py --config c3 --seg 0 --data dense --reps 3 >> $out/${tag}_c3_dense.jsonl 2>>$out/${tag}.err
done
cat $out/${tag}_*.jsonl | cut -c1-260
tail -5 $out/${tag}.err
