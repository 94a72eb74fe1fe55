#!/bin/bash
# compute-sanitizer over the new code paths + the forced tile-ordering parity test + the N1 device-functor tests
out=gpurun_out; tag=r02_w
python -m pytest tests/test_cuda_parity_scale.py tests/test_cuda_parity_aux.py -x -q -m gpu -k "tile_ordering or device_resident" > $out/${tag}_pytest_new.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest_new.log
tail -5 $out/${tag}_pytest_new.log
for tool in memcheck racecheck; do
CLAW_TUNE=32 timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python scripts/sanitize_probe.py > $out/${tag}_sanitizer_$tool.log 2>&1; echo "$tool exit $?" >> $out/${tag}_sanitizer_$tool.log
tail -6 $out/${tag}_sanitizer_$tool.log
done
