#!/bin/bash
# round 2, call h: ncu --set full of the Euler kernel with the uniform-region fast path (ring 8)
out=gpurun_out; tag=r02_h
CLAW_VARIANT=72 ncu --set full --clock-control none --import-source on -k regex:sweep2_unsplit_ws -s 6 -c 1 -o $out/${tag}_euler_quiet8 -f \
   python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-other-configs --no-parity > $out/${tag}_ncu.log 2>&1
tail -3 $out/${tag}_ncu.log; ls -la $out/${tag}_euler_quiet8.ncu-rep
