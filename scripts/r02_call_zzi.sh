#!/bin/bash
# dense-data switch with probes (test), then the whole suite + bench once more on the committed build
out=gpurun_out; tag=r02_zzi
python -m pytest tests -x -q -m gpu > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest_gpu.log
tail -4 $out/${tag}_pytest_gpu.log
python bench.py --steps 20 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
python -c "
import json
d=json.loads([l for l in open('$out/${tag}_bench.json').read().splitlines() if l.startswith('{')][-1])
r=d['roofline']; print('c2', d['value'], d['ms_per_step'], r['bound'], r['frac'], r['traffic'], r.get('ncu_stale'), r['dense']['kernel_ms'], r['dense']['frac'], d['e2e']['value'], d['gpu_launches'])
for o in d['other_configs']: print(o['config'], o['value'], o['ms_per_step'], o['roofline']['bound'], o['roofline']['frac'], o['roofline'].get('ncu_stale'))
"
tail -2 $out/${tag}_bench.err
