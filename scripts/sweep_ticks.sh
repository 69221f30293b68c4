#!/bin/bash
# sweep max_sims_per_tick: sims/s, ms/step, tower/heads/advance kernel ms
for m in 1 2 4 8 16; do
  timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --max-sims-per-tick $m 2>&1 | tail -1 > /tmp/b.json
  python -c "
import json
d=json.load(open('/tmp/b.json')); r=d['roofline']
print($m, round(d['value']), round(d['ms_per_step'],1), round(r['kernel_ms'],4), round(r['heads_kernel_ms'],4), round(r['advance_kernel_ms'],4))"
done
