#!/bin/bash
# A/B of library builds on the headline loop: sims/s, ms/step, tower/heads/advance kernel ms.  usage: ab_bench.sh lib1 lib2 ...
for L in "$@"; do
  [ "$L" = default ] && P="" || P=$L
  SPX_LIB_PATH=$P timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 > /tmp/b.json
  python -c "
import json
d=json.load(open('/tmp/b.json')); r=d['roofline']
print('$L', round(d['value']), round(d['ms_per_step'],1), 'tower', round(r['kernel_ms'],4), 'heads', r['heads_kernel_ms'], 'advance', round(r['advance_kernel_ms'],4), 'search-only', round(d['search_roofline']['advance_ms_per_launch'],4), round(d['search_roofline']['with_16384_games']['advance_ms_per_launch'],4), d['clocks']['sm_mhz'])"
done
