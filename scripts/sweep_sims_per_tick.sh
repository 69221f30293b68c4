for m in 2 4 8 16 32; do
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-aux-rooflines --max-sims-per-tick $m 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print($m, round(d['value']), round(d['ms_per_step'],2), round(d['leaf_evals_per_sec']), d['clocks']['sm_mhz'])"
done
