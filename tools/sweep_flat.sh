#!/bin/bash
# Tuning aid (run under gpurun): 1M-box step time for each flat-primal configuration (threads per block, blocks per SM).
for v in 1286 1285 1284; do
  export AVBD_FLAT=$v
  echo "== $v"; timeout 300 python bench.py --no-cpu-baseline --steps 10 --warmup 3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print('ms/step', round(d['ms_per_step'], 3), 'primal avg launch ms', round(r['avg_launch_ms'], 4), 'frac', round(r['frac'], 3), 'primal_dual ms', round(d['stage_ms']['primal_dual'], 3), 'contacts', d['config']['contacts'], 'maxPen', d['diagnostics']['maxPen'])
"
done
