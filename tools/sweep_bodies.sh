#!/bin/bash
# Runs ON the GPU box: headline step time for each configuration of the one-body-per-thread sweep kernel (AVBD_BODIES).
out=gpurun_out/${1:-sweep}_bodies.log; : > $out
for cfg in ${CFGS:-12841 12840 12831 12851 12850 6481 6461 25621}; do
  echo "== AVBD_BODIES=$cfg" >> $out
  AVBD_BODIES=$cfg python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline 2>>$out | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('ms_per_step', round(d['ms_per_step'], 3), 'stage_ms', {k: round(v, 3) for k, v in d['stage_ms'].items()}, 'primal frac', round(d['roofline']['frac'], 3), 'colours', d['workload_stats']['colours'])" >> $out
done
cat $out
