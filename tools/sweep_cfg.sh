#!/bin/bash
# Runs ON the GPU box: headline step time for each value of an environment variable.  usage: sweep_cfg.sh <tag> <VAR> <values...>
tag=$1; var=$2; shift 2
out=gpurun_out/${tag}_${var}.log; : > $out
for cfg in "$@"; do
  echo "== $var=$cfg" >> $out
  env $var=$cfg python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline 2>>$out | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('ms_per_step', round(d['ms_per_step'], 3), 'stage_ms', {k: round(v, 3) for k, v in d['stage_ms'].items()}, 'primal frac', round(d['roofline']['frac'], 3), 'colours', d['workload_stats']['colours'], 'e2e', round(d['e2e']['ms_per_step'],3))" >> $out
done
cat $out
