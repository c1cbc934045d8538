#!/bin/bash
# Runs ON the GPU box: headline step time for each A/B build under avbd-demo3d_b200/variants (make variant NAME=.. SOLVE_DEFS=..).
# usage: sweep_libs.sh <tag> [extra env assignments...]
tag=$1; shift
out=gpurun_out/${tag}_libs.log; : > $out
for lib in avbd-demo3d_b200/variants/*.so; do
  echo "== $lib $@" >> $out
  env AVBD_B200_LIB=$PWD/$lib "$@" python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline 2>>$out | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('ms_per_step', round(d['ms_per_step'], 3), 'primal', round(d['stage_ms']['primal'], 3), 'frac', round(d['roofline']['frac'], 3), 'graph', round(d['stage_ms']['graph'], 3))" >> $out
done
cat $out
