#!/bin/bash
# Runs ON the GPU box (under gpurun): launch list of one 1M-box step + ncu --set full of every stage's top kernel.
#   gpurun --timeout 1500 -- 'bash profiles/capture.sh r01 [solve|collide|launches ...]'
# Outputs land in gpurun_out/; summarise them here with profiles/summarize.py and commit the summaries.
tag=${1:-r01}; shift
what=${@:-launches solve collide}
mkdir -p gpurun_out
export AVBD_PROFILE_RANGE=1
T="python tools/profile_target.py grid100 1"
NCU="ncu --clock-control none --profile-from-start off"
for w in $what; do case $w in
launches)  # every launch of ONE steady-state step with its device time
  timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/${tag}_launches.csv $T > gpurun_out/${tag}_launches.log 2>&1;;
solve)     # first iteration of the step: every colour's {visit sums, block solve} + the dual pass
  timeout 900 $NCU --set full --import-source on -k 'regex:primal_|dual_contacts' -c ${SOLVE_COUNT:-9} -o gpurun_out/${tag}_solve -f $T > gpurun_out/${tag}_solve.log 2>&1;;
collide)   # broadphase sweep, SAT cull, manifold build, graph kernels of the same step
  timeout 900 $NCU --set full --import-source on -k 'regex:bp_|np_|visit_fill|entry_fill|colour_round|velocity_bodies|predict_bodies' -c 16 -o gpurun_out/${tag}_collide -f $T > gpurun_out/${tag}_collide.log 2>&1;;
esac; done
ls -la gpurun_out
