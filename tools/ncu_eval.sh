#!/bin/bash
# ncu --set full of the evaluation / walk kernels on cfg3: tools/ncu_eval.sh tag "kernel regex" [extra python args]
tag=$1; rx=$2; shift 2
ncu --set full --clock-control none --import-source on -k regex:"$rx" --launch-skip 6 -c 3 -f -o gpurun_out/prof_$tag python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline "$@" > gpurun_out/ncu_$tag.log 2>&1
tail -3 gpurun_out/ncu_$tag.log
