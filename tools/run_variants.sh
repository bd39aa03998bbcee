#!/bin/bash
# tools/run_variants.sh tag [reps] [k=v,k=v ...]: tools/exp_opts.py on cfg3 with every library under build_variants/;
# output in gpurun_out/<tag>.log
tag=$1; reps=${2:-30}; shift 2
mkdir -p gpurun_out
for f in build_variants/*.so; do RT_B200_LIB=$PWD/$f timeout 150 python tools/exp_opts.py cfg3 $reps "$@" 2>&1 | grep -v "^$" | tail -12; done > gpurun_out/$tag.log 2>&1
cat gpurun_out/$tag.log
