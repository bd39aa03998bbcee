#!/bin/bash
# run exp_phases (minimal) with every library under build_variants/
for f in build_variants/*.so; do echo "== $f"; RT_B200_LIB=$PWD/$f RT_EXP_MIN=1 python tools/exp_phases.py cfg3 2>&1 | grep -E "${1:-default|two-stage  |eval_ms}"; done
