#!/bin/bash
# tools/run_final_checks.sh tag : the round-end sequence on one B200 -- GPU tests, bench line, reference arm, ncu launch list and
# --set full capture of the three hot kernels (summarise afterwards with `python profiles/summarize.py <tag>`)
tag=${1:-final}
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_$tag.log 2>&1; tail -2 gpurun_out/pytest_$tag.log
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo bench=$?; cut -c1-300 gpurun_out/bench_$tag.json
python bench.py --impl reference > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; cut -c1-200 gpurun_out/bench_ref_$tag.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_seed|k_march|k_eval3" --launch-skip 9 -c 3 -f -o gpurun_out/prof_$tag python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_$tag.log 2>&1
tail -1 gpurun_out/ncu_full_$tag.log
