python bench.py > gpurun_out/bench_r1_v.json 2> gpurun_out/bench_r1_v.err; echo bench=$?; cat gpurun_out/bench_r1_v.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_v.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench_v.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_seed|k_march|k_eval3" --launch-skip 9 -c 3 -f -o gpurun_out/prof_r1_v python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_v.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r1_v.json 2>/dev/null; cat gpurun_out/bench_ref_r1_v.json
