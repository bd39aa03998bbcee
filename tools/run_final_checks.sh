python bench.py > gpurun_out/bench_r1_w.json 2> gpurun_out/bench_r1_w.err; echo bench=$?; cat gpurun_out/bench_r1_w.json | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_w.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench_w.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_seed|k_march|k_eval3" --launch-skip 9 -c 3 -f -o gpurun_out/prof_r1_w python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_w.log 2>&1
python tools/exp_all.py > gpurun_out/exp_all_w.log 2>&1; tail -4 gpurun_out/exp_all_w.log
python tools/exp_big.py cfg4 3 1 > gpurun_out/exp_big_cfg4_w.log 2>&1; tail -1 gpurun_out/exp_big_cfg4_w.log | cut -c1-300
