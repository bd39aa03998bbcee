python tools/exp_all.py > gpurun_out/exp_all_v.log 2>&1; tail -5 gpurun_out/exp_all_v.log
python tools/exp_big.py cfg4 3,0 1 > gpurun_out/exp_big_cfg4_v.log 2>&1; tail -3 gpurun_out/exp_big_cfg4_v.log
timeout 1200 python tools/exp_big.py cfg5 3,0 1 > gpurun_out/exp_big_cfg5_v.log 2>&1; tail -4 gpurun_out/exp_big_cfg5_v.log
