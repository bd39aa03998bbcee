python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_parity.py::test_cfg4_full_size_pipelines_agree 2>&1 | tail -5
python tools/exp_single.py cfg3 2>&1 | head -9
python tools/exp_cfg4.py cfg4 1.5e9 2>&1 | tail -20
ncu --set full --clock-control none --import-source on -k regex:"k_march|k_eval3" --launch-skip 6 -c 2 -f -o gpurun_out/prof_r1_s python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --pipeline 3 > gpurun_out/ncu_full_s.log 2>&1
