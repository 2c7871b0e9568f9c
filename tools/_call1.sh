mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_psroi_gpu.py -x -q -m gpu > gpurun_out/c1_psroi_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/c1_psroi_tests.log
timeout 400 python tools/psroi_sweep.py --quick --out gpurun_out/psroi_sweep_r1b.json > gpurun_out/psroi_sweep_r1b.log 2>&1
timeout 300 python bench.py --workload psroi_sweep_top --steps 20 --warmup 3 > gpurun_out/bench_psroi_r1b.json 2> gpurun_out/bench_psroi_r1b.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:psroi_fwd_select -s 2 -c 1 -f -o gpurun_out/psroi_select_r1b python bench.py --workload psroi_sweep_top --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_sel.log 2>&1
tail -3 gpurun_out/c1_psroi_tests.log; tail -20 gpurun_out/psroi_sweep_r1b.log; cat gpurun_out/bench_psroi_r1b.json
