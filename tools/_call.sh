mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_mode_gpu.py -q -m gpu > gpurun_out/c3_parity_tests.log 2>&1; echo "parity rc=$?" >> gpurun_out/c3_parity_tests.log
timeout 600 python -m pytest tests/test_psroi_gpu.py -x -q -m gpu > gpurun_out/c3_psroi_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/c3_psroi_tests.log
timeout 300 python bench.py --workload psroi_sweep_top --steps 20 --warmup 3 > gpurun_out/bench_psroi_r1d.json 2> gpurun_out/bench_psroi_r1d.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_psroi_r1d.csv python bench.py --workload psroi_sweep_top --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
tail -40 gpurun_out/c3_parity_tests.log; tail -3 gpurun_out/c3_psroi_tests.log; cat gpurun_out/bench_psroi_r1d.json
