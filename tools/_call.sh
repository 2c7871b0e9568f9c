mkdir -p gpurun_out
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_net_tuned.json 2> gpurun_out/bench_net_tuned.err; tail -3 gpurun_out/bench_net_tuned.err; cat gpurun_out/bench_net_tuned.json | cut -c1-400
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-autotune > gpurun_out/bench_net_untuned.json 2>> gpurun_out/bench_net_tuned.err; cat gpurun_out/bench_net_untuned.json | cut -c1-250
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py -q -m gpu -x 2>&1 | tail -3
