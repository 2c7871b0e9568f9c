mkdir -p gpurun_out/final
cd /root/repo
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/final/bench_net.json 2> gpurun_out/final/bench_net.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final/bench_net_reference_arm.json 2>> gpurun_out/final/bench_net.err
timeout 900 python bench.py --workload psroi_sweep_top --steps 20 --warmup 3 > gpurun_out/final/bench_psroi.json 2> gpurun_out/final/bench_psroi.err
timeout 900 python bench.py --workload lighthead_xception_800 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/final/bench_xception800.json 2> gpurun_out/final/bench_x.err
timeout 900 python bench.py --workload lighthead_resnet50_train --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/final/bench_train.json 2> gpurun_out/final/bench_train.err
timeout 600 python tools/psroi_sweep.py --out gpurun_out/final/psroi_sweep.json > gpurun_out/final/psroi_sweep.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final/launches_net.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-autotune > gpurun_out/final/ncu_net.log 2>&1
timeout 300 python tools/net_profile.py > gpurun_out/final/net_kernel_breakdown.txt 2>&1
timeout 300 python tools/net_profile.py --train > gpurun_out/final/train_kernel_breakdown.txt 2>&1
for f in gpurun_out/final/bench_*.json; do echo $f; cut -c1-300 $f; done; tail -3 gpurun_out/final/*.err
