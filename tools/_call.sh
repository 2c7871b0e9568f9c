timeout 600 python -m pytest tests/test_conv_gpu.py -q -m gpu -k depthwise 2>&1 | tail -2
for d in "32 100 100 728" "32 200 200 256" "8 30 30 728" "32 50 50 1024"; do timeout 120 python tools/dw_one.py $d | tail -1; done
timeout 300 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active --clock-control none -k regex:depthwise3x3_rows -s 4 -c 1 python tools/dw_one.py 32 100 100 728 2>&1 | grep -E "inst_executed|duration|issue_active|pipe_"
timeout 600 python -m pytest tests/test_model_gpu.py -q -m gpu 2>&1 | tail -2
timeout 900 python bench.py --workload lighthead_xception_800 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_x800_dw.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/bench_x800_dw.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'])"
