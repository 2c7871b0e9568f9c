mkdir -p gpurun_out/final
( timeout 1500 python -m pytest tests -x -q -m gpu ) 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/final/bench_net_p.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/final/bench_net_p.json'));print('net',d['value'],d['ms_per_step'],d['e2e']['value'])"
timeout 900 python bench.py --workload lighthead_resnet50_train --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/final/bench_train_p.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/final/bench_train_p.json'));print('train',d['value'],d['ms_per_step'],d['e2e']['value'])"
timeout 300 python tools/net_profile.py 2>&1 | grep -E "nms_|rpn_topk|kernel time"
timeout 300 python tools/net_profile.py --train 2>&1 | grep -E "nms_|rpn_topk|kernel time"
