timeout 600 python -m pytest tests/test_psroi_gpu.py -q -m gpu -x 2>&1 | tail -2
timeout 300 python tools/net_profile.py 2>&1 | grep -E "psroi_fwd|kernel time"
timeout 300 python bench.py --workload psroi_sweep_top --steps 30 --warmup 3 --no-cpu-baseline | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(d['value'],d['ms_per_step'])"
