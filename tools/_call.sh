timeout 900 python -m pytest tests/test_detections_gpu.py tests/test_psroi_gpu.py -q -m gpu -x 2>&1 | tail -6
timeout 300 python tools/net_profile.py --train > gpurun_out/final/train_kernel_breakdown.txt 2>&1; head -12 gpurun_out/final/train_kernel_breakdown.txt
