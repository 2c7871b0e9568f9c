mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_mode_gpu.py tests/test_detections_gpu.py tests/test_proposals_gpu.py -q -m gpu > gpurun_out/c6_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c6_tests.log
grep -n "^E  \|passed\|failed\|rc=\|Error" gpurun_out/c6_tests.log | cut -c1-300 | head -40
