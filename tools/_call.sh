mkdir -p gpurun_out/final
( time timeout 1500 python -m pytest tests -x -q -m gpu ) 2>&1 | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python tools/net_profile.py --train > gpurun_out/final/train_kernel_breakdown.txt 2>&1; head -14 gpurun_out/final/train_kernel_breakdown.txt | cut -c1-120
