timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k first_conv > gpurun_out/pytest_fc.log 2>&1; tail -5 gpurun_out/pytest_fc.log
if ! grep -q " passed" gpurun_out/pytest_fc.log || grep -q "failed" gpurun_out/pytest_fc.log; then
  PFB_FC_DESC_SWAP=1 timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k first_conv > gpurun_out/pytest_fc_swap.log 2>&1; tail -5 gpurun_out/pytest_fc_swap.log
fi
timeout 300 python -m pytest tests/test_gpu_e2e.py -m gpu -x -q 2>&1 | tail -2
for V in 0 1; do
PFB_NATIVE_CONV1=$V timeout 120 python tools/time_config.py --batch 8 --height 436 --width 1024 --iters 12 --steps 10
done
