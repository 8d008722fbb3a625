set -x
mkdir -p gpurun_out
PFB_CAPTURE_EXCLUSIVE=0 timeout 300 python tools/stress_capture.py --rounds 14 > gpurun_out/stress_gate0.log 2>&1
tail -5 gpurun_out/stress_gate0.log
PFB_CAPTURE_EXCLUSIVE=1 timeout 300 python tools/stress_capture.py --rounds 14 > gpurun_out/stress_gate1.log 2>&1
tail -3 gpurun_out/stress_gate1.log
for i in 1 2 3 4 5 6; do
  ( timeout 200 python -m pytest tests/test_gpu_e2e.py -m gpu -q -p no:cacheprovider -k "pipeline" 2>&1 | tail -3 ) >> gpurun_out/pytest_r02p_pipeline.log 2>&1
done
grep -E "passed|failed" gpurun_out/pytest_r02p_pipeline.log
true
