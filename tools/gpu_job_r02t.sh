set -x
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_corr_block.py -m gpu -q -p no:cacheprovider -k "onthefly_tensor_core" 2>&1 | tail -5 ) > gpurun_out/pytest_r02t.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_r02t.log
PFB_SANITIZE_ONLY_R2=1 PFB_SANITIZE_TIMEOUT=200 bash tools/sanitize.sh > gpurun_out/sanitize_r2.log 2>&1
grep -E "SANITIZER|ERROR SUMMARY|passed|failed" gpurun_out/sanitize_r2.log | tail -12
true
