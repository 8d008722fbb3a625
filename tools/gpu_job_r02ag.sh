set -x
mkdir -p gpurun_out
PFB_SANITIZE_ONLY_SIMT=1 PFB_SANITIZE_TIMEOUT=150 bash tools/sanitize.sh > gpurun_out/sanitize_simt_final.log 2>&1
grep -E "SANITIZER|ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitize_simt_final.log | tail -10
true
