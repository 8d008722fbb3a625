set -x
mkdir -p gpurun_out
PFB_CAPTURE_EXCLUSIVE=1 timeout 300 python tools/stress_capture.py --rounds 16 > gpurun_out/stress_gate1.log 2>&1
grep -c "ok, max" gpurun_out/stress_gate1.log; tail -3 gpurun_out/stress_gate1.log
PFB_CAPTURE_EXCLUSIVE=0 timeout 300 python tools/stress_capture.py --rounds 16 > gpurun_out/stress_gate0.log 2>&1
grep -c "ok, max" gpurun_out/stress_gate0.log; tail -1 gpurun_out/stress_gate0.log
rm -f gpurun_out/pytest_r02q_pipeline.log
for i in 1 2 3 4; do
  ( timeout 200 python -m pytest tests/test_gpu_e2e.py -m gpu -q -p no:cacheprovider -k "pipeline" 2>&1 | tail -3 ) >> gpurun_out/pytest_r02q_pipeline.log 2>&1
done
grep -E "passed|failed" gpurun_out/pytest_r02q_pipeline.log
( timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -15 ) > gpurun_out/pytest_r02q.log 2>&1
grep -E "passed|failed|error" gpurun_out/pytest_r02q.log | tail -3
Q="--no-comparators --no-cpu-baseline --protocol-samples 0 --sustained-seconds 0"
timeout 300 python bench.py $Q --inflight 2 > gpurun_out/bench_r02q_inflight2.json 2> gpurun_out/bench_r02q_inflight2.log
head -c 300 gpurun_out/bench_r02q_inflight2.json
true
