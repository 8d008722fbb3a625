set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
( time python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -60 ) > gpurun_out/pytest_r02a.log 2>&1
python bench.py > gpurun_out/bench_r02a.json 2> gpurun_out/bench_r02a.log
python bench.py --inflight 2 --no-comparators --no-cpu-baseline --no-parity --protocol-samples 0 --sustained-seconds 3 > gpurun_out/bench_r02a_inflight2.json 2> gpurun_out/bench_r02a_inflight2.log
python bench.py --cuda-graph 0 --no-comparators --no-cpu-baseline --no-parity --sustained-seconds 0 > gpurun_out/bench_r02a_eager.json 2> gpurun_out/bench_r02a_eager.log
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02a.csv python tools/profile_step.py > gpurun_out/profile_step.log 2>&1
PFB_SANITIZE_TIMEOUT=240 bash tools/sanitize.sh > gpurun_out/sanitize_r02a.log 2>&1
tail -5 gpurun_out/pytest_r02a.log
cat gpurun_out/bench_r02a.json | head -c 3000
