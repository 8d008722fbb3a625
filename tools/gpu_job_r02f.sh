set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -30 ) > gpurun_out/pytest_r02f.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_r02f.log | tail -2
Q="--no-comparators --no-cpu-baseline --no-parity --protocol-samples 0 --sustained-seconds 0"
timeout 600 python bench.py $Q > gpurun_out/bench_r02f.json 2> gpurun_out/bench_r02f.log
timeout 600 python bench.py $Q > gpurun_out/bench_r02f_2.json 2> gpurun_out/bench_r02f_2.log
timeout 300 python bench.py --inflight 2 $Q > gpurun_out/bench_r02f_inflight2.json 2> gpurun_out/bench_r02f_inflight2.log
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02f.csv python tools/profile_step.py > gpurun_out/profile_step_f.log 2>&1
for f in gpurun_out/bench_r02f*.json; do echo $f; head -c 300 $f; echo; done
