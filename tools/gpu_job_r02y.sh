set -x
mkdir -p gpurun_out
( PFB_FORK_FLOW=1 timeout 600 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_configs.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -8 ) > gpurun_out/pytest_r02y_fork.log 2>&1
grep -E "passed|failed|error" gpurun_out/pytest_r02y_fork.log | tail -3
Q="--no-comparators --no-cpu-baseline --protocol-samples 0 --sustained-seconds 0 --steps 10 --warmup 3"
for v in 1 0 1 0; do
PFB_FORK_FLOW=$v timeout 300 python bench.py $Q > gpurun_out/bench_r02y_fork_$v.json 2> gpurun_out/bench_r02y_fork_$v.log
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_r02y_fork_$v.json").read().strip().splitlines()[-1])
print("fork = $v:", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "parity", d["parity"]["max_abs_px"])
PY
done
true
