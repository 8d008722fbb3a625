set -x
mkdir -p gpurun_out
( PFB_FORK_FLOW=1 PFB_FORK_ENCODERS=1 timeout 600 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_configs.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -8 ) > gpurun_out/pytest_r02z_fork.log 2>&1
grep -E "passed|failed|error" gpurun_out/pytest_r02z_fork.log | tail -3
Q="--no-comparators --no-cpu-baseline --protocol-samples 0 --sustained-seconds 0 --steps 10 --warmup 3 --no-parity"
for rep in 1 2; do
for v in "0 0" "1 0" "0 1" "1 1"; do
set -- $v
PFB_FORK_FLOW=$1 PFB_FORK_ENCODERS=$2 timeout 300 python bench.py $Q > gpurun_out/bench_r02z_$1$2.json 2> gpurun_out/bench_r02z_$1$2.log
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_r02z_$1$2.json").read().strip().splitlines()[-1])
print("fork flow=$1 enc=$2:", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"])
PY
done
done
true
