set -x
mkdir -p gpurun_out
( PFB_ENCODER_LANES=3 timeout 600 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_configs.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -8 ) > gpurun_out/pytest_r02ab_lanes3.log 2>&1
grep -E "passed|failed|error" gpurun_out/pytest_r02ab_lanes3.log | tail -3
Q="--no-comparators --no-cpu-baseline --protocol-samples 0 --sustained-seconds 0 --steps 20 --warmup 5 --no-parity"
for rep in 1 2; do
for v in "2 1" "3 1" "1 0"; do
set -- $v
PFB_ENCODER_LANES=$1 PFB_FORK_FLOW=$2 timeout 300 python bench.py $Q > gpurun_out/bench_r02ab_$1$2.json 2> gpurun_out/bench_r02ab_$1$2.log
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_r02ab_$1$2.json").read().strip().splitlines()[-1])
print("lanes=$1 fork_flow=$2:", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"])
PY
done
done
true
