set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -15 ) > gpurun_out/pytest_r02u.log 2>&1
grep -E "passed|failed|error" gpurun_out/pytest_r02u.log | tail -3
Q="--no-comparators --no-cpu-baseline --protocol-samples 0 --sustained-seconds 0"
timeout 300 python bench.py --model raft_small --batch 1 --height 128 --width 256 --iters 4 --dtype fp32 $Q > gpurun_out/bench_r02u_cfg1.json 2> gpurun_out/bench_r02u_cfg1.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r02u_cfg1.json").read().strip().splitlines()[-1])
print("cfg1", d["value"], d["ms_per_step"], {k:v.get("ms_per_step") for k,v in d["kernels"].items()}, d.get("parity"))
PY
true
