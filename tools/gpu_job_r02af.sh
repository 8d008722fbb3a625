set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -8 ) > gpurun_out/pytest_r02af.log 2>&1
grep -E "passed|failed|error" gpurun_out/pytest_r02af.log | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02af.log 2>&1; tail -2 gpurun_out/smoke_r02af.log
Q="--no-comparators --no-cpu-baseline --protocol-samples 0 --sustained-seconds 0"
timeout 300 python bench.py $Q > gpurun_out/bench_r02af.json 2> gpurun_out/bench_r02af.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r02af.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["parity"]["max_abs_px"])
PY
true
