set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -8 ) > gpurun_out/pytest_r02ac.log 2>&1
grep -E "passed|failed|error" gpurun_out/pytest_r02ac.log | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02ac.log 2>&1; tail -2 gpurun_out/smoke_r02ac.log
timeout 900 python bench.py > gpurun_out/bench_r02ac_default.json 2> gpurun_out/bench_r02ac_default.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r02ac_default.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "protocol", d["protocol"]["value"], "sustained", d["sustained"]["value"], d["clocks"], d["roofline"]["frac"], d["corr_hbm_roofline"]["frac_of_hbm_peak"], d["parity"]["max_abs_px"], d["gpu_launches"])
PY
Q="--no-comparators --no-cpu-baseline --protocol-samples 0 --sustained-seconds 0 --no-parity"
timeout 300 python bench.py --model gma --dtype bf16 --batch 4 $Q > gpurun_out/bench_r02ac_gma.json 2> gpurun_out/bench_r02ac_gma.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r02ac_gma.json").read().strip().splitlines()[-1])
print("gma bf16 b4:", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"])
PY
true
