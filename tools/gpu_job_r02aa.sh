set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -8 ) > gpurun_out/pytest_r02aa.log 2>&1
grep -E "passed|failed|error" gpurun_out/pytest_r02aa.log | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02aa.log 2>&1; tail -2 gpurun_out/smoke_r02aa.log
PFB_CAPTURE_EXCLUSIVE=1 timeout 300 python tools/stress_capture.py --rounds 8 > gpurun_out/stress_forks.log 2>&1; tail -2 gpurun_out/stress_forks.log
START=$(date +%s)
timeout 900 python bench.py > gpurun_out/bench_r02aa_default.json 2> gpurun_out/bench_r02aa_default.log
echo "default bench wall seconds: $(( $(date +%s) - START ))"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r02aa_default.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "protocol", d["protocol"]["value"], "sustained", d["sustained"]["value"], d["clocks"], d["roofline"]["frac"], d["corr_hbm_roofline"]["frac_of_hbm_peak"], d["parity"]["max_abs_px"], d["config"].get("stream_forks"))
PY
Q="--no-comparators --no-cpu-baseline --protocol-samples 0 --sustained-seconds 0 --steps 10 --warmup 3 --no-parity"
timeout 300 python bench.py $Q > gpurun_out/bench_r02aa_quick10.json 2> gpurun_out/bench_r02aa_quick10.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r02aa_quick10.json").read().strip().splitlines()[-1])
print("10 steps:", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"])
PY
true
