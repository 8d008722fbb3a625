set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -8 ) > gpurun_out/pytest_r02x.log 2>&1
grep -E "passed|failed|error" gpurun_out/pytest_r02x.log | tail -3
Q="--no-comparators --no-cpu-baseline --protocol-samples 0 --sustained-seconds 0 --steps 10 --warmup 3"
for v in 1 0; do
PFB_NATIVE_CONV2=$v timeout 300 python bench.py $Q > gpurun_out/bench_r02x_conv2_$v.json 2> gpurun_out/bench_r02x_conv2_$v.log
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_r02x_conv2_$v.json").read().strip().splitlines()[-1])
k=d["kernels"]
print("native conv2 = $v:", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "conv", k["conv"]["ms_per_step"], "affine", k["enc_affine"]["ms_per_step"], "other", k["_not_this_library"]["ms_per_step"], "parity", d["parity"]["max_abs_px"])
PY
done
true
