set -x
mkdir -p gpurun_out
Q="--no-comparators --no-cpu-baseline --protocol-samples 0 --sustained-seconds 0 --no-parity"
for c in 0 2 4 8; do
  PFB_ENCODER_CHUNK=$c timeout 300 python bench.py $Q > gpurun_out/bench_r02s_chunk$c.json 2> gpurun_out/bench_r02s_chunk$c.log
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_r02s_chunk$c.json").read().strip().splitlines()[-1])
k=d["kernels"]
print("chunk $c", d["value"], d["ms_per_step"], "affine", k["enc_affine"]["ms_per_step"], "stats", k["enc_stats"]["ms_per_step"], "other", k["_not_this_library"]["ms_per_step"])
PY
done
true
