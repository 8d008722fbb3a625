set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r02ad_n2.json 2> gpurun_out/bench_r02ad_n2.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r02ad_n2.json").read().strip().splitlines()[-1])
print("N=2:", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d.get("strong_scaling"), d["n_gpus"])
PY
true
