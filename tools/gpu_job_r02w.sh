set -x
mkdir -p gpurun_out
START=$(date +%s)
timeout 900 python bench.py > gpurun_out/bench_r02w_default.json 2> gpurun_out/bench_r02w_default.log
echo "default bench wall seconds: $(( $(date +%s) - START ))"
head -c 700 gpurun_out/bench_r02w_default.json
true
