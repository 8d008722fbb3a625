set -x
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_corr_block.py tests/test_gpu_e2e.py tests/test_gpu_configs.py -m gpu -q -p no:cacheprovider -k "onthefly or altcorr" 2>&1 | tail -30 ) > gpurun_out/pytest_r02l.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_r02l.log | tail -2
timeout 300 python tools/time_config4.py > gpurun_out/config4_lookup_l.json 2> gpurun_out/config4_lookup_l.log
cat gpurun_out/config4_lookup_l.json
Q="--no-comparators --no-cpu-baseline --protocol-samples 0 --sustained-seconds 0"
timeout 300 python bench.py --batch 1 --height 1080 --width 1920 --iters 32 --alternate-corr $Q > gpurun_out/bench_r02l_cfg4_onthefly_tc.json 2> gpurun_out/bench_r02l_cfg4_onthefly_tc.log
head -c 300 gpurun_out/bench_r02l_cfg4_onthefly_tc.json
true
