set -x
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_corr_block.py -m gpu -q -p no:cacheprovider -k "onthefly" 2>&1 | tail -40 ) > gpurun_out/pytest_r02j_otf.log 2>&1
tail -15 gpurun_out/pytest_r02j_otf.log
( timeout 600 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_configs.py tests/test_gpu_ops.py -m gpu -q -p no:cacheprovider -k "altcorr or onthefly or alt_corr" 2>&1 | tail -30 ) > gpurun_out/pytest_r02j_alt.log 2>&1
tail -6 gpurun_out/pytest_r02j_alt.log
timeout 300 python tools/time_config4.py > gpurun_out/config4_lookup_j.json 2> gpurun_out/config4_lookup_j.log
cat gpurun_out/config4_lookup_j.json; tail -3 gpurun_out/config4_lookup_j.log
Q="--no-comparators --no-cpu-baseline --no-parity --protocol-samples 0 --sustained-seconds 0"
timeout 300 python bench.py --batch 1 --height 1080 --width 1920 --iters 32 $Q > gpurun_out/bench_r02j_cfg4_volume.json 2> gpurun_out/bench_r02j_cfg4_volume.log
sed -i 's/XXX/XXX/' /dev/null
