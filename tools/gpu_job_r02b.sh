set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_corr_block.py tests/test_gpu_e2e.py tests/test_gpu_configs.py tests/test_gpu_ops.py -m gpu -q -p no:cacheprovider 2>&1 | tail -40 ) > gpurun_out/pytest_r02b.log 2>&1
tail -8 gpurun_out/pytest_r02b.log
Q="--no-comparators --no-cpu-baseline --no-parity --protocol-samples 0 --sustained-seconds 0"
timeout 600 python bench.py --no-comparators --no-cpu-baseline > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.log
timeout 300 python bench.py --inflight 2 $Q > gpurun_out/bench_r02b_inflight2.json 2> gpurun_out/bench_r02b_inflight2.log
PFB_VOLUME_TILED=0 timeout 300 python bench.py $Q > gpurun_out/bench_r02b_dense.json 2> gpurun_out/bench_r02b_dense.log
PFB_ENCODER_CHUNK=8 timeout 300 python bench.py $Q > gpurun_out/bench_r02b_chunk8.json 2> gpurun_out/bench_r02b_chunk8.log
PFB_ENCODER_CHUNK=4 timeout 300 python bench.py $Q > gpurun_out/bench_r02b_chunk4.json 2> gpurun_out/bench_r02b_chunk4.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'corr_lookup_tiled|corr_volume_tiled' -c 3 -o gpurun_out/r02b_corr python tools/profile_step.py > gpurun_out/ncu_corr.log 2>&1
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02b.csv python tools/profile_step.py > gpurun_out/profile_step_b.log 2>&1
for f in gpurun_out/bench_r02b*.json; do echo $f; head -c 400 $f; echo; done
tail -3 gpurun_out/bench_r02b_inflight2.log
