set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -40 ) > gpurun_out/pytest_r02c.log 2>&1
tail -8 gpurun_out/pytest_r02c.log
Q="--no-comparators --no-cpu-baseline --no-parity --protocol-samples 0 --sustained-seconds 0"
timeout 600 python bench.py --no-comparators --no-cpu-baseline > gpurun_out/bench_r02c.json 2> gpurun_out/bench_r02c.log
PFB_CUDNN_FUSED_RELU=1 timeout 300 python bench.py $Q > gpurun_out/bench_r02c_fusedrelu.json 2> gpurun_out/bench_r02c_fusedrelu.log
PFB_MERGE_C2F2=1 timeout 300 python bench.py $Q > gpurun_out/bench_r02c_merge.json 2> gpurun_out/bench_r02c_merge.log
PFB_GRU_CTX_SPLIT=0 timeout 300 python bench.py $Q > gpurun_out/bench_r02c_nosplit.json 2> gpurun_out/bench_r02c_nosplit.log
timeout 300 python bench.py --inflight 2 $Q > gpurun_out/bench_r02c_inflight2.json 2> gpurun_out/bench_r02c_inflight2.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'corr_lookup_tiled|corr_volume_tiled' -c 2 -o gpurun_out/r02c_corr python tools/profile_step.py > gpurun_out/ncu_corr_c.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_umma -s 14 -c 11 -o gpurun_out/r02c_conv python tools/profile_step.py > gpurun_out/ncu_conv_c.log 2>&1
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02c.csv python tools/profile_step.py > gpurun_out/profile_step_c.log 2>&1
for f in gpurun_out/bench_r02c*.json; do echo $f; head -c 300 $f; echo; done
