set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -40 ) > gpurun_out/pytest_r02d.log 2>&1
tail -8 gpurun_out/pytest_r02d.log
Q="--no-comparators --no-cpu-baseline --no-parity --protocol-samples 0 --sustained-seconds 0"
timeout 600 python bench.py --no-comparators --no-cpu-baseline > gpurun_out/bench_r02d.json 2> gpurun_out/bench_r02d.log
PFB_CONV_TMA_STORE=0 timeout 300 python bench.py $Q > gpurun_out/bench_r02d_notmastore.json 2> gpurun_out/bench_r02d_notmastore.log
timeout 300 python bench.py --inflight 2 $Q > gpurun_out/bench_r02d_inflight2.json 2> gpurun_out/bench_r02d_inflight2.log
rm -f gpurun_out/conv_trace_d.jsonl
PFB_CONV_TRACE=gpurun_out/conv_trace_d.jsonl PFB_CUDA_GRAPH=0 timeout 300 python tools/profile_step.py --iters 1 --warmup 1 > gpurun_out/trace_d.log 2>&1
python tools/conv_trace_report.py gpurun_out/conv_trace_d.jsonl 1.9 > gpurun_out/conv_trace_d.txt 2>&1
timeout 300 python tools/time_config4.py > gpurun_out/config4_lookup_d.json 2> gpurun_out/config4_lookup_d.log
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02d.csv python tools/profile_step.py > gpurun_out/profile_step_d.log 2>&1
for f in gpurun_out/bench_r02d*.json; do echo $f; head -c 300 $f; echo; done
cat gpurun_out/config4_lookup_d.json
