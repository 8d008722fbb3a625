set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -30 ) > gpurun_out/pytest_r02e.log 2>&1
tail -8 gpurun_out/pytest_r02e.log
Q="--no-comparators --no-cpu-baseline --no-parity --protocol-samples 0 --sustained-seconds 0"
timeout 600 python bench.py $Q > gpurun_out/bench_r02e.json 2> gpurun_out/bench_r02e.log
PFB_CONV_TMA_STORE=0 timeout 300 python bench.py $Q > gpurun_out/bench_r02e_notmastore.json 2> gpurun_out/bench_r02e_notmastore.log
PFB_INORM_FUSED=1 timeout 300 python bench.py $Q > gpurun_out/bench_r02e_inormfused.json 2> gpurun_out/bench_r02e_inormfused.log
timeout 300 python bench.py --inflight 2 $Q > gpurun_out/bench_r02e_inflight2.json 2> gpurun_out/bench_r02e_inflight2.log
rm -f gpurun_out/conv_trace_e.jsonl
PFB_CONV_TRACE=gpurun_out/conv_trace_e.jsonl PFB_CUDA_GRAPH=0 timeout 300 python tools/profile_step.py --iters 1 --warmup 1 > gpurun_out/trace_e.log 2>&1
python tools/conv_trace_report.py gpurun_out/conv_trace_e.jsonl 1.9 > gpurun_out/conv_trace_e.txt 2>&1
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02e.csv python tools/profile_step.py > gpurun_out/profile_step_e.log 2>&1
for f in gpurun_out/bench_r02e*.json; do echo $f; head -c 300 $f; echo; done
