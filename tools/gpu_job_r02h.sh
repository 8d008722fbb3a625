set -x
mkdir -p gpurun_out
Q="--no-comparators --no-cpu-baseline --no-parity --protocol-samples 0 --sustained-seconds 0"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02h.log 2>&1; tail -5 gpurun_out/smoke_r02h.log
PFB_CONV_VHALO=0 timeout 300 python bench.py $Q > gpurun_out/bench_r02h_novhalo.json 2> gpurun_out/bench_r02h_novhalo.log
PFB_CONV_VHALO=0 timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02h_novhalo.csv python tools/profile_step.py > gpurun_out/profile_step_h.log 2>&1
timeout 300 python bench.py --model gma --dtype bf16 --batch 4 $Q > gpurun_out/bench_r02h_gma_bf16_b4.json 2> gpurun_out/bench_r02h_gma_bf16_b4.log
PFB_GMA_BATCHED=0 timeout 300 python bench.py --model gma --dtype bf16 --batch 4 $Q > gpurun_out/bench_r02h_gma_bf16_b4_loop.json 2> gpurun_out/bench_r02h_gma_bf16_b4_loop.log
timeout 300 python bench.py --model raft_small --dtype fp32 --batch 1 --height 128 --width 256 --iters 4 $Q > gpurun_out/bench_r02h_cfg1.json 2> gpurun_out/bench_r02h_cfg1.log
timeout 300 python bench.py --batch 1 --height 1080 --width 1920 --iters 32 $Q > gpurun_out/bench_r02h_cfg4_volume.json 2> gpurun_out/bench_r02h_cfg4_volume.log
timeout 300 python bench.py --fp32-context $Q > gpurun_out/bench_r02h_fp32ctx.json 2> gpurun_out/bench_r02h_fp32ctx.log
for f in gpurun_out/bench_r02h*.json; do echo $f; head -c 250 $f; echo; done
