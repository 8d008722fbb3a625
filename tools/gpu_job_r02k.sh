set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -30 ) > gpurun_out/pytest_r02k.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_r02k.log | tail -2
Q="--no-comparators --no-cpu-baseline --protocol-samples 0 --sustained-seconds 0"
timeout 300 python bench.py --batch 1 --height 1080 --width 1920 --iters 32 --alternate-corr $Q > gpurun_out/bench_r02k_cfg4_onthefly_tc.json 2> gpurun_out/bench_r02k_cfg4_onthefly_tc.log
PFB_ONTHEFLY_TC=0 timeout 300 python bench.py --batch 1 --height 1080 --width 1920 --iters 32 --alternate-corr $Q > gpurun_out/bench_r02k_cfg4_onthefly_simt.json 2> gpurun_out/bench_r02k_cfg4_onthefly_simt.log
timeout 300 python bench.py --batch 1 --height 1080 --width 1920 --iters 32 $Q > gpurun_out/bench_r02k_cfg4_volume.json 2> gpurun_out/bench_r02k_cfg4_volume.log
timeout 300 python bench.py $Q > gpurun_out/bench_r02k.json 2> gpurun_out/bench_r02k.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:corr_onthefly_umma -c 2 -o gpurun_out/r02k_otf python tools/profile_step.py --batch 1 --height 1080 --width 1920 --iters 2 --alternate-corr > gpurun_out/ncu_otf_k.log 2>&1
for f in gpurun_out/bench_r02k*.json; do echo $f; head -c 300 $f; echo; done
true
