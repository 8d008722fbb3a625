set -x
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_v16.log 2>&1; tail -3 gpurun_out/pytest_v16.log
for P in 0 1; do
PFB_PDL=$P timeout 120 python tools/time_config.py --batch 8 --height 436 --width 1024 --iters 12 --steps 10 > gpurun_out/t16_pdl$P.json 2>&1; cat gpurun_out/t16_pdl$P.json
done
PFB_CONV_TAP_GROUP=0 timeout 120 python tools/time_config.py --batch 8 --height 436 --width 1024 --iters 12 --steps 10
PFB_CONV_TAP_GROUP=0 PFB_CONV_VHALO=0 timeout 120 python tools/time_config.py --batch 8 --height 436 --width 1024 --iters 12 --steps 10
