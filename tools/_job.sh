timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "flow_conv7x7 or first_conv or instance_norm" 2>&1 | tail -5
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_v18.log 2>&1; tail -3 gpurun_out/pytest_v18.log
for V in 0 1; do
PFB_FLOW_CONV_UMMA=$V timeout 120 python tools/time_config.py --batch 8 --height 436 --width 1024 --iters 12 --steps 10
done
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v18.csv python tools/profile_step.py > /dev/null 2>&1
