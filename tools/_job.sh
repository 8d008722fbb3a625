timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_e2e.py -m gpu -x -q 2>&1 | tail -3
timeout 120 python tools/time_config.py --batch 8 --height 436 --width 1024 --iters 12 --steps 10
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v19.csv python tools/profile_step.py > /dev/null 2>&1
