timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_v20.log 2>&1; tail -3 gpurun_out/pytest_v20.log
timeout 120 python tools/time_config.py --batch 8 --height 436 --width 1024 --iters 12 --steps 10
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v20.csv python tools/profile_step.py > /dev/null 2>&1
