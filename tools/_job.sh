timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_v17.log 2>&1; tail -2 gpurun_out/pytest_v17.log
for C in 3 4 8; do
PFB_FRAME_CHANNELS=$C timeout 120 python tools/time_config.py --batch 8 --height 436 --width 1024 --iters 12 --steps 10
done
PFB_FRAME_CHANNELS=4 timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v17_c4.csv python tools/profile_step.py > /dev/null 2>&1
PFB_FRAME_CHANNELS=8 timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v17_c8.csv python tools/profile_step.py > /dev/null 2>&1
