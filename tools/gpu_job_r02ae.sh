set -x
mkdir -p gpurun_out
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02ae.csv python tools/profile_step.py > gpurun_out/profile_step_ae.log 2>&1
wc -l gpurun_out/launches_r02ae.csv
timeout 300 ncu --set full --clock-control none --import-source on -k regex:corr_onthefly_umma -c 2 -o gpurun_out/r02ae_otf -f python tools/profile_step.py --alternate-corr --batch 1 --height 1080 --width 1920 --iters 2 --cuda-graph 0 > gpurun_out/ncu_r02ae_otf.log 2>&1
ls -la gpurun_out/r02ae_otf.ncu-rep
true
