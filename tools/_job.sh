timeout 200 python tools/time_config.py --model gma --batch 4 --height 436 --width 1024 --iters 12 --dtype bf16 --steps 6 | tee gpurun_out/cfg3_gma_v22.json
timeout 200 python tools/time_config.py --model raft --batch 1 --height 1080 --width 1920 --iters 32 --dtype fp16 --steps 4 | tee gpurun_out/cfg4_vol_v22.json
timeout 300 python tools/time_config.py --model raft --batch 1 --height 1080 --width 1920 --iters 32 --dtype fp16 --steps 3 --alternate-corr | tee gpurun_out/cfg4_alt_v22.json
timeout 200 python tools/time_config.py --model raft_small --batch 1 --height 128 --width 256 --iters 4 --dtype fp32 --steps 20 | tee gpurun_out/cfg1_small_v22.json
