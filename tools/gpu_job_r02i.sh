set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_r02i_n8.json 2> gpurun_out/bench_r02i_n8.log ) 2> gpurun_out/bench_r02i_n8.time
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/bench_r02i_n4.json 2> gpurun_out/bench_r02i_n4.log ) 2> gpurun_out/bench_r02i_n4.time
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --steps 10 --warmup 3 --model gma --dtype bf16 --batch 4 --no-parity --sustained-seconds 0 > gpurun_out/bench_r02i_gma_bf16_b32_n8.json 2> gpurun_out/bench_r02i_gma_n8.log ) 2> gpurun_out/bench_r02i_gma_n8.time
cat gpurun_out/bench_r02i_n8.time; head -c 500 gpurun_out/bench_r02i_n8.json; echo; head -c 400 gpurun_out/bench_r02i_gma_bf16_b32_n8.json
