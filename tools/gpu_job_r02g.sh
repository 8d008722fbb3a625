set -x
mkdir -p gpurun_out
nvidia-smi -L
( time python bench.py > gpurun_out/bench_r02g_full.json 2> gpurun_out/bench_r02g_full.log ) 2> gpurun_out/bench_r02g_full.time
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r02g_n2.json 2> gpurun_out/bench_r02g_n2.log ) 2> gpurun_out/bench_r02g_n2.time
( time python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r02g_ref.json 2> gpurun_out/bench_r02g_ref.log ) 2> gpurun_out/bench_r02g_ref.time
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02g.log 2>&1
tail -3 gpurun_out/smoke_r02g.log
cat gpurun_out/bench_r02g_full.time gpurun_out/bench_r02g_n2.time
head -c 600 gpurun_out/bench_r02g_n2.json
