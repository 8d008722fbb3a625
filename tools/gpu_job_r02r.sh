set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02r.log 2>&1; tail -3 gpurun_out/smoke_r02r.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r02r_n2.json 2> gpurun_out/bench_r02r_n2.log
head -c 400 gpurun_out/bench_r02r_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_r02r_ref_n2.json 2> gpurun_out/bench_r02r_ref_n2.log
head -c 300 gpurun_out/bench_r02r_ref_n2.json
true
