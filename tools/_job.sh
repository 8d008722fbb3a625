timeout 600 python bench.py > gpurun_out/bench_v23.json 2> gpurun_out/bench_v23.err; tail -1 gpurun_out/bench_v23.json | cut -c1-220; grep "resident\|e2e:\|warm" gpurun_out/bench_v23.err | cut -c1-150
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_v23.log 2>&1; tail -2 gpurun_out/pytest_v23.log
