timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_v22.log 2>&1; tail -2 gpurun_out/pytest_v22.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_v22.json 2> gpurun_out/bench_v22.err; tail -1 gpurun_out/bench_v22.json | cut -c1-200; grep "resident\|e2e:" gpurun_out/bench_v22.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
