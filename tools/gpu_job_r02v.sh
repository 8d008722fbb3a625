set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -8 ) > gpurun_out/pytest_r02v.log 2>&1
grep -E "passed|failed|error" gpurun_out/pytest_r02v.log | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02v.log 2>&1; tail -2 gpurun_out/smoke_r02v.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02v_reference.json 2> gpurun_out/bench_r02v_reference.log
head -c 250 gpurun_out/bench_r02v_reference.json; echo
/usr/bin/time -v timeout 900 python bench.py > gpurun_out/bench_r02v_default.json 2> gpurun_out/bench_r02v_default.log
grep -E "Elapsed" gpurun_out/bench_r02v_default.log
head -c 600 gpurun_out/bench_r02v_default.json
true
