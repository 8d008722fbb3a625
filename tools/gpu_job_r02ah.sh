set -x
mkdir -p gpurun_out
( timeout 500 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -5 ) > gpurun_out/pytest_r02ah.log 2>&1
grep -E "passed|failed|error" gpurun_out/pytest_r02ah.log | tail -3
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02ah.log 2>&1; tail -1 gpurun_out/smoke_r02ah.log
true
