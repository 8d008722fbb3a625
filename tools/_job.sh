PFB_BENCH_DEBUG=1 timeout 600 python bench.py --no-cpu-baseline 2>&1 | grep "debug\|resident\|e2e:"
PFB_BENCH_DEBUG=1 PFB_PDL=0 timeout 600 python bench.py --no-cpu-baseline 2>&1 | grep "debug\|resident\|e2e:"
