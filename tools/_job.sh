for P in 0 1 0 1; do PFB_PIPE_PRIORITY=$P timeout 300 python bench.py --no-cpu-baseline --steps 20 2>&1 | grep "resident\|e2e:" | cut -c1-60; done
