set -x
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_corr_block.py tests/test_gpu_e2e.py tests/test_gpu_configs.py -m gpu -q -p no:cacheprovider -k "onthefly or altcorr or tiled or pipeline" 2>&1 | tail -30 ) > gpurun_out/pytest_r02o_subset.log 2>&1
grep -E "passed|failed|error" gpurun_out/pytest_r02o_subset.log | tail -3
rm -f gpurun_out/otf_trace.jsonl
PFB_OTF_TRACE=gpurun_out/otf_trace.jsonl timeout 200 python - <<'PY' > gpurun_out/otf_trace_run.log 2>&1
import torch, sys, os
sys.path.insert(0, os.getcwd())
from ptlflow_b200 import ops
dev = "cuda:0"
B, H, W, C, L, R = 1, 135, 240, 256, 4, 4
torch.manual_seed(0)
f1 = torch.randn(B, H, W, C, device=dev).half(); f2 = torch.randn(B, H, W, C, device=dev).half()
pyr = ops.feature_pyramid(f2, L)
ys, xs = torch.meshgrid(torch.arange(H, device=dev, dtype=torch.float32), torch.arange(W, device=dev, dtype=torch.float32), indexing="ij")
for sig in (1.0, 4.0):
    cs = (torch.stack([xs, ys], -1)[None] + sig * torch.randn(B, H, W, 2, device=dev)).contiguous()
    for _ in range(2):
        ops.corr_lookup_onthefly_tc(f1, pyr, cs, R)
    torch.cuda.synchronize()
print("ok")
PY
tail -3 gpurun_out/otf_trace_run.log
python tools/otf_trace_report.py gpurun_out/otf_trace.jsonl > gpurun_out/otf_trace_report_o.txt 2>&1; head -26 gpurun_out/otf_trace_report_n.txt
timeout 300 python tools/time_config4.py > gpurun_out/config4_lookup_o.json 2> gpurun_out/config4_lookup_o.log
cat gpurun_out/config4_lookup_o.json
( timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -15 ) > gpurun_out/pytest_r02o.log 2>&1
grep -E "passed|failed|error" gpurun_out/pytest_r02o.log | tail -3
Q="--no-comparators --no-cpu-baseline --protocol-samples 0 --sustained-seconds 0"
timeout 300 python bench.py --batch 1 --height 1080 --width 1920 --iters 32 --alternate-corr $Q > gpurun_out/bench_r02o_cfg4_onthefly_tc.json 2> gpurun_out/bench_r02o_cfg4_onthefly_tc.log
head -c 300 gpurun_out/bench_r02o_cfg4_onthefly_tc.json
timeout 300 python bench.py $Q > gpurun_out/bench_r02o_quick.json 2> gpurun_out/bench_r02o_quick.log
head -c 300 gpurun_out/bench_r02o_quick.json
true
