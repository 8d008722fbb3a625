set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -15 ) > gpurun_out/pytest_r02m.log 2>&1
grep -E "passed|failed|error" gpurun_out/pytest_r02m.log | tail -3
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
python tools/otf_trace_report.py gpurun_out/otf_trace.jsonl > gpurun_out/otf_trace_report.txt 2>&1; cat gpurun_out/otf_trace_report.txt
Q="--no-comparators --no-cpu-baseline --protocol-samples 0 --sustained-seconds 0"
timeout 300 python bench.py $Q > gpurun_out/bench_r02m_quick.json 2> gpurun_out/bench_r02m_quick.log
head -c 400 gpurun_out/bench_r02m_quick.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:corr_lookup_tiled -c 2 -o gpurun_out/r02m_lookup -f python tools/profile_step.py --cuda-graph 0 > gpurun_out/ncu_r02m.log 2>&1
ls -la gpurun_out/r02m_lookup.ncu-rep
true
