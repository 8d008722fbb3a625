#!/usr/bin/env python
"""Time one forward configuration (parity-case configs of BASELINE.json that are not the bench line)."""
import argparse, os, sys, time, json
from argparse import Namespace
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ptlflow_b200 as pb

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="raft"); ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--height", type=int, default=1080); ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--iters", type=int, default=32); ap.add_argument("--dtype", default="fp16")
ap.add_argument("--alternate-corr", action="store_true"); ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
dtype = {"fp16": torch.float16, "bf16": torch.bfloat16, "fp32": torch.float32}[a.dtype]
torch.manual_seed(1234)
m = pb.get_model(a.model, args=Namespace(model=Namespace(iters=a.iters, alternate_corr=a.alternate_corr))).eval().cuda().to(dtype)
x = torch.rand(a.batch, 2, 3, a.height, a.width, device="cuda", dtype=dtype)
with torch.no_grad():
    for _ in range(2): m({"images": x})
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(a.steps): m({"images": x})
    e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
print(json.dumps({"model": a.model, "batch": a.batch, "hw": [a.height, a.width], "iters": a.iters, "dtype": a.dtype, "alternate_corr": a.alternate_corr,
                  "ms_per_forward": round(ms, 3), "pairs_per_s": round(a.batch / ms * 1e3, 2), "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 2)}))
