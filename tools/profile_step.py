#!/usr/bin/env python
"""One forward of the bench workload between cudaProfilerStart/Stop, for ncu:

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tools/profile_step.py
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_umma -c 4 \
      -o gpurun_out/conv_umma python tools/profile_step.py
"""
import argparse
import os
import sys
from argparse import Namespace

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import ptlflow_b200 as pb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="raft")
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--height", type=int, default=436)
ap.add_argument("--width", type=int, default=1024)
ap.add_argument("--iters", type=int, default=12)
ap.add_argument("--dtype", default="fp16")
ap.add_argument("--kernel-impl", type=int, default=0)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--alternate-corr", action="store_true")
ap.add_argument("--cuda-graph", type=int, default=0, help="0 (default): eager launches, so every kernel shows up by name under ncu")
a = ap.parse_args()

dtype = {"fp16": torch.float16, "bf16": torch.bfloat16, "fp32": torch.float32}[a.dtype]
torch.manual_seed(1234)
mkw = dict(iters=a.iters)
if a.alternate_corr:
    mkw["alternate_corr"] = True
model = pb.get_model(a.model, args=Namespace(model=Namespace(**mkw))).eval().cuda().to(dtype)
model.kernel_impl = a.kernel_impl
model.use_cuda_graph = bool(a.cuda_graph)
x = torch.rand(a.batch, 2, 3, a.height, a.width, device="cuda", dtype=dtype)
with torch.no_grad():
    for _ in range(a.warmup):
        model({"images": x})
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    model({"images": x})
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("done")
