#!/usr/bin/env python
"""Throughput of S concurrent forward passes (one CUDA stream + one host thread each) of the bench workload."""
import argparse, json, os, sys, threading, time
from argparse import Namespace
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ptlflow_b200 as pb

ap = argparse.ArgumentParser()
ap.add_argument("--streams", type=int, default=2); ap.add_argument("--steps", type=int, default=12); ap.add_argument("--batch", type=int, default=8)
a = ap.parse_args()
torch.manual_seed(1234)
m = pb.get_model("raft", args=Namespace(model=Namespace(iters=12))).eval().cuda().half()
xs = [torch.rand(a.batch, 2, 3, 436, 1024, device="cuda", dtype=torch.float16) for _ in range(a.streams)]
streams = [torch.cuda.Stream() for _ in range(a.streams)]

bar = threading.Barrier(a.streams + 1)

def worker(i, n):
    # cuDNN's autotune cache is thread-local in torch: warm up in the thread that runs the timed steps
    with torch.no_grad(), torch.cuda.stream(streams[i]):
        for _ in range(3):
            m({"images": xs[i]})
        streams[i].synchronize()
        bar.wait()
        for _ in range(n):
            m({"images": xs[i]})
        streams[i].synchronize()
        bar.wait()

with torch.no_grad():
    for _ in range(2): m({"images": xs[0]})
torch.cuda.synchronize()
ts = [threading.Thread(target=worker, args=(i, a.steps)) for i in range(a.streams)]
for t in ts: t.start()
bar.wait()
t0 = time.perf_counter()
bar.wait()
dt = time.perf_counter() - t0
for t in ts: t.join()
total = a.steps * a.streams
print(json.dumps({"streams": a.streams, "steps_total": total, "ms_per_step": round(dt / total * 1e3, 3), "pairs_per_s": round(total * a.batch / dt, 1)}))
