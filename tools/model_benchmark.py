#!/usr/bin/env python
"""The reference's model_benchmark.py protocol, run against this backend.

The reference script (model_benchmark.py:230-330, :421-466) needs jsonargparse / lightning / plotly for its CLI and
plots; what it *measures* needs none of them.  This tool keeps the measuring part call for call:

  * per trial: ``ptlflow.get_model(name, args=...)`` -> ``.eval()`` -> ``.cuda()`` -> ``.half()`` (fp16), ``count_parameters``;
  * per sample: a fresh ``torch.rand(batch, 2, 3, H, W)`` made on the CPU, moved / converted OUTSIDE the timed region,
    ``Timer.tic()`` (``cuda.synchronize`` + clock) ... ``model(inputs)`` ... ``Timer.toc()`` (``synchronize`` + clock);
  * the first forward of every trial and the whole first trial are discarded; the result is the median over the rest,
    reported as ms per frame pair (time / batch) like the reference's CSV column, and as pairs/s.

``ptlflow`` is this package under the reference's name: ``sys.modules["ptlflow"] = ptlflow_b200`` (tests/test_gpu_dropin.py
runs exactly that).  In a ptlflow checkout the same effect is one line in ``ptlflow/models/raft/__init__.py``
(INTEGRATION.md section 1).

    python tools/model_benchmark.py --model raft --model.iters 12 --input_size 436 1024 --datatypes fp16 --batch_size 8
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
from argparse import Namespace

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import ptlflow_b200  # noqa: E402

sys.modules.setdefault("ptlflow", ptlflow_b200)
import ptlflow  # noqa: E402  (this backend under the reference's package name)
from ptlflow_b200.utils.timer import Timer  # noqa: E402
from ptlflow_b200.utils.utils import count_parameters  # noqa: E402


@torch.no_grad()
def estimate_inference_time(args: Namespace, model, input_size, dtype_str: str):
    """Seconds per frame pair for ``num_samples`` forwards (one more is run first and dropped).  model_benchmark.py:421-466."""
    timer = Timer("inference")
    time_vals = []
    for i in range(args.num_samples + 1):
        inputs = {"images": torch.rand(args.batch_size, 2, 3, input_size[0], input_size[1])}
        if torch.cuda.is_available():
            inputs["images"] = inputs["images"].cuda()
            if dtype_str == "fp16":
                inputs["images"] = inputs["images"].half()
        if i > 0:
            timer.reset()
            timer.tic()
        model(inputs)
        if i > 0:
            timer.toc()
            time_vals.append(timer.total() / args.batch_size)
    return time_vals


def benchmark(args: Namespace):
    rows = []
    model_args = Namespace(model=Namespace(**args.model_kwargs))
    for isize in range(0, len(args.input_size), 2):
        input_size = args.input_size[isize : isize + 2]
        for dtype_str in args.datatypes:
            all_times = []
            params = 0
            for irep in range(args.num_trials + 1):
                torch.cuda.empty_cache()
                model = ptlflow.get_model(args.model, args=model_args).eval()
                if torch.cuda.is_available():
                    model = model.cuda()
                    if dtype_str == "fp16":
                        model = model.half()
                params = count_parameters(model)
                times = estimate_inference_time(args, model, input_size, dtype_str)
                if irep > 0:  # first trial dropped (model_benchmark.py:284)
                    all_times.extend(times)
                model = None
            med = statistics.median(all_times)
            rows.append({"model": args.model, "params": params, "input_size": list(input_size), "dtype": dtype_str, "batch_size": args.batch_size,
                         "samples": len(all_times), "time_ms_per_pair": round(1e3 * med, 4), "pairs_per_s": round(1.0 / med, 2),
                         "time_ms_min": round(1e3 * min(all_times), 4), "time_ms_max": round(1e3 * max(all_times), 4)})
    return rows


def parse_args(argv=None) -> Namespace:
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="raft")
    ap.add_argument("--input_size", type=int, nargs="+", default=[436, 1024])
    ap.add_argument("--datatypes", nargs="+", default=["fp16"], choices=["fp32", "fp16"])
    ap.add_argument("--batch_size", type=int, default=1)
    ap.add_argument("--num_samples", type=int, default=10)
    ap.add_argument("--num_trials", type=int, default=1)
    args, rest = ap.parse_known_args(argv)
    kw = {}
    it = iter(rest)
    for tok in it:  # --model.<kw> <value> pairs, as the reference's CLI names them
        if not tok.startswith("--model."):
            raise SystemExit(f"unknown argument {tok}")
        val = next(it)
        for cast in (int, float):
            try:
                val = cast(val)
                break
            except ValueError:
                pass
        if val in ("true", "True", "false", "False"):
            val = val in ("true", "True")
        kw[tok[len("--model."):]] = val
    args.model_kwargs = kw
    return args


if __name__ == "__main__":
    for row in benchmark(parse_args()):
        print(json.dumps(row))
