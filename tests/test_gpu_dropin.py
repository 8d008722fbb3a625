"""Script-level drop-in (north_star: "infer.py, validate.py and model_benchmark.py are drop-in"):
the model-facing calls of the reference's scripts, run against this backend installed under the reference's package name.

The scripts' own CLI / data modules (jsonargparse, lightning, plotly) are outside the hot path and absent on the GPU box;
what is exercised here is everything they do with a *model*: ``ptlflow.get_model`` -> ``.eval().cuda().half()`` ->
``estimate_inference_time`` (model_benchmark.py:421-466) and ``IOAdapter.prepare_inputs`` -> ``model(inputs)`` ->
``io_adapter.unscale`` -> ``tensor_dict_to_numpy`` (infer.py:141-196), ``validation_step`` / ``test_step``
(validate.py:411, base_model.py:366-430).
"""
import importlib
import os
import sys
from argparse import Namespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def ptlflow():
    """``import ptlflow`` resolves to this backend, submodules included (what a maintainer's binding achieves)."""
    import ptlflow_b200

    saved = {k: v for k, v in sys.modules.items() if k == "ptlflow" or k.startswith("ptlflow.")}
    for k in saved:
        del sys.modules[k]
    sys.modules["ptlflow"] = ptlflow_b200
    for name, mod in list(sys.modules.items()):
        if name.startswith("ptlflow_b200."):
            sys.modules["ptlflow." + name[len("ptlflow_b200."):]] = mod
    try:
        yield importlib.import_module("ptlflow")
    finally:
        for k in [k for k in sys.modules if k == "ptlflow" or k.startswith("ptlflow.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_model_benchmark_protocol(ptlflow):
    """estimate_inference_time-shaped loop: fresh torch.rand per sample, synchronise around each forward, first forward
    and first trial dropped, median -- on raft f16 at a small size (the BASELINE size is what bench.py's `protocol` runs)."""
    from ptlflow.models.base_model.base_model import BaseModel  # the import lines of model_benchmark.py:34-39
    from ptlflow.utils.timer import Timer  # noqa: F401
    from ptlflow.utils.utils import count_parameters

    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import model_benchmark as mb

    args = mb.parse_args(["--model", "raft", "--model.iters", "4", "--input_size", "184", "320", "--datatypes", "fp16", "fp32",
                          "--batch_size", "2", "--num_samples", "3", "--num_trials", "1"])
    rows = mb.benchmark(args)
    assert [r["dtype"] for r in rows] == ["fp16", "fp32"]
    for r in rows:
        assert r["params"] == 5257536  # the reference's "Params" column for raft (docs: 5.258 M)
        assert r["samples"] == 3 and r["time_ms_per_pair"] > 0
    model = ptlflow.get_model("raft", args=Namespace(model=Namespace(iters=2)))
    assert isinstance(model, BaseModel) and count_parameters(model) == 5257536


def test_infer_call_shape(ptlflow):
    """infer.py:141-196: numpy BGR frames -> IOAdapter -> model(inputs) -> unscale -> numpy; fp16 like ``--fp16``."""
    from ptlflow.utils.io_adapter import IOAdapter
    from ptlflow.utils.utils import tensor_dict_to_numpy

    model = ptlflow.get_model("raft_small", args=Namespace(model=Namespace(iters=3))).eval().cuda().half()
    rng = np.random.default_rng(0)
    img1 = rng.integers(0, 255, (130, 170, 3), dtype=np.uint8)
    img2 = np.roll(img1, 2, axis=1)
    io_adapter = IOAdapter(output_stride=model.output_stride, input_size=img1.shape[:2], target_size=None, cuda=True, fp16=True)
    with torch.no_grad():
        inputs = io_adapter.prepare_inputs([img1, img2])
        assert inputs["images"].shape == (1, 2, 3, 130, 170) and inputs["images"].dtype == torch.float16 and inputs["images"].is_cuda
        keep = inputs["images"].clone()
        preds = model(inputs)
        assert torch.equal(inputs["images"], keep)  # infer.py:190 re-uses inputs["images"] afterwards
        preds["images"] = inputs["images"]
        preds = io_adapter.unscale(preds)
    assert preds["flows"].shape == (1, 1, 2, 130, 170) and preds["flows"].dtype == torch.float16
    assert preds["flow_small"].shape == (1, 2, 17, 22)
    npy = tensor_dict_to_numpy(preds)
    assert npy["flows"].shape == (130, 170, 2) and np.isfinite(npy["flows"]).all()
    # a rescaled run (--input_size): flows come back at the original size
    io2 = IOAdapter(output_stride=model.output_stride, input_size=img1.shape[:2], target_size=(160, 192), cuda=True, fp16=True)
    with torch.no_grad():
        p2 = io2.unscale(model(io2.prepare_inputs([img1, img2])))
    assert p2["flows"].shape == (1, 1, 2, 130, 170)


def test_validation_step_warm_start_order(ptlflow):
    """validate.py:411 -> base_model.validation_step: prev_preds goes in unconditionally and is dropped AFTER the forward of a
    batch that starts a sequence (base_model.py:396-430)."""
    model = ptlflow.get_model("raft_small", args=Namespace(model=Namespace(iters=2, warm_start=True))).eval().cuda()
    img = torch.rand(1, 2, 3, 128, 160, device="cuda")
    gt = torch.zeros(1, 1, 2, 128, 160, device="cuda")
    with torch.no_grad():
        o1 = model.validation_step({"images": img, "flows": gt, "meta": {"is_seq_start": [True]}}, 0)
        assert model.prev_preds is None  # dropped after a sequence start
        o2 = model.validation_step({"images": img, "flows": gt, "meta": {"is_seq_start": [False]}}, 1)
        assert model.prev_preds is not None and "flow_small" in model.prev_preds
        o3 = model.validation_step({"images": img, "flows": gt, "meta": {"is_seq_start": [False]}}, 2)
    d12 = (o1["preds"]["flows"] - o2["preds"]["flows"]).abs().max().item()
    d23 = (o2["preds"]["flows"] - o3["preds"]["flows"]).abs().max().item()
    assert d12 < 1e-4, d12  # both started cold (instance-norm sums use atomics: equal up to fp32 summation order)
    assert d23 > 1e-3, d23  # the third is warm-started
    assert "val/epe" in o3["metrics"]
