"""GPU parity tests, end to end: ptlflow_b200.get_model(...)(inputs) against vectors produced by the
reference model (tests/golden/e2e_*.npz) and against the oracle at other shapes.

Gates (BASELINE.json north_star): fp32 <= 1e-3 max-abs on the predicted flow.  f16/bf16 are compared to the
*fp32* reference.  north_star's 1e-2 is met by the MEAN-abs error in f16 and missed by the max-abs error by up to
2.3x (measured 0.009-0.023 px; bf16 0.05-0.16 px); the reference's own half model is 0.16-0.28 px (f16) / 1.7-4.2 px
(bf16) away from its fp32 output.  tools/f16_error_budget.py attributes > 90 % of the residual to single-pass f16
operand rounding of the context encoder and the GRU weights (static perturbations seen identically by every
iteration); halving it needs hi/lo split operands = two tensor-core passes.  The gates are the measured bounds with
~1.7x margin, not the old 0.25 / 2.0.
"""
import json
import os
from argparse import Namespace

import numpy as np
import pytest
import torch

from helpers import E2E, e2e_inputs, load_golden
from oracle import raft_oracle as O
from oracle import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
REPORT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "parity_report.jsonl")


def _report(**kw):
    try:
        os.makedirs(os.path.dirname(REPORT), exist_ok=True)
        with open(REPORT, "a") as f:
            f.write(json.dumps(kw) + "\n")
    except OSError:
        pass


def _build(variant, kwargs, sd, dtype=torch.float32, impl=0):
    import ptlflow_b200 as pb

    model = pb.get_model(variant, args=Namespace(model=Namespace(**kwargs)))
    missing = model.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    model = model.eval().to(DEV)
    if dtype != torch.float32:
        model = model.to(dtype)
    model.kernel_impl = impl
    return model


@pytest.mark.parametrize("name", E2E)
def test_fp32_matches_reference_vectors(name):
    recipe, g = load_golden(name)
    sd, img, kw = e2e_inputs(recipe)
    model = _build(recipe["variant"], kw, sd)
    with torch.no_grad():
        out = model({"images": img.to(DEV)})
    assert out["flows"].shape == g["flows"].shape
    assert out["flow_small"].shape == g["flow_small"].shape
    err_small = np.abs(out["flow_small"].cpu().numpy() - g["flow_small"]).max()
    err = np.abs(out["flows"].cpu().numpy() - g["flows"]).max()
    _report(test="fp32_golden", case=name, err_flow=float(err), err_flow_small=float(err_small), max_flow=float(np.abs(g["flows"]).max()))
    assert err < 1e-3, f"{name}: max-abs flow error {err}"
    assert err_small < 1e-3


@pytest.mark.parametrize("name", ["e2e_raft_noise", "e2e_raft_smooth_b2", "e2e_raft_altcorr", "e2e_gma"])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_half_against_fp32_reference(name, dtype):
    """Storage in f16/bf16, coordinates / accumulators / gates in fp32.  The bound asserted here is
    what this backend achieves against the reference's FP32 output; the reference's own half model
    is 0.16-0.28 px (f16) and 1.7-4.2 px (bf16) away from it."""
    recipe, g = load_golden(name)
    sd, img, kw = e2e_inputs(recipe)
    model = _build(recipe["variant"], kw, sd, dtype)
    with torch.no_grad():
        out = model({"images": img.to(DEV, dtype)})
    assert out["flows"].dtype == dtype
    err = np.abs(out["flows_fp32"].cpu().numpy() - g["flows"]).max()
    mean = np.abs(out["flows_fp32"].cpu().numpy() - g["flows"]).mean()
    _report(test="half_vs_fp32_ref", case=name, dtype=str(dtype), err_flow=float(err), mean_err=float(mean), max_flow=float(np.abs(g["flows"]).max()))
    bound, mean_bound = (4e-2, 1e-2) if dtype == torch.float16 else (3e-1, 6e-2)
    assert err < bound, f"{name} {dtype}: max-abs flow error {err}"
    assert mean < mean_bound, f"{name} {dtype}: mean-abs flow error {mean}"


@pytest.mark.parametrize("variant,iters,b,h,w", [("raft", 3, 1, 436, 1024), ("raft_small", 2, 2, 200, 328)])
def test_fp32_matches_oracle_at_larger_shapes(variant, iters, b, h, w):
    """Config-2 image size (436x1024 -> 55x128 grid, odd height: pooling floors, 4 px padding)."""
    shapes = O.state_dict_shapes(variant)
    sd = synth.synth_state_dict(shapes, 77)
    img = torch.from_numpy(synth.synth_images(b, h, w, 78, "smooth"))
    ref = O.raft_forward(sd, img, variant, iters=iters)
    model = _build(variant, dict(iters=iters), sd)
    with torch.no_grad():
        out = model({"images": img.to(DEV)})
    err = (out["flows"].cpu() - ref["flows"]).abs().max().item()
    _report(test="fp32_oracle_large", case=f"{variant}_{h}x{w}", err_flow=err, max_flow=ref["flows"].abs().max().item())
    assert out["flows"].shape == (b, 1, 2, h, w)
    assert err < 1e-3


def test_warm_start_and_input_not_mutated():
    shapes = O.state_dict_shapes("raft_small")
    sd = synth.synth_state_dict(shapes, 5)
    img = torch.from_numpy(synth.synth_images(1, 128, 160, 6, "smooth")).to(DEV)
    keep = img.clone()
    model = _build("raft_small", dict(iters=2), sd)
    with torch.no_grad():
        out1 = model({"images": img})
        out2 = model({"images": img, "prev_preds": {"flow_small": out1["flow_small"]}})
    assert torch.equal(img, keep), "caller's images were modified (base_model.py:210-214 works on a copy)"
    # warm start == the oracle started from the same forward-interpolated flow
    from ptlflow_b200.utils.warm_start import forward_interpolate_batch

    init = forward_interpolate_batch(out1["flow_small"].cpu())
    ref = O.raft_forward(sd, img.cpu(), "raft_small", iters=2, flow_init=init)
    assert (out2["flows"].cpu() - ref["flows"]).abs().max().item() < 1e-3


def test_simt_and_auto_paths_agree_in_half():
    """kernel_impl=1 (SIMT, fp32 accumulate) vs auto (tcgen05 where available) on identical f16 inputs."""
    recipe, g = load_golden("e2e_raft_noise")
    sd, img, kw = e2e_inputs(recipe)
    outs = []
    for impl in (1, 0):
        model = _build(recipe["variant"], kw, sd, torch.float16, impl=impl)
        with torch.no_grad():
            outs.append(model({"images": img.to(DEV, torch.float16)})["flows_fp32"].cpu())
    d = (outs[0] - outs[1]).abs().max().item()
    _report(test="simt_vs_auto_f16", err=d)
    assert d < 0.1


def test_pipeline_matches_sequential():
    """Two batches in flight on two streams / host threads (ptlflow_b200.pipeline) give the results of sequential calls:
    scratch buffers are per stream, nothing is shared but read-only weights."""
    from ptlflow_b200.pipeline import FramePipeline

    shapes = O.state_dict_shapes("raft")
    sd = synth.synth_state_dict(shapes, 91)
    model = _build("raft", dict(iters=4), sd, torch.float16)
    imgs = [torch.from_numpy(synth.synth_images(2, 128, 192, 300 + k, "smooth" if k % 2 else "noise")).to(DEV, torch.float16) for k in range(4)]
    with torch.no_grad():
        ref = [model({"images": x})["flows"].float().cpu() for x in imgs]
    host_in = imgs[1].cpu().pin_memory()
    host_out = torch.empty(ref[1].shape, dtype=torch.float16).pin_memory()
    with FramePipeline(model, depth=2) as pipe:
        res = [pipe.submit({"images": imgs[k % 4]}) for k in range(12)]
        hres = pipe.submit({"images": host_in}, host_out=host_out)  # pinned host frames in, flow copied back on the slot's stream
        outs = [r.get()["flows"].float().cpu() for r in res]
        hres.get()
    for k, o in enumerate(outs):
        d = (o - ref[k % 4]).abs().max().item()
        assert d < 2e-2, f"batch {k}: pipelined result differs from the sequential one by {d} px"
    assert (host_out.float() - ref[1]).abs().max().item() < 2e-2
