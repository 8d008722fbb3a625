"""This library's plugin entry point against the REFERENCE'S OWN native kernel (oracle/_ref/alt_cuda_corr.so, built from
/root/reference/ptlflow/utils/external/alt_cuda_corr by oracle/build_ref.py and shipped prebuilt to the GPU box):
``alt_cuda_corr.forward(fmap1, fmap2, coords, radius)`` -- same tensors in, same tensor out (correlation.cpp:23-33).
Skipped when the prebuilt file is absent (it is git-ignored; __graft_entry__.build() creates it where the reference is)."""
import pytest
import torch

from oracle import build_ref
from oracle import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ref_plugin():
    mod = build_ref.load()
    if mod is None:
        pytest.skip("oracle/_ref/alt_cuda_corr.so not built (needs /root/reference at build time)")
    return mod


@pytest.mark.parametrize("b,c,h1,w1,h2,w2,r", [(1, 256, 16, 24, 16, 24, 4), (2, 128, 17, 29, 8, 14, 4), (1, 64, 9, 12, 9, 12, 3), (1, 256, 55, 128, 27, 64, 4)])
def test_forward_matches_the_reference_kernel(ref_plugin, b, c, h1, w1, h2, w2, r):
    from ptlflow_b200 import alt_cuda_corr

    f1 = torch.from_numpy(synth.synth_normal("rp/f1", (b, h1, w1, c), 21)).to(DEV)
    f2 = torch.from_numpy(synth.synth_normal("rp/f2", (b, h2, w2, c), 21)).to(DEV)
    grid = torch.stack(torch.meshgrid(torch.arange(w1, dtype=torch.float32), torch.arange(h1, dtype=torch.float32), indexing="xy"), dim=-1)  # [h1,w1,2] (x, y)
    coords = (grid[None, None] * (w2 / w1) + torch.from_numpy(synth.synth_normal("rp/c", (b, 1, h1, w1, 2), 21, scale=3.0))).contiguous().to(DEV)
    (ours,) = alt_cuda_corr.forward(f1, f2, coords, r)
    (ref,) = ref_plugin.forward(f1, f2, coords, r)
    assert ours.shape == ref.shape and ours.dtype == ref.dtype == torch.float32
    scale = max(1.0, ref.abs().max().item())
    assert (ours - ref).abs().max().item() < 2e-5 * scale, (ours - ref).abs().max().item()
