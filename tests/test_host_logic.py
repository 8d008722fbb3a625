"""CPU: the Python boundary mirrors the reference's surface (names, hparams, state_dict keys,
padding rules, error behaviour) and the sharding helpers work under gloo with world_size 2."""
import json
import os
from argparse import Namespace

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import GOLDEN
from oracle import raft_oracle as O


def test_registry_and_get_model():
    import ptlflow_b200 as pb

    assert {"raft", "raft_small"} <= set(pb.get_model_names())
    assert pb.get_model_reference("raft").__name__ == "raft"
    with pytest.raises(ValueError):
        pb.get_model_reference("not_a_model")
    m = pb.get_model("raft", args=Namespace(model=Namespace(iters=12, corr_radius=3, alternate_corr=True)))
    assert (m.iters, m.corr_radius, m.alternate_corr, m.output_stride) == (12, 3, True, 8)
    assert m.hparams.iters == 12 and m.hparams.corr_radius == 3 and not hasattr(m.hparams, "loss_fn")
    assert m.update_block.encoder.convc1.weight.shape == (256, 4 * 49, 1, 1)
    m2 = pb.get_model("raft_small", args={"model": {"init_args": {"iters": 4}}})
    assert m2.iters == 4 and m2.hidden_dim == 96


@pytest.mark.parametrize("variant", ["raft", "raft_small", "gma"])
def test_state_dict_keys_equal_reference(variant):
    import ptlflow_b200 as pb

    with open(os.path.join(GOLDEN, f"state_shapes_{variant}.json")) as f:
        ref = {k: tuple(v) for k, v in json.load(f).items()}
    mine = {k: tuple(v.shape) for k, v in pb.get_model(variant).state_dict().items()}
    assert mine == ref


def test_checkpoint_roundtrip(tmp_path):
    import ptlflow_b200 as pb

    m = pb.get_model("raft_small")
    path = tmp_path / "m.ckpt"
    torch.save({"state_dict": m.state_dict(), "hyper_parameters": {"train_size": [368, 496], "extra_params": {"a": 1}}}, path)
    m2 = pb.get_model("raft_small", ckpt_path=str(path))
    assert m2.train_size == [368, 496] and m2.extra_params == {"a": 1}
    for (k1, v1), (k2, v2) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)
    with pytest.raises(ValueError):
        pb.get_model("raft_small", ckpt_path="not_a_checkpoint_name")


@pytest.mark.parametrize("h,w", [(436, 1024), (128, 256), (132, 164), (1080, 1920), (37, 41)])
def test_padding_rule(h, w):
    from ptlflow_b200.utils.utils import InputPadder

    x = torch.arange(2 * 2 * 3 * h * w, dtype=torch.float32).reshape(2, 2, 3, h, w)
    p = InputPadder(x.shape, stride=8)
    l, r, t, b = O.pad_amounts(h, w)
    assert p._pad == [l, r, t, b] and p.pad_top_left == (t, l)
    y = p.fill(x)
    assert y.shape[-2] % 8 == 0 and y.shape[-1] % 8 == 0
    ref, _ = O.preprocess(x / x.max())
    assert y.shape == ref.shape
    assert torch.equal(p.unfill(y), x)
    if (l, r, t, b) != (0, 0, 0, 0):
        assert p.unfill(x) is x  # already un-padded tensors pass through (utils.py:87-90)


def test_preprocess_matches_oracle_and_keeps_input():
    import ptlflow_b200 as pb

    m = pb.get_model("raft")
    img = torch.rand(1, 2, 3, 37, 50)
    keep = img.clone()
    x, resizer = m.preprocess_images(img, bgr_add=-0.5, bgr_mult=2.0, bgr_to_rgb=True, resize_mode="pad", pad_mode="replicate", pad_two_side=True)
    ref, pads = O.preprocess(img)
    assert torch.equal(img, keep)
    assert torch.allclose(x, ref, atol=0, rtol=0)


def test_no_cpu_fallback():
    import ptlflow_b200 as pb
    from ptlflow_b200 import ops

    m = pb.get_model("raft_small").eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        m({"images": torch.rand(1, 2, 3, 64, 64)})
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.corr_volume_build(torch.zeros(1, 8, 8, 16), torch.zeros(1, 8, 8, 16), 1)
    with pytest.raises(RuntimeError):
        m.update_block(None, None, None, None)  # parameter container, not a PyTorch implementation


def test_shard_range_tiles_exactly():
    from ptlflow_b200.sharding import shard_range

    for n in (0, 1, 7, 8, 9, 32, 1041):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _gloo_worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from ptlflow_b200 import sharding

    assert sharding.init_process_group("gloo")
    lo, hi = sharding.shard_range(9, rank, world)
    sharding.barrier()
    slowest = sharding.max_over_ranks(10.0 + rank)
    total = sharding.sum_over_ranks(hi - lo)
    out.put((rank, lo, hi, slowest, total))
    dist.destroy_process_group()


def test_gloo_world_size_2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 5), (5, 9)]
    assert all(r[3] == 11.0 and r[4] == 9.0 for r in res)
