"""Plain torch-fp32 CPU restatement of the reference's RAFT inference path.

TEST INFRASTRUCTURE (see oracle/__init__.py): the checker, never the product.

Every function restates one row of SURVEY.md section 8(a) and cites the reference
file:line (relative to the reference checkout, ``ptlflow/...``) it follows.  It is
written functionally over a ``state_dict`` (reference parameter names) and deliberately
avoids the reference's own building blocks where an independent formulation is cheap:
the bilinear lookup is an explicit 4-tap gather rather than ``F.grid_sample``, the convex
upsample is an explicit 9-tap loop rather than ``F.unfold``.  ``tests/test_oracle_golden.py``
pins it against vectors produced by the real reference (oracle/make_golden.py).

Parity status: PINNED against tests/golden/*.npz (reference outputs generated in the
build container).  Tolerance used by the pin: 2e-4 max-abs on flows for <= 12 iterations.

The functions follow the device of their inputs, so the same restatement also runs on CUDA tensors
(fp32, TF32 off: ``fp32_strict()`` below): that is how the parity tests check the BASELINE-sized shapes
in seconds and how bench.py times a same-GPU PyTorch comparator.  It stays the checker in both roles.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


class fp32_strict:
    """Context manager: true fp32 on CUDA (no TF32 in matmul / cuDNN), restored on exit."""

    def __enter__(self):
        self._saved = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        return self

    def __exit__(self, *exc):
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = self._saved


# --------------------------------------------------------------------------------------
# boundary pre/post-processing
# --------------------------------------------------------------------------------------
def pad_amounts(h: int, w: int, stride: int = 8) -> Tuple[int, int, int, int]:
    """(left, right, top, bottom) replicate padding, split evenly.
    ptlflow/utils/external/raft.py:46-70 (two_side_pad=True)."""
    ph = (((h // stride) + 1) * stride - h) % stride
    pw = (((w // stride) + 1) * stride - w) % stride
    return pw // 2, pw - pw // 2, ph // 2, ph - ph // 2


def preprocess(images: Tensor) -> Tuple[Tensor, Tuple[int, int, int, int]]:
    """images [B,2,3,H,W] BGR in [0,1] -> RGB in [-1,1], replicate-padded to a multiple of 8.
    ptlflow/models/raft/raft.py:127-135, ptlflow/models/base_model/base_model.py:206-246."""
    x = (images + (-0.5)) * 2.0
    x = torch.flip(x, dims=[-3])
    pads = pad_amounts(x.shape[-2], x.shape[-1])
    b, n = x.shape[:2]
    x = F.pad(x.reshape(b * n, *x.shape[2:]), pads, mode="replicate")
    return x.reshape(b, n, *x.shape[1:]).contiguous(), pads


def unpad(x: Tensor, pads: Tuple[int, int, int, int]) -> Tensor:
    """ptlflow/utils/external/raft.py:83-86."""
    l, r, t, b = pads
    h, w = x.shape[-2:]
    return x[..., t : h - b, l : w - r]


# --------------------------------------------------------------------------------------
# encoders (context for end-to-end parity)
# --------------------------------------------------------------------------------------
def _conv(x: Tensor, sd: SD, name: str, stride=1, padding=0) -> Tensor:
    return F.conv2d(x, sd[name + ".weight"], sd.get(name + ".bias"), stride=stride, padding=padding)


def _norm(x: Tensor, sd: SD, name: str, kind: str) -> Tensor:
    if kind == "instance":  # nn.InstanceNorm2d defaults: no affine, no running stats
        return F.instance_norm(x, eps=1e-5)
    if kind == "batch":  # eval mode: running statistics
        return F.batch_norm(
            x, sd[name + ".running_mean"], sd[name + ".running_var"], sd[name + ".weight"], sd[name + ".bias"],
            training=False, eps=1e-5,
        )
    if kind == "none":
        return x
    raise ValueError(kind)


def _residual_block(x: Tensor, sd: SD, p: str, kind: str, stride: int) -> Tensor:
    """ptlflow/models/raft/extractor.py:6-58."""
    y = torch.relu(_norm(_conv(x, sd, p + "conv1", stride=stride, padding=1), sd, p + "norm1", kind))
    y = torch.relu(_norm(_conv(y, sd, p + "conv2", padding=1), sd, p + "norm2", kind))
    if stride != 1:
        x = _norm(_conv(x, sd, p + "downsample.0", stride=stride), sd, p + "downsample.1", kind)  # same module as norm3; this key is loaded last
    return torch.relu(x + y)


def _bottleneck_block(x: Tensor, sd: SD, p: str, kind: str, stride: int) -> Tensor:
    """ptlflow/models/raft/extractor.py:61-119."""
    y = torch.relu(_norm(_conv(x, sd, p + "conv1"), sd, p + "norm1", kind))
    y = torch.relu(_norm(_conv(y, sd, p + "conv2", stride=stride, padding=1), sd, p + "norm2", kind))
    y = torch.relu(_norm(_conv(y, sd, p + "conv3"), sd, p + "norm3", kind))
    if stride != 1:
        x = _norm(_conv(x, sd, p + "downsample.0", stride=stride), sd, p + "downsample.1", kind)  # same module as norm4
    return torch.relu(x + y)


def encoder(x: Tensor, sd: SD, prefix: str, kind: str, small: bool) -> Tensor:
    """BasicEncoder / SmallEncoder forward, eval mode.
    ptlflow/models/raft/extractor.py:171-194 and :246-267."""
    block = _bottleneck_block if small else _residual_block
    x = torch.relu(_norm(_conv(x, sd, prefix + "conv1", stride=2, padding=3), sd, prefix + "norm1", kind))
    for li, stride in ((1, 1), (2, 2), (3, 2)):
        x = block(x, sd, f"{prefix}layer{li}.0.", kind, stride)
        x = block(x, sd, f"{prefix}layer{li}.1.", kind, 1)
    return _conv(x, sd, prefix + "conv2")


# --------------------------------------------------------------------------------------
# a1 / a2: all-pairs correlation volume and its pooled pyramid
# --------------------------------------------------------------------------------------
def corr_volume(fmap1: Tensor, fmap2: Tensor) -> Tensor:
    """fmap [B,C,H,W] -> volume [B*H*W, 1, H, W] = <f1(q), f2(t)> / sqrt(C).
    ptlflow/models/raft/corr.py:56-64 and the reshape at :21-22."""
    b, c, h, w = fmap1.shape
    a = fmap1.reshape(b, c, h * w).transpose(1, 2)
    v = torch.bmm(a, fmap2.reshape(b, c, h * w)) / math.sqrt(c)
    return v.reshape(b * h * w, 1, h, w)


def corr_pyramid(volume: Tensor, num_levels: int) -> List[Tensor]:
    """2x2 mean pooling over the *target* axes, floor sizes. ptlflow/models/raft/corr.py:24-27."""
    pyr = [volume]
    for _ in range(num_levels - 1):
        v = pyr[-1]
        hh, ww = v.shape[-2] // 2, v.shape[-1] // 2
        v = v[..., : 2 * hh, : 2 * ww]
        v = 0.25 * (v[..., 0::2, 0::2] + v[..., 0::2, 1::2] + v[..., 1::2, 0::2] + v[..., 1::2, 1::2])
        pyr.append(v.contiguous())
    return pyr


def _bilinear_zero(img: Tensor, x: Tensor, y: Tensor) -> Tensor:
    """img [N,H,W]; x,y [N,K] pixel coordinates -> [N,K].  Bilinear, align_corners=True
    (pixel coordinate == index), taps outside the map contribute zero.
    ptlflow/models/raft/utils.py:67-81 (grid_sample semantics), SURVEY 8(a) note 3."""
    n, h, w = img.shape
    x0, y0 = torch.floor(x), torch.floor(y)
    wx, wy = x - x0, y - y0
    flat = img.reshape(n, h * w)
    out = torch.zeros_like(x)
    for dy, dx, wgt in ((0, 0, (1 - wx) * (1 - wy)), (0, 1, wx * (1 - wy)), (1, 0, (1 - wx) * wy), (1, 1, wx * wy)):
        xi, yi = (x0 + dx).long(), (y0 + dy).long()
        ok = (xi >= 0) & (xi < w) & (yi >= 0) & (yi < h)
        idx = yi.clamp(0, h - 1) * w + xi.clamp(0, w - 1)
        out = out + torch.gather(flat, 1, idx) * (wgt * ok.to(img.dtype))
    return out


def corr_lookup(pyramid: Sequence[Tensor], coords: Tensor, radius: int) -> Tensor:
    """a3.  coords [B,2,H,W] (channel 0 = x) -> [B, L*(2r+1)^2, H, W].
    Window entry (i,j) samples (x + (i-r), y + (j-r)) of level l at coords / 2**l and lands in
    channel l*(2r+1)^2 + i*(2r+1) + j  -- x-major.  ptlflow/models/raft/corr.py:29-54."""
    b, _, h, w = coords.shape
    n = b * h * w
    cx = coords[:, 0].reshape(n, 1, 1)
    cy = coords[:, 1].reshape(n, 1, 1)
    d = torch.arange(-radius, radius + 1, dtype=coords.dtype, device=coords.device)
    k = 2 * radius + 1
    outs = []
    for lvl, vol in enumerate(pyramid):
        px = (cx / 2**lvl + d.view(1, k, 1)).expand(n, k, k).reshape(n, k * k)
        py = (cy / 2**lvl + d.view(1, 1, k)).expand(n, k, k).reshape(n, k * k)
        outs.append(_bilinear_zero(vol[:, 0], px, py).view(b, h, w, k * k))
    return torch.cat(outs, dim=-1).permute(0, 3, 1, 2).contiguous()


def alt_corr_lookup(fmap1: Tensor, fmap2: Tensor, coords: Tensor, radius: int, num_levels: int) -> Tensor:
    """a4 / a5: same values as corr_lookup(corr_pyramid(corr_volume(...))) without keeping the
    volume: level l correlates full-resolution fmap1 with fmap2 average-pooled l times, the
    1/sqrt(C) scale applied to the looked-up values.  ptlflow/models/raft/corr.py:67-101,
    kernel semantics ptlflow/utils/external/alt_cuda_corr/correlation_kernel.cu:59-116,
    pure-torch equivalent ptlflow/utils/correlation.py:581-615.
    (Restated through linearity: sampling bilinearly commutes with the channel dot product.)"""
    b, c, h, w = fmap1.shape
    a = fmap1.reshape(b, c, h * w).transpose(1, 2)
    pyr, f2 = [], fmap2
    for lvl in range(num_levels):
        if lvl > 0:
            f2 = F.avg_pool2d(f2, 2, stride=2)
        v = torch.bmm(a, f2.reshape(b, c, -1)).reshape(b * h * w, 1, f2.shape[-2], f2.shape[-1])
        pyr.append(v)
    return corr_lookup(pyr, coords, radius) / math.sqrt(c)


def alt_cuda_corr_forward(fmap1: Tensor, fmap2: Tensor, coords: Tensor, radius: int) -> Tensor:
    """The reference's native plugin entry point, one level, no scaling.
    fmap1 [B,H1,W1,C], fmap2 [B,H2,W2,C] (NHWC), coords [B,1,H1,W1,2] -> [B,1,(2r+1)^2,H1,W1].
    ptlflow/utils/external/alt_cuda_corr/correlation.cpp:23-33, correlation_kernel.cu:18-119."""
    b, h1, w1, c = fmap1.shape
    h2, w2 = fmap2.shape[1:3]
    v = torch.bmm(fmap1.reshape(b, h1 * w1, c), fmap2.reshape(b, h2 * w2, c).transpose(1, 2))
    v = v.reshape(b * h1 * w1, 1, h2, w2)
    cc = coords[:, 0].permute(0, 3, 1, 2)  # [B,2,H1,W1]
    return corr_lookup([v], cc, radius).unsqueeze(1)


def coords_grid(b: int, h: int, w: int, dtype=torch.float32, device=None) -> Tensor:
    """[B,2,H,W], channel 0 = x, channel 1 = y.  ptlflow/models/raft/utils.py:84-91."""
    ys, xs = torch.meshgrid(torch.arange(h, dtype=dtype, device=device), torch.arange(w, dtype=dtype, device=device), indexing="ij")
    return torch.stack([xs, ys], dim=0)[None].repeat(b, 1, 1, 1)


# --------------------------------------------------------------------------------------
# a6 - a9, a12: update blocks
# --------------------------------------------------------------------------------------
def motion_encoder_basic(flow: Tensor, corr: Tensor, sd: SD, p: str = "update_block.encoder.") -> Tensor:
    """ptlflow/models/raft/update.py:94-112."""
    cor = torch.relu(_conv(corr, sd, p + "convc1"))
    cor = torch.relu(_conv(cor, sd, p + "convc2", padding=1))
    flo = torch.relu(_conv(flow, sd, p + "convf1", padding=3))
    flo = torch.relu(_conv(flo, sd, p + "convf2", padding=1))
    out = torch.relu(_conv(torch.cat([cor, flo], 1), sd, p + "conv", padding=1))
    return torch.cat([out, flow], 1)


def motion_encoder_small(flow: Tensor, corr: Tensor, sd: SD, p: str = "update_block.encoder.") -> Tensor:
    """ptlflow/models/raft/update.py:76-91."""
    cor = torch.relu(_conv(corr, sd, p + "convc1"))
    flo = torch.relu(_conv(flow, sd, p + "convf1", padding=3))
    flo = torch.relu(_conv(flo, sd, p + "convf2", padding=1))
    out = torch.relu(_conv(torch.cat([cor, flo], 1), sd, p + "conv", padding=1))
    return torch.cat([out, flow], 1)


def _gru_half(h: Tensor, x: Tensor, sd: SD, p: str, suffix: str, padding) -> Tensor:
    hx = torch.cat([h, x], 1)
    z = torch.sigmoid(_conv(hx, sd, p + "convz" + suffix, padding=padding))
    r = torch.sigmoid(_conv(hx, sd, p + "convr" + suffix, padding=padding))
    q = torch.tanh(_conv(torch.cat([r * h, x], 1), sd, p + "convq" + suffix, padding=padding))
    return (1 - z) * h + z * q


def sep_conv_gru(h: Tensor, x: Tensor, sd: SD, p: str = "update_block.gru.") -> Tensor:
    """Horizontal (1x5) then vertical (5x1) gated update.  ptlflow/models/raft/update.py:58-73."""
    h = _gru_half(h, x, sd, p, "1", (0, 2))
    return _gru_half(h, x, sd, p, "2", (2, 0))


def conv_gru(h: Tensor, x: Tensor, sd: SD, p: str = "update_block.gru.") -> Tensor:
    """3x3 ConvGRU of raft_small.  ptlflow/models/raft/update.py:24-32."""
    return _gru_half(h, x, sd, p, "", 1)


def flow_head(net: Tensor, sd: SD, p: str = "update_block.flow_head.") -> Tensor:
    """ptlflow/models/raft/update.py:13-14."""
    return _conv(torch.relu(_conv(net, sd, p + "conv1", padding=1)), sd, p + "conv2", padding=1)


def mask_head(net: Tensor, sd: SD, p: str = "update_block.mask.") -> Tensor:
    """0.25 * conv1x1(relu(conv3x3(net))).  ptlflow/models/raft/update.py:138-142,:152."""
    return 0.25 * _conv(torch.relu(_conv(net, sd, p + "0", padding=1)), sd, p + "2")


def basic_update_block(net, inp, corr, flow, sd: SD):
    """-> (net, mask, delta_flow).  ptlflow/models/raft/update.py:144-153."""
    motion = motion_encoder_basic(flow, corr, sd)
    net = sep_conv_gru(net, torch.cat([inp, motion], 1), sd)
    return net, mask_head(net, sd), flow_head(net, sd)


def small_update_block(net, inp, corr, flow, sd: SD):
    """-> (net, None, delta_flow).  ptlflow/models/raft/update.py:122-128."""
    motion = motion_encoder_small(flow, corr, sd)
    net = conv_gru(net, torch.cat([inp, motion], 1), sd)
    return net, None, flow_head(net, sd)


# --------------------------------------------------------------------------------------
# a10: upsampling
# --------------------------------------------------------------------------------------
def convex_upsample(flow: Tensor, mask: Tensor) -> Tensor:
    """flow [B,2,H,W], mask [B,576,H,W] -> [B,2,8H,8W].  Mask channel = tap*64 + sy*8 + sx with
    tap = 3*(dy+1) + (dx+1); softmax over the 9 taps; neighbours of 8*flow, zero outside.
    ptlflow/models/raft/raft.py:112-123."""
    b, _, h, w = flow.shape
    m = torch.softmax(mask.view(b, 9, 8, 8, h, w), dim=1)
    f = F.pad(8.0 * flow, (1, 1, 1, 1))
    out = torch.zeros(b, 2, 8, 8, h, w, dtype=flow.dtype, device=flow.device)
    for tap in range(9):
        dy, dx = tap // 3, tap % 3
        nb = f[:, :, dy : dy + h, dx : dx + w]  # [B,2,H,W]
        out = out + m[:, tap][:, None] * nb[:, :, None, None]
    # out[b,c,sy,sx,y,x] -> [b,c,8y+sy,8x+sx]
    return out.permute(0, 1, 4, 2, 5, 3).reshape(b, 2, 8 * h, 8 * w)


def upflow8(flow: Tensor) -> Tensor:
    """8 * bilinear (align_corners=True) 8x resize.  ptlflow/models/raft/utils.py:94-96."""
    h, w = flow.shape[-2:]
    return 8.0 * F.interpolate(flow, size=(8 * h, 8 * w), mode="bilinear", align_corners=True)


# --------------------------------------------------------------------------------------
# a11: the forward loop
# --------------------------------------------------------------------------------------
VARIANTS = {
    # name: (small, hidden, context, fnet_dim, cnet_norm, default_radius)
    "raft": (False, 128, 128, 256, "batch", 4),
    "raft_small": (True, 96, 64, 128, "none", 3),
    "gma": (False, 128, 128, 256, "batch", 4),
}


# --------------------------------------------------------------------------------------
# a13: GMA extras (content-only attention, num_heads = 1: the registered default)
# --------------------------------------------------------------------------------------
def gma_attention(inp: Tensor, sd: SD, p: str = "att.") -> Tensor:
    """softmax_j( scale * q_i . k_j ), scale = dim_head^-1/2, q|k = to_qk(inp).  -> [B, N, N]
    ptlflow/models/gma/gma_utils.py:58-76 (position_only = position_and_content = False)."""
    b, c, h, w = inp.shape
    qk = F.conv2d(inp, sd[p + "to_qk.weight"])
    q, k = qk[:, :c].reshape(b, c, h * w), qk[:, c:].reshape(b, c, h * w)
    sim = torch.bmm((q * c ** -0.5).transpose(1, 2), k)
    return torch.softmax(sim, dim=-1)


def gma_aggregate(attn: Tensor, fmap: Tensor, sd: SD, p: str = "update_block.aggregator.") -> Tensor:
    """fmap + gamma * (attn @ to_v(fmap)).  ptlflow/models/gma/gma_utils.py:101-113 (heads = 1: no project)."""
    b, c, h, w = fmap.shape
    v = F.conv2d(fmap, sd[p + "to_v.weight"]).reshape(b, c, h * w)
    out = torch.bmm(attn, v.transpose(1, 2)).transpose(1, 2).reshape(b, c, h, w)
    return fmap + sd[p + "gamma"] * out


def gma_update_block(net, inp, corr, flow, attn, sd: SD):
    """-> (net, mask, delta_flow).  ptlflow/models/gma/update.py:148-160."""
    motion = motion_encoder_basic(flow, corr, sd)
    motion_global = gma_aggregate(attn, motion, sd)
    net = sep_conv_gru(net, torch.cat([inp, motion, motion_global], 1), sd)
    return net, mask_head(net, sd), flow_head(net, sd)


def raft_forward(
    sd: SD,
    images: Tensor,
    variant: str = "raft",
    iters: int = 12,
    corr_levels: int = 4,
    corr_radius: Optional[int] = None,
    alternate_corr: bool = False,
    flow_init: Optional[Tensor] = None,
    trace: Optional[dict] = None,
) -> Dict[str, Tensor]:
    """Eval-mode RAFT.forward.  ptlflow/models/raft/raft.py:125-194.
    ``trace`` (if a dict) receives per-stage tensors for operator-level comparisons."""
    small, hdim, cdim, _fdim, cnorm, r_default = VARIANTS[variant]
    radius = r_default if corr_radius is None else corr_radius
    sd = {k: v.float() for k, v in sd.items() if v.is_floating_point()}
    x, pads = preprocess(images.float())
    img1, img2 = x[:, 0], x[:, 1]
    b = img1.shape[0]

    fmaps = encoder(torch.cat([img1, img2], 0), sd, "fnet.", "instance", small)
    fmap1, fmap2 = fmaps[:b], fmaps[b:]
    cnet = encoder(img1, sd, "cnet.", cnorm, small)
    net, inp = torch.tanh(cnet[:, :hdim]), torch.relu(cnet[:, hdim : hdim + cdim])

    pyramid = None if alternate_corr else corr_pyramid(corr_volume(fmap1, fmap2), corr_levels)
    h8, w8 = fmap1.shape[-2:]
    coords0 = coords_grid(b, h8, w8, device=fmap1.device)
    coords1 = coords0.clone() if flow_init is None else coords0 + flow_init
    if trace is not None:
        trace.update(fmap1=fmap1, fmap2=fmap2, net0=net, inp=inp, lookups=[], nets=[], deltas=[])
        if pyramid is not None:
            trace["pyramid"] = pyramid

    block = small_update_block if small else basic_update_block
    attn = gma_attention(inp, sd) if variant == "gma" else None  # ptlflow/models/gma/gma.py:181
    if trace is not None and attn is not None:
        trace["attention"] = attn
    mask = None
    for _ in range(iters):
        if alternate_corr:
            corr = alt_corr_lookup(fmap1, fmap2, coords1, radius, corr_levels)
        else:
            corr = corr_lookup(pyramid, coords1, radius)
        if attn is not None:
            net, mask, delta = gma_update_block(net, inp, corr, coords1 - coords0, attn, sd)
        else:
            net, mask, delta = block(net, inp, corr, coords1 - coords0, sd)
        coords1 = coords1 + delta
        if trace is not None:
            trace["lookups"].append(corr)
            trace["nets"].append(net)
            trace["deltas"].append(delta)

    flow_small = coords1 - coords0
    up = upflow8(flow_small) if mask is None else convex_upsample(flow_small, mask)
    if trace is not None:
        trace["mask"] = mask
    return {"flows": unpad(up, pads)[:, None], "flow_small": flow_small}


def state_dict_shapes(variant: str, corr_levels: int = 4, corr_radius: Optional[int] = None) -> Dict[str, Tuple[int, ...]]:
    """Ordered name -> shape of the reference model's state_dict (what restore_model loads,
    ptlflow/__init__.py:282), derived from the architecture so the GPU box can rebuild
    synth weights without the reference.  Checked against the real reference in
    tests/test_oracle_golden.py via tests/golden/state_shapes_*.json."""
    small, hdim, cdim, fdim, cnorm, r_default = VARIANTS[variant]
    radius = r_default if corr_radius is None else corr_radius
    cor_planes = corr_levels * (2 * radius + 1) ** 2
    s: Dict[str, Tuple[int, ...]] = {}

    def conv(name, cout, cin, kh, kw):
        s[name + ".weight"] = (cout, cin, kh, kw)
        s[name + ".bias"] = (cout,)

    def bn(name, c):
        s[name + ".weight"] = (c,)
        s[name + ".bias"] = (c,)
        s[name + ".running_mean"] = (c,)
        s[name + ".running_var"] = (c,)
        s[name + ".num_batches_tracked"] = ()

    def enc(prefix, out_dim, kind):
        dims = (32, 32, 64, 96) if small else (64, 64, 96, 128)
        if kind == "batch":
            bn(prefix + "norm1", dims[0])
        conv(prefix + "conv1", dims[0], 3, 7, 7)
        cin = dims[0]
        for li, (dim, stride) in enumerate(zip(dims[1:], (1, 2, 2)), start=1):
            for bi in range(2):
                p = f"{prefix}layer{li}.{bi}."
                st = stride if bi == 0 else 1
                if small:
                    conv(p + "conv1", dim // 4, cin, 1, 1)
                    conv(p + "conv2", dim // 4, dim // 4, 3, 3)
                    conv(p + "conv3", dim, dim // 4, 1, 1)
                    norms = [("norm1", dim // 4), ("norm2", dim // 4), ("norm3", dim)]
                    extra = "norm4"
                else:
                    conv(p + "conv1", dim, cin, 3, 3)
                    conv(p + "conv2", dim, dim, 3, 3)
                    norms = [("norm1", dim), ("norm2", dim)]
                    extra = "norm3"
                if kind == "batch":
                    for nn_, c in norms:
                        bn(p + nn_, c)
                    if st != 1:
                        bn(p + extra, dim)
                if st != 1:
                    conv(p + "downsample.0", dim, cin, 1, 1)
                    if kind == "batch":
                        bn(p + "downsample.1", dim)
                cin = dim
        conv(prefix + "conv2", out_dim, dims[3], 1, 1)

    enc("fnet.", fdim, "instance")
    enc("cnet.", hdim + cdim, cnorm)
    e, g, fh = "update_block.encoder.", "update_block.gru.", "update_block.flow_head."
    if small:
        conv(e + "convc1", 96, cor_planes, 1, 1)
        conv(e + "convf1", 64, 2, 7, 7)
        conv(e + "convf2", 32, 64, 3, 3)
        conv(e + "conv", 80, 128, 3, 3)
        gin = hdim + 82 + 64
        for nm in ("convz", "convr", "convq"):
            conv(g + nm, hdim, gin, 3, 3)
        conv(fh + "conv1", 128, hdim, 3, 3)
        conv(fh + "conv2", 2, 128, 3, 3)
    else:
        conv(e + "convc1", 256, cor_planes, 1, 1)
        conv(e + "convc2", 192, 256, 3, 3)
        conv(e + "convf1", 128, 2, 7, 7)
        conv(e + "convf2", 64, 128, 3, 3)
        conv(e + "conv", 126, 256, 3, 3)
        gin = hdim + 128 + hdim + (128 if variant == "gma" else 0)
        for sfx, (kh, kw) in (("1", (1, 5)), ("2", (5, 1))):
            for nm in ("convz", "convr", "convq"):
                conv(g + nm + sfx, hdim, gin, kh, kw)
        conv(fh + "conv1", 256, hdim, 3, 3)
        conv(fh + "conv2", 2, 256, 3, 3)
        conv("update_block.mask.0", 256, 128, 3, 3)
        conv("update_block.mask.2", 576, 256, 1, 1)
    if variant == "gma":
        s["update_block.aggregator.gamma"] = (1,)
        s["update_block.aggregator.to_v.weight"] = (128, 128, 1, 1)
        s["att.to_qk.weight"] = (256, 128, 1, 1)
        s["att.pos_emb.rel_ind"] = (160, 160)
        s["att.pos_emb.rel_height.weight"] = (319, 128)
        s["att.pos_emb.rel_width.weight"] = (319, 128)
    return s
