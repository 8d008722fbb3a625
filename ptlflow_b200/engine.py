"""RaftEngine: packed weights + workspaces + one C call for the whole refinement loop.

This is the host-side glue between the kept ``RAFT.forward`` structure
(ptlflow/models/raft/raft.py:125-194) and ``pfb_raft_refine``.  It owns nothing numerical:
packing, buffers, pointer structs, CUDA-graph capture.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib, ops
from ._lib import check, dtype_code, load, ptr_array, stream_ptr


_TILED = bool(int(os.environ.get("PFB_VOLUME_TILED", "1")))


class RaftEngine:
    """Built lazily at first forward and rebuilt when the parameters' dtype / device / storage
    change (callers do ``model.eval().cuda().half()`` *after* construction and after
    ``load_state_dict`` -- model_benchmark.py:274-279, infer.py:148-152)."""

    def __init__(self, update_block: torch.nn.Module, variant: int, hidden_dim: int, context_dim: int,
                 corr_levels: int, corr_radius: int, dtype: torch.dtype, device: torch.device, impl: int = 0,
                 attention_module: Optional[torch.nn.Module] = None):
        self.variant, self.hidden_dim, self.context_dim = variant, hidden_dim, context_dim
        self.corr_levels, self.corr_radius = corr_levels, corr_radius
        self.dtype, self.device, self.impl = dtype, device, impl
        ub = update_block
        enc, gru, fh = ub.encoder, ub.gru, ub.flow_head
        planes = corr_levels * (2 * corr_radius + 1) ** 2
        hd, cd = hidden_dim, context_dim

        def P(srcs, *convs):  # srcs: channel counts of the concatenated inputs, in order (tcgen05 K-major pack)
            return ops.PackedConv(convs, dtype, device, src_channels=srcs)

        layers: Dict[int, ops.PackedConv] = {}
        layers[_lib.L_CONVF1] = P(None, enc.convf1)  # 7x7 on the 2-channel fp32 flow: dedicated kernel
        if variant in (0, 2) and dtype != torch.float32 and tuple(enc.convf1.weight.shape) == (128, 2, 7, 7):
            # tensor-core form (csrc/first_conv.cu): the K-major slot of the layer carries the overlapping-window tiles
            pc = layers[_lib.L_CONVF1]
            pc.weight_k = ops.pack_flow_conv(enc.convf1.weight, dtype).to(device)
            pc.Cin_pad, pc.Cout_pad_k = 64, 128
        if variant in (0, 2):
            layers[_lib.L_CONVC1] = P([planes], enc.convc1)
            layers[_lib.L_CONVC2] = P([256], enc.convc2)
            layers[_lib.L_CONVF2] = P([128], enc.convf2)
            layers[_lib.L_CONV] = P([256], enc.conv)  # cat[cor(192), flo(64)] lives in one 256-channel buffer
            # cat[h or r*h, inp, motion (, motion_global)] -- three tensor maps, no concat copy (gma keeps
            # motion | motion_global in one 256-channel buffer)
            gsrc = [hd, cd, 256 if variant == 2 else 128]
            layers[_lib.L_GRU_ZR1] = P(gsrc, gru.convz1, gru.convr1)  # z | r share the input: one GEMM, N = 2*hidden
            layers[_lib.L_GRU_Q1] = P(gsrc, gru.convq1)
            layers[_lib.L_GRU_ZR2] = P(gsrc, gru.convz2, gru.convr2)
            layers[_lib.L_GRU_Q2] = P(gsrc, gru.convq2)
            if dtype != torch.float32:
                # round 2: (a) the context columns of the four GRU convolutions become layers of their own, evaluated once per
                # forward with the bias folded in, and the per-iteration layers lose them (-1/3 of the GRU's K); (b) convc2 |
                # convf2 as one block-diagonal N = 256 layer.  Pure repacking of the same parameters.
                class _View:
                    def __init__(self, weight, bias):
                        self.weight, self.bias = weight, bias

                def cols(conv, lo, hi, with_bias):
                    return _View(conv.weight.detach()[:, lo:hi].contiguous(), conv.bias if with_bias else None)

                def rest(conv):  # [h | inp | motion...] without the inp columns
                    w = conv.weight.detach()
                    return _View(torch.cat([w[:, :hd], w[:, hd + cd:]], dim=1).contiguous(), None)

                xsrc = [hd, gsrc[2]]
                for ctx_id, x_id, convs in ((_lib.L_CTX_ZR1, _lib.L_GRUX_ZR1, (gru.convz1, gru.convr1)), (_lib.L_CTX_Q1, _lib.L_GRUX_Q1, (gru.convq1,)),
                                            (_lib.L_CTX_ZR2, _lib.L_GRUX_ZR2, (gru.convz2, gru.convr2)), (_lib.L_CTX_Q2, _lib.L_GRUX_Q2, (gru.convq2,))):
                    layers[ctx_id] = P([cd], *[cols(cv, hd, hd + cd, True) for cv in convs])
                    layers[x_id] = P(xsrc, *[rest(cv) for cv in convs])
                wc, wf = enc.convc2.weight.detach(), enc.convf2.weight.detach()
                if tuple(wc.shape[2:]) == tuple(wf.shape[2:]) and wc.shape[0] + wf.shape[0] == 256:
                    bd = torch.zeros((256, wc.shape[1] + wf.shape[1]) + tuple(wc.shape[2:]), dtype=wc.dtype, device=wc.device)
                    bd[: wc.shape[0], : wc.shape[1]] = wc
                    bd[wc.shape[0]:, wc.shape[1]:] = wf
                    layers[_lib.L_CONVC2F2] = P([wc.shape[1], wf.shape[1]], _View(bd, torch.cat([enc.convc2.bias.detach(), enc.convf2.bias.detach()])))
            layers[_lib.L_FLOW1] = P([hd], fh.conv1)
            layers[_lib.L_FLOW2] = P([256], fh.conv2)
            if dtype != torch.float32:
                # conv2 as a 1x1 layer producing the 9 x 2 per-tap products (row = tap*2 + o); bias is added by the gather
                class _Taps:
                    def __init__(self, w):
                        self.weight, self.bias = w.detach().permute(2, 3, 0, 1).reshape(18, w.shape[1], 1, 1).contiguous(), None

                layers[_lib.L_FLOW2T] = P([256], _Taps(fh.conv2.weight))
            layers[_lib.L_MASK1] = P([hd], ub.mask[0])
            layers[_lib.L_MASK2] = P([256], ub.mask[2])
        else:  # raft_small: odd channel counts (96 / 82 / 146) -> SIMT kernels only
            layers[_lib.L_CONVC1] = P(None, enc.convc1)
            layers[_lib.L_CONVF2] = P(None, enc.convf2)
            layers[_lib.L_CONV] = P(None, enc.conv)
            layers[_lib.L_GRU_ZR1] = P(None, gru.convz, gru.convr)
            layers[_lib.L_GRU_Q1] = P(None, gru.convq)
            layers[_lib.L_FLOW1] = P(None, fh.conv1)
            layers[_lib.L_FLOW2] = P(None, fh.conv2)
        self.agg_gamma = 0.0
        self.att_q = self.att_k = None
        if variant == 2:
            agg = ub.aggregator
            layers[_lib.L_AGG_V] = P([128], agg.to_v)
            self.agg_gamma = float(agg.gamma.detach().float().cpu().item())

            class _Half:  # q / k halves of Attention.to_qk as separate 1x1 layers (contiguous outputs for the GEMM)
                def __init__(self, w):
                    self.weight, self.bias = w, None

            wqk = attention_module.to_qk.weight
            c = wqk.shape[0] // 2
            self.att_q = ops.PackedConv([_Half(wqk[:c])], dtype, device, src_channels=[wqk.shape[1]])
            self.att_k = ops.PackedConv([_Half(wqk[c:])], dtype, device, src_channels=[wqk.shape[1]])
        self.layers = layers
        self.weights = _lib.RaftWeights()
        for k, v in layers.items():
            self.weights.layers[k] = v.layer_struct()
        self._workspaces: Dict[Tuple, torch.Tensor] = {}
        self.signature = self.param_signature(update_block)
        # the pack kernels ran on the constructing thread's stream; other streams / host threads (pipeline slots) may use
        # the packed weights as soon as the engine is published, so finish them first (one-time cost)
        torch.cuda.current_stream(device).synchronize()

    # -- cache invalidation --------------------------------------------------------------------
    @staticmethod
    def param_signature(update_block: torch.nn.Module):
        return tuple((p.data_ptr(), p._version, p.dtype, str(p.device)) for p in update_block.parameters())

    # -- run --------------------------------------------------------------------------------
    def make_cfg(self, B: int, H: int, W: int, iters: int, out_hw, pad, alternate_corr: bool, feat_dim: int, volume_layout: int = 0) -> _lib.RaftCfg:
        return _lib.RaftCfg(self.variant, dtype_code(self.dtype), B, H, W, feat_dim, self.corr_levels, self.corr_radius,
                            self.hidden_dim, self.context_dim, iters, int(alternate_corr), out_hw[0], out_hw[1], pad[0], pad[1],
                            self.impl, 0 if alternate_corr else int(volume_layout), int(getattr(self, "fork_flow", False)))

    def build_volume(self, fmap1: torch.Tensor, fmap2: torch.Tensor, impl: int = 0):
        """a1 + a2 for the refinement loop of this engine.  f16 / bf16 with tensor-core-shaped features get the tiled
        pyramid (64-byte tiles, csrc/corr_tiled.cu); everything else the dense one.  refine() reads the layout back from
        ``self.volume_layout`` (set here, per call)."""
        self.volume_layout = 0
        if _TILED and impl != 1 and self.corr_radius in (3, 4) and ops.tiled_supported(fmap1, self.corr_levels) \
                and (self.corr_radius, self.corr_levels) in ((4, 4), (4, 3), (4, 2), (4, 1), (3, 4), (3, 3)):
            self.volume_layout = 1
            return ops.corr_volume_build_tiled(fmap1, fmap2, self.corr_levels)
        return ops.corr_volume_build(fmap1, fmap2, self.corr_levels, impl=impl)

    def workspace(self, cfg: _lib.RaftCfg, scratch: Optional[dict] = None) -> torch.Tensor:
        if scratch is not None:
            # the caller owns the scratch memory (a CUDA graph keeps the workspace it was captured with alive)
            key = ("raft_ws", cfg.B, cfg.H, cfg.W)
            ws = scratch.get(key)
            if ws is None:
                nbytes = load().pfb_raft_workspace_bytes(C.byref(cfg))
                if nbytes == 0:
                    check(-1, "raft_workspace_bytes")
                ws = scratch[key] = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            return ws
        # one workspace per CUDA stream (batches in flight on different streams must not share scratch memory,
        # ptlflow_b200/pipeline.py) and, per stream, one shape resident (bounded memory)
        sid = torch.cuda.current_stream(self.device).cuda_stream
        key = (cfg.B, cfg.H, cfg.W)
        ent = self._workspaces.get(sid)
        if ent is None or ent[0] != key:
            nbytes = load().pfb_raft_workspace_bytes(C.byref(cfg))
            if nbytes == 0:
                check(-1, "raft_workspace_bytes")
            ent = (key, torch.empty(nbytes, dtype=torch.uint8, device=self.device))
            if len(self._workspaces) >= 8:
                # streams come and go: do not grow without bound.  A dropped tensor returns to torch's caching allocator,
                # which keeps the block reserved for the stream it was allocated on until that stream's queued work is
                # done -- an evicted workspace with work still in flight is therefore not reused under that work.
                self._workspaces.pop(next(iter(self._workspaces)))
            self._workspaces[sid] = ent
        return ent[1]

    def refine(self, pyramid: Sequence[torch.Tensor], net: torch.Tensor, inp: torch.Tensor, coords: torch.Tensor,
               iters: int, out_hw, pad, fmap1: Optional[torch.Tensor] = None, attention: Optional[torch.Tensor] = None,
               scratch: Optional[dict] = None):
        """Runs the loop in place on (net, coords); returns (flow_up fp32 [B,2,oh,ow], flow_small fp32 [B,2,H,W])."""
        B, H, W, _ = net.shape
        alt = fmap1 is not None
        cfg = self.make_cfg(B, H, W, iters, out_hw, pad, alt, fmap1.shape[-1] if alt else 0, getattr(self, "volume_layout", 0))
        ws = self.workspace(cfg, scratch)
        flow_up = torch.empty((B, 2, out_hw[0], out_hw[1]), dtype=torch.float32, device=self.device)
        flow_small = torch.empty((B, 2, H, W), dtype=torch.float32, device=self.device)
        pyr = ptr_array(pyramid)
        buf = _lib.RaftBuffers(C.cast(pyr, C.POINTER(C.c_void_p)), fmap1.data_ptr() if alt else None, net.data_ptr(),
                               inp.data_ptr(), coords.data_ptr(), flow_up.data_ptr(), flow_small.data_ptr(),
                               ws.data_ptr(), ws.numel(), attention.data_ptr() if attention is not None else None, self.agg_gamma)
        with torch.cuda.device(self.device):
            check(load().pfb_raft_refine(C.byref(cfg), C.byref(self.weights), C.byref(buf), stream_ptr(self.device)), "raft_refine")
        return flow_up, flow_small

    def update_iter(self, net: torch.Tensor, inp: torch.Tensor, coords: torch.Tensor, corr: Optional[torch.Tensor] = None,
                    pyramid: Optional[Sequence[torch.Tensor]] = None, want_mask: bool = False):
        """One update-block evaluation (operator-level tests).  corr: pixel-major [B,H,W,planes]."""
        B, H, W, _ = net.shape
        cfg = self.make_cfg(B, H, W, 1, (8 * H, 8 * W), (0, 0), False, 0)
        ws = self.workspace(cfg)
        mask = torch.empty((B, H, W, 576), dtype=self.dtype, device=self.device) if (want_mask and self.variant == 0) else None
        pyr = ptr_array(pyramid) if pyramid is not None else None
        buf = _lib.RaftBuffers(C.cast(pyr, C.POINTER(C.c_void_p)) if pyr is not None else None, None, net.data_ptr(), inp.data_ptr(),
                               coords.data_ptr(), None, None, ws.data_ptr(), ws.numel(), None, 0.0)
        with torch.cuda.device(self.device):
            check(load().pfb_raft_update_iter(C.byref(cfg), C.byref(self.weights), C.byref(buf),
                                              corr.data_ptr() if corr is not None else None,
                                              mask.data_ptr() if mask is not None else None, stream_ptr(self.device)), "raft_update_iter")
        return mask
