"""B200-native pretrain step for the backbone: forward + backward over the C-ABI kernels, gradient all-reduce over
NCCL (bucketed, overlapped with the remaining backward on a side stream), global-norm clipping and fused AdamW on flat
buffers, optionally captured into one CUDA graph.

Mirrors the per-iteration body of Multi-Task_Pretrain/main_pretrain.py:701-832 for the encoder:
``encoder(cat(x1,x2,x3))`` -> losses -> ``backward`` (DDP all-reduce) -> ``clip_grad_norm_(5)`` -> ``AdamW.step`` ->
``CosineAnnealingLR.step``.  The three decoders are third-party (mmseg/mmdet/mmrotate, absent here) and out of scope;
``heads`` is a callable taking the four NCHW maps and returning (loss, cotangents) — the default is the synthetic
objective the parity tests use.
"""
from __future__ import annotations

import os

import math
from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist

from . import _lib as L
from . import engine, engine_bwd, ops
from .flat import FlatLayout, is_gemm_weight as _is_big

BF16, F32 = torch.bfloat16, torch.float32


def layer_decay_group(name: str, shape, num_layers: int, prefix: str = "encoder."):
    """(layer_id, no_decay) as LayerDecayOptimizerConstructor_ViT assigns them
    (mmcv_custom/layer_decay_optimizer_constructor_vit.py:7-16,41-49).  In the pretrain script parameter names start with
    ``encoder.`` while the constructor looks for ``backbone.``, so every parameter lands in the last layer (scale 1);
    ``prefix='backbone.'`` gives the intended layer-wise decay."""
    full = prefix + name
    no_decay = len(shape) == 1 or name.endswith(".bias") or "pos_embed" in name
    if full in ("backbone.cls_token", "backbone.mask_token", "backbone.pos_embed") or full.startswith("backbone.patch_embed"):
        lid = 0
    elif full.startswith("backbone.blocks"):
        lid = int(full.split(".")[2]) + 1
    else:
        lid = num_layers - 1
    return lid, no_decay


def synthetic_heads(feats: Sequence[torch.Tensor], grads: Optional[Sequence[torch.Tensor]] = None, loss: Optional[torch.Tensor] = None,
                    weight: float = 1.0):
    """Stand-in objective: weight * 0.5 * sum_k mean(f_k^2); returns (loss, d loss / d f_k).  ``grads`` / ``loss``: write the
    cotangents into the given tensors (views of a larger batch) and add into an existing loss scalar."""
    fused = all(f.is_cuda and f.dtype == torch.bfloat16 and f.is_contiguous() and f.numel() % 8 == 0 and f.data_ptr() % 16 == 0 for f in feats)
    if grads is not None:
        fused = fused and all(g.is_contiguous() and g.dtype == torch.bfloat16 and g.data_ptr() % 16 == 0 for g in grads)
    if fused:
        if loss is None:
            loss = torch.zeros((), device=feats[0].device, dtype=F32)       # one fused pass per map (mtp_sqloss_fwd_bwd)
        if grads is None:
            grads = [torch.empty_like(f) for f in feats]
        for f, g in zip(feats, grads):
            L.call("mtp_sqloss_fwd_bwd_w", f.data_ptr(), g.data_ptr(), loss.data_ptr(), f.numel(), float(weight), ops._stream())
        return loss, grads
    out = []
    for k, f in enumerate(feats):
        ff = f.float()
        l = (ff * ff).mean() * (0.5 * weight)
        loss = l if loss is None else loss + l
        g = (ff * (weight / f.numel())).to(f.dtype)
        if grads is not None:
            grads[k].copy_(g)
            g = grads[k]
        out.append(g)
    return loss, out


class ThreeTaskHeads:
    """The fan-out of MTP.forward (Multi-Task_Pretrain/models.py:327-335): ONE encoder call on ``cat(x1, x2, x3)``; the four feature
    maps are split on-device ``[:b1] | [b1:b1+b2] | [b1+b2:]`` and handed to the semantic-segmentation, rotated-detection and
    instance-segmentation decoders, whose cotangents land in the matching batch slices of one gradient tensor per map.
    ``heads[k](feats_k, grads_k, loss)`` must write d loss_k / d feats_k into ``grads_k`` (views) and add loss_k into ``loss``.
    The decoders themselves are third-party (mmseg / mmrotate / mmdet, out of scope); the default stand-ins consume the same 4-map
    contract with distinct weights so that the three slices carry different cotangents."""

    def __init__(self, split: Sequence[int], heads: Optional[Sequence[Callable]] = None):
        assert len(split) == 3 and all(b > 0 for b in split)
        self.split = tuple(int(b) for b in split)
        if heads is None:
            heads = [lambda f, g, l, w=w: synthetic_heads(f, g, l, weight=w) for w in (1.0, 0.5, 2.0)]
        assert len(heads) == 3
        self.heads = list(heads)

    def __call__(self, feats: Sequence[torch.Tensor]):
        assert feats[0].shape[0] == sum(self.split), f"batch {feats[0].shape[0]} != split {self.split}"
        loss = torch.zeros((), device=feats[0].device, dtype=F32)
        grads = [torch.empty_like(f) for f in feats]
        lo = 0
        for b, head in zip(self.split, self.heads):
            fk = [f[lo:lo + b] for f in feats]          # batch slices of contiguous NCHW maps are contiguous views: no copies
            gk = [g[lo:lo + b] for g in grads]
            head(fk, gk, loss)
            lo += b
        return loss, grads


class PretrainStep:
    def __init__(self, model, lr=6e-5, weight_decay=0.05, betas=(0.9, 0.999), eps=1e-8, max_norm=5.0, t_max=0, eta_min=0.0,
                 layer_decay_rate=0.9, name_prefix="encoder.", heads: Optional[Callable] = None, process_group=None,
                 bucket_blocks=4, use_cuda_graph=False, comm_sms=16, grad_comm="bf16"):
        self.model = model
        self.heads = heads or synthetic_heads
        self.lr, self.eta_min, self.t_max = lr, eta_min, t_max
        self.betas, self.eps, self.max_norm = betas, eps, max_norm
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        self.bucket_blocks = bucket_blocks
        # Gradient all-reduce dtype of the GEMM-weight region (world > 1).  "bf16" (0.635 GB instead of the reference DDP's
        # 1.27 GB fp32): every bucket is cast once (fp32 wgrad output -> bf16), summed over the ranks in bf16 by NCCL and read by AdamW
        # as bf16; the small accumulated region (biases, LayerNorm, tables, sampling heads, pos_embed) always travels as fp32.
        # "fp32" reproduces the reference's DDP arithmetic.
        assert grad_comm in ("bf16", "fp32")
        self.grad_comm = grad_comm if self.world > 1 else "fp32"
        # SMs left to the NCCL all-reduce kernels while the backward runs beside them (set NCCL_MAX_CTAS to the same number):
        # the persistent GEMM grids of the backward are sized for the remaining SMs instead of spilling into a second wave
        self.comm_sms = comm_sms if self.world > 1 else 0
        dev = next(model.parameters()).device
        assert dev.type == "cuda", "PretrainStep needs the model on a CUDA device"
        self.dev = dev
        # ---- flat layout: [small parameters | GEMM weights], every tensor 64-element aligned
        named = list(model.named_parameters())
        self.layout = FlatLayout([(n, tuple(p.shape)) for n, p in named])
        pmap = dict(named)
        order = [(n, pmap[n]) for n in self.layout.order]
        self.offsets, self.small_end, total = self.layout.offsets, self.layout.small_end, self.layout.total
        self.total = total
        self.flat_p = torch.zeros(total, device=dev, dtype=F32)
        self.flat_g = torch.zeros(total, device=dev, dtype=F32)
        self.flat_m = torch.zeros(total, device=dev, dtype=F32)
        self.flat_v = torch.zeros(total, device=dev, dtype=F32)
        self.flat_g16 = torch.zeros(total, device=dev, dtype=BF16) if self.grad_comm == "bf16" else None
        num_layers = len(model.blocks) + 2
        groups, chunk_group = {}, torch.zeros(total // 64, dtype=torch.uint8)
        self._convt = {"fpn1.0.weight": "fpn1_0", "fpn1.3.weight": "fpn1_3", "fpn2.0.weight": "fpn2_0"}
        for n, p in order:
            o = self.offsets[n]
            if n in self._convt:
                # ConvTranspose2d(k2,s2) weights live in the flat buffers in GEMM layout [4*Cout, Cin] (row = (dy*2+dx)*Cout + co): the
                # bf16 mirror slice IS the GEMM operand and the wgrad GEMM writes the flat gradient in place.  The nn.Parameter keeps
                # its (Cin, Cout, 2, 2) shape as a permuted (non-contiguous) view, so state_dict / load_state_dict are unchanged.
                cin, cout = p.shape[0], p.shape[1]
                self.flat_p[o:o + p.numel()].copy_(p.data.permute(2, 3, 1, 0).reshape(-1))
                p.data = self.flat_p[o:o + p.numel()].view(2, 2, cout, cin).permute(3, 2, 0, 1)
            else:
                self.flat_p[o:o + p.numel()].copy_(p.data.reshape(-1))
                p.data = self.flat_p[o:o + p.numel()].view(p.shape)
            lid, nodecay = layer_decay_group(n, p.shape, num_layers, name_prefix)
            key = (layer_decay_rate ** (num_layers - lid - 1), 0.0 if nodecay else weight_decay)
            gid = groups.setdefault(key, len(groups))
            chunk_group[o // 64:(o + (p.numel() + 63) // 64 * 64) // 64] = gid
        assert len(groups) < 256
        self.chunk_group = chunk_group.to(dev)
        self.group_lr = torch.tensor([k[0] for k in groups], device=dev, dtype=F32)
        self.group_wd = torch.tensor([k[1] for k in groups], device=dev, dtype=F32)
        self.state = torch.zeros(2, device=dev, dtype=F32)
        # ---- bf16 mirror kept current by the optimizer kernel; GEMM weights read it directly
        self.flat_p16 = ops.cast_f32_bf16(self.flat_p)
        st = model._engine_state
        C = model.embed_dim
        for n, p in order:
            if not _is_big(n):
                continue
            o = self.offsets[n]
            v16 = self.flat_p16[o:o + p.numel()]
            if n == "patch_embed.proj.weight":
                st.pinned["pe_w"] = v16.view(C, -1)
            elif n in self._convt:
                st.pinned[self._convt[n]] = v16.view(4 * p.shape[1], p.shape[0])
            elif n.startswith("blocks."):
                i = int(n.split(".")[1])
                nm = n.split(".")[-2]
                st.pinned[f"b{i}.{nm}"] = v16.view(p.shape)
        st.listeners.append(self.refresh_mirror)       # checkpoint loading / load_state_dict / manual p.data edits: engine invalidate()
        # ---- gradient store bound to the flat gradient buffer (same layout)
        self.G = engine_bwd.GradStore.__new__(engine_bwd.GradStore)
        self.G.names = [n for n, _ in named]
        self.G.flat = self.flat_g
        self.G.views, self.G.packed = {}, {}
        for n, p in named:
            gslice = self.flat_g[self.offsets[n]:self.offsets[n] + p.numel()]
            if n in self._convt:
                self.G.packed[n] = gslice.view(4 * p.shape[1], p.shape[0])
                self.G.views[n] = gslice.view(2, 2, p.shape[1], p.shape[0]).permute(3, 2, 0, 1)
            else:
                self.G.views[n] = gslice.view(p.shape)
        self.G.touched = set()
        # world size 1: the wgrad GEMM epilogues accumulate the squared norm of the weight gradients they store, so the clip
        # needs a reduction pass over the small region only (verified once against the full pass, see _forward_backward)
        self.fused_norm = self.world == 1 and bool(self.max_norm and self.max_norm > 0) and os.environ.get("MTP_FUSED_NORM", "1") != "0"
        self._norm_checked = False
        self.G.sumsq = self.state[1:2] if self.fused_norm else None
        # ---- all-reduce buckets over the big region (in backward order) + one bucket for the small region
        self.comm_stream = torch.cuda.Stream(device=dev) if self.world > 1 else None
        self.graph = None
        self.use_cuda_graph = use_cuda_graph
        self._static_x = None
        self._static_loss = None

    # ------------------------------------------------------------------------------------------------------------
    def _bucket_ranges(self):
        return self.layout.block_buckets(len(self.model.blocks), self.bucket_blocks)

    def _stage_range(self, lo, hi):
        """On the compute stream (inside the captured graph piece): the bf16 copy of a final GEMM-weight gradient range."""
        if self.world > 1 and self.grad_comm == "bf16" and lo >= self.small_end:
            L.call("mtp_cast_f32_bf16", self.flat_g.data_ptr() + 4 * lo, self.flat_g16.data_ptr() + 2 * lo, hi - lo, ops._stream())

    def _allreduce_range(self, lo, hi):
        if self.world == 1:
            return
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self.comm_stream.wait_event(ev)
        with torch.cuda.stream(self.comm_stream):
            if self.grad_comm == "bf16" and lo >= self.small_end:
                dist.all_reduce(self.flat_g16[lo:hi], op=dist.ReduceOp.SUM, group=self.pg)
            else:
                dist.all_reduce(self.flat_g[lo:hi], op=dist.ReduceOp.SUM, group=self.pg)

    def _forward_backward(self, x, on_bucket):
        """forward + heads + backward; ``on_bucket(ranges)`` is called when flat-gradient ranges have become final."""
        m = self.model
        keep = engine._draw_keep(m, x.shape[0], x.device)
        outs, ctx = engine._forward_impl(m, x, keep, save=True)
        loss, douts = self.heads(outs)
        # gradients: zero only the small (accumulated) region; GEMM weight gradients are overwritten
        self.flat_g[:self.small_end].zero_()
        self.G.touched = set()
        buckets = self._bucket_ranges() if self.world > 1 else {}
        if self.fused_norm:                          # ++step, squared norm = 0 BEFORE the wgrad epilogues add to it
            L.call("mtp_optim_step_begin", self.state.data_ptr(), ops._stream())
        if self.comm_sms:
            L.call("mtp_set_sm_limit", max(8, L.load().mtp_num_sms() - self.comm_sms))
        fpn_ranges, rest = self.layout.split_tail(len(m.blocks), self.bucket_blocks) if self.world > 1 else ([], [])

        def emit(ranges):
            for lo, hi in ranges:
                self._stage_range(lo, hi)
            on_bucket(ranges)
        try:
            engine_bwd.backward_impl(m, x, ctx, douts, grad_store=self.G,
                                     after_block=(lambda i: emit([buckets[i]]) if i in buckets else None),
                                     after_fpn=((lambda: emit(fpn_ranges)) if (self.world > 1 and fpn_ranges) else None))
        finally:
            if self.comm_sms:
                L.call("mtp_set_sm_limit", 0)
        if self.world > 1:
            # remaining pieces: the small (accumulated) region and the patch-embed weight, final only now
            emit(rest)
        if self.fused_norm and not self._norm_checked and not torch.cuda.is_current_stream_capturing():
            self._norm_checked = True
            fused = float(self.state[1].item()) + float((self.flat_g[:self.small_end].double() ** 2).sum().item())
            full = float((self.flat_g.double() ** 2).sum().item())
            if abs(fused - full) > 1e-3 * max(full, 1e-30):      # some weight gradient did not come out of a wgrad epilogue
                self.fused_norm, self.G.sumsq = False, None
                self.state[0] -= 1.0                 # the optimizer's own step_begin counts this step and recomputes the norm
        return loss

    def _optimizer(self):
        stream = ops._stream()
        if self.fused_norm:
            L.call("mtp_sumsq_f32", self.flat_g.data_ptr(), self.small_end, self.state.data_ptr() + 4, stream)
        else:
            L.call("mtp_optim_step_begin", self.state.data_ptr(), stream)
            if self.max_norm and self.max_norm > 0:
                if self.grad_comm == "bf16":
                    L.call("mtp_sumsq_f32", self.flat_g.data_ptr(), self.small_end, self.state.data_ptr() + 4, stream)
                    L.call("mtp_sumsq_bf16", self.flat_g16.data_ptr() + 2 * self.small_end, self.total - self.small_end, self.state.data_ptr() + 4, stream)
                else:
                    L.call("mtp_sumsq_f32", self.flat_g.data_ptr(), self.total, self.state.data_ptr() + 4, stream)
        g16 = self.flat_g16.data_ptr() if self.grad_comm == "bf16" else 0
        L.call("mtp_adamw_step_mixed", self.flat_p.data_ptr(), self.flat_g.data_ptr(), g16, self.small_end, self.flat_m.data_ptr(),
               self.flat_v.data_ptr(), self.flat_p16.data_ptr(), self.chunk_group.data_ptr(), self.group_lr.data_ptr(), self.group_wd.data_ptr(),
               self.state.data_ptr(), self.total, float(self.lr), float(self.eta_min), int(self.t_max), float(self.betas[0]),
               float(self.betas[1]), float(self.eps), float(self.max_norm or 0.0), 1.0 / self.world, stream)

    # ---- training-state checkpointing (the reference saves optimizer + scheduler state: main_pretrain.py:825-832) --------------
    def refresh_mirror(self):
        """Re-derive the bf16 weight mirror from the fp32 masters (after weights were loaded / edited outside the optimizer) and
        invalidate captured graphs' assumptions about nothing (they read the same buffers)."""
        L.call("mtp_cast_f32_bf16", self.flat_p.data_ptr(), self.flat_p16.data_ptr(), self.total, ops._stream())
        self.model._engine_state.cache.clear()

    def state_dict(self):
        """Optimizer + schedule state keyed by parameter name (layout independent): exp_avg / exp_avg_sq per parameter in the
        parameter's own shape, the step counter that drives the bias correction and the cosine schedule."""
        out = {"step": float(self.state[0].item()), "exp_avg": {}, "exp_avg_sq": {}}
        for n, p in self.model.named_parameters():
            o = self.offsets[n]
            for key, flat in (("exp_avg", self.flat_m), ("exp_avg_sq", self.flat_v)):
                v = flat[o:o + p.numel()]
                if n in self._convt:
                    v = v.view(2, 2, p.shape[1], p.shape[0]).permute(3, 2, 0, 1)
                else:
                    v = v.view(p.shape)
                out[key][n] = v.detach().clone().contiguous()
        return out

    def load_state_dict(self, sd):
        with torch.no_grad():
            for n, p in self.model.named_parameters():
                o = self.offsets[n]
                for key, flat in (("exp_avg", self.flat_m), ("exp_avg_sq", self.flat_v)):
                    src = sd[key][n].to(self.dev, F32)
                    if n in self._convt:
                        src = src.permute(2, 3, 1, 0)
                    flat[o:o + p.numel()].copy_(src.reshape(-1))
            self.state[0] = float(sd["step"])
        self.refresh_mirror()

    def _reduce(self, ranges):
        for lo, hi in ranges:
            self._allreduce_range(lo, hi)

    def _step_body(self, x):
        loss = self._forward_backward(x, self._reduce)
        if self.world > 1:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        self._optimizer()
        return loss

    # ---- CUDA-graph execution -------------------------------------------------------------------------------------
    # world == 1: the whole step is ONE graph.  world > 1: NCCL stays out of the graphs — the step is cut at the bucket
    # boundaries into a chain of graphs sharing one memory pool; after each piece the bucket's all-reduce is launched
    # eagerly on the comm stream, so it overlaps with the replay of the following pieces.
    def _capture(self, x):
        self._static_x = x.clone()
        snap = [t.clone() for t in (self.flat_p, self.flat_m, self.flat_v, self.flat_p16, self.state)]
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):                      # warm-up (kernel attributes, allocator, NCCL communicators)
                self._step_body(self._static_x)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        if self.world == 1:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._static_loss = self._step_body(self._static_x)
            self.graphs, self.reduce_after = [g], [[]]
        else:
            pool = torch.cuda.graph_pool_handle()
            self.graphs, self.reduce_after = [], []
            cap = torch.cuda.Stream(device=self.dev)
            cap.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(cap):
                cur = [torch.cuda.CUDAGraph()]
                cur[0].capture_begin(pool=pool)

                def cut(ranges):                    # close the current piece, remember what to reduce after it, open the next
                    cur[0].capture_end()
                    self.graphs.append(cur[0])
                    self.reduce_after.append(list(ranges))
                    cur[0] = torch.cuda.CUDAGraph()
                    cur[0].capture_begin(pool=pool)
                self._static_loss = self._forward_backward(self._static_x, cut)
                self._optimizer()                   # last piece: clip + AdamW (replayed after the comm stream has been joined)
                cur[0].capture_end()
                self.graphs.append(cur[0])
                self.reduce_after.append([])
            torch.cuda.current_stream().wait_stream(cap)
        for dst, src in zip((self.flat_p, self.flat_m, self.flat_v, self.flat_p16, self.state), snap):
            dst.copy_(src)                          # warm-up must not advance the training state
        del snap
        torch.cuda.synchronize()

    def step(self, x: torch.Tensor) -> torch.Tensor:
        """One training step on a device-resident batch; returns the (device) scalar loss."""
        if not self.use_cuda_graph:
            return self._step_body(x)
        if self.graph is None:
            self._capture(x)
            self.graph = True
        if x.data_ptr() != self._static_x.data_ptr():
            self._static_x.copy_(x, non_blocking=True)
        last = len(self.graphs) - 1
        for k, g in enumerate(self.graphs):
            if k == last and self.world > 1:
                torch.cuda.current_stream().wait_stream(self.comm_stream)
            g.replay()
            self._reduce(self.reduce_after[k])
        return self._static_loss

    def step_from_host(self, x_host_pinned) -> float:
        """End-to-end step: host (pinned) batch -> device, step, loss back to the host.  A tuple / list of three pinned batches
        (the three task streams of models.py:327-329) is concatenated on the way: each stream is copied straight into its batch slice
        of the step's input buffer (``x = torch.cat((x1, x2, x3), 0)`` without the extra device copy)."""
        if isinstance(x_host_pinned, (tuple, list)):
            parts = list(x_host_pinned)
            n = sum(p.shape[0] for p in parts)
            buf = self._static_x if (self.use_cuda_graph and self._static_x is not None and self._static_x.shape[0] == n
                                     and self._static_x.dtype == parts[0].dtype) else \
                torch.empty((n,) + tuple(parts[0].shape[1:]), device=self.dev, dtype=parts[0].dtype)
            lo = 0
            for p in parts:
                buf[lo:lo + p.shape[0]].copy_(p, non_blocking=True)
                lo += p.shape[0]
            return float(self.step(buf).item())
        x = x_host_pinned.to(self.dev, non_blocking=True)
        return float(self.step(x).item())
