"""Backward pass of the backbone over the C-ABI kernels: the autograd of ViT_Win_RVSA_V3_WSZ7.forward_features
([V]:787-813) written out by hand, block by block in reverse.

Per block (cotangent of the fp32 residual stream comes in as ``dx``):
  MLP branch   g = bf16(keep * dx) [+ fc2 bias grad]  ->  fc2 wgrad (MN,MN GEMM)  ->  fc2 dgrad fused with GELU' (K,MN GEMM)
               -> fc1 bias grad, fc1 wgrad, fc1 dgrad -> LayerNorm backward fused with the residual-path add
  attn branch  g = bf16(keep * dx) [+ proj bias grad] ->  proj wgrad / dgrad -> attention backward (RVSA or dense)
               -> qkv bias grad, wgrad, dgrad (+ pooled sampling-head path) -> LayerNorm backward + residual add
Feature-pyramid cotangents are folded into ``dx`` lazily at the block they were tapped from (their last kernel
accumulates straight into the residual-stream gradient).
"""
from __future__ import annotations

import os

import torch

from . import _lib as L
from . import engine, ops

BF16, F32 = torch.bfloat16, torch.float32
_TAIL_FUSE = os.environ.get("MTP_TAIL_FUSE", "1") != "0"      # A/B: the fused follow-up launch of the window blocks' backward
_POOL_FUSE = os.environ.get("MTP_POOL_FUSE", "1") != "0"      # A/B switch: 0 = separate rvsa_pool_bwd_add kernel


class GradStore:
    """fp32 gradient tensors, one per parameter, carved out of a single flat buffer (16-byte aligned views)."""

    def __init__(self, module, device):
        self.names, self.views = [], {}
        offs, total = [], 0
        plist = list(module.named_parameters())
        for name, p in plist:
            offs.append(total)
            total += (p.numel() + 63) // 64 * 64
        self.flat = torch.zeros(total, device=device, dtype=F32)
        for (name, p), o in zip(plist, offs):
            self.names.append(name)
            self.views[name] = self.flat[o:o + p.numel()].view(p.shape)
        self.touched = set()
        self.packed = {}

    packed = {}        # ConvTranspose2d weights: name -> [4*Cout, Cin] fp32 tensor ALIASING the gradient in GEMM layout (trainer)
    sumsq = None       # optional fp32 scalar: the wgrad GEMMs add the sum of squares of the weight gradients they store (trainer)

    def g(self, name):
        self.touched.add(name)
        return self.views[name]


def _linear_bwd(g_bf16, x_bf16, w_bf16, dW, T, n_out, n_in, *, dgrad_mode=L.EPI_BF16, aux=None, colsum=None, sumsq=None):
    """y = x W^T + b with x [T, n_in], W [n_out, n_in], cotangent g [T, n_out].
    dW = g^T x (both operands MN-major, K = T) and dx = g W (B operand MN-major) only share the input g: they run as ONE
    grouped launch whose tiles are spread over the SMs by a common schedule.  ``colsum``: fp32 [n_in] that receives the column
    sums of dx (the bias gradient of the Linear below) from the dgrad epilogue."""
    dx = torch.empty(T, n_in, device=g_bf16.device, dtype=BF16)
    # b_static: the weights and the activation saved by the forward are not written by the kernel just before this launch
    ops.gemm_dual(dict(A=g_bf16, B=w_bf16, M=T, N=n_in, K=n_out, out=dx, b_mn=True, mode=dgrad_mode, aux=aux, lda=n_out, ldb=n_in,
                       colsum=colsum, b_static=True),
                  dict(A=g_bf16, B=x_bf16, M=n_out, N=n_in, K=T, out=dW, a_mn=True, b_mn=True, mode=L.EPI_F32, lda=n_out, ldb=n_in, ldo=n_in,
                       b_static=True, sumsq=sumsq))
    return dx


def _block_backward(i, d, s, dx2, g2, G, B, gh, gw, nH, keep, nxt=None):
    """``g2``: bf16(keep_mlp * dx2) when the kernel that produced dx2 already emitted it (else None).  ``nxt`` = (row_scale,
    colsum) of the block below: its g2 is emitted by this block's last LayerNorm backward.  Returns (dx0, g2 of the block below)."""
    pre = f"blocks.{i}."
    T, C = dx2.shape
    N = gh * gw
    hid = d["fc1_w"].shape[0]
    km = keep[i, 1] if keep is not None else None
    ka = keep[i, 0] if keep is not None else None
    # ---- MLP branch ([V]:509)
    if g2 is None:
        g2 = ops.scale_cast_bf16(dx2, km, N, colsum=G.g(pre + "mlp.fc2.bias"))
    dh = _linear_bwd(g2, s["a"], d["fc2_w"], G.g(pre + "mlp.fc2.weight"), T, C, hid, dgrad_mode=L.EPI_BF16_DGELU, aux=s["hpre"],
                     colsum=G.g(pre + "mlp.fc1.bias"), sumsq=G.sumsq)
    dy2 = _linear_bwd(dh, s["y2"], d["fc1_w"], G.g(pre + "mlp.fc1.weight"), T, hid, C, sumsq=G.sumsq)
    # ---- attention branch ([V]:508): its cotangent bf16(keep_attn * dx1) and the proj bias gradient come out of the LN backward
    dx1, g1 = ops.layernorm_bwd(dy2, s["x1"], s["mean2"], s["rstd2"], d["norm2_w"], None, dx2, G.g(pre + "norm2.weight"),
                                G.g(pre + "norm2.bias"), cast=(ka, N, G.g(pre + "attn.proj.bias")))
    do = _linear_bwd(g1, s["o"], d["proj_w"], G.g(pre + "attn.proj.weight"), T, C, C, sumsq=G.sumsq)
    pool = None
    fused_tail = d["window"] and _POOL_FUSE and _TAIL_FUSE and nH % 2 == 0
    if fused_tail:      # attention backward + sampling heads' backward; their four follow-up kernels share one launch
        a = pre + "attn.sampling_"
        dqkv, dpooled = ops.rvsa_attn_bwd_fused(s["qkv"], s["params"], d["rel_h"], d["rel_w"], d["table"], s["lse"], do,
                                                G.g(pre + "attn.rel_pos_h"), G.g(pre + "attn.rel_pos_w"),
                                                G.g(pre + "attn.relative_position_bias_table"), G.g(pre + "attn.qkv.bias"),
                                                s["pooled"], d["off_w"], d["sc_w"], d["ang_w"],
                                                G.g(a + "offsets.2.weight"), G.g(a + "offsets.2.bias"), G.g(a + "scales.2.weight"),
                                                G.g(a + "scales.2.bias"), G.g(a + "angles.2.weight"), G.g(a + "angles.2.bias"), B, gh, gw, nH)
        pool = (dpooled, gh, gw)
    elif d["window"]:
        dqkv, dparams = ops.rvsa_attn_bwd(s["qkv"], s["params"], d["rel_h"], d["rel_w"], d["table"], s["lse"], do,
                                          G.g(pre + "attn.rel_pos_h"), G.g(pre + "attn.rel_pos_w"),
                                          G.g(pre + "attn.relative_position_bias_table"), B, gh, gw, nH,
                                          d_qkv_bias=G.g(pre + "attn.qkv.bias"))
    else:
        has_rel = d["rel_h"] is not None
        dqkv = ops.full_attn_bwd(s["qkv"], d["rel_h"], d["rel_w"], s["lse"], s["o"], do,
                                 G.g(pre + "attn.full_attn_rel_pos_h") if has_rel else None,
                                 G.g(pre + "attn.full_attn_rel_pos_w") if has_rel else None, B, gh, gw, nH)
        ops.colsum_bf16(dqkv, G.g(pre + "attn.qkv.bias"))
    if d["window"] and _POOL_FUSE and not fused_tail:        # the sampling heads' backward only needs dparams: it runs BEFORE the qkv GEMMs (light kernels after
        #                                   light kernels); its pooled (AvgPool) path joins dy1 inside the LayerNorm backward
        a = pre + "attn.sampling_"
        dpooled = ops.rvsa_sampling_bwd(dparams, s["pooled"], d["off_w"], d["sc_w"], d["ang_w"],
                                        G.g(a + "offsets.2.weight"), G.g(a + "offsets.2.bias"), G.g(a + "scales.2.weight"),
                                        G.g(a + "scales.2.bias"), G.g(a + "angles.2.weight"), G.g(a + "angles.2.bias"), None, B, gh, gw, nH)
        pool = (dpooled, gh, gw)
    dy1 = _linear_bwd(dqkv, s["y1"], d["qkv_w"], G.g(pre + "attn.qkv.weight"), T, 3 * C, C, sumsq=G.sumsq)
    if d["window"] and not _POOL_FUSE:
        a = pre + "attn.sampling_"
        ops.rvsa_sampling_bwd(dparams, s["pooled"], d["off_w"], d["sc_w"], d["ang_w"],
                              G.g(a + "offsets.2.weight"), G.g(a + "offsets.2.bias"), G.g(a + "scales.2.weight"),
                              G.g(a + "scales.2.bias"), G.g(a + "angles.2.weight"), G.g(a + "angles.2.bias"), dy1, B, gh, gw, nH)
    if nxt is None:
        dx0 = ops.layernorm_bwd(dy1, s["x0"], s["mean1"], s["rstd1"], d["norm1_w"], None, dx1, G.g(pre + "norm1.weight"), G.g(pre + "norm1.bias"),
                                pool_add=pool)
        return dx0, None
    return ops.layernorm_bwd(dy1, s["x0"], s["mean1"], s["rstd1"], d["norm1_w"], None, dx1, G.g(pre + "norm1.weight"),
                             G.g(pre + "norm1.bias"), cast=(nxt[0], N, nxt[1]), pool_add=pool)


def _convt_wgrad(G, wname, bname, dy, x_in, T, C):
    """Gradients of a ConvTranspose2d(k2,s2) run as the GEMM y[T, 4*Cout] = x[T, Cin] Wp^T, Wp [4*Cout, Cin] (row = (dy*2+dx)*Cout + co):
    dWp = dy^T x straight into the gradient storage when it is kept in GEMM layout (trainer), else via a temporary and one permuted copy
    to (Cin, Cout, 2, 2); the bias gradient is the column sum over the [4T, Cout] view of dy (the 4 sub-pixels share one bias)."""
    dWp = G.packed.get(wname)
    tmp = dWp is None
    if tmp:
        dWp = torch.empty(4 * C, C, device=dy.device, dtype=F32)
    else:
        G.touched.add(wname)
    ops.gemm(dy, x_in, 4 * C, C, T, dWp, a_mn=True, b_mn=True, mode=L.EPI_F32, lda=4 * C, ldb=C, sumsq=G.sumsq)
    if tmp:
        G.g(wname).copy_(dWp.view(2, 2, C, C).permute(3, 2, 0, 1))
    ops.colsum_bf16(dy.view(-1, C), G.g(bname))


def _fpn_tap_backward(m, W, S, k, grad, dx, G, B, gh, gw):
    """Fold the cotangent of pyramid level k into dx (fp32 [T, C], accumulated in place)."""
    C = W.C
    T = B * gh * gw
    grad = grad.contiguous()
    dev = dx.device
    if not m.apply_fpn or k == 2:
        ops.nchw_to_tok(grad, dx, B, gh, gw, C, 0, accumulate=True)
        return
    F_, sv = W.fpn, S["fpn"]
    if k == 3:
        ho, wo = gh // 2, gw // 2
        dp = torch.empty(B * ho * wo, C, device=dev, dtype=F32)
        ops.nchw_to_tok(grad, dp, B, ho, wo, C, 0)
        ops.maxpool2_tok_bwd(sv["f3"], dp, dx, B, gh, gw, C)
        return
    if k == 1:
        dv1 = torch.empty(T, 4 * C, device=dev, dtype=BF16)
        ops.nchw_to_tok(grad, dv1, B, gh, gw, C, 1)
        _convt_wgrad(G, "fpn2.0.weight", "fpn2.0.bias", dv1, sv["a1"], T, C)
        ops.gemm(dv1, F_["fpn2_0_w"], T, C, 4 * C, dx, b_mn=True, mode=L.EPI_F32, accumulate=True, lda=4 * C, ldb=C)
        return
    # k == 0: ConvT -> LN -> GELU -> ConvT
    du2 = torch.empty(4 * T, 4 * C, device=dev, dtype=BF16)
    ops.nchw_to_tok(grad, du2, B, gh, gw, C, 2)
    _convt_wgrad(G, "fpn1.3.weight", "fpn1.3.bias", du2, sv["z"], 4 * T, C)
    dz = torch.empty(4 * T, C, device=dev, dtype=BF16)
    ops.gemm(du2, F_["fpn1_3_w"], 4 * T, C, 4 * C, dz, b_mn=True, lda=4 * C, ldb=C)
    du1 = ops.layernorm_bwd(dz, sv["u1"].view(4 * T, C), sv["mean"], sv["rstd"], F_["ln_w"], F_["ln_b"], None,
                            G.g("fpn1.1.ln.weight"), G.g("fpn1.1.ln.bias"), gelu=True).view(T, 4 * C)
    _convt_wgrad(G, "fpn1.0.weight", "fpn1.0.bias", du1, sv["a0"], T, C)
    ops.gemm(du1, F_["fpn1_0_w"], T, C, 4 * C, dx, b_mn=True, mode=L.EPI_F32, accumulate=True, lda=4 * C, ldb=C)


def backward_impl(m, x, S, grad_outs, grad_store=None, after_block=None, after_fpn=None):
    """Returns fp32 gradients in ``m.parameters()`` order (None where a parameter does not take part).
    ``grad_store``: reuse a persistent (pre-zeroed where accumulated) GradStore; ``after_block(i)`` is called when the
    gradients of block i are final (used to launch the bucketed all-reduce while the backward continues).
    ``after_fpn``: when given, the whole pyramid tail (fpn1..4) is differentiated FIRST -- its cotangents w.r.t. the tapped features
    wait in per-tap buffers until the block loop reaches their block -- and ``after_fpn()`` is called as soon as the ConvTranspose2d
    weight gradients are final, so that their all-reduce overlaps the entire block backward instead of the last third of it."""
    W = S["W"]
    B, gh, gw, keep = S["B"], S["gh"], S["gw"], S["keep"]
    C, nH = W.C, W.nH
    T = B * gh * gw
    dev = x.device
    G = grad_store if grad_store is not None else GradStore(m, dev)
    grad_outs = list(grad_outs)

    dx = None
    if m.feature_mode == "last_norm":
        acc = torch.zeros(T, C, device=dev, dtype=F32)
        for k, g in enumerate(grad_outs):
            if g is not None:
                _fpn_tap_backward(m, W, S, k, g, acc, G, B, gh, gw)
        fin = S["final"]
        dyl = ops.scale_cast_bf16(acc)
        dx = ops.layernorm_bwd(dyl, fin["x"], fin["mean"], fin["rstd"], W.norm_w, None, None, G.g("norm.weight"), G.g("norm.bias"))
        taps = {}
    else:
        taps = {blk: k for k, blk in enumerate(m.out_indices)}

    g16 = None                            # bf16(keep_mlp * dx) of the block about to be processed, when already emitted
    tapped = lambda j: j in taps and grad_outs[taps[j]] is not None
    hoisted = {}
    if after_fpn is not None and m.feature_mode != "last_norm":
        for blk, k in sorted(taps.items(), key=lambda kv: -kv[0]):
            if grad_outs[k] is not None:
                buf = torch.zeros(T, C, device=dev, dtype=F32)
                _fpn_tap_backward(m, W, S, k, grad_outs[k], buf, G, B, gh, gw)
                hoisted[blk] = buf
        after_fpn()
    for i in range(W.depth - 1, -1, -1):
        if tapped(i):
            if i in hoisted:
                if dx is None:
                    dx = hoisted.pop(i)
                else:
                    L.call("mtp_add_f32", hoisted.pop(i).data_ptr(), dx.data_ptr(), dx.numel(), ops._stream())
            else:
                if dx is None:
                    dx = torch.zeros(T, C, device=dev, dtype=F32)
                _fpn_tap_backward(m, W, S, taps[i], grad_outs[taps[i]], dx, G, B, gh, gw)
        if dx is None:
            continue                      # blocks above the highest tapped block receive no gradient
        s = S["blocks"][i]
        if S["ckpt"]:                     # activation checkpointing: recompute this block's activations ([V]:799-800)
            ka = keep[i, 0] if keep is not None else None
            km = keep[i, 1] if keep is not None else None
            _, s = engine._block_forward(W.blocks[i], s["x0"], B, gh, gw, nH, ka, km, save=True)
        # the block below gets its MLP-branch cotangent from this block's last kernel unless a pyramid tap still adds into dx
        nxt = None
        if i > 0 and not tapped(i - 1):
            nxt = (keep[i - 1, 1] if keep is not None else None, G.g(f"blocks.{i - 1}.mlp.fc2.bias"))
        dx, g16 = _block_backward(i, W.blocks[i], s, dx, g16, G, B, gh, gw, nH, keep, nxt)
        S["blocks"][i] = None             # release activations as we go
        if after_block is not None:
            after_block(i)

    if dx is not None:
        # patch embed + pos_embed ([V]:790-794)
        g = ops.scale_cast_bf16(dx, None, 0, colsum=G.g("patch_embed.proj.bias"))
        if W.pos is not None:
            dpos = G.g("pos_embed").view(-1)
            scratch = torch.empty(B, gh * gw * C, device=dev, dtype=BF16)
            L.call("mtp_scale_cast_bf16", dx.data_ptr(), 0, 0, scratch.data_ptr(), dpos.data_ptr(), B, gh * gw * C, ops._stream())
        K0 = S["patches"].shape[1]
        ops.gemm(g, S["patches"], C, K0, T, G.g("patch_embed.proj.weight").view(C, K0), a_mn=True, b_mn=True, mode=L.EPI_F32,
                 lda=C, ldb=K0, ldo=K0, sumsq=G.sumsq)

    out = []
    for (name, p) in m.named_parameters():
        out.append(G.views[name] if (p.requires_grad and name in G.touched) else None)
    return out
