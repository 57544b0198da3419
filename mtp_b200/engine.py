"""Host-side orchestration of the backbone forward / backward over the C-ABI kernels (libmtp_b200.so).

Data layout in HBM (T = B*Hp*Wp tokens, image-major):
  residual stream            fp32 [T, C]           (never rounded; every block reads and writes it once per branch)
  LN outputs, qkv, attention bf16 [T, C|3C|4C]     (GEMM operands; q|k|v head-major so a K/V tap is one 128-B line)
  weights                    bf16 copies of the fp32 nn.Parameters, refreshed when a parameter's version changes
  gradients                  fp32 for parameters and the residual stream, bf16 for activation cotangents

Mirrors ViT_Win_RVSA_V3_WSZ7.forward_features ([V]:787-813) and Block.forward ([V]:506-513).  There is no fallback
path: every op below is a kernel of the library; a missing library or a CPU tensor raises.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import _lib as L
from . import ops

BF16, F32 = torch.bfloat16, torch.float32


class EngineState:
    """Per-module cache of bf16 / packed weight copies keyed on (parameter version, storage pointer)."""

    def __init__(self):
        self.cache: Dict[str, tuple] = {}
        self.pinned: Dict[str, torch.Tensor] = {}      # always-current bf16 views maintained by the fused optimizer
        self.consts: Dict[tuple, torch.Tensor] = {}    # small device constants (DropPath keep probabilities)
        self.listeners: List = []                      # callbacks run by invalidate() (PretrainStep.refresh_mirror)

    def get(self, key, param, fn):
        pin = self.pinned.get(key)
        if pin is not None:
            return pin
        ver = (param._version, param.data_ptr())
        ent = self.cache.get(key)
        if ent is None or ent[0] != ver:
            with torch.no_grad():
                ent = (ver, fn(param.detach()))
            self.cache[key] = ent
        return ent[1]

    def clear(self):
        self.cache.clear()

    def invalidate(self):
        """Weights were changed behind autograd's back (``p.data`` writes, ``load_state_dict``, checkpoint loading): drop the cached
        bf16 copies and let the owners of always-current mirrors (the fused-optimizer trainer) refresh theirs."""
        self.cache.clear()
        for fn in list(self.listeners):
            fn()


def _as_bf16(w):
    w = w.contiguous()
    if w.dtype == F32 and w.numel() % 4 == 0:
        return ops.cast_f32_bf16(w)
    return w.to(BF16)


def _pack_convt(w):
    """ConvTranspose2d weight (Cin, Cout, 2, 2) -> GEMM B operand [4*Cout, Cin], row = (dy*2+dx)*Cout + co."""
    return _as_bf16(w.permute(2, 3, 1, 0).reshape(4 * w.shape[1], w.shape[0]).contiguous())


def _f32(p, name):
    if p.dtype != F32 or not p.is_cuda or not p.is_contiguous():
        raise RuntimeError(f"mtp_b200: parameter {name} must be a contiguous fp32 CUDA tensor (got {p.dtype}, {p.device})")
    return p.detach()


class _W:
    """Resolved kernel-ready views of one module's parameters for the current step."""

    def __init__(self, m):
        st = m._engine_state
        C = m.embed_dim
        self.C, self.nH, self.depth = C, m.num_heads, m.depth
        pe = m.patch_embed.proj
        self.pe_w = st.get("pe_w", pe.weight, lambda w: _as_bf16(w.reshape(C, -1)))
        self.pe_b = _f32(pe.bias, "patch_embed.proj.bias")
        self.pos = _f32(m.pos_embed, "pos_embed").reshape(-1, C) if m.pos_embed is not None else None
        self.blocks = []
        for i, blk in enumerate(m.blocks):
            d = {"window": blk.window}
            for nm, lin in (("qkv", blk.attn.qkv), ("proj", blk.attn.proj), ("fc1", blk.mlp.fc1), ("fc2", blk.mlp.fc2)):
                d[nm + "_w"] = st.get(f"b{i}.{nm}", lin.weight, _as_bf16)
                if lin.bias is None:
                    raise RuntimeError("mtp_b200: qkv_bias=False is not used by any MTP config and is unsupported")
                d[nm + "_b"] = _f32(lin.bias, f"blocks.{i}.{nm}.bias")
            for nm in ("norm1", "norm2"):
                ln = getattr(blk, nm)
                d[nm + "_w"], d[nm + "_b"] = _f32(ln.weight, nm), _f32(ln.bias, nm)
            a = blk.attn
            if blk.window:
                d["rel_h"], d["rel_w"] = _f32(a.rel_pos_h, "rel_pos_h"), _f32(a.rel_pos_w, "rel_pos_w")
                d["table"] = _f32(a.relative_position_bias_table, "relative_position_bias_table")
                for nm, seq in (("off", a.sampling_offsets), ("sc", a.sampling_scales), ("ang", a.sampling_angles)):
                    conv = seq[2]
                    d[nm + "_w"] = _f32(conv.weight, "sampling").reshape(conv.weight.shape[0], C)
                    d[nm + "_b"] = _f32(conv.bias, "sampling")
            elif m.full_attn_rel_pos:
                d["rel_h"], d["rel_w"] = _f32(a.full_attn_rel_pos_h, "full_attn_rel_pos_h"), _f32(a.full_attn_rel_pos_w, "full_attn_rel_pos_w")
            else:
                d["rel_h"] = d["rel_w"] = None
            self.blocks.append(d)
        if m.feature_mode == "last_norm":
            self.norm_w, self.norm_b = _f32(m.norm.weight, "norm.weight"), _f32(m.norm.bias, "norm.bias")
        if m.apply_fpn:
            self.fpn = {}
            for key, conv in (("fpn1_0", m.fpn1[0]), ("fpn1_3", m.fpn1[3]), ("fpn2_0", m.fpn2[0])):
                self.fpn[key + "_w"] = st.get(key, conv.weight, _pack_convt)
                self.fpn[key + "_b"] = _f32(conv.bias, key + ".bias")      # [Cout]; the GEMM epilogue applies it with period Cout
            self.fpn["ln_w"], self.fpn["ln_b"] = _f32(m.fpn1[1].ln.weight, "fpn1.1.ln"), _f32(m.fpn1[1].ln.bias, "fpn1.1.ln")


# ------------------------------------------------------------------------------------------------------- forward
def _block_forward(d, x0, B, gh, gw, nH, keep_a, keep_m, save):
    """One Block.forward ([V]:506-513).  x0: fp32 [T, C] residual stream.  Returns (x2, saved-dict or None)."""
    T, C = x0.shape
    N = gh * gw
    y1, mean1, rstd1 = ops.layernorm_fwd(x0, d["norm1_w"], d["norm1_b"], save_stats=save)
    params = pooled = None
    if d["window"]:      # the sampling heads only need y1: launched BEFORE the qkv GEMM
        params, pooled = ops.rvsa_sampling_fwd(y1, d["off_w"], d["off_b"], d["sc_w"], d["sc_b"], d["ang_w"], d["ang_b"], B, gh, gw, nH,
                                               save_pooled=save)
    qkv = torch.empty(T, 3 * C, device=x0.device, dtype=BF16)
    ops.gemm(y1, d["qkv_w"], T, 3 * C, C, qkv, bias=d["qkv_b"], b_static=True)
    if d["window"]:
        o, lse = ops.rvsa_attn_fwd(qkv, params, d["rel_h"], d["rel_w"], d["table"], B, gh, gw, nH, save_lse=save)
    else:
        o, lse = ops.full_attn_fwd(qkv, d["rel_h"], d["rel_w"], B, gh, gw, nH, save_lse=save)
    x1 = torch.empty_like(x0) if save else x0
    ops.gemm(o, d["proj_w"], T, C, C, x1, mode=L.EPI_F32_RESID, bias=d["proj_b"], aux=x0, row_scale=keep_a, rows_per_group=N, b_static=True)
    y2, mean2, rstd2 = ops.layernorm_fwd(x1, d["norm2_w"], d["norm2_b"], save_stats=save)
    hid = d["fc1_w"].shape[0]
    a = torch.empty(T, hid, device=x0.device, dtype=BF16)
    hpre = torch.empty(T, hid, device=x0.device, dtype=BF16) if save else None
    ops.gemm(y2, d["fc1_w"], T, hid, C, a, mode=L.EPI_BF16_GELU, bias=d["fc1_b"], out2=hpre, b_static=True)
    x2 = torch.empty_like(x0) if save else x1
    ops.gemm(a, d["fc2_w"], T, C, hid, x2, mode=L.EPI_F32_RESID, bias=d["fc2_b"], aux=x1, row_scale=keep_m, rows_per_group=N, b_static=True)
    saved = None
    if save:
        saved = dict(x0=x0, mean1=mean1, rstd1=rstd1, y1=y1, qkv=qkv, params=params, pooled=pooled, lse=lse, o=o, x1=x1,
                     mean2=mean2, rstd2=rstd2, y2=y2, hpre=hpre, a=a)
    return x2, saved


def _fpn_forward(m, W, feats, B, gh, gw, out_dtype, save):
    """fpn1..fpn4 ([V]:640-654, 807-811) on token-major features; returns the four NCHW maps (+ saved tensors)."""
    C = W.C
    T = B * gh * gw
    outs, saved = [], {}
    if not m.apply_fpn:
        for f in feats:
            outs.append(ops.tok_to_nchw(f, B, gh, gw, C, 0, out_dtype))
        return outs, saved
    F_ = W.fpn
    # fpn1: ConvT -> Norm2d(LN over C) -> GELU -> ConvT ; rows of u1 viewed as [4T, C] are the 2x-upsampled pixels
    a0 = feats[0] if feats[0].dtype == BF16 else ops.cast_f32_bf16(feats[0])
    u1 = torch.empty(T, 4 * C, device=a0.device, dtype=BF16)
    ops.gemm(a0, F_["fpn1_0_w"], T, 4 * C, C, u1, bias=F_["fpn1_0_b"], ps=(0, 0, C), b_static=True)
    z, mean, rstd = ops.layernorm_fwd(u1.view(4 * T, C), F_["ln_w"], F_["ln_b"], gelu=True, save_stats=save)
    u2 = torch.empty(4 * T, 4 * C, device=a0.device, dtype=BF16)
    ops.gemm(z, F_["fpn1_3_w"], 4 * T, 4 * C, C, u2, bias=F_["fpn1_3_b"], ps=(0, 0, C), b_static=True)
    outs.append(ops.tok_to_nchw(u2, B, gh, gw, C, 2, out_dtype))
    # fpn2: ConvT
    a1 = feats[1] if feats[1].dtype == BF16 else ops.cast_f32_bf16(feats[1])
    v1 = torch.empty(T, 4 * C, device=a0.device, dtype=BF16)
    ops.gemm(a1, F_["fpn2_0_w"], T, 4 * C, C, v1, bias=F_["fpn2_0_b"], ps=(0, 0, C), b_static=True)
    outs.append(ops.tok_to_nchw(v1, B, gh, gw, C, 1, out_dtype))
    # fpn3: identity
    outs.append(ops.tok_to_nchw(feats[2], B, gh, gw, C, 0, out_dtype))
    # fpn4: MaxPool2d(2, 2)
    f3 = feats[3] if feats[3].dtype == F32 else feats[3].float()
    pooled = ops.maxpool2_tok_fwd(f3, B, gh, gw, C)
    outs.append(ops.tok_to_nchw(pooled, B, gh // 2, gw // 2, C, 0, out_dtype))
    if save:
        saved = dict(a0=a0, u1=u1, mean=mean, rstd=rstd, z=z, a1=a1, f3=f3)
    return outs, saved


def _check_input(m, x):
    if not x.is_cuda:
        raise RuntimeError("mtp_b200: the backbone runs only on a CUDA (sm_100a) device; there is no CPU fallback")
    if x.device.index != torch.cuda.current_device():
        raise RuntimeError(f"mtp_b200: input lives on {x.device} but the current CUDA device is cuda:{torch.cuda.current_device()} "
                           "(kernels are enqueued on the current device's stream): wrap the call in torch.cuda.device(x.device)")
    pre = getattr(m, "input_preprocess", None)
    hwc = x.dtype == torch.uint8 and pre is not None and pre.layout == "hwc"
    cdim, hdim, wdim = (3, 1, 2) if hwc else (1, 2, 3)
    if x.dim() != 4 or x.shape[cdim] != m.in_chans:
        raise ValueError(f"expected {'(B, H, W, %d)' % m.in_chans if hwc else '(B, %d, H, W)' % m.in_chans}, got {tuple(x.shape)}")
    gh, gw = m.patch_embed.patch_shape
    if x.shape[hdim] != gh * 16 or x.shape[wdim] != gw * 16:
        raise ValueError(f"input {(x.shape[hdim], x.shape[wdim])} must equal img_size {(gh * 16, gw * 16)} (fixed pos_embed / rel-pos tables, [V]:103,629)")
    if x.dtype == torch.uint8:
        if pre is None:
            raise TypeError("uint8 input needs module.input_preprocess = ImagePreprocess(mean, std, ...) (the fused MTP_DataPreprocessor)")
    elif x.dtype not in (F32, BF16):
        raise TypeError("input must be float32, bfloat16, or uint8 with input_preprocess set")
    return gh, gw


def _forward_impl(m, x, keep, save):
    gh, gw = _check_input(m, x)
    W = _W(m)
    B, C, nH = x.shape[0], W.C, W.nH
    T = B * gh * gw
    x = x.contiguous()
    if x.dtype == torch.uint8:          # MTP_DataPreprocessor (BGR->RGB, mean/std) folded into the patch gather
        pre = m.input_preprocess
        patches = ops.patchify_u8(x, pre.mean, pre.std, pre.flip_channels, pre.layout == "hwc")
        out_dtype = pre.out_dtype
    else:
        patches = ops.patchify(x)
        out_dtype = x.dtype
    xres = torch.empty(T, C, device=x.device, dtype=F32)
    K0 = patches.shape[1]
    if W.pos is not None:
        ops.gemm(patches, W.pe_w, T, C, K0, xres, mode=L.EPI_F32_POS, bias=W.pe_b, aux=W.pos, pos_rows=gh * gw, b_static=True)
    else:
        ops.gemm(patches, W.pe_w, T, C, K0, xres, mode=L.EPI_F32, bias=W.pe_b, b_static=True)
    ckpt = save and m.use_checkpoint
    feats, blocks_saved = [], []
    for i, d in enumerate(W.blocks):
        ka = keep[i, 0] if keep is not None else None
        km = keep[i, 1] if keep is not None else None
        if ckpt:
            blocks_saved.append(dict(x0=xres))                     # activation checkpointing ([V]:799-800): keep the input only
            xres, _ = _block_forward(d, xres.clone(), B, gh, gw, nH, ka, km, save=False)
        else:
            xin = xres
            if not save and i > 0 and m.feature_mode == "multi" and (i - 1) in m.out_indices:
                xin = xres.clone()                                 # keep the tapped feature intact in in-place inference mode
            xres, sv = _block_forward(d, xin, B, gh, gw, nH, ka, km, save)
            blocks_saved.append(sv)
        if m.feature_mode == "multi" and i in m.out_indices:
            feats.append(xres)
    final = None
    if m.feature_mode == "last_norm":
        yl, meanl, rstdl = ops.layernorm_fwd(xres, W.norm_w, W.norm_b, save_stats=save)
        feats = [yl, yl, yl, yl]
        final = dict(x=xres, mean=meanl, rstd=rstdl)
    outs, fpn_saved = _fpn_forward(m, W, feats, B, gh, gw, out_dtype, save)
    ctx = None
    if save:
        ctx = dict(W=W, B=B, gh=gh, gw=gw, patches=patches, blocks=blocks_saved, fpn=fpn_saved, final=final, keep=keep,
                   feats=feats, out_dtype=out_dtype, ckpt=ckpt)
    return outs, ctx


def _draw_keep(m, B, device):
    """DropPath multipliers, timm semantics ([V]:31-39): per sample bernoulli(keep)/keep, one draw per branch per block."""
    probs = [blk.drop_path_prob for blk in m.blocks]
    if not m.training or all(p == 0.0 for p in probs):
        return None
    key = ("keep_prob", tuple(probs), B, str(device))
    kp = m._engine_state.consts.get(key)
    if kp is None:      # built once (an H2D copy is not capturable in a CUDA graph); bernoulli itself is graph-safe
        kp = torch.tensor([1.0 - p for p in probs], dtype=F32).view(-1, 1, 1).expand(len(probs), 2, B).contiguous().to(device)
        m._engine_state.consts = {key: kp}
    return torch.bernoulli(kp) / kp


def backbone_apply(m, x, keep=None):
    """Entry point used by ViT_Win_RVSA_V3_WSZ7.forward_features."""
    L.load()
    if keep is None:
        keep = _draw_keep(m, x.shape[0], x.device)
    params = [p for p in m.parameters()]
    needs_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params))
    if getattr(m, "precision", "bf16") == "fp32x3":
        if needs_grad:
            raise RuntimeError("mtp_b200: precision='fp32x3' is forward-only (inference / verification); wrap the call in torch.no_grad()")
        _check_input(m, x)
        from . import precise
        with torch.no_grad():
            return precise.forward(m, x.float() if x.dtype != torch.uint8 else m.input_preprocess.reference(x).contiguous())
    if not needs_grad:
        with torch.no_grad():
            outs, _ = _forward_impl(m, x, keep, save=False)
        return outs
    from .autograd import BackboneFunction
    return list(BackboneFunction.apply(m, x, keep, *params))
