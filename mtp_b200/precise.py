"""fp32-class forward of the backbone (``precision="fp32x3"``): the same kernels' arithmetic with every bf16 tensor of the fast path
carried as hi | lo bf16 word pairs -- GEMMs as three tcgen05 passes (hi*hi + hi*lo + lo*hi, fp32 accumulate), attention, LayerNorm and
the pyramid in fp32 (csrc/precise.cu, HILO instantiations of the SIMT attention kernels).  Forward only.

Why it exists: BASELINE.json's north star asks for the ViT-L forward "within 1e-3 rel" of the reference, whose arithmetic is fp32
(main_pretrain.py never enters autocast).  With bf16 storage between kernels no implementation can be closer than ~3e-3 .. 5e-3 at
depth 12 / 24 (the reference's own bf16-autocast run: 3.5e-3 .. 6.6e-3); this mode is ~1e-5, at roughly 3x the GEMM time.
It doubles as the full-depth logic check that rounding noise cannot mask (tests/test_precise_gpu.py).   Mirrors [V]:787-817.
"""
from __future__ import annotations

import torch

from . import _lib as L
from . import ops
from .engine import _f32

BF16, F32 = torch.bfloat16, torch.float32


def split_hilo(x2d: torch.Tensor) -> torch.Tensor:
    """fp32 [rows, K] -> bf16 [rows, 2K] (hi | lo)."""
    x2d = x2d.contiguous()
    rows, K = x2d.shape
    out = torch.empty(rows, 2 * K, device=x2d.device, dtype=BF16)
    L.call("mtp_split_hilo", x2d.data_ptr(), K, out.data_ptr(), rows, K, ops._stream())
    return out


def _ln(x, gamma, beta, rows, C, *, hilo_in=None, gelu=False, eps=1e-6):
    y = torch.empty(rows, 2 * C, device=x.device, dtype=BF16)
    ld, lo, sub = hilo_in if hilo_in else (0, 0, 1)
    L.call("mtp_layernorm_fwd_hilo", x.data_ptr(), int(hilo_in is not None), ld, lo, sub, gamma.data_ptr(), beta.data_ptr(), y.data_ptr(),
           rows, C, float(eps), int(gelu), ops._stream())
    return y


def _weights(m):
    """hi | lo copies of the GEMM weights, cached on the module's engine state (dropped by EngineState.invalidate())."""
    st = m._engine_state
    C = m.embed_dim

    def get(key, p, fn):
        return st.get("hl." + key, p, lambda w: split_hilo(fn(w).float()))
    W = {"pe": get("pe", m.patch_embed.proj.weight, lambda w: w.reshape(C, -1))}
    for i, blk in enumerate(m.blocks):
        for nm, lin in (("qkv", blk.attn.qkv), ("proj", blk.attn.proj), ("fc1", blk.mlp.fc1), ("fc2", blk.mlp.fc2)):
            W[f"b{i}.{nm}"] = get(f"b{i}.{nm}", lin.weight, lambda w: w)
    if m.apply_fpn:
        for key, conv in (("fpn1_0", m.fpn1[0]), ("fpn1_3", m.fpn1[3]), ("fpn2_0", m.fpn2[0])):
            W[key] = get(key, conv.weight, lambda w: w.permute(2, 3, 1, 0).reshape(4 * w.shape[1], w.shape[0]))
    return W


def forward(m, x: torch.Tensor):
    """Returns the four fp32 NCHW maps.  ``x``: fp32 (B, in_chans, H, W) on the current CUDA device; eval semantics (no DropPath)."""
    L.load()
    if x.dtype != F32:
        raise TypeError("precision='fp32x3' takes float32 images")
    if m.training and any(blk.drop_path_prob > 0 for blk in m.blocks):
        raise RuntimeError("precision='fp32x3' is an inference / verification mode: call .eval() first (DropPath is not implemented here)")
    if m.feature_mode != "multi":
        raise NotImplementedError("precision='fp32x3' supports the multi-tap feature mode of [V] only")
    gh, gw = m.patch_embed.patch_shape
    B, C, nH = x.shape[0], m.embed_dim, m.num_heads
    N = gh * gw
    T = B * N
    dev = x.device
    W = _weights(m)
    st = ops._stream()
    x = x.contiguous()
    K0 = m.in_chans * 256
    patches = torch.empty(T, 2 * K0, device=dev, dtype=BF16)
    L.call("mtp_patchify_hilo", x.data_ptr(), patches.data_ptr(), B, m.in_chans, x.shape[2], x.shape[3], st)
    xres = torch.empty(T, C, device=dev, dtype=F32)
    pe_b = _f32(m.patch_embed.proj.bias, "patch_embed.proj.bias")
    if m.pos_embed is not None:
        ops.gemm(patches, W["pe"], T, C, K0, xres, mode=L.EPI_F32_POS, bias=pe_b, aux=_f32(m.pos_embed, "pos_embed").reshape(-1, C),
                 pos_rows=N, hilo=True, lda=2 * K0, ldb=2 * K0)
    else:
        ops.gemm(patches, W["pe"], T, C, K0, xres, mode=L.EPI_F32, bias=pe_b, hilo=True, lda=2 * K0, ldb=2 * K0)
    feats = []
    for i, blk in enumerate(m.blocks):
        a = blk.attn
        y1 = _ln(xres, _f32(blk.norm1.weight, "norm1"), _f32(blk.norm1.bias, "norm1"), T, C)
        qkv = torch.empty(T, 6 * C, device=dev, dtype=BF16)
        ops.gemm(y1, W[f"b{i}.qkv"], T, 3 * C, C, qkv, bias=_f32(a.qkv.bias, "qkv.bias"), hilo=True, lda=2 * C, ldb=2 * C, ldo=6 * C, out_lo=3 * C)
        o = torch.empty(T, 2 * C, device=dev, dtype=BF16)
        if blk.window:
            nwin = ((gh + 6) // 7) * ((gw + 6) // 7)
            params = torch.empty(B * nwin, nH, 8, device=dev, dtype=F32)
            pooled = torch.empty(B * nwin, C, device=dev, dtype=F32)
            cw = lambda seq: _f32(seq[2].weight, "sampling").reshape(seq[2].weight.shape[0], C)
            L.call("mtp_rvsa_sampling_fwd_hilo", y1.data_ptr(), cw(a.sampling_offsets).data_ptr(), _f32(a.sampling_offsets[2].bias, "s").data_ptr(),
                   cw(a.sampling_scales).data_ptr(), _f32(a.sampling_scales[2].bias, "s").data_ptr(), cw(a.sampling_angles).data_ptr(),
                   _f32(a.sampling_angles[2].bias, "s").data_ptr(), pooled.data_ptr(), params.data_ptr(), B, gh, gw, C, nH, st)
            L.call("mtp_rvsa_attn_fwd_hilo", qkv.data_ptr(), params.data_ptr(), _f32(a.rel_pos_h, "rel").data_ptr(), _f32(a.rel_pos_w, "rel").data_ptr(),
                   _f32(a.relative_position_bias_table, "table").data_ptr(), o.data_ptr(), B, gh, gw, C, nH, st)
        else:
            rh = _f32(a.full_attn_rel_pos_h, "rel").data_ptr() if m.full_attn_rel_pos else 0
            rw = _f32(a.full_attn_rel_pos_w, "rel").data_ptr() if m.full_attn_rel_pos else 0
            L.call("mtp_full_attn_fwd_hilo", qkv.data_ptr(), rh, rw, o.data_ptr(), B, gh, gw, C, nH, st)
        ops.gemm(o, W[f"b{i}.proj"], T, C, C, xres, mode=L.EPI_F32_RESID, bias=_f32(a.proj.bias, "proj.bias"), aux=xres, hilo=True,
                 lda=2 * C, ldb=2 * C)
        y2 = _ln(xres, _f32(blk.norm2.weight, "norm2"), _f32(blk.norm2.bias, "norm2"), T, C)
        hid = blk.mlp.fc1.weight.shape[0]
        act = torch.empty(T, 2 * hid, device=dev, dtype=BF16)
        ops.gemm(y2, W[f"b{i}.fc1"], T, hid, C, act, mode=L.EPI_BF16_GELU, bias=_f32(blk.mlp.fc1.bias, "fc1.bias"), hilo=True, lda=2 * C, ldb=2 * C,
                 ldo=2 * hid, out_lo=hid)
        ops.gemm(act, W[f"b{i}.fc2"], T, C, hid, xres, mode=L.EPI_F32_RESID, bias=_f32(blk.mlp.fc2.bias, "fc2.bias"), aux=xres, hilo=True,
                 lda=2 * hid, ldb=2 * hid)
        if i in m.out_indices:
            feats.append(xres.clone())
    if not m.apply_fpn:
        return [ops.tok_to_nchw(f, B, gh, gw, C, 0, F32) for f in feats]
    outs = []

    def to_nchw_hilo(tok, level):
        out = torch.empty(B, C, gh << level, gw << level, device=dev, dtype=F32)
        L.call("mtp_tok_to_nchw_hilo", tok.data_ptr(), 8 * C, 4 * C, out.data_ptr(), B, gh, gw, C, level, st)
        return out
    # fpn1: ConvT -> Norm2d (LN over C) -> GELU -> ConvT    ([V]:640-646)
    u1 = torch.empty(T, 8 * C, device=dev, dtype=BF16)
    ops.gemm(split_hilo(feats[0]), W["fpn1_0"], T, 4 * C, C, u1, bias=_f32(m.fpn1[0].bias, "fpn"), ps=(0, 0, C), hilo=True, lda=2 * C, ldb=2 * C,
             ldo=8 * C, out_lo=4 * C)
    z = _ln(u1, _f32(m.fpn1[1].ln.weight, "fpn.ln"), _f32(m.fpn1[1].ln.bias, "fpn.ln"), 4 * T, C, hilo_in=(8 * C, 4 * C, 4), gelu=True)
    u2 = torch.empty(4 * T, 8 * C, device=dev, dtype=BF16)
    ops.gemm(z, W["fpn1_3"], 4 * T, 4 * C, C, u2, bias=_f32(m.fpn1[3].bias, "fpn"), ps=(0, 0, C), hilo=True, lda=2 * C, ldb=2 * C, ldo=8 * C,
             out_lo=4 * C)
    outs.append(to_nchw_hilo(u2, 2))
    v1 = torch.empty(T, 8 * C, device=dev, dtype=BF16)
    ops.gemm(split_hilo(feats[1]), W["fpn2_0"], T, 4 * C, C, v1, bias=_f32(m.fpn2[0].bias, "fpn"), ps=(0, 0, C), hilo=True, lda=2 * C, ldb=2 * C,
             ldo=8 * C, out_lo=4 * C)
    outs.append(to_nchw_hilo(v1, 1))
    outs.append(ops.tok_to_nchw(feats[2], B, gh, gw, C, 0, F32))
    pooled = ops.maxpool2_tok_fwd(feats[3], B, gh, gw, C)
    outs.append(ops.tok_to_nchw(pooled, B, gh // 2, gw // 2, C, 0, F32))
    return outs
