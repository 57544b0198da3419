"""Thin torch-tensor wrappers over the C ABI (pointers + sizes + the current CUDA stream).  No autograd here."""
import ctypes

import torch

from . import _lib as L

BF16, F32 = torch.bfloat16, torch.float32


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return 0 if t is None else t.data_ptr()


def _chk(t, dtype, name):
    assert t.is_cuda and t.dtype == dtype and t.is_contiguous(), f"{name}: need contiguous cuda {dtype}, got {t.dtype} {t.device} contiguous={t.is_contiguous()}"


def gemm(A, B, M, N, K, out, *, a_mn=False, b_mn=False, mode=L.EPI_BF16, bias=None, out2=None, aux=None, row_scale=None,
         rows_per_group=0, pos_rows=0, accumulate=False, ldo=None, lda=None, ldb=None, ps=None, force_bn=0, colsum=None, b_static=False, sumsq=None,
         hilo=False, out_lo=0):
    """acc[m,n] = sum_k A[m,k] B[n,k] with fused epilogue; see include/mtp_b200.h."""
    ep = L.Epilogue()
    ep.mode = mode
    ep.colsum = _p(colsum)
    ep.sumsq = _p(sumsq)
    ep.b_static = int(bool(b_static))
    ep.ldo = int(ldo if ldo is not None else out.shape[-1])
    ep.bias, ep.out, ep.out2, ep.aux, ep.row_scale = _p(bias), _p(out), _p(out2), _p(aux), _p(row_scale)
    ep.rows_per_group, ep.pos_rows, ep.accumulate = int(rows_per_group), int(pos_rows), int(bool(accumulate))
    ep.hilo, ep.out_lo_offset = int(bool(hilo)), int(out_lo)
    if ps is not None:
        ep.ps_h, ep.ps_w, ep.ps_cout = ps
    lda = int(lda if lda is not None else A.shape[-1])
    ldb = int(ldb if ldb is not None else B.shape[-1])
    L.call("mtp_gemm_bf16", A.data_ptr(), lda, int(a_mn), B.data_ptr(), ldb, int(b_mn), int(M), int(N), int(K),
           ep, int(force_bn), _stream())
    return out


def _desc(A, B, M, N, K, out, *, a_mn=False, b_mn=False, mode=L.EPI_BF16, bias=None, out2=None, aux=None, row_scale=None,
          rows_per_group=0, pos_rows=0, accumulate=False, ldo=None, lda=None, ldb=None, colsum=None, b_static=False, sumsq=None):
    ep = L.Epilogue()
    ep.mode = mode
    ep.colsum = _p(colsum)
    ep.sumsq = _p(sumsq)
    ep.b_static = int(bool(b_static))
    ep.ldo = int(ldo if ldo is not None else out.shape[-1])
    ep.bias, ep.out, ep.out2, ep.aux, ep.row_scale = _p(bias), _p(out), _p(out2), _p(aux), _p(row_scale)
    ep.rows_per_group, ep.pos_rows, ep.accumulate = int(rows_per_group), int(pos_rows), int(bool(accumulate))
    d = L.GemmDesc()
    d.A, d.lda, d.a_mn_major = A.data_ptr(), int(lda if lda is not None else A.shape[-1]), int(a_mn)
    d.B, d.ldb, d.b_mn_major = B.data_ptr(), int(ldb if ldb is not None else B.shape[-1]), int(b_mn)
    d.M, d.N, d.K = int(M), int(N), int(K)
    d.ep = ctypes.pointer(ep)
    d._keep = ep
    return d


def gemm_dual(g0, g1, force_bn=0):
    """Two independent GEMMs (dicts of gemm() arguments) in one grouped persistent launch."""
    d0, d1 = _desc(**g0), _desc(**g1)
    L.call("mtp_gemm_bf16_dual", ctypes.byref(d0), ctypes.byref(d1), int(force_bn), _stream())


def layernorm_fwd(x, gamma, beta, eps=1e-6, gelu=False, save_stats=True):
    rows, C = x.shape
    y = torch.empty(rows, C, device=x.device, dtype=BF16)
    mean = torch.empty(rows, device=x.device, dtype=F32) if save_stats else None
    rstd = torch.empty(rows, device=x.device, dtype=F32) if save_stats else None
    L.call("mtp_layernorm_fwd", x.data_ptr(), int(x.dtype == BF16), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(),
           _p(mean), _p(rstd), rows, C, float(eps), int(gelu), _stream())
    return y, mean, rstd


def layernorm_bwd(dy, x, mean, rstd, gamma, beta, dres, dgamma, dbeta, gelu=False, cast=None, pool_add=None):
    """``cast=(row_scale|None, rows_per_group, colsum|None)``: also return bf16(row_scale * dx) (+ its column sums into colsum).
    ``pool_add=(dpooled, h, w)``: fold the AvgPool backward of the RVSA sampling heads into dy."""
    rows, C = x.shape
    dx = torch.empty(rows, C, device=x.device, dtype=x.dtype)
    g16 = None
    sc, rpg, cs = (None, 0, None)
    if cast is not None:
        sc, rpg, cs = cast
        g16 = torch.empty(rows, C, device=x.device, dtype=BF16)
    L.call("mtp_layernorm_bwd", dy.data_ptr(), x.data_ptr(), int(x.dtype == BF16), mean.data_ptr(), rstd.data_ptr(),
           gamma.data_ptr(), _p(beta), _p(dres), dx.data_ptr(), int(dx.dtype == BF16), dgamma.data_ptr(), dbeta.data_ptr(),
           _p(sc), int(rpg), _p(g16), _p(cs), _p(pool_add[0]) if pool_add else 0, int(pool_add[1]) if pool_add else 0,
           int(pool_add[2]) if pool_add else 0, rows, C, int(gelu), _stream())
    return dx if cast is None else (dx, g16)


def scale_cast_bf16(x, row_scale=None, rows_per_group=0, colsum=None):
    rows, C = x.shape
    out = torch.empty(rows, C, device=x.device, dtype=BF16)
    L.call("mtp_scale_cast_bf16", x.data_ptr(), _p(row_scale), int(rows_per_group), out.data_ptr(), _p(colsum), rows, C, _stream())
    return out


def colsum_bf16(x, colsum, ld=None):
    rows, C = x.shape
    L.call("mtp_colsum_bf16", x.data_ptr(), int(ld if ld is not None else x.stride(0)), colsum.data_ptr(), rows, C, _stream())
    return colsum


def cast_f32_bf16(x):
    out = torch.empty(x.shape, device=x.device, dtype=BF16)
    L.call("mtp_cast_f32_bf16", x.data_ptr(), out.data_ptr(), x.numel(), _stream())
    return out


def add_bf16_into_f32(src, dst):
    L.call("mtp_add_bf16_into_f32", src.data_ptr(), dst.data_ptr(), src.numel(), _stream())
    return dst


# ---------------------------------------------------------------------------------------------- attention
def rvsa_sampling_fwd(yn, w_off, b_off, w_sc, b_sc, w_ang, b_ang, B, h, w, nH, save_pooled=True):
    C = yn.shape[-1]
    nwin = ((h + 6) // 7) * ((w + 6) // 7)
    params = torch.empty(B * nwin, nH, 8, device=yn.device, dtype=F32)       # all 8 slots are written by the kernel
    pooled = torch.empty(B * nwin, C, device=yn.device, dtype=F32)       # written by the pooling kernel, read by the conv kernel
    L.call("mtp_rvsa_sampling_fwd", yn.data_ptr(), w_off.data_ptr(), b_off.data_ptr(), w_sc.data_ptr(), b_sc.data_ptr(),
           w_ang.data_ptr(), b_ang.data_ptr(), _p(pooled), params.data_ptr(), B, h, w, C, nH, _stream())
    return params, pooled


def rvsa_attn_fwd(qkv, params, rel_h, rel_w, table, B, h, w, nH, save_lse=True):
    C = qkv.shape[-1] // 3
    out = torch.empty(qkv.shape[0], C, device=qkv.device, dtype=BF16)
    lse = torch.empty(params.shape[0] * nH, 49, device=qkv.device, dtype=F32) if save_lse else None
    L.call("mtp_rvsa_attn_fwd", qkv.data_ptr(), params.data_ptr(), rel_h.data_ptr(), rel_w.data_ptr(), table.data_ptr(),
           out.data_ptr(), _p(lse), B, h, w, C, nH, _stream())
    return out, lse


def full_attn_fwd(qkv, rel_h, rel_w, B, gh, gw, nH, save_lse=True):
    C = qkv.shape[-1] // 3
    out = torch.empty(qkv.shape[0], C, device=qkv.device, dtype=BF16)
    lse = torch.empty(B, nH, gh * gw, device=qkv.device, dtype=F32) if save_lse else None
    L.call("mtp_full_attn_fwd", qkv.data_ptr(), _p(rel_h), _p(rel_w), out.data_ptr(), _p(lse), B, gh, gw, C, nH, _stream())
    return out, lse


# ---------------------------------------------------------------------------------------------- layout
def patchify(img):
    B, cin, H, W = img.shape
    assert img.is_contiguous() and img.dtype in (F32, BF16)
    out = torch.empty(B * (H // 16) * (W // 16), cin * 256, device=img.device, dtype=BF16)
    L.call("mtp_patchify", img.data_ptr(), int(img.dtype == BF16), out.data_ptr(), B, cin, H, W, _stream())
    return out


def patchify_u8(img, mean, std, flip_channels, hwc):
    """uint8 images (B,cin,H,W) or (B,H,W,cin) -> normalised bf16 patch rows (MTP_DataPreprocessor + im2col in one pass)."""
    assert img.is_cuda and img.dtype == torch.uint8 and img.is_contiguous()
    if hwc:
        B, H, W, cin = img.shape
    else:
        B, cin, H, W = img.shape
    assert len(mean) == cin and len(std) == cin
    out = torch.empty(B * (H // 16) * (W // 16), cin * 256, device=img.device, dtype=BF16)
    fa = (ctypes.c_float * cin)
    L.call("mtp_patchify_u8", img.data_ptr(), int(bool(hwc)), int(bool(flip_channels)), fa(*[float(v) for v in mean]),
           fa(*[float(v) for v in std]), out.data_ptr(), B, cin, H, W, _stream())
    return out


def tok_to_nchw(tok, B, h, w, C, level, out_dtype):
    out = torch.empty(B, C, h << level, w << level, device=tok.device, dtype=out_dtype)
    L.call("mtp_tok_to_nchw", tok.data_ptr(), int(tok.dtype == BF16), int(tok.shape[-1]), out.data_ptr(), int(out_dtype == BF16),
           B, h, w, C, level, _stream())
    return out


def nchw_to_tok(grad, tok, B, h, w, C, level, accumulate=False):
    assert grad.is_contiguous()
    L.call("mtp_nchw_to_tok", grad.data_ptr(), int(grad.dtype == BF16), tok.data_ptr(), int(tok.dtype == BF16), int(tok.shape[-1]),
           int(accumulate), B, h, w, C, level, _stream())
    return tok


def maxpool2_tok_fwd(x, B, h, w, C):
    y = torch.empty(B * (h // 2) * (w // 2), C, device=x.device, dtype=F32)
    L.call("mtp_maxpool2_tok_fwd", x.data_ptr(), y.data_ptr(), B, h, w, C, _stream())
    return y


def maxpool2_tok_bwd(x, dy, dx, B, h, w, C):
    L.call("mtp_maxpool2_tok_bwd", x.data_ptr(), dy.data_ptr(), dx.data_ptr(), B, h, w, C, _stream())
    return dx


# ---------------------------------------------------------------------------------------------- attention backward
def _workspace(nbytes, device):
    return torch.empty((nbytes + 3) // 4, device=device, dtype=F32)


def rvsa_attn_bwd(qkv, params, rel_h, rel_w, table, lse, dout, d_rel_h, d_rel_w, d_table, B, h, w, nH, d_qkv_bias=None):
    C = qkv.shape[-1] // 3
    dqkv = torch.empty_like(qkv)
    dparams = torch.empty_like(params)
    # scratch_zeroed = 0: the per-call memset of the fp32 scatter scratch doubles as its L2 prefetch -- keeping the scratch zeroed
    # between calls instead (scratch_zeroed = 1) measured 0.4 ms / step SLOWER: the red.global.add of the tap scatter then hit
    # lines that had been evicted to HBM since the previous block
    ws = _workspace(L.load().mtp_rvsa_bwd_workspace_bytes(B, h, w, C, nH), qkv.device)
    zeroed = 0
    L.call("mtp_rvsa_attn_bwd", qkv.data_ptr(), params.data_ptr(), rel_h.data_ptr(), rel_w.data_ptr(), table.data_ptr(), lse.data_ptr(),
           dout.data_ptr(), dqkv.data_ptr(), dparams.data_ptr(), d_rel_h.data_ptr(), d_rel_w.data_ptr(), d_table.data_ptr(),
           _p(d_qkv_bias), ws.data_ptr(), zeroed, B, h, w, C, nH, _stream())
    return dqkv, dparams


def rvsa_attn_bwd_fused(qkv, params, rel_h, rel_w, table, lse, dout, d_rel_h, d_rel_w, d_table, d_qkv_bias, pooled, w_off, w_sc, w_ang,
                        dw_off, db_off, dw_sc, db_sc, dw_ang, db_ang, B, h, w, nH):
    """rvsa_attn_bwd + rvsa_sampling_bwd(dyn=None) with their follow-up kernels in one launch; returns (dqkv, dpooled [B*nWin, C] fp32)."""
    C = qkv.shape[-1] // 3
    dqkv = torch.empty_like(qkv)
    dparams = torch.empty_like(params)
    lib = L.load()
    n1 = lib.mtp_rvsa_bwd_workspace_bytes(B, h, w, C, nH)
    n2 = lib.mtp_rvsa_sampling_bwd_workspace_bytes(B, h, w, C, nH)
    n1 = (n1 + 255) // 256 * 256
    ws = _workspace(n1 + n2, qkv.device)
    L.call("mtp_rvsa_attn_bwd_fused", qkv.data_ptr(), params.data_ptr(), rel_h.data_ptr(), rel_w.data_ptr(), table.data_ptr(), lse.data_ptr(),
           dout.data_ptr(), dqkv.data_ptr(), dparams.data_ptr(), d_rel_h.data_ptr(), d_rel_w.data_ptr(), d_table.data_ptr(), _p(d_qkv_bias),
           ws.data_ptr(), pooled.data_ptr(), w_off.data_ptr(), w_sc.data_ptr(), w_ang.data_ptr(), dw_off.data_ptr(), db_off.data_ptr(),
           dw_sc.data_ptr(), db_sc.data_ptr(), dw_ang.data_ptr(), db_ang.data_ptr(), ws.data_ptr() + n1, B, h, w, C, nH, _stream())
    n_bw = pooled.shape[0]
    sws = ws[n1 // 4:]      # fp32 workspace: the sampling part starts n1 bytes in
    return dqkv, sws[n_bw * 5 * nH: n_bw * 5 * nH + n_bw * C].view(n_bw, C)


def rvsa_sampling_bwd(dparams, pooled, w_off, w_sc, w_ang, dw_off, db_off, dw_sc, db_sc, dw_ang, db_ang, dyn, B, h, w, nH):
    """``dyn=None``: returns dpooled [B*nWin, C] (fp32) for ``layernorm_bwd(pool_add=...)`` instead of adding it into dyn."""
    C = pooled.shape[-1]
    ws = _workspace(L.load().mtp_rvsa_sampling_bwd_workspace_bytes(B, h, w, C, nH), pooled.device)
    L.call("mtp_rvsa_sampling_bwd", dparams.data_ptr(), pooled.data_ptr(), w_off.data_ptr(), w_sc.data_ptr(), w_ang.data_ptr(),
           dw_off.data_ptr(), db_off.data_ptr(), dw_sc.data_ptr(), db_sc.data_ptr(), dw_ang.data_ptr(), db_ang.data_ptr(),
           _p(dyn), ws.data_ptr(), B, h, w, C, nH, _stream())
    if dyn is not None:
        return dyn
    n_bw = pooled.shape[0]
    return ws[n_bw * 5 * nH: n_bw * 5 * nH + n_bw * C].view(n_bw, C)


def full_attn_bwd(qkv, rel_h, rel_w, lse, out, dout, d_rel_h, d_rel_w, B, gh, gw, nH):
    C = qkv.shape[-1] // 3
    dqkv = torch.empty_like(qkv)
    ws = _workspace(L.load().mtp_full_attn_bwd_workspace_bytes(B, gh, gw, nH), qkv.device)
    L.call("mtp_full_attn_bwd", qkv.data_ptr(), _p(rel_h), _p(rel_w), lse.data_ptr(), out.data_ptr(), dout.data_ptr(), dqkv.data_ptr(),
           _p(d_rel_h), _p(d_rel_w), ws.data_ptr(), B, gh, gw, C, nH, _stream())
    return dqkv
