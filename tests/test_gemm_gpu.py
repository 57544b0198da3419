"""tcgen05 GEMM vs fp32 torch on identical bf16-rounded inputs (all operand layouts, tile widths and epilogues)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, device="cuda", generator=g) * scale).to(torch.bfloat16)


def _ref(A, B, a_mn, b_mn):
    a = A.float().t() if a_mn else A.float()
    b = B.float().t() if b_mn else B.float()
    return a @ b.t()


def _tol(K):
    return 2e-2 * math.sqrt(K / 64.0) * 0.05 + 1e-2


@pytest.mark.parametrize("layout", [(False, False), (False, True), (True, True), (True, False)])
@pytest.mark.parametrize("bn", [64, 128, 192, 256])
def test_gemm_layouts_tiles(layout, bn):
    from mtp_b200 import ops, _lib as L
    a_mn, b_mn = layout
    M, N, K = 392, 320, 456          # ragged in M, N (vs every BN) and K (not a multiple of 64)
    A = _mk((K, M) if a_mn else (M, K), seed=1)
    B = _mk((K, N) if b_mn else (N, K), seed=2)
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float32)
    ops.gemm(A, B, M, N, K, out, a_mn=a_mn, b_mn=b_mn, mode=L.EPI_F32, force_bn=bn)
    torch.cuda.synchronize()
    ref = _ref(A, B, a_mn, b_mn)
    err = (out - ref).abs().max().item()
    assert err < 1e-2 * math.sqrt(K), f"max err {err}"
    assert (out - ref).norm().item() / ref.norm().item() < 1e-5      # fp32 accumulate of exact bf16 products


@pytest.mark.parametrize("layout", [(False, False), (False, True), (True, True), (True, False)])
@pytest.mark.parametrize("bn", [64, 128, 192, 256])
@pytest.mark.parametrize("M", [392, 704])          # odd (4) and even (6) numbers of m-tiles: the odd case runs a dummy partner tile
def test_gemm_cluster_multicast(layout, bn, M):
    """2-CTA clusters with TMA-multicast of the shared B tile (force_bn = 1000 + width)."""
    from mtp_b200 import ops, _lib as L
    a_mn, b_mn = layout
    if b_mn and bn % 128 != 0:
        pytest.skip("MN-major B is multicast in 64-column boxes: needs an even number of boxes")
    N, K = 448, 456
    A = _mk((K, M) if a_mn else (M, K), seed=21)
    B = _mk((K, N) if b_mn else (N, K), seed=22)
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float32)
    ops.gemm(A, B, M, N, K, out, a_mn=a_mn, b_mn=b_mn, mode=L.EPI_F32, force_bn=1000 + bn)
    torch.cuda.synchronize()
    ref = _ref(A, B, a_mn, b_mn)
    assert (out - ref).norm().item() / ref.norm().item() < 1e-5


@pytest.mark.parametrize("shape", [(1568, 3072, 1024), (1568, 1024, 4096), (6272, 2304, 768), (128, 64, 64), (1, 8, 8)])
def test_gemm_model_shapes_bias_bf16(shape):
    from mtp_b200 import ops, _lib as L
    M, N, K = shape
    A, B = _mk((M, K), seed=3), _mk((N, K), 0.05, seed=4)
    bias = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(A, B, M, N, K, out, bias=bias)
    ref = A.float() @ B.float().t() + bias
    assert ((out.float() - ref).abs() <= 8e-3 * ref.abs() + 1e-2).all()


def test_epilogue_gelu_and_dgelu():
    from mtp_b200 import ops, _lib as L
    M, N, K = 300, 512, 256
    A, B = _mk((M, K), seed=5), _mk((N, K), 0.1, seed=6)
    bias = torch.randn(N, device="cuda") * 0.1
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    pre = torch.empty_like(out)
    ops.gemm(A, B, M, N, K, out, mode=L.EPI_BF16_GELU, bias=bias, out2=pre)
    ref_pre = A.float() @ B.float().t() + bias
    assert ((pre.float() - ref_pre).abs() <= 8e-3 * ref_pre.abs() + 1e-2).all()
    ref = torch.nn.functional.gelu(ref_pre)
    assert ((out.float() - ref).abs() <= 8e-3 * ref.abs() + 1e-2).all()
    # dgelu: out = acc * gelu'(aux)
    h = _mk((M, N), seed=7)
    dg = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(A, B, M, N, K, dg, mode=L.EPI_BF16_DGELU, aux=h)
    hf = h.float().requires_grad_(True)
    torch.nn.functional.gelu(hf).sum().backward()
    ref = (A.float() @ B.float().t()) * hf.grad
    assert ((dg.float() - ref).abs() <= 1e-2 * ref.abs() + 2e-2).all()


def test_epilogue_residual_pos_accumulate():
    from mtp_b200 import ops, _lib as L
    B_, ntok, N, K = 3, 100, 256, 128
    M = B_ * ntok
    A, W = _mk((M, K), seed=8), _mk((N, K), 0.1, seed=9)
    bias = torch.randn(N, device="cuda")
    resid = torch.randn(M, N, device="cuda")
    keep = torch.tensor([0.0, 1.25, 1.0], device="cuda")
    out = torch.empty(M, N, device="cuda")
    ops.gemm(A, W, M, N, K, out, mode=L.EPI_F32_RESID, bias=bias, aux=resid, row_scale=keep, rows_per_group=ntok)
    ref = resid + keep.repeat_interleave(ntok)[:, None] * (A.float() @ W.float().t() + bias)
    assert (out - ref).abs().max().item() < 1e-3
    # in-place on the residual stream
    r2 = resid.clone()
    ops.gemm(A, W, M, N, K, r2, mode=L.EPI_F32_RESID, bias=bias, aux=r2, row_scale=None)
    assert (r2 - (resid + A.float() @ W.float().t() + bias)).abs().max().item() < 1e-3
    pos = torch.randn(ntok, N, device="cuda")
    ops.gemm(A, W, M, N, K, out, mode=L.EPI_F32_POS, bias=bias, aux=pos, pos_rows=ntok)
    ref = A.float() @ W.float().t() + bias + pos.repeat(B_, 1)
    assert (out - ref).abs().max().item() < 1e-3
    acc = torch.ones(M, N, device="cuda")
    ops.gemm(A, W, M, N, K, acc, mode=L.EPI_F32, accumulate=True)
    assert (acc - (1 + A.float() @ W.float().t())).abs().max().item() < 1e-3


def test_linear_fwd_dgrad_wgrad_consistency():
    """The three layouts reproduce autograd of y = x W^T on bf16-rounded operands."""
    from mtp_b200 import ops, _lib as L
    T, Cin, Cout = 1000, 256, 384
    x, W = _mk((T, Cin), seed=10), _mk((Cout, Cin), 0.1, seed=11)
    dy = _mk((T, Cout), seed=12)
    dx = torch.empty(T, Cin, device="cuda", dtype=torch.bfloat16)
    ops.gemm(dy, W, T, Cin, Cout, dx, b_mn=True)                      # dX = dY W
    dW = torch.empty(Cout, Cin, device="cuda")
    ops.gemm(dy, x, Cout, Cin, T, dW, a_mn=True, b_mn=True, mode=L.EPI_F32)   # dW = dY^T X
    ref_dx = dy.float() @ W.float()
    ref_dW = dy.float().t() @ x.float()
    assert ((dx.float() - ref_dx).abs() <= 8e-3 * ref_dx.abs() + 1e-2).all()
    assert (dW - ref_dW).norm().item() / ref_dW.norm().item() < 1e-5


def test_gemm_rejects_bad_args():
    from mtp_b200 import ops, _lib as L
    A, B = _mk((16, 20)), _mk((16, 20))
    out = torch.empty(16, 16, device="cuda")
    with pytest.raises(L.MtpError):
        ops.gemm(A, B, 16, 16, 20, out, mode=L.EPI_F32)       # lda not a multiple of 8


def test_grouped_dual_launch_matches_separate():
    """dgrad + wgrad of one Linear in a single grouped launch (host LPT schedule over both problems)."""
    from mtp_b200 import ops, _lib as L
    T, Cin, Cout = 1568, 1024, 256
    x, W = _mk((T, Cin), seed=31), _mk((Cout, Cin), 0.1, seed=32)
    dy = _mk((T, Cout), seed=33)
    dx = torch.empty(T, Cin, device="cuda", dtype=torch.bfloat16)
    dW = torch.empty(Cout, Cin, device="cuda")
    for force in (0, 128, 1128, 256):
        dx.zero_(); dW.zero_()
        ops.gemm_dual(dict(A=dy, B=W, M=T, N=Cin, K=Cout, out=dx, b_mn=True, lda=Cout, ldb=Cin),
                      dict(A=dy, B=x, M=Cout, N=Cin, K=T, out=dW, a_mn=True, b_mn=True, mode=L.EPI_F32, lda=Cout, ldb=Cin), force_bn=force)
        torch.cuda.synchronize()
        ref_dx = dy.float() @ W.float()
        ref_dW = dy.float().t() @ x.float()
        assert ((dx.float() - ref_dx).abs() <= 8e-3 * ref_dx.abs() + 1e-2).all(), force
        assert (dW - ref_dW).norm().item() / ref_dW.norm().item() < 1e-5, force


@pytest.mark.parametrize("bn", [0, 64, 192, 1256])
def test_epilogue_column_sums(bn):
    """colsum of the BF16 / BF16_DGELU epilogues = column sums of the stored bf16 output (bias gradient of the Linear below)."""
    from mtp_b200 import ops, _lib as L
    M, N, K = 1568, 1096, 264          # ragged M (12.25 tiles) and N
    A, Bm = _mk((M, K), seed=31), _mk((K, N), 0.1, seed=32)
    h = _mk((M, N), seed=33)
    for mode, aux in ((L.EPI_BF16, None), (L.EPI_BF16_DGELU, h)):
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        cs = torch.zeros(N, device="cuda")
        ops.gemm(A, Bm, M, N, K, out, b_mn=True, mode=mode, aux=aux, colsum=cs, force_bn=bn)
        ref = out.float().sum(0)
        assert (cs - ref).abs().max().item() <= 2e-4 * max(1.0, ref.abs().max().item())
        out2 = torch.empty_like(out)
        ops.gemm(A, Bm, M, N, K, out2, b_mn=True, mode=mode, aux=aux, force_bn=bn)
        assert torch.equal(out, out2)


@pytest.mark.parametrize("bn", [0, 64, 128, 256, 1128, 1256])
@pytest.mark.parametrize("b_mn", [False, True])
def test_b_static_prefetch_matches(bn, b_mn):
    """b_static only moves the first B loads ahead of the dependency wait: results are bit-identical, also for K shorter than the ring."""
    from mtp_b200 import ops, _lib as L
    for K in (72, 456, 1024):
        M, N = 704, 512
        A = _mk((M, K), seed=41)
        B = _mk((K, N) if b_mn else (N, K), seed=42)
        o0 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        o1 = torch.empty_like(o0)
        ops.gemm(A, B, M, N, K, o0, b_mn=b_mn, force_bn=bn)
        ops.gemm(A, B, M, N, K, o1, b_mn=b_mn, force_bn=bn, b_static=True)
        torch.cuda.synchronize()
        assert torch.equal(o0, o1)
    # grouped launch: dgrad-like (K-major A, MN-major B) + wgrad-like (both MN-major) sharing A
    G = _mk((704, 512), seed=45)
    Wt, X = _mk((512, 512), seed=44), _mk((704, 512), seed=43)
    outs = []
    for st in (False, True):
        dx = torch.empty(704, 512, device="cuda", dtype=torch.bfloat16)
        dW = torch.empty(512, 512, device="cuda")
        ops.gemm_dual(dict(A=G, B=Wt, M=704, N=512, K=512, out=dx, b_mn=True, b_static=st),
                      dict(A=G, B=X, M=512, N=512, K=704, out=dW, a_mn=True, b_mn=True, mode=L.EPI_F32, b_static=st), force_bn=bn)
        outs.append((dx, dW))
    torch.cuda.synchronize()
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("case", ["qkv_fwd", "fc1_fwd_gelu", "fc2_fwd", "fc1_dgrad", "fc2_wgrad", "fc2_dual", "ragged", "pixbias"])
def test_gemm_variant2_matches_persistent_kernel(case):
    """Variant 2 (one tile per CTA, two CTAs per SM) against the persistent kernel: all operand layouts, the fused epilogues, the grouped launch."""
    from mtp_b200 import ops, _lib as L
    T, C = 1568, 1024
    dev = "cuda"
    bf = lambda *s: (torch.randn(*s, device=dev) * 0.5).to(torch.bfloat16)

    def run(variant):
        L.call("mtp_gemm_set_variant", variant)
        torch.manual_seed(11)
        if case == "qkv_fwd":
            A, B = bf(T, C), bf(3 * C, C)
            out = torch.empty(T, 3 * C, device=dev, dtype=torch.bfloat16)
            ops.gemm(A, B, T, 3 * C, C, out, bias=torch.randn(3 * C, device=dev), b_static=True)
            outs = [out]
        elif case == "fc1_fwd_gelu":
            A, B = bf(T, C), bf(4 * C, C)
            out = torch.empty(T, 4 * C, device=dev, dtype=torch.bfloat16)
            out2 = torch.empty_like(out)
            ops.gemm(A, B, T, 4 * C, C, out, mode=L.EPI_BF16_GELU, bias=torch.randn(4 * C, device=dev), out2=out2)
            outs = [out, out2]
        elif case == "fc2_fwd":
            A, B = bf(T, 4 * C), bf(C, 4 * C)
            out = torch.empty(T, C, device=dev)
            ops.gemm(A, B, T, C, 4 * C, out, mode=L.EPI_F32_RESID, bias=torch.randn(C, device=dev), aux=torch.randn(T, C, device=dev),
                     row_scale=torch.rand(8, device=dev), rows_per_group=196)
            outs = [out]
        elif case == "fc1_dgrad":
            g, w = bf(T, 4 * C), bf(4 * C, C)
            out = torch.empty(T, C, device=dev, dtype=torch.bfloat16)
            cs = torch.zeros(C, device=dev)
            ops.gemm(g, w, T, C, 4 * C, out, b_mn=True, lda=4 * C, ldb=C, colsum=cs)
            outs = [out, cs]
        elif case == "fc2_wgrad":
            g, x = bf(T, C), bf(T, 4 * C)
            out = torch.empty(C, 4 * C, device=dev)
            ss = torch.zeros(1, device=dev)
            ops.gemm(g, x, C, 4 * C, T, out, a_mn=True, b_mn=True, mode=L.EPI_F32, lda=C, ldb=4 * C, ldo=4 * C, sumsq=ss)
            outs = [out, ss]
        elif case == "fc2_dual":
            g, w, x, hp = bf(T, C), bf(C, 4 * C), bf(T, 4 * C), bf(T, 4 * C)
            dx = torch.empty(T, 4 * C, device=dev, dtype=torch.bfloat16)
            dW = torch.empty(C, 4 * C, device=dev)
            cs = torch.zeros(4 * C, device=dev)
            ops.gemm_dual(dict(A=g, B=w, M=T, N=4 * C, K=C, out=dx, b_mn=True, mode=L.EPI_BF16_DGELU, aux=hp, lda=C, ldb=4 * C, colsum=cs),
                          dict(A=g, B=x, M=C, N=4 * C, K=T, out=dW, a_mn=True, b_mn=True, mode=L.EPI_F32, lda=C, ldb=4 * C, ldo=4 * C))
            outs = [dx, dW, cs]
        elif case == "pixbias":
            A, B = bf(T, C), bf(4 * C, C)
            out = torch.empty(T, 4 * C, device=dev, dtype=torch.bfloat16)
            ops.gemm(A, B, T, 4 * C, C, out, bias=torch.randn(C, device=dev), ps=(0, 0, C))
            outs = [out]
        else:       # ragged M / N / K with an accumulate epilogue
            M, N, K = 1000, 840, 1992
            A, B = bf(M, K), bf(N, K)
            out = torch.randn(M, N, device=dev)
            ops.gemm(A, B, M, N, K, out, mode=L.EPI_F32, accumulate=True)
            outs = [out]
        torch.cuda.synchronize()
        return [o.float().clone() for o in outs], L.load().mtp_gemm_last_config()
    try:
        ref, cfg1 = run(1)
        got, cfg2 = run(2)
    finally:
        L.call("mtp_gemm_set_variant", 1)
    assert cfg1 < 3000 <= cfg2, (cfg1, cfg2)
    for a, b in zip(got, ref):
        assert torch.isfinite(a).all()
        err = float((a - b).norm() / b.norm().clamp_min(1e-20))
        assert err < 1e-5, err          # same k order per tile: identical up to the atomics of the column sums / sumsq


@pytest.mark.parametrize("bn", [0, 64, 192, 256, 1128, 1256])
def test_specialised_epilogues_match_generic(bn):
    """The per-mode tile epilogues (kernel template parameter ESET, picked per launch) against the generic epilogue (debug mode 30):
    bit-identical outputs for every mode they cover, single and grouped launches, ragged M and a tile-ragged N."""
    from mtp_b200 import ops, _lib as L
    M, N, K = 1568, 1120, 264          # 12.25 row tiles; N = 35 chunks of 32 columns: the last tile of every width is partial
    A, Bm = _mk((M, K), seed=51), _mk((K, N), 0.1, seed=52)
    h = _mk((M, N), seed=53)
    bias = torch.randn(N, device="cuda")
    resid = torch.randn(M, N, device="cuda")
    keep = torch.tensor([0.0, 1.25, 1.0, 2.0], device="cuda")
    X, G = _mk((M, 512), seed=54), _mk((M, N), seed=55)

    def run(generic):
        L.call("mtp_gemm_set_debug_mode", 30 if generic else 0)
        res = []
        try:
            o = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16); cs = torch.zeros(N, device="cuda")
            ops.gemm(A, Bm, M, N, K, o, b_mn=True, bias=bias, colsum=cs, force_bn=bn); res += [o, cs]
            o = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16); pre = torch.zeros_like(o)
            ops.gemm(A, Bm, M, N, K, o, b_mn=True, mode=L.EPI_BF16_GELU, bias=bias, out2=pre, force_bn=bn); res += [o, pre]
            o = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16); cs = torch.zeros(N, device="cuda")
            ops.gemm(A, Bm, M, N, K, o, b_mn=True, mode=L.EPI_BF16_DGELU, aux=h, colsum=cs, force_bn=bn); res += [o, cs]
            o = torch.zeros(M, N, device="cuda"); sq = torch.zeros(1, device="cuda")
            ops.gemm(A, Bm, M, N, K, o, b_mn=True, mode=L.EPI_F32, sumsq=sq, force_bn=bn); res += [o, sq]
            o = torch.zeros(M, N, device="cuda")
            ops.gemm(A, Bm, M, N, K, o, b_mn=True, mode=L.EPI_F32_RESID, bias=bias, aux=resid, row_scale=keep, rows_per_group=392, force_bn=bn); res += [o]
            r2 = resid.clone()                   # in place on the residual stream
            ops.gemm(A, Bm, M, N, K, r2, b_mn=True, mode=L.EPI_F32_RESID, aux=r2, force_bn=bn); res += [r2]
            # grouped: a dgrad-like bf16 (or GELU') problem + a wgrad-like fp32 problem with the gradient-norm partial sum
            for mode, aux in ((L.EPI_BF16, None), (L.EPI_BF16_DGELU, _mk((M, 512), seed=56))):
                dx = torch.zeros(M, 512, device="cuda", dtype=torch.bfloat16); dW = torch.zeros(N, 512, device="cuda")
                cs = torch.zeros(512, device="cuda"); sq = torch.zeros(1, device="cuda")
                W = _mk((N, 512), 0.1, seed=57)
                ops.gemm_dual(dict(A=G, B=W, M=M, N=512, K=N, out=dx, b_mn=True, mode=mode, aux=aux, colsum=cs),
                              dict(A=G, B=X, M=N, N=512, K=M, out=dW, a_mn=True, b_mn=True, mode=L.EPI_F32, sumsq=sq), force_bn=bn)
                res += [dx, dW, cs, sq]
            torch.cuda.synchronize()
        finally:
            L.call("mtp_gemm_set_debug_mode", 0)
        return res

    fast, gen = run(False), run(True)
    assert len(fast) == len(gen)
    for i, (a, b) in enumerate(zip(fast, gen)):
        if a.numel() == 1 or (a.dim() == 1 and a.dtype == torch.float32):      # atomically accumulated sums: order differs
            assert (a - b).abs().max().item() <= 1e-4 * max(1.0, b.abs().max().item()), (bn, i)
        elif a.dtype == torch.float32:       # scale * acc + residual may or may not be contracted into an FMA: one fp32 ulp
            assert (a - b).abs().max().item() <= 2e-6 * max(1.0, b.abs().max().item()), (bn, i)
        else:
            assert torch.equal(a, b), (bn, i, (a.float() - b.float()).abs().max().item())
    ref = A.float() @ Bm.float() + bias
    assert ((fast[0].float() - ref).abs() <= 8e-3 * ref.abs() + 1e-2).all()
