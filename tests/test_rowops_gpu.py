"""LayerNorm / cast / column-sum kernels vs fp32 torch."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("C", [128, 768, 1024])
@pytest.mark.parametrize("rows", [1, 37, 1568])
def test_layernorm_fwd_bwd(C, rows):
    from mtp_b200 import ops
    torch.manual_seed(0)
    x = torch.randn(rows, C, device="cuda") * 2 + 0.5
    g = torch.randn(C, device="cuda") * 0.2 + 1
    b = torch.randn(C, device="cuda") * 0.2
    y, mean, rstd = ops.layernorm_fwd(x, g, b)
    xr = x.clone().requires_grad_(True)
    gr, br = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr, (C,), gr, br, 1e-6)
    assert ((y.float() - ref).abs() <= 8e-3 * ref.abs() + 1e-3).all()
    assert (mean - x.mean(1)).abs().max().item() < 1e-5
    dy = torch.randn(rows, C, device="cuda").to(torch.bfloat16)
    dres = torch.randn(rows, C, device="cuda")
    ref.backward(dy.float())
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    dx = ops.layernorm_bwd(dy, x, mean, rstd, g, b, dres, dg, db)
    assert (dx - (xr.grad + dres)).abs().max().item() < 1e-3
    assert (dg - gr.grad).abs().max().item() < 1e-3 * max(1.0, gr.grad.abs().max().item())
    assert (db - br.grad).abs().max().item() < 1e-3 * max(1.0, br.grad.abs().max().item())


def test_layernorm_gelu_bf16_variant():
    from mtp_b200 import ops
    torch.manual_seed(1)
    rows, C = 500, 128
    x = (torch.randn(rows, C, device="cuda") * 2).to(torch.bfloat16)
    g = torch.randn(C, device="cuda") * 0.2 + 1
    b = torch.randn(C, device="cuda") * 0.2
    y, mean, rstd = ops.layernorm_fwd(x, g, b, gelu=True)
    xr = x.float().requires_grad_(True)
    gr, br = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = torch.nn.functional.gelu(torch.nn.functional.layer_norm(xr, (C,), gr, br, 1e-6))
    assert ((y.float() - ref).abs() <= 8e-3 * ref.abs() + 2e-3).all()
    dy = torch.randn(rows, C, device="cuda").to(torch.bfloat16)
    ref.backward(dy.float())
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    dx = ops.layernorm_bwd(dy, x, mean, rstd, g, b, None, dg, db, gelu=True)
    assert ((dx.float() - xr.grad).abs() <= 1e-2 * xr.grad.abs() + 5e-3).all()
    assert (dg - gr.grad).abs().max().item() < 2e-3 * max(1.0, gr.grad.abs().max().item())
    assert (db - br.grad).abs().max().item() < 2e-3 * max(1.0, br.grad.abs().max().item())


def test_casts_and_colsums():
    from mtp_b200 import ops
    torch.manual_seed(2)
    rows, C, ntok = 300, 256, 100
    x = torch.randn(rows, C, device="cuda")
    keep = torch.tensor([1.0, 0.0, 2.0], device="cuda")
    cs = torch.zeros(C, device="cuda")
    y = ops.scale_cast_bf16(x, keep, ntok, cs)
    ref = x * keep.repeat_interleave(ntok)[:, None]
    assert ((y.float() - ref).abs() <= 4e-3 * ref.abs() + 1e-6).all()
    assert (cs - ref.sum(0)).abs().max().item() < 1e-3
    cs2 = torch.zeros(C, device="cuda")
    ops.colsum_bf16(y, cs2)
    assert (cs2 - y.float().sum(0)).abs().max().item() < 1e-3
    z = ops.cast_f32_bf16(x)
    assert torch.equal(z, x.to(torch.bfloat16))
    acc = torch.ones(rows, C, device="cuda")
    ops.add_bf16_into_f32(z, acc)
    assert torch.equal(acc, 1 + z.float())


@pytest.mark.parametrize("rows", [37, 1568, 6272])
def test_layernorm_bwd_fused_cast_matches_scale_cast(rows):
    """The cast/colsum outputs of mtp_layernorm_bwd equal a separate mtp_scale_cast_bf16 of its fp32 result."""
    from mtp_b200 import ops
    torch.manual_seed(5)
    C, ntok = 1024, 196 if rows % 196 == 0 else 37
    x = torch.randn(rows, C, device="cuda")
    g = torch.randn(C, device="cuda") * 0.2 + 1
    _, mean, rstd = ops.layernorm_fwd(x, g, torch.zeros_like(g))
    dy = torch.randn(rows, C, device="cuda").to(torch.bfloat16)
    dres = torch.randn(rows, C, device="cuda")
    keep = (torch.rand(rows // ntok, device="cuda") > 0.3).float() / 0.7
    dg, db, cs = (torch.zeros(C, device="cuda") for _ in range(3))
    dx, g16 = ops.layernorm_bwd(dy, x, mean, rstd, g, None, dres, dg, db, cast=(keep, ntok, cs))
    dg2, db2, cs2 = (torch.zeros(C, device="cuda") for _ in range(3))
    dx2 = ops.layernorm_bwd(dy, x, mean, rstd, g, None, dres, dg2, db2)
    ref16 = ops.scale_cast_bf16(dx2, keep, ntok, cs2)
    assert torch.equal(dx, dx2)
    assert torch.equal(g16, ref16)
    assert (dg - dg2).abs().max().item() <= 1e-4 * dg2.abs().max().item()
    assert (cs - cs2).abs().max().item() <= 1e-4 * max(1.0, cs2.abs().max().item())
    # without a row scale
    cs3 = torch.zeros(C, device="cuda")
    _, g16b = ops.layernorm_bwd(dy, x, mean, rstd, g, None, dres, dg, db, cast=(None, 0, cs3))
    assert torch.equal(g16b, dx2.to(torch.bfloat16))


@pytest.mark.parametrize("grid", [14, 10])
def test_layernorm_bwd_pool_add(grid):
    """pool_add folds dpooled[window(t)] / 49 (AvgPool backward of the RVSA sampling heads, zero-padded 7x7 windows) into dy."""
    from mtp_b200 import ops
    torch.manual_seed(7)
    B, C = 2, 256
    h = w = grid
    rows = B * h * w
    pad = (7 - grid % 7) % 7
    pt = pad // 2
    nwin = (grid + pad) // 7
    x = torch.randn(rows, C, device="cuda")
    g = torch.randn(C, device="cuda") * 0.2 + 1
    _, mean, rstd = ops.layernorm_fwd(x, g, torch.zeros_like(g))
    dy = torch.randn(rows, C, device="cuda").to(torch.bfloat16)
    dp = torch.randn(B * nwin * nwin, C, device="cuda")
    t = torch.arange(rows, device="cuda")
    xx, yy, bb = t % w, (t // w) % h, t // (w * h)
    win = (bb * nwin + (yy + pt) // 7) * nwin + (xx + pt) // 7
    dy_eff = dy.float() + dp[win] / 49.0
    xr = x.clone().requires_grad_(True)
    gr = g.clone().requires_grad_(True)
    torch.nn.functional.layer_norm(xr, (C,), gr, torch.zeros_like(g), 1e-6).backward(dy_eff)
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    dx = ops.layernorm_bwd(dy, x, mean, rstd, g, None, None, dg, db, pool_add=(dp, h, w))
    assert (dx - xr.grad).abs().max().item() < 1e-3
    assert (dg - gr.grad).abs().max().item() < 1e-3 * max(1.0, gr.grad.abs().max().item())
    assert (db - dy_eff.sum(0)).abs().max().item() < 1e-3 * max(1.0, dy_eff.sum(0).abs().max().item())
