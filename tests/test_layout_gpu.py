"""Layout kernels: patchify, token<->NCHW with pixel-shuffle levels, 2x2 max pool."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_patchify_matches_conv_unfold():
    from mtp_b200 import ops
    torch.manual_seed(0)
    x = torch.randn(2, 3, 64, 96, device="cuda")
    p = ops.patchify(x)
    ref = x.reshape(2, 3, 4, 16, 6, 16).permute(0, 2, 4, 1, 3, 5).reshape(2 * 4 * 6, 768)
    assert torch.equal(p, ref.to(torch.bfloat16))
    pb = ops.patchify(x.to(torch.bfloat16))
    assert torch.equal(pb, ref.to(torch.bfloat16))


@pytest.mark.parametrize("level", [0, 1, 2])
@pytest.mark.parametrize("dt_in,dt_out", [(torch.bfloat16, torch.float32), (torch.float32, torch.bfloat16), (torch.bfloat16, torch.bfloat16)])
def test_tok_nchw_roundtrip(level, dt_in, dt_out):
    from mtp_b200 import ops
    B, h, w, C = 2, 5, 7, 96
    torch.manual_seed(level)
    # reference: build the NCHW map first, derive the token matrix from the definition of the nested transposed convs
    Ho, Wo = h << level, w << level
    nchw = torch.randn(B, C, Ho, Wo, device="cuda").to(dt_in).float()
    if level == 0:
        tok = nchw.permute(0, 2, 3, 1).reshape(B * h * w, C)
    elif level == 1:
        tok = nchw.reshape(B, C, h, 2, w, 2).permute(0, 2, 4, 3, 5, 1).reshape(B * h * w, 4 * C)
    else:
        # rows (b,y,x,g1) cols (g2,c): Y = 4y + 2*dy1 + dy2
        tok = nchw.reshape(B, C, h, 2, 2, w, 2, 2).permute(0, 2, 5, 3, 6, 4, 7, 1).reshape(B * h * w * 4, 4 * C)
    tok = tok.contiguous().to(dt_in)
    out = ops.tok_to_nchw(tok, B, h, w, C, level, dt_out)
    assert torch.equal(out, nchw.to(dt_out))
    back = torch.zeros_like(tok)
    ops.nchw_to_tok(out, back, B, h, w, C, level)
    assert torch.equal(back, tok.to(dt_out).to(dt_in))
    acc = torch.ones(tok.shape, device="cuda")
    ops.nchw_to_tok(out, acc, B, h, w, C, level, accumulate=True)
    assert torch.equal(acc, 1 + tok.to(dt_out).float())


def test_maxpool_fwd_bwd():
    from mtp_b200 import ops
    B, h, w, C = 2, 6, 10, 64
    torch.manual_seed(3)
    x = torch.randn(B, h, w, C, device="cuda")
    y = ops.maxpool2_tok_fwd(x.reshape(-1, C), B, h, w, C)
    xr = x.permute(0, 3, 1, 2).clone().requires_grad_(True)
    ref = torch.nn.functional.max_pool2d(xr, 2, 2)
    assert torch.equal(y.reshape(B, h // 2, w // 2, C).permute(0, 3, 1, 2), ref)
    dy = torch.randn_like(ref)
    ref.backward(dy)
    dx = torch.ones(B * h * w, C, device="cuda")
    ops.maxpool2_tok_bwd(x.reshape(-1, C), dy.permute(0, 2, 3, 1).reshape(-1, C).contiguous(), dx, B, h, w, C)
    assert (dx.reshape(B, h, w, C).permute(0, 3, 1, 2) - 1 - xr.grad).abs().max().item() < 1e-6
