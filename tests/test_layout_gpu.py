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
@pytest.mark.parametrize("shape", [(2, 5, 7, 96), (2, 6, 14, 128), (1, 3, 10, 192)])      # the 2nd / 3rd take the 16-byte bf16 kernels where Wo % 8 == 0
def test_tok_nchw_roundtrip(level, dt_in, dt_out, shape):
    from mtp_b200 import ops
    B, h, w, C = shape
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


@pytest.mark.parametrize("layout,flip", [("chw", True), ("hwc", True), ("chw", False)])
def test_patchify_u8_fused_preprocess_bit_exact(layout, flip):
    """MTP_DataPreprocessor (BGR->RGB, mean/std, preprocessing.py:145-187 / mmengine ImgDataPreprocessor) folded into the patch gather:
    bit-exact against the torch restatement of that arithmetic followed by the fp32 patchify."""
    from mtp_b200 import ops
    from mtp_b200.preprocess import ImagePreprocess
    g = torch.Generator().manual_seed(11)
    B, H, W = 3, 64, 96
    shape = (B, H, W, 3) if layout == "hwc" else (B, 3, H, W)
    x = torch.randint(0, 256, shape, dtype=torch.uint8, generator=g).cuda()
    pre = ImagePreprocess(bgr_to_rgb=flip, layout=layout)
    got = ops.patchify_u8(x, pre.mean, pre.std, pre.flip_channels, layout == "hwc")
    want = ops.patchify(pre.reference(x).contiguous())
    assert torch.equal(got, want)


def test_backbone_accepts_uint8_with_input_preprocess():
    from mtp_b200.preprocess import ImagePreprocess
    from tests.test_backbone_gpu import build_module
    from tests.helpers import load_golden
    g = load_golden("tiny160")
    m = build_module("tiny160")
    m.load_state_dict(g["sd"])
    m = m.cuda().eval()
    x = torch.randint(0, 256, (2, 3, 160, 160), dtype=torch.uint8, generator=torch.Generator().manual_seed(5)).cuda()
    with pytest.raises(TypeError):
        m(x)
    m.input_preprocess = ImagePreprocess()
    with torch.no_grad():
        a = m(x)
        b = m(m.input_preprocess.reference(x).contiguous())
    for u, v in zip(a, b):
        assert u.dtype == torch.float32 and torch.equal(u, v)
