"""fp32-class forward mode (precision="fp32x3": hi | lo bf16 word pairs, three-pass tcgen05 GEMMs, fp32 attention) against the fp32
CPU oracle.  BASELINE.json north star: "matches the reference forward within 1e-3 rel" -- asserted here at the real depths (ViT-B config 1,
ViT-L headline) and on the padded-window grids; the same inputs through the bf16 fast path are printed next to it."""
import pytest
import torch

from oracle import rvsa_oracle as O
from tests.helpers import build_backbone

pytestmark = pytest.mark.gpu

NORTH_STAR = 1e-3


def _rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / b.norm())


def _run(embed, depth, heads, interval, out_idx, img, B, seed, cfg):
    m, sd = build_backbone(embed, depth, heads, interval, out_idx, seed=seed, img_size=img)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if ".sampling_" in n:
                p.mul_(3.0)                   # move the taps off the pixel centres
        sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    x = torch.randn(B, 3, img, img, generator=torch.Generator().manual_seed(seed))
    with torch.no_grad():
        ref = O.backbone_forward(sd, cfg, x)
    m = m.cuda().eval()
    with torch.no_grad():
        fast = m(x.cuda())
        m.precision = "fp32x3"
        prec = m(x.cuda())
    torch.cuda.synchronize()
    e_fast = [_rel(a, b) for a, b in zip(fast, ref)]
    e_prec = [_rel(a, b) for a, b in zip(prec, ref)]
    return e_fast, e_prec, m, x


def test_gemm_hilo_three_pass_matches_fp64():
    from mtp_b200 import ops, _lib as L
    from mtp_b200.precise import split_hilo
    torch.manual_seed(0)
    M, N, K = 300, 200, 256
    A = torch.randn(M, K, device="cuda")
    Bw = torch.randn(N, K, device="cuda") * 0.05
    bias = torch.randn(N, device="cuda")
    out = torch.empty(M, 2 * N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(split_hilo(A), split_hilo(Bw), M, N, K, out, bias=bias, hilo=True, lda=2 * K, ldb=2 * K, ldo=2 * N, out_lo=N)
    torch.cuda.synchronize()
    got = out[:, :N].double() + out[:, N:].double()
    want = A.double() @ Bw.double().t() + bias.double()
    err = float((got - want).norm() / want.norm())
    plain = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(A.to(torch.bfloat16), Bw.to(torch.bfloat16), M, N, K, plain, bias=bias)
    err_plain = float((plain.double() - want).norm() / want.norm())
    print(f"hilo GEMM rel-L2 {err:.2e} (plain bf16 path {err_plain:.2e})")
    assert err < 2e-5 and err_plain > 50 * err


@pytest.mark.parametrize("img,B", [(160, 2), (224, 2), (320, 1)])
def test_precise_forward_small_padded_grids(img, B):
    cfg = O.OracleConfig(img_size=img, embed_dim=256, depth=4, num_heads=4, interval=2, out_indices=(0, 1, 2, 3))
    e_fast, e_prec, _, _ = _run(256, 4, 4, 2, [0, 1, 2, 3], img, B, 40 + img, cfg)
    print(f"img {img}: bf16 path {['%.2e' % e for e in e_fast]}  fp32x3 {['%.2e' % e for e in e_prec]}")
    assert max(e_prec) < 1e-4, e_prec


def test_precise_forward_vit_b_config1():
    e_fast, e_prec, m, x = _run(768, 12, 12, 3, [3, 5, 7, 11], 224, 1, 0, O.vit_b_config(224))
    print(f"ViT-B 1x224: bf16 path {['%.2e' % e for e in e_fast]}  fp32x3 {['%.2e' % e for e in e_prec]}")
    assert max(e_prec) < NORTH_STAR, e_prec
    with pytest.raises(RuntimeError):                 # forward-only: asking for gradients fails loudly
        m(x.cuda())


def test_precise_forward_vit_l_headline():
    e_fast, e_prec, _, _ = _run(1024, 24, 16, 6, [7, 11, 15, 23], 224, 2, 1, O.vit_l_config(224))
    print(f"ViT-L 2x224: bf16 path {['%.2e' % e for e in e_fast]}  fp32x3 {['%.2e' % e for e in e_prec]}")
    assert max(e_prec) < NORTH_STAR, e_prec
