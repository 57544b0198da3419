"""Pin the oracle against the LIVE reference module.  Runs only where /root/reference is mounted (build container)."""
import pytest
import torch

from oracle import rvsa_oracle as O
from oracle import ref_import

pytestmark = pytest.mark.skipif(not ref_import.reference_available(), reason="/root/reference not mounted")


def _kw(img, C, depth, nH, interval, oi, dpr=0.1):
    return dict(img_size=img, patch_size=16, embed_dim=C, depth=depth, num_heads=nH, mlp_ratio=4, qkv_bias=True,
                use_abs_pos_emb=True, interval=interval, out_indices=list(oi), drop_path_rate=dpr, use_rel_pos_bias=True)


def test_config1_vit_b_224_forward():
    """BASELINE.json configs[0]: ViT-B backbone forward, 1x3x224x224 on CPU."""
    m = ref_import.build_reference(_kw(224, 768, 12, 12, 3, (3, 5, 7, 11)), seed=0)
    cfg = O.vit_b_config(224)
    torch.manual_seed(0)
    x = torch.randn(1, 3, 224, 224)
    with torch.no_grad():
        r = m(x)
        o = O.backbone_forward(m.state_dict(), cfg, x)
    for a, b in zip(r, o):
        assert a.shape == b.shape and float((a - b).abs().max()) < 1e-5


@pytest.mark.parametrize("img", [160, 320, 512])        # pad cases Hp 10->14, 20->21, 32->35
def test_padded_grids(img):
    m = ref_import.build_reference(_kw(img, 128, 4, 2, 2, (0, 1, 2, 3)), seed=3)
    cfg = O.OracleConfig(img_size=img, embed_dim=128, depth=4, num_heads=2, interval=2, out_indices=(0, 1, 2, 3))
    torch.manual_seed(0)
    x = torch.randn(2, 3, img, img)
    with torch.no_grad():
        r = m(x)
        o = O.backbone_forward(m.state_dict(), cfg, x)
    for a, b in zip(r, o):
        assert float((a - b).abs().max()) < 2e-5


def test_backward_all_params():
    m = ref_import.build_reference(_kw(160, 128, 4, 2, 2, (0, 1, 2, 3)), seed=5)
    cfg = O.OracleConfig(img_size=160, embed_dim=128, depth=4, num_heads=2, interval=2, out_indices=(0, 1, 2, 3))
    torch.manual_seed(0)
    x = torch.randn(2, 3, 160, 160)
    O.synthetic_loss(m(x)).backward()
    P = {k: (v.detach().clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in m.state_dict().items()}
    O.synthetic_loss(O.backbone_forward(P, cfg, x)).backward()
    for k, p in m.named_parameters():
        if p.grad is None:
            assert k.startswith("norm.")
            continue
        err = float((P[k].grad - p.grad).norm() / p.grad.norm().clamp_min(1e-20))
        assert err < 2e-4, (k, err)


def test_drop_path_train_mode():
    """timm drop_path semantics: x * bernoulli(keep)/keep per sample, separate draws for attn and MLP branches."""
    kw = _kw(160, 128, 4, 2, 2, (0, 1, 2, 3), dpr=0.5)
    m = ref_import.build_reference(kw, seed=7).train()
    cfg = O.OracleConfig(img_size=160, embed_dim=128, depth=4, num_heads=2, interval=2, out_indices=(0, 1, 2, 3))
    B = 4
    torch.manual_seed(0)
    x = torch.randn(B, 3, 160, 160)
    rates = [r.item() for r in torch.linspace(0, 0.5, 4)]
    g = torch.Generator().manual_seed(11)
    keep = torch.ones(4, 2, B)
    for i, r in enumerate(rates):
        if r > 0:
            keep[i] = torch.bernoulli(torch.full((2, B), 1 - r), generator=g) / (1 - r)
    # block 0 has drop_prob 0 -> nn.Identity, no shim call
    ref_import.KEEP_QUEUE[:] = [keep[i, j] for i in range(1, 4) for j in range(2)]
    with torch.no_grad():
        r = m(x)
        o = O.backbone_forward(m.state_dict(), cfg, x, keep=keep)
    assert not ref_import.KEEP_QUEUE
    for a, b in zip(r, o):
        assert float((a - b).abs().max()) < 2e-5
