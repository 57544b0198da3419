"""Why the full-depth parity tests compare error LEVELS and not elements -- demonstrated on the CPU oracle.

A network whose activations are stored in bf16 is not a continuous function of its inputs: a perturbation d << ulp in front of a
rounding flips a fraction d/ulp of the roundings, each by a whole ulp, i.e. it comes out as an rms change of sqrt(d * ulp).  Fed
through a few more roundings (sqrt(sqrt(d ulp) ulp), ...) it converges to the bf16 rounding error itself.  Consequence: two correct
bf16 implementations of the backbone (the CUDA path and any emulation of it) whose pre-rounding values differ by 1e-6 are, after
12 blocks, as far from each other as either is from the fp32 result.  The smooth fp32 network amplifies the same perturbation
by a factor ~3 only."""
import dataclasses

import torch

from oracle import rvsa_oracle as O


def _vit_b_state(seed=0):
    from mtp_b200 import ViT_Win_RVSA_V3_WSZ7
    torch.manual_seed(seed)
    m = ViT_Win_RVSA_V3_WSZ7(img_size=224, patch_size=16, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4, qkv_bias=True,
                             use_abs_pos_emb=True, interval=3, out_indices=[3, 5, 7, 11], drop_path_rate=0.1, use_rel_pos_bias=True)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "rel_pos" in n:
                p.normal_(0, 0.02)
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def _rel(a, b):
    return float((a - b).norm() / b.norm())


def test_bf16_rounding_decorrelates_two_correct_implementations(monkeypatch):
    sd = _vit_b_state()
    x = torch.randn(1, 3, 224, 224)
    cfg = O.vit_b_config(224)
    cfg16 = dataclasses.replace(cfg, emulate_bf16=True)
    eps = 1e-6
    with torch.no_grad():
        ref32 = O.backbone_forward(sd, cfg, x)
        emu = O.backbone_forward(sd, cfg16, x)
        # (a) the same perturbation WITHOUT rounding: the network is smooth
        monkeypatch.setattr(O, "_ste_bf16", lambda t: t * (1 + eps * torch.randn_like(t)))
        smooth = O.backbone_forward(sd, cfg16, x)
        # (b) a second bf16 implementation: identical arithmetic, values perturbed by 1e-6 before every rounding
        monkeypatch.setattr(O, "_ste_bf16", lambda t: (t * (1 + eps * torch.randn_like(t))).to(torch.bfloat16).to(t.dtype))
        emu2 = O.backbone_forward(sd, cfg16, x)
    bf16_err = [_rel(a, b) for a, b in zip(emu, ref32)]            # ~ 2.8e-3 .. 5.4e-3
    smooth_amp = [_rel(a, b) / eps for a, b in zip(smooth, ref32)]   # ~ 1.7 .. 3.3
    pair = [_rel(a, b) for a, b in zip(emu2, emu)]                  # ~ 1.4e-3 .. 4.3e-3
    print("bf16 emulation vs fp32:", bf16_err, "\nfp32 amplification of a 1e-6 perturbation:", smooth_amp,
          "\ntwo bf16 emulations 1e-6 apart:", pair)
    assert max(smooth_amp) < 10.0
    for p, e in zip(pair, bf16_err):
        assert p > 0.25 * e, (p, e)          # decorrelated to the order of the bf16 error itself, not ~1e-6
        assert p < 1.5 * e, (p, e)
    # both emulations sit at the same distance from the fp32 result: the error LEVEL is reproducible, the elements are not
    for a, b, r in zip(emu2, emu, ref32):
        assert abs(_rel(a, r) / _rel(b, r) - 1.0) < 0.15


def test_emulated_backward_runs_and_is_close_to_fp32_at_small_depth():
    """emulate_bf16_grad is a perturbation of the fp32 gradients of the size of bf16 rounding, not a different function."""
    from tests.helpers import load_golden
    g = load_golden("tiny160")
    cfg = dataclasses.replace(g["cfg"], emulate_bf16=True, emulate_bf16_grad=True)
    grads = []
    for c in (g["cfg"], cfg):
        P = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in g["sd"].items()}
        O.synthetic_loss(O.backbone_forward(P, c, g["x"])).backward()
        grads.append({k: v.grad for k, v in P.items() if v.is_floating_point() and v.grad is not None})
    assert grads[0].keys() == grads[1].keys()
    errs = {k: _rel(grads[1][k], grads[0][k]) for k in grads[0] if "sampling" not in k and "norm1" not in k}
    worst = max(errs.values())
    print("tiny160: emulated-bf16 backward vs fp32, worst non-coordinate gradient:", worst)
    assert 1e-4 < worst < 6e-2
