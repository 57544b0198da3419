"""Fused pretrain step (forward + backward + clip + AdamW on flat buffers) vs autograd + torch.optim.AdamW."""
import copy

import pytest
import torch

from oracle import rvsa_oracle as O
from tests.helpers import load_golden
from tests.test_backbone_gpu import build_module

pytestmark = pytest.mark.gpu


def _torch_reference_step(m, x, lr, wd, max_norm):
    decay, no_decay = [], []
    for n, p in m.named_parameters():
        (no_decay if (p.dim() == 1 or n.endswith(".bias") or "pos_embed" in n) else decay).append(p)
    opt = torch.optim.AdamW([{"params": decay, "weight_decay": wd}, {"params": no_decay, "weight_decay": 0.0}], lr=lr, betas=(0.9, 0.999), eps=1e-8)
    loss = O.synthetic_loss(m(x))
    loss.backward()
    torch.nn.utils.clip_grad_norm_([p for p in m.parameters() if p.grad is not None], max_norm)
    opt.step()
    return loss


@pytest.mark.parametrize("graph", [False, True])
def test_fused_step_matches_torch_adamw(graph):
    from mtp_b200.trainer import PretrainStep
    g = load_golden("tiny160")
    m1 = build_module("tiny160")
    m1.load_state_dict(g["sd"])
    m1 = m1.cuda().eval()
    m2 = copy.deepcopy(m1)
    p0 = {n: p.detach().clone() for n, p in m1.named_parameters()}
    x = g["x"].cuda()
    lr, wd, max_norm = 1e-3, 0.05, 0.05        # small max_norm so the clip is active
    l2 = _torch_reference_step(m2, x, lr, wd, max_norm)
    tr = PretrainStep(m1, lr=lr, weight_decay=wd, max_norm=max_norm, use_cuda_graph=graph)
    l1 = tr.step(x)
    torch.cuda.synchronize()
    l1v = l1.item()          # graph mode returns the same static loss tensor every step: read it now
    assert abs(l1v - l2.item()) < 1e-4
    worst = 0.0
    for (n, p1), (_, p2) in zip(m1.named_parameters(), m2.named_parameters()):
        d1, d2 = p1.detach() - p0[n], p2.detach() - p0[n]
        if d2.abs().max().item() == 0:
            assert d1.abs().max().item() == 0, n
            continue
        # Adam's first step is lr * sign(g) (up to eps): compare element-wise where the gradient is not ~0
        err = (d1 - d2).abs().max().item() / lr
        worst = max(worst, err)
        assert err < 0.35, (n, err)
        assert ((d1 - d2).norm() / d2.norm()).item() < 0.05, n
    # the bf16 mirror that the next forward reads is in sync with the fp32 masters
    assert torch.equal(tr.flat_p16, tr.flat_p.to(torch.bfloat16))
    # second step runs (graph replay path) and changes the loss
    l3 = tr.step(x)
    assert l3.item() != l1v


def test_fused_stand_in_heads_match_torch():
    """mtp_sqloss_fwd_bwd (one pass per map) = the torch formulation of the stand-in objective."""
    from mtp_b200.trainer import synthetic_heads
    torch.manual_seed(3)
    feats = [torch.randn(2, 64, s, s, device="cuda").to(torch.bfloat16) for s in (56, 28, 14, 7)]
    loss, grads = synthetic_heads(feats)
    ref_loss = sum((f.float() ** 2).mean() * 0.5 for f in feats)
    assert abs(loss.item() - ref_loss.item()) <= 1e-4 * abs(ref_loss.item())
    for f, g in zip(feats, grads):
        assert torch.equal(g, (f.float() * (1.0 / f.numel())).to(torch.bfloat16))


def test_fused_gradient_norm_matches_full_pass():
    """World size 1: squared gradient norm = wgrad-epilogue partial sums + small-region pass (mtp_epilogue.sumsq)."""
    from mtp_b200 import _lib as L, ops
    from mtp_b200.trainer import PretrainStep
    g = load_golden("tiny224")
    m = build_module("tiny224")
    m.load_state_dict(g["sd"], strict=True)
    m = m.cuda().train()
    x = g["x"].cuda().to(torch.bfloat16)
    tr = PretrainStep(m, lr=1e-4, max_norm=5.0)
    assert tr.fused_norm
    tr._forward_backward(x, lambda r: None)
    assert tr.fused_norm, "the one-time check against the full reduction disabled the fused norm"
    L.call("mtp_sumsq_f32", tr.flat_g.data_ptr(), tr.small_end, tr.state.data_ptr() + 4, ops._stream())
    full = float((tr.flat_g.double() ** 2).sum().item())
    assert abs(float(tr.state[1].item()) - full) <= 1e-4 * full


def _reference_steps(m, x, lr, wd, max_norm, steps):
    decay, no_decay = [], []
    for n, p in m.named_parameters():
        (no_decay if (p.dim() == 1 or n.endswith(".bias") or "pos_embed" in n) else decay).append(p)
    opt = torch.optim.AdamW([{"params": decay, "weight_decay": wd}, {"params": no_decay, "weight_decay": 0.0}], lr=lr, betas=(0.9, 0.999), eps=1e-8)
    losses = []
    for _ in range(steps):
        opt.zero_grad(set_to_none=True)
        loss = O.synthetic_loss(m(x))
        loss.backward()
        torch.nn.utils.clip_grad_norm_([p for p in m.parameters() if p.grad is not None], max_norm)
        opt.step()
        losses.append(loss.item())
    return losses


@pytest.mark.parametrize("graph", [False, True])
def test_three_steps_track_torch_adamw(graph):
    """Every forward after step 1 must see the UPDATED weights and biases (incl. the ConvTranspose2d weights / biases of the pyramid,
    which the GEMM reads in a re-packed layout): losses of steps 2 and 3 and the final feature maps follow autograd + torch AdamW."""
    from mtp_b200.trainer import PretrainStep
    g = load_golden("tiny160")
    m1 = build_module("tiny160")
    m1.load_state_dict(g["sd"])
    m1 = m1.cuda().eval()
    m2 = copy.deepcopy(m1)
    p0 = {n: p.detach().clone() for n, p in m1.named_parameters()}
    x = g["x"].cuda()
    lr, wd, max_norm = 2e-3, 0.05, 1.0        # a large lr so three steps move the loss visibly
    ref_losses = _reference_steps(m2, x, lr, wd, max_norm, 3)
    tr = PretrainStep(m1, lr=lr, weight_decay=wd, max_norm=max_norm, use_cuda_graph=graph)
    losses = []
    for _ in range(3):
        l = tr.step(x)
        torch.cuda.synchronize()
        losses.append(l.item())
    print("losses", losses, "reference", ref_losses)
    assert abs(ref_losses[2] - ref_losses[0]) > 20 * 2e-3 * abs(ref_losses[0]), "the reference loss must move for this test to mean anything"
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) < 2e-3 * abs(b), (losses, ref_losses)
    for n in ("fpn1.0.bias", "fpn1.3.bias", "fpn2.0.bias", "fpn1.0.weight", "fpn2.0.weight", "blocks.0.mlp.fc1.weight", "pos_embed"):
        d1 = dict(m1.named_parameters())[n].detach() - p0[n]
        d2 = dict(m2.named_parameters())[n].detach() - p0[n]
        assert d2.abs().max().item() > 0
        assert ((d1 - d2).norm() / d2.norm()).item() < 0.08, n
    with torch.no_grad():
        o1, o2 = m1(x), m2(x)
    for a, b in zip(o1, o2):      # three Adam steps at lr 2e-3 move every weight by ~10 %: sign flips of ~0 gradients show up here
        assert ((a - b).norm() / b.norm()).item() < 2.5e-2


def test_trainer_state_dict_round_trip_and_weight_reload():
    """Optimizer state survives state_dict()/load_state_dict(); loading model weights AFTER the trainer exists refreshes the bf16
    mirror the kernels read."""
    from mtp_b200.trainer import PretrainStep
    g = load_golden("tiny160")
    x = g["x"].cuda()

    def fresh():
        m = build_module("tiny160")
        m.load_state_dict(g["sd"])
        return m.cuda().eval()
    ma = fresh()
    ta = PretrainStep(ma, lr=1e-3, max_norm=1.0)
    for _ in range(2):
        ta.step(x)
    opt_sd = {k: (v if not isinstance(v, dict) else {n: t.clone() for n, t in v.items()}) for k, v in ta.state_dict().items()}
    model_sd = {k: v.detach().clone() for k, v in ma.state_dict().items()}
    la = ta.step(x).item()
    # resume in a new trainer built on a freshly initialised module
    mb = fresh()
    tb = PretrainStep(mb, lr=1e-3, max_norm=1.0)
    mb.load_state_dict(model_sd)                       # goes through the parameters' .data views -> flat_p; mirror refreshed by the hook
    assert torch.equal(tb.flat_p16, tb.flat_p.to(torch.bfloat16))
    tb.load_state_dict(opt_sd)
    lb = tb.step(x).item()
    assert abs(la - lb) <= 1e-5 * abs(la), (la, lb)
    for (n, pa), (_, pb) in zip(ma.named_parameters(), mb.named_parameters()):
        # one Adam step moves a weight by <= lr = 1e-3; the two runs differ by the summation order of the atomics in the gradient kernels,
        # which Adam amplifies where g ~ 0 (m / (sqrt(v) + eps)): allow 5 % of a step on the worst element, 0.5 % in the norm
        assert float((pa - pb).abs().max()) <= 5e-5, n
        assert float((pa - pb).norm()) <= 5e-6 * pa.numel() ** 0.5, n


def test_three_task_fan_out_matches_one_head_per_slice():
    """models.py:327-335: one encoder call on cat(x1,x2,x3), maps split b1|b2|b3, three heads, cotangents in the matching slices."""
    from mtp_b200.trainer import ThreeTaskHeads, synthetic_heads
    torch.manual_seed(4)
    feats = [torch.randn(8, 64, s, s, device="cuda").to(torch.bfloat16) for s in (56, 28, 14, 7)]
    heads = ThreeTaskHeads((3, 3, 2))
    loss, grads = heads(feats)
    ref = 0.0
    lo = 0
    for b, w in zip((3, 3, 2), (1.0, 0.5, 2.0)):
        for f, g in zip(feats, grads):
            fk = f[lo:lo + b].float()
            ref = ref + (fk ** 2).mean() * 0.5 * w
            assert torch.equal(g[lo:lo + b], (fk * (w / fk.numel())).to(torch.bfloat16))
        lo += b
    assert abs(loss.item() - ref.item()) <= 1e-4 * abs(ref.item())


def test_three_stream_uint8_step_from_host():
    """The e2e shape of the reference step: three pinned uint8 batches -> one encoder call with the fused preprocessor -> three heads."""
    from mtp_b200.preprocess import ImagePreprocess
    from mtp_b200.trainer import PretrainStep, ThreeTaskHeads
    g = load_golden("tiny160")
    m = build_module("tiny160")
    m.load_state_dict(g["sd"])
    m = m.cuda().train()
    m.input_preprocess = ImagePreprocess(out_dtype=torch.bfloat16)
    gen = torch.Generator().manual_seed(9)
    parts = [torch.randint(0, 256, (b, 3, 160, 160), dtype=torch.uint8, generator=gen).pin_memory() for b in (3, 3, 2)]
    for graph in (False, True):
        tr = PretrainStep(m, lr=1e-4, max_norm=5.0, heads=ThreeTaskHeads((3, 3, 2)), use_cuda_graph=graph)
        l0 = tr.step_from_host(parts)
        l1 = tr.step_from_host(parts)
        assert l0 == l0 and l1 == l1 and l0 > 0 and l1 != l0
