"""RVSA window attention and dense rel-pos attention kernels vs the oracle's pieces (fp32, same bf16-rounded qkv).

The tensor-core kernels stage K~ / V~ / P / dS as bf16 MMA operands (fp32 accumulate); tolerances reflect that."""
import pytest
import torch

from oracle import rvsa_oracle as O

pytestmark = pytest.mark.gpu


def _attn_params(C, nH, gh, seed, big_sampling=1.0):
    g = torch.Generator().manual_seed(seed)
    P = {}
    pre = "a."
    P[pre + "qkv.weight"] = torch.randn(3 * C, C, generator=g) * 0.05
    P[pre + "qkv.bias"] = torch.randn(3 * C, generator=g) * 0.1
    P[pre + "proj.weight"] = torch.eye(C)
    P[pre + "proj.bias"] = torch.zeros(C)
    P[pre + "rel_pos_h"] = torch.randn(13, 64, generator=g) * 0.1
    P[pre + "rel_pos_w"] = torch.randn(13, 64, generator=g) * 0.1
    P[pre + "relative_position_bias_table"] = torch.randn(169, nH, generator=g) * 0.2
    for name, oc in (("sampling_offsets", 2 * nH), ("sampling_scales", 2 * nH), ("sampling_angles", nH)):
        P[pre + name + ".2.weight"] = torch.randn(oc, C, 1, 1, generator=g) * 0.05 * big_sampling
        P[pre + name + ".2.bias"] = torch.randn(oc, generator=g) * 0.1 * big_sampling
    P[pre + "full_attn_rel_pos_h"] = torch.randn(2 * gh - 1, 64, generator=g) * 0.1
    P[pre + "full_attn_rel_pos_w"] = torch.randn(2 * gh - 1, 64, generator=g) * 0.1
    return P


def _bf16_round(t):
    return t.to(torch.bfloat16).float()


@pytest.mark.parametrize("grid,B,nH,big", [(14, 2, 2, 1.0), (10, 3, 2, 1.0), (20, 1, 4, 1.0), (14, 2, 2, 6.0), (32, 1, 2, 3.0), (10, 1, 16, 2.0), (20, 1, 16, 1.0),
                                            (14, 2, 12, 1.0)])
def test_rvsa_attention_vs_oracle(grid, B, nH, big):
    """Covers no-pad (14), pad 2+2 (10->14), pad 0+1 (20->21), 1+2 (32->35) and large offsets that push taps out of the image."""
    from mtp_b200 import ops
    C = nH * 64
    P = _attn_params(C, nH, grid, seed=grid + nH, big_sampling=big)
    torch.manual_seed(0)
    xn = _bf16_round(torch.randn(B, grid * grid, C))
    # the oracle runs on the same bf16-rounded qkv the kernel reads
    qkv = _bf16_round(xn @ P["a.qkv.weight"].t() + P["a.qkv.bias"])
    want = _oracle_rvsa_from_qkv(xn, qkv, P, "a.", grid, grid, nH)
    dev = "cuda"
    d = lambda t: t.to(dev).contiguous()
    params, pooled = ops.rvsa_sampling_fwd(d(xn.reshape(-1, C)).to(torch.bfloat16),
                                           d(P["a.sampling_offsets.2.weight"].reshape(-1, C)), d(P["a.sampling_offsets.2.bias"]),
                                           d(P["a.sampling_scales.2.weight"].reshape(-1, C)), d(P["a.sampling_scales.2.bias"]),
                                           d(P["a.sampling_angles.2.weight"].reshape(-1, C)), d(P["a.sampling_angles.2.bias"]),
                                           B, grid, grid, nH)
    out, lse = ops.rvsa_attn_fwd(d(qkv.reshape(-1, 3 * C)).to(torch.bfloat16), params, d(P["a.rel_pos_h"]), d(P["a.rel_pos_w"]),
                                 d(P["a.relative_position_bias_table"]), B, grid, grid, nH)
    torch.cuda.synchronize()
    got = out.float().cpu().reshape(B, grid * grid, C)
    err = (got - want).abs()
    # bf16 output rounding (2^-9 relative) on top of the bf16 K~ / V~ / P staging: bounded relative to the output scale
    assert err.max().item() < 1e-2 * max(2.0, want.abs().max().item())
    assert (err.norm() / want.norm()).item() < 4e-3
    # sampling params against the oracle
    pt, pb, pl, pr = O.window_padding(grid, grid)
    xg = torch.nn.functional.pad(xn.reshape(B, grid, grid, C), (0, 0, pl, pr, pt, pb))
    ox, oy, sx, sy, th = O.sampling_params(xg, P, "a.", nH, grid, grid)
    ref = torch.stack([ox, oy, sx, sy, th], -1).reshape(-1, nH, 5)
    assert (params[:, :, :5].cpu() - ref).abs().max().item() < 1e-4


def _oracle_rvsa_from_qkv(xn, qkv, P, pre, h, w, nH):
    """oracle.rvsa_attention with the qkv projection replaced by a given (bf16-rounded) qkv tensor."""
    B, N, C = xn.shape
    import torch.nn.functional as F
    hd = C // nH
    pt, pb, pl, pr = O.window_padding(h, w)
    Hq, Wq = h + pt + pb, w + pl + pr
    nh, nw = Hq // 7, Wq // 7
    xg = F.pad(xn.reshape(B, h, w, C), (0, 0, pl, pr, pt, pb))
    ox, oy, sx, sy, th = O.sampling_params(xg, P, pre, nH, h, w)
    px, py = O.rvsa_coords(ox, oy, sx, sy, th, Hq, Wq)
    q4 = F.pad(qkv.reshape(B, h, w, 3, nH, hd), (0, 0, 0, 0, 0, 0, pl, pr, pt, pb))
    q, k, v = (q4[:, :, :, i].permute(0, 3, 1, 2, 4) for i in range(3))
    G = B * nH
    pxf, pyf = px.reshape(G, Hq * Wq), py.reshape(G, Hq * Wq)
    ks = O.bilinear_gather(k.reshape(G, Hq, Wq, hd), pxf, pyf).reshape(B, nH, nh, 7, nw, 7, hd)
    vs = O.bilinear_gather(v.reshape(G, Hq, Wq, hd), pxf, pyf).reshape(B, nH, nh, 7, nw, 7, hd)
    win = lambda t: t.permute(0, 2, 4, 1, 3, 5, 6).reshape(B, nh, nw, nH, 49, hd)
    qw, kw, vw = win(q.reshape(B, nH, nh, 7, nw, 7, hd)), win(ks), win(vs)
    S = hd ** -0.5 * (qw @ kw.transpose(-1, -2))
    iy = torch.arange(7).repeat_interleave(7)
    ix = torch.arange(7).repeat(7)
    Rh = P[pre + "rel_pos_h"][(iy[:, None] - torch.arange(7)[None, :]) + 6]
    Rw = P[pre + "rel_pos_w"][(ix[:, None] - torch.arange(7)[None, :]) + 6]
    S = S + torch.einsum("...qc,qkc->...qk", qw, Rh)[..., :, iy] + torch.einsum("...qc,qkc->...qk", qw, Rw)[..., :, ix]
    idx = (iy[:, None] - iy[None, :] + 6) * 13 + (ix[:, None] - ix[None, :] + 6)
    S = S + P[pre + "relative_position_bias_table"][idx.reshape(-1)].reshape(49, 49, nH).permute(2, 0, 1)
    Oo = torch.softmax(S, -1) @ vw
    Oo = Oo.reshape(B, nh, nw, nH, 7, 7, hd).permute(0, 1, 4, 2, 5, 3, 6).reshape(B, Hq, Wq, C)
    return Oo[:, pt:pt + h, pl:pl + w].reshape(B, N, C)


@pytest.mark.parametrize("grid,B,nH,rel", [(14, 2, 2, True), (10, 3, 3, True), (14, 1, 2, False), (16, 1, 2, True), (13, 1, 2, True), (32, 1, 2, True),
                                           (32, 2, 16, True), (40, 1, 2, True), (19, 1, 2, False), (64, 1, 2, True)])
def test_full_attention_vs_oracle(grid, B, nH, rel):
    from mtp_b200 import ops
    C = nH * 64
    P = _attn_params(C, nH, grid, seed=7 + grid)
    torch.manual_seed(1)
    N = grid * grid
    qkv = _bf16_round(torch.randn(B, N, 3 * C))
    q, k, v = (qkv.reshape(B, N, 3, nH, 64).permute(2, 0, 3, 1, 4)[i] for i in range(3))
    q = q * 0.125
    S = q @ k.transpose(-1, -2)
    if rel:
        ty = torch.arange(grid).repeat_interleave(grid)
        tx = torch.arange(grid).repeat(grid)
        Rh = P["a.full_attn_rel_pos_h"][(ty[:, None] - torch.arange(grid)[None, :]) + grid - 1]
        Rw = P["a.full_attn_rel_pos_w"][(tx[:, None] - torch.arange(grid)[None, :]) + grid - 1]
        S = S + torch.einsum("bnqc,qkc->bnqk", q, Rh)[..., :, ty] + torch.einsum("bnqc,qkc->bnqk", q, Rw)[..., :, tx]
    want = (torch.softmax(S, -1) @ v).transpose(1, 2).reshape(B, N, C)
    want_lse = torch.logsumexp(S, -1)
    d = lambda t: t.cuda().contiguous()
    out, lse = ops.full_attn_fwd(d(qkv.reshape(-1, 3 * C)).to(torch.bfloat16),
                                 d(P["a.full_attn_rel_pos_h"]) if rel else None, d(P["a.full_attn_rel_pos_w"]) if rel else None,
                                 B, grid, grid, nH)
    torch.cuda.synchronize()
    got = out.float().cpu().reshape(B, N, C)
    assert ((got - want).norm() / want.norm()).item() < 4e-3
    assert (lse.cpu() - want_lse).abs().max().item() < 1e-3


def _close(got, want, rel, name):
    err = (got - want).norm().item() / max(want.norm().item(), 1e-20)
    assert err < rel, f"{name}: rel-L2 {err:.3e} >= {rel}"
    return err


@pytest.mark.parametrize("grid,B,nH,big", [(14, 2, 2, 1.0), (10, 2, 2, 2.0), (20, 1, 4, 1.0), (14, 1, 2, 6.0), (10, 1, 16, 2.0), (20, 1, 16, 1.0)])
def test_rvsa_backward_vs_oracle_autograd(grid, B, nH, big):
    from mtp_b200 import ops
    C = nH * 64
    P = _attn_params(C, nH, grid, seed=100 + grid + nH, big_sampling=big)
    torch.manual_seed(5)
    N = grid * grid
    xn = _bf16_round(torch.randn(B, N, C)).requires_grad_(True)
    qkv = _bf16_round(torch.randn(B, N, 3 * C)).requires_grad_(True)
    leaf_names = ["a.rel_pos_h", "a.rel_pos_w", "a.relative_position_bias_table"] + \
        [f"a.sampling_{k}.2.{w}" for k in ("offsets", "scales", "angles") for w in ("weight", "bias")]
    Pl = dict(P)
    for k in leaf_names:
        Pl[k] = P[k].clone().requires_grad_(True)
    out = _oracle_rvsa_from_qkv(xn, qkv, Pl, "a.", grid, grid, nH)
    gout = _bf16_round(torch.randn(B, N, C))
    (out * gout).sum().backward()

    d = lambda t: t.detach().cuda().contiguous()
    params, pooled = ops.rvsa_sampling_fwd(d(xn.reshape(-1, C)).to(torch.bfloat16),
                                           d(P["a.sampling_offsets.2.weight"].reshape(-1, C)), d(P["a.sampling_offsets.2.bias"]),
                                           d(P["a.sampling_scales.2.weight"].reshape(-1, C)), d(P["a.sampling_scales.2.bias"]),
                                           d(P["a.sampling_angles.2.weight"].reshape(-1, C)), d(P["a.sampling_angles.2.bias"]),
                                           B, grid, grid, nH)
    qkv_d = d(qkv.reshape(-1, 3 * C)).to(torch.bfloat16)
    rel_h, rel_w, table = d(P["a.rel_pos_h"]), d(P["a.rel_pos_w"]), d(P["a.relative_position_bias_table"])
    _, lse = ops.rvsa_attn_fwd(qkv_d, params, rel_h, rel_w, table, B, grid, grid, nH)
    z = lambda t: torch.zeros_like(t)
    g_rel_h, g_rel_w, g_table = z(rel_h), z(rel_w), z(table)
    g_qkv_bias = torch.zeros(3 * C, device="cuda")
    dqkv, dparams = ops.rvsa_attn_bwd(qkv_d, params, rel_h, rel_w, table, lse, d(gout.reshape(-1, C)).to(torch.bfloat16),
                                      g_rel_h, g_rel_w, g_table, B, grid, grid, nH, d_qkv_bias=g_qkv_bias)
    # fused qkv-bias gradient = column sums of the stored bf16 dqkv; unused parameter slots are zero
    ref_bias = dqkv.float().sum(0)
    assert (g_qkv_bias - ref_bias).abs().max().item() <= 1e-3 * max(1.0, ref_bias.abs().max().item())
    assert (dparams[..., 5:] == 0).all() and (params[..., 5:] == 0).all()
    w = {k: d(P[f"a.sampling_{k}.2.weight"].reshape(-1, C)) for k in ("offsets", "scales", "angles")}
    gw = {k: z(v) for k, v in w.items()}
    gb = {k: torch.zeros(v.shape[0], device="cuda") for k, v in w.items()}
    dyn = torch.zeros(B * N, C, device="cuda", dtype=torch.bfloat16)
    ops.rvsa_sampling_bwd(dparams, pooled, w["offsets"], w["scales"], w["angles"], gw["offsets"], gb["offsets"], gw["scales"], gb["scales"],
                          gw["angles"], gb["angles"], dyn, B, grid, grid, nH)
    torch.cuda.synchronize()
    errs = {}
    errs["dqkv"] = _close(dqkv.float().cpu().reshape(B, N, 3 * C), qkv.grad, 1.5e-2, "dqkv")
    errs["rel_h"] = _close(g_rel_h.cpu(), Pl["a.rel_pos_h"].grad, 1e-2, "d rel_pos_h")
    errs["rel_w"] = _close(g_rel_w.cpu(), Pl["a.rel_pos_w"].grad, 1e-2, "d rel_pos_w")
    errs["table"] = _close(g_table.cpu(), Pl["a.relative_position_bias_table"].grad, 1e-2, "d bias table")
    for k in ("offsets", "scales", "angles"):
        errs[k + "_w"] = _close(gw[k].cpu(), Pl[f"a.sampling_{k}.2.weight"].grad.reshape(-1, C), 1.5e-2, f"d sampling_{k}.weight")
        errs[k + "_b"] = _close(gb[k].cpu(), Pl[f"a.sampling_{k}.2.bias"].grad, 1.5e-2, f"d sampling_{k}.bias")
    errs["dyn"] = _close(dyn.float().cpu().reshape(B, N, C), xn.grad, 1.5e-2, "pooled-path grad")
    print("rvsa bwd", grid, {k: "%.1e" % v for k, v in errs.items()})
    # the fused form (one launch for the four follow-up kernels) computes the same thing: compare against the separate launches above
    f_rel_h, f_rel_w, f_table, f_bias = z(rel_h), z(rel_w), z(table), torch.zeros(3 * C, device="cuda")
    fw = {k: z(v) for k, v in w.items()}
    fb = {k: torch.zeros(v.shape[0], device="cuda") for k, v in w.items()}
    dqkv_f, dpooled_f = ops.rvsa_attn_bwd_fused(qkv_d, params, rel_h, rel_w, table, lse, d(gout.reshape(-1, C)).to(torch.bfloat16),
                                                f_rel_h, f_rel_w, f_table, f_bias, pooled, w["offsets"], w["scales"], w["angles"],
                                                fw["offsets"], fb["offsets"], fw["scales"], fb["scales"], fw["angles"], fb["angles"], B, grid, grid, nH)
    dpooled_s = ops.rvsa_sampling_bwd(dparams, pooled, w["offsets"], w["scales"], w["angles"], z(w["offsets"]), z(gb["offsets"]), z(w["scales"]),
                                      z(gb["scales"]), z(w["angles"]), z(gb["angles"]), None, B, grid, grid, nH).clone()
    torch.cuda.synchronize()

    def same(a, b, what, tol=2e-3):          # the fp32 scatter sums are accumulated with atomics: order-dependent in the last bits
        assert (a.float() - b.float()).norm().item() <= tol * max(b.float().norm().item(), 1e-6), what
    same(dqkv_f, dqkv, "fused dqkv")
    same(f_rel_h, g_rel_h, "fused d rel_pos_h"); same(f_rel_w, g_rel_w, "fused d rel_pos_w"); same(f_table, g_table, "fused d table")
    same(f_bias, g_qkv_bias, "fused d qkv bias")
    for k in ("offsets", "scales", "angles"):
        same(fw[k], gw[k], f"fused d sampling_{k}.weight"); same(fb[k], gb[k], f"fused d sampling_{k}.bias")
    same(dpooled_f, dpooled_s, "fused dpooled")


@pytest.mark.parametrize("grid,B,nH,rel", [(14, 2, 2, True), (10, 1, 3, True), (14, 1, 2, False), (16, 1, 2, True), (13, 1, 2, True), (20, 1, 2, True), (14, 2, 16, True),
                                           (32, 1, 2, True), (32, 2, 4, True), (40, 1, 2, True), (19, 1, 2, False), (64, 1, 1, True)])
def test_full_attention_backward_vs_autograd(grid, B, nH, rel):
    from mtp_b200 import ops
    C = nH * 64
    P = _attn_params(C, nH, grid, seed=200 + grid)
    torch.manual_seed(6)
    N = grid * grid
    qkv = _bf16_round(torch.randn(B, N, 3 * C)).requires_grad_(True)
    Rh_p = P["a.full_attn_rel_pos_h"].clone().requires_grad_(True)
    Rw_p = P["a.full_attn_rel_pos_w"].clone().requires_grad_(True)
    q, k, v = (qkv.reshape(B, N, 3, nH, 64).permute(2, 0, 3, 1, 4)[i] for i in range(3))
    q = q * 0.125
    S = q @ k.transpose(-1, -2)
    if rel:
        ty = torch.arange(grid).repeat_interleave(grid)
        tx = torch.arange(grid).repeat(grid)
        Rh = Rh_p[(ty[:, None] - torch.arange(grid)[None, :]) + grid - 1]
        Rw = Rw_p[(tx[:, None] - torch.arange(grid)[None, :]) + grid - 1]
        S = S + torch.einsum("bnqc,qkc->bnqk", q, Rh)[..., :, ty] + torch.einsum("bnqc,qkc->bnqk", q, Rw)[..., :, tx]
    out = (torch.softmax(S, -1) @ v).transpose(1, 2).reshape(B, N, C)
    gout = _bf16_round(torch.randn(B, N, C))
    (out * gout).sum().backward()
    d = lambda t: t.detach().cuda().contiguous()
    qkv_d = d(qkv.reshape(-1, 3 * C)).to(torch.bfloat16)
    rh, rw = (d(Rh_p), d(Rw_p)) if rel else (None, None)
    o, lse = ops.full_attn_fwd(qkv_d, rh, rw, B, grid, grid, nH)
    g_rh = torch.zeros_like(rh) if rel else None
    g_rw = torch.zeros_like(rw) if rel else None
    dqkv = ops.full_attn_bwd(qkv_d, rh, rw, lse, o, d(gout.reshape(-1, C)).to(torch.bfloat16), g_rh, g_rw, B, grid, grid, nH)
    torch.cuda.synchronize()
    e = _close(dqkv.float().cpu().reshape(B, N, 3 * C), qkv.grad, 1.5e-2, "dqkv")
    if rel:
        _close(g_rh.cpu(), Rh_p.grad, 1e-2, "d full_attn_rel_pos_h")
        _close(g_rw.cpu(), Rw_p.grad, 1e-2, "d full_attn_rel_pos_w")
    print("full bwd", grid, "%.1e" % e)
