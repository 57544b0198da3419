"""Shared test utilities: golden fixture loading and oracle plumbing (tests only)."""
import os

import numpy as np
import torch

from oracle import rvsa_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CFGS = {
    "tiny160": dict(img_size=160, embed_dim=128, depth=4, num_heads=2, interval=2, out_indices=(0, 1, 2, 3)),
    "tiny224": dict(img_size=224, embed_dim=128, depth=4, num_heads=2, interval=2, out_indices=(0, 1, 2, 3)),
}
GRAD_SAMPLES = 256


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {"x": torch.from_numpy(z["x"]), "loss": float(z["loss"]),
         "outs": [torch.from_numpy(z[f"out{i}"]) for i in range(4)],
         "sd": {}, "gnorm": {}, "gfull": {}, "gsamp": {}}
    for k in z.files:
        for grp in ("sd", "gnorm", "gfull", "gsamp"):
            if k.startswith(grp + "/"):
                v = z[k]
                g[grp][k[len(grp) + 1:]] = torch.from_numpy(v) if v.ndim else float(v)
    g["cfg"] = O.OracleConfig(**GOLDEN_CFGS[name])
    return g


def sample_idx(n):
    return np.unique(np.linspace(0, n - 1, min(n, GRAD_SAMPLES)).astype(np.int64))


def rel_l2(a, b):
    a = a.double().reshape(-1)
    b = b.double().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def check_grads_against_golden(named_grads, g, tol, skip=()):
    """named_grads: dict name -> tensor (cpu).  Compares against golden norms + samples / full tensors."""
    worst = 0.0
    for k, gn in g["gnorm"].items():
        if k in skip:
            continue
        assert k in named_grads and named_grads[k] is not None, f"missing grad {k}"
        got = named_grads[k].detach().double().cpu()
        denom = max(gn, 1e-12)
        if k in g["gfull"]:
            ref = g["gfull"][k].double()
            err = float((got - ref).norm()) / denom
        else:
            ref = g["gsamp"][k].double()
            flat = got.reshape(-1)
            idx = torch.from_numpy(sample_idx(flat.numel()))
            # sampled entries: scale the error to the whole-tensor norm via the sampling fraction
            err = float((flat[idx] - ref).norm()) / max(float(ref.norm()), 1e-12 * denom + 1e-30)
            nerr = abs(float(got.norm()) - gn) / denom
            err = max(err, nerr)
        worst = max(worst, err)
        assert err <= tol, f"grad {k}: rel err {err:.3e} > {tol:.1e}"
    return worst
