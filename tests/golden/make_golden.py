"""Generate the golden fixtures from the LIVE reference module (run in the build container only).

    python tests/golden/make_golden.py

For each tiny configuration the unmodified reference class ([V], imported by oracle/ref_import.py) is built with a
fixed seed, run forward (eval) and forward+backward on a fixed input, and the results are written to
``tests/golden/<name>.npz``:  the state_dict, the input, the four output maps, the loss
(oracle.synthetic_loss) and, per parameter, the gradient's L2 norm plus GRAD_SAMPLES evenly spaced entries
(full tensors for parameters with <= FULL_GRAD_MAX elements).  The fixtures travel to the GPU box, where
/root/reference does not exist.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_import import build_reference  # noqa: E402
from oracle.rvsa_oracle import synthetic_loss  # noqa: E402

GRAD_SAMPLES = 256
FULL_GRAD_MAX = 4096

CONFIGS = {
    # name: (img_size, embed_dim, depth, heads, interval, out_indices, batch)
    "tiny160": dict(img_size=160, embed_dim=128, depth=4, num_heads=2, interval=2, out_indices=[0, 1, 2, 3], batch=2),
    "tiny224": dict(img_size=224, embed_dim=128, depth=4, num_heads=2, interval=2, out_indices=[0, 1, 2, 3], batch=1),
}


def ref_kwargs(c):
    return dict(img_size=c["img_size"], patch_size=16, embed_dim=c["embed_dim"], depth=c["depth"],
                num_heads=c["num_heads"], mlp_ratio=4, qkv_bias=True, use_abs_pos_emb=True, interval=c["interval"],
                out_indices=list(c["out_indices"]), drop_path_rate=0.1, use_rel_pos_bias=True)


def sample_idx(n):
    return np.unique(np.linspace(0, n - 1, min(n, GRAD_SAMPLES)).astype(np.int64))


def main():
    for name, c in CONFIGS.items():
        model = build_reference(ref_kwargs(c), seed=0)
        torch.manual_seed(1234)
        x = torch.randn(c["batch"], 3, c["img_size"], c["img_size"])
        with torch.no_grad():
            outs = model(x)
        model.zero_grad()
        loss = synthetic_loss(model(x))
        loss.backward()
        blob = {"x": x.numpy(), "loss": np.float64(loss.item())}
        for i, o in enumerate(outs):
            blob[f"out{i}"] = o.numpy()
        for k, v in model.state_dict().items():
            blob["sd/" + k] = v.numpy()
        for k, p in model.named_parameters():
            if p.grad is None:          # encoder.norm never participates ([V]:638)
                continue
            g = p.grad.reshape(-1).double().numpy()
            blob["gnorm/" + k] = np.float64(np.sqrt((g * g).sum()))
            if g.size <= FULL_GRAD_MAX:
                blob["gfull/" + k] = p.grad.numpy()
            else:
                blob["gsamp/" + k] = g[sample_idx(g.size)].astype(np.float32)
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), name + ".npz")
        np.savez_compressed(path, **blob)
        print(name, "loss", loss.item(), "->", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
