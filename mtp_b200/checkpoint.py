"""Checkpoint loading for the drop-in backbone: the host-side logic of ``init_weights``.

``variant='pretrain'`` follows ViT_Win_RVSA_V3_WSZ7.init_weights ([V]:693-778): prefix stripping, ``num_extra_tokens = 1``
(MAE checkpoints carry a cls token), bicubic pos-embed resize, ``load_state_dict(strict=False)``.
``variant='finetune'`` follows the registered twins (e.g. RS_Tasks_Finetune/Semantic_Segmentation/mmseg/models/backbones/
vit_rvsa_mtp.py:684-807): additionally bicubic-resizes ``full_attn_rel_pos_{h,w}`` to the new ``2*Hp-1`` and drops the cls
token only if the checkpoint has one.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


def _reinit(m):
    if isinstance(m, nn.Linear):
        nn.init.trunc_normal_(m.weight, std=.02, a=-2.0, b=2.0)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
    elif isinstance(m, nn.LayerNorm):
        nn.init.constant_(m.bias, 0)
        nn.init.constant_(m.weight, 1.0)


def convert_state_dict(model, checkpoint, variant="pretrain", verbose=False):
    """Pure function: checkpoint object -> state_dict adapted to ``model`` (no loading)."""
    if "state_dict" in checkpoint:
        sd = checkpoint["state_dict"]
    elif "model" in checkpoint:
        sd = checkpoint["model"]
    else:
        sd = checkpoint
    sd = dict(sd)
    keys = list(sd.keys())
    if keys and keys[0].startswith("module."):
        sd = {k[7:]: v for k, v in sd.items()}
    if sd and sorted(sd.keys())[0].startswith("encoder"):
        sd = {k.replace("encoder.", ""): v for k, v in sd.items() if k.startswith("encoder.")}
    if variant == "pretrain" and model.in_chans != 3:
        for k in list(sd.keys()):
            if "patch_embed.proj" in k:
                del sd[k]
    if variant == "finetune":
        target = None
        for name, p in model.named_parameters():
            if "attn.full_attn_rel_pos_h" in name:
                target = tuple(p.shape)
                break
        if target is not None:
            for k in list(sd.keys()):
                if "full_attn_rel_pos_h" in k or "full_attn_rel_pos_w" in k:
                    old = sd[k]
                    new = F.interpolate(old.reshape(1, 1, *old.shape).float(), size=target, mode="bicubic", align_corners=False)
                    sd[k] = new.squeeze()
    if "pos_embed" in sd:
        pe = sd["pos_embed"]
        emb = pe.shape[-1]
        H, W = model.patch_embed.patch_shape
        num_patches = model.patch_embed.num_patches
        extra = 1 if variant == "pretrain" else (1 if "cls_token" in sd else 0)
        orig = int((pe.shape[-2] - extra) ** 0.5)
        new = int(num_patches ** 0.5)
        if orig != new:
            if verbose:
                print("Position interpolate from %dx%d to %dx%d" % (orig, orig, H, W))
            tok = pe[:, extra:].reshape(-1, orig, orig, emb).permute(0, 3, 1, 2)
            tok = F.interpolate(tok.float(), size=(H, W), mode="bicubic", align_corners=False)
            sd["pos_embed"] = tok.permute(0, 2, 3, 1).flatten(1, 2)
        else:
            sd["pos_embed"] = pe[:, extra:]
    return sd


def init_weights(model, pretrained=None, variant="pretrain", verbose=False):
    pretrained = pretrained or model.pretrained
    if isinstance(pretrained, str):
        model.apply(_reinit)
        ckpt = torch.load(pretrained, map_location="cpu")
        sd = convert_state_dict(model, ckpt, variant, verbose)
        msg = model.load_state_dict(sd, strict=False)
        model._engine_state.invalidate()
        if verbose:
            print(msg)
        return msg
    if pretrained is None:
        model.apply(_reinit)
        model._engine_state.invalidate()
        return None
    raise TypeError("pretrained must be a str or None")
