"""Drop-in ViT + RVSA backbone for MTP: same constructor kwargs, state_dict keys, outputs and factories as the
reference class ``ViT_Win_RVSA_V3_WSZ7`` (Multi-Task_Pretrain/backbone/vit_win_rvsa_v3_wsz7.py:587-865, "[V]") and its
registered finetune twins ``RVSA_MTP`` / ``RVSA_MTP_branches`` (RS_Tasks_Finetune/*/…/backbones/vit_rvsa_mtp*.py),
with every per-block computation running in the hand-written sm_100a kernels of libmtp_b200.so.

The nn.Module tree below only *holds parameters* under the reference's names (so checkpoints, layer-decay optimizer
constructors and DDP see exactly the reference layout); none of the holder modules' own ``forward`` methods is ever
called.  ``forward`` runs :mod:`mtp_b200.engine`.  There is no PyTorch/CPU fallback: without a CUDA device or without
the built library, ``forward`` raises.
"""
from __future__ import annotations

import math
from functools import partial
from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from . import engine as _engine

WS = 7


def _trunc_normal_(t, std=0.02):
    return nn.init.trunc_normal_(t, mean=0.0, std=std, a=-2.0, b=2.0)      # timm.trunc_normal_ defaults ([V]:22)


class PatchEmbed(nn.Module):
    """Parameter holder for the 16x16/16 patch projection ([V]:515-540)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768):
        super().__init__()
        img_size = (img_size, img_size) if isinstance(img_size, int) else tuple(img_size)
        patch_size = (patch_size, patch_size) if isinstance(patch_size, int) else tuple(patch_size)
        self.patch_shape = (img_size[0] // patch_size[0], img_size[1] // patch_size[1])
        self.num_patches = self.patch_shape[0] * self.patch_shape[1]
        self.img_size, self.patch_size = img_size, patch_size
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.fc2 = nn.Linear(hidden_features, in_features)


class Attention(nn.Module):
    """Dense attention block parameters ([V]:65-88).  ``rel_pos=False`` mirrors the mmdet/mmrotate twins, which comment
    the decomposed rel-pos tables out."""

    def __init__(self, dim, num_heads, qkv_bias, window_size, rel_pos=True):
        super().__init__()
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.window_size = window_size
        if rel_pos:
            self.full_attn_rel_pos_h = nn.Parameter(torch.zeros(2 * window_size[0] - 1, head_dim))
            self.full_attn_rel_pos_w = nn.Parameter(torch.zeros(2 * window_size[1] - 1, head_dim))
        self.proj = nn.Linear(dim, dim)


class RotatedVariedSizeWindowAttention(nn.Module):
    """RVSA block parameters ([V]:195-285): qkv/proj, decomposed rel-pos (13 x hd), bias table (169 x nH) + index buffer,
    and the three pooled 1x1-conv sampling heads (offsets 2nH, scales 2nH, angles nH)."""

    def __init__(self, dim, num_heads, qkv_bias):
        super().__init__()
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.rel_pos_h = nn.Parameter(torch.zeros(2 * WS - 1, head_dim))
        self.rel_pos_w = nn.Parameter(torch.zeros(2 * WS - 1, head_dim))

        def head(out_ch):
            return nn.Sequential(nn.AvgPool2d(kernel_size=WS, stride=WS), nn.LeakyReLU(), nn.Conv2d(dim, out_ch, kernel_size=1, stride=1))
        self.sampling_offsets = head(num_heads * 2)
        self.sampling_scales = head(num_heads * 2)
        self.sampling_angles = head(num_heads)
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * WS - 1) * (2 * WS - 1), num_heads))
        iy = torch.arange(WS).repeat_interleave(WS)
        ix = torch.arange(WS).repeat(WS)
        index = (iy[:, None] - iy[None, :] + WS - 1) * (2 * WS - 1) + (ix[:, None] - ix[None, :] + WS - 1)   # [V]:272-282
        self.register_buffer("relative_position_index", index)
        _trunc_normal_(self.relative_position_bias_table, std=0.02)


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio, qkv_bias, drop_path, norm_layer, window, patch_shape, full_rel_pos):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.window = window
        if window:
            self.attn = RotatedVariedSizeWindowAttention(dim, num_heads, qkv_bias)
        else:
            self.attn = Attention(dim, num_heads, qkv_bias, patch_shape, rel_pos=full_rel_pos)
        self.drop_path_prob = float(drop_path)
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))


class Norm2d(nn.Module):
    def __init__(self, embed_dim):
        super().__init__()
        self.ln = nn.LayerNorm(embed_dim, eps=1e-6)


class ViT_Win_RVSA_V3_WSZ7(nn.Module):
    """B200-native drop-in for [V]:587-817.  Extra keyword-only switches select the finetune-twin behaviours
    (SURVEY.md §2.2); the defaults reproduce the pretrain class."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=80, embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop_rate=0., attn_drop_rate=0.,
                 drop_path_rate=0., hybrid_backbone=None, norm_layer=None, init_values=None, use_checkpoint=False,
                 use_abs_pos_emb=False, use_rel_pos_bias=False, use_shared_rel_pos_bias=False,
                 out_indices=[11], interval=3, pretrained=None, restart_regression=True, *,
                 full_attn_rel_pos=True, feature_mode="multi", apply_fpn=True, final_norm=True, return_tuple=False,
                 frozen_stages=-1, precision="bf16"):
        super().__init__()
        if hybrid_backbone is not None:
            raise NotImplementedError("hybrid_backbone (HybridEmbed, [V]:542-574) is unused by every MTP config")
        if patch_size != 16:
            raise NotImplementedError("only patch_size=16 is used by the MTP configs ([V]:640-654)")
        if init_values is not None:
            raise NotImplementedError("init_values (gamma_1/gamma_2 layer scale, [V]:500-504) is unused by MTP")
        if drop_rate != 0. or attn_drop_rate != 0.:
            raise NotImplementedError("drop_rate / attn_drop_rate are 0 in every MTP config")
        if qk_scale is not None:
            raise NotImplementedError("qk_scale override is unused by MTP (scale = head_dim ** -0.5)")
        if embed_dim % num_heads != 0 or embed_dim // num_heads != 64 or embed_dim % 128 != 0:
            raise NotImplementedError("kernels are built for head_dim 64 (ViT-B 768/12, ViT-L 1024/16)")
        assert feature_mode in ("multi", "last_norm")
        norm_layer = norm_layer or partial(nn.LayerNorm, eps=1e-6)
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        self.in_chans = in_chans
        self.num_heads = num_heads
        self.depth = depth
        self.patch_embed = PatchEmbed(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim)
        num_patches = self.patch_embed.num_patches
        self.out_indices = list(out_indices)
        if feature_mode == "multi" and len(self.out_indices) != 4:
            raise ValueError("out_indices must name four blocks (one per pyramid level, [V]:804-811)")
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches, embed_dim)) if use_abs_pos_emb else None
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, depth, device='cpu')]          # [V]:619
        self.use_rel_pos_bias = use_rel_pos_bias
        self.use_checkpoint = use_checkpoint
        self.blocks = nn.ModuleList([
            Block(embed_dim, num_heads, mlp_ratio, qkv_bias, dpr[i], norm_layer, window=((i + 1) % interval != 0),
                  patch_shape=self.patch_embed.patch_shape, full_rel_pos=full_attn_rel_pos)
            for i in range(depth)])                                                  # [V]:625-631
        self.interval = interval
        if self.pos_embed is not None:
            _trunc_normal_(self.pos_embed, std=.02)
        if final_norm:
            self.norm = norm_layer(embed_dim)                                        # allocated but unused in [V] ([V]:638)
        self.fpn1 = nn.Sequential(nn.ConvTranspose2d(embed_dim, embed_dim, kernel_size=2, stride=2), Norm2d(embed_dim),
                                  nn.GELU(), nn.ConvTranspose2d(embed_dim, embed_dim, kernel_size=2, stride=2))
        self.fpn2 = nn.Sequential(nn.ConvTranspose2d(embed_dim, embed_dim, kernel_size=2, stride=2))
        self.fpn3 = nn.Identity()
        self.fpn4 = nn.MaxPool2d(kernel_size=2, stride=2)
        self.apply(self._init_weights)
        self.fix_init_weight()
        self.pretrained = pretrained
        self.out_channels = [embed_dim, embed_dim, embed_dim, embed_dim]
        self.full_attn_rel_pos = full_attn_rel_pos
        self.feature_mode = feature_mode
        self.apply_fpn = apply_fpn
        self.return_tuple = return_tuple
        self.frozen_stages = frozen_stages
        if precision not in ("bf16", "fp32x3"):
            raise ValueError("precision must be 'bf16' (training / fast path) or 'fp32x3' (fp32-class forward, inference / verification)")
        self.precision = precision        # may also be switched on an existing module: m.precision = "fp32x3"
        self._engine_state = _engine.EngineState()
        self.input_preprocess = None     # optional mtp_b200.preprocess.ImagePreprocess: accept uint8 images, normalise in the patch gather
        # weights replaced wholesale: drop cached bf16 copies / refresh always-current mirrors (engine.EngineState.invalidate)
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._engine_state.invalidate())

    def invalidate_weight_cache(self):
        """Call after editing parameters through ``p.data`` (EMA / weight averaging / manual rescale): such writes do not bump the
        autograd version counter the bf16 weight cache is keyed on."""
        self._engine_state.invalidate()

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)           # .to() / .cuda() / .float(): storages move
        if hasattr(self, "_engine_state"):
            self._engine_state.invalidate()
        return out

    # ---- init (mirrors [V]:676-691) -------------------------------------------------------------------------
    def fix_init_weight(self):
        for layer_id, layer in enumerate(self.blocks):
            layer.attn.proj.weight.data.div_(math.sqrt(2.0 * (layer_id + 1)))
            layer.mlp.fc2.weight.data.div_(math.sqrt(2.0 * (layer_id + 1)))

    @staticmethod
    def _init_weights(m):
        if isinstance(m, nn.Linear):
            _trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def init_weights(self, pretrained=None):
        from .checkpoint import init_weights
        return init_weights(self, pretrained)

    def get_num_layers(self):
        return len(self.blocks)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {'pos_embed', 'cls_token'}

    # ---- forward -----------------------------------------------------------------------------------------------
    def forward_features(self, x):
        return _engine.backbone_apply(self, x)

    def forward(self, x):
        feats = self.forward_features(x)
        return tuple(feats) if self.return_tuple else feats


def vit_b_rvsa(args, inchannels=3):
    """[V]:819-841.  ``args`` needs ``image_size`` and ``use_ckpt`` ('True'/'False'), as main_pretrain.py passes them."""
    return ViT_Win_RVSA_V3_WSZ7(img_size=args.image_size, in_chans=inchannels, patch_size=16, drop_path_rate=0.1,
                                out_indices=[3, 5, 7, 11], embed_dim=768, depth=12, num_heads=12, mlp_ratio=4, qkv_bias=True,
                                qk_scale=None, drop_rate=0., attn_drop_rate=0., use_checkpoint=(args.use_ckpt == 'True'),
                                use_abs_pos_emb=True, interval=3, use_rel_pos_bias=True)


def vit_l_rvsa(args, inchannels=3):
    """[V]:843-865."""
    return ViT_Win_RVSA_V3_WSZ7(img_size=args.image_size, in_chans=inchannels, patch_size=16, drop_path_rate=0.1,
                                out_indices=[7, 11, 15, 23], embed_dim=1024, depth=24, num_heads=16, mlp_ratio=4, qkv_bias=True,
                                qk_scale=None, drop_rate=0., attn_drop_rate=0., use_checkpoint=(args.use_ckpt == 'True'),
                                use_abs_pos_emb=True, interval=6, use_rel_pos_bias=True)
