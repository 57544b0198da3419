"""Registration of the drop-in classes under the reference's registry keys.

The finetune overlays register the backbone as ``RVSA_MTP`` / ``RVSA_MTP_branches`` in each OpenMMLab toolkit's
``MODELS`` registry (mmcv-1.x mmrotate 0.3.4: ``ROTATED_BACKBONES``) — e.g. mmseg ``vit_rvsa_mtp.py:577-578``.  The nine
twins differ from the pretrain class only in four switches; ``FLAVOURS`` records them per toolkit.
``register_all()`` registers into every toolkit registry that is importable; ``MODELS`` below is a local registry with
the same ``register_module() / build(cfg)`` surface for environments without the toolkits (this image has none).
"""
from __future__ import annotations

import importlib
from typing import Dict

from .backbone import ViT_Win_RVSA_V3_WSZ7
from .checkpoint import init_weights as _init_weights

# toolkit -> class name -> constructor switches (diffed against the nine registered files)
FLAVOURS: Dict[str, Dict[str, dict]] = {
    "mmseg": {"RVSA_MTP": dict(final_norm=False, return_tuple=True)},
    "mmpretrain": {"RVSA_MTP": dict(apply_fpn=False, return_tuple=True)},
    "opencd": {"RVSA_MTP": dict(apply_fpn=False, return_tuple=True)},
    "mmdet": {"RVSA_MTP": dict(full_attn_rel_pos=False, feature_mode="last_norm", return_tuple=True),
              "RVSA_MTP_branches": dict(full_attn_rel_pos=False, return_tuple=True)},
    "mmrotate": {"RVSA_MTP": dict(full_attn_rel_pos=False, feature_mode="last_norm", return_tuple=True),
                 "RVSA_MTP_branches": dict(full_attn_rel_pos=False, return_tuple=True)},
}
_REGISTRY_PATHS = {
    "mmseg": ("mmseg.registry", "MODELS"), "mmpretrain": ("mmpretrain.registry", "MODELS"), "opencd": ("opencd.registry", "MODELS"),
    "mmdet": ("mmdet.registry", "MODELS"), "mmrotate": ("mmrotate.registry", "MODELS"),
    "mmrotate0.3.4": ("mmrotate.models.builder", "ROTATED_BACKBONES"),
}


def make_class(toolkit: str, name: str):
    """Build the twin class for one toolkit: same kwargs as the reference file, finetune-style ``init_weights()``."""
    switches = FLAVOURS["mmrotate" if toolkit == "mmrotate0.3.4" else toolkit][name]

    class _Twin(ViT_Win_RVSA_V3_WSZ7):
        def __init__(self, *args, **kwargs):
            for k, v in switches.items():
                kwargs.setdefault(k, v)
            super().__init__(*args, **kwargs)

        def init_weights(self, pretrained=None):           # the twins take no argument ([Vseg]:684); accept one for convenience
            return _init_weights(self, pretrained, variant="finetune")

    _Twin.__name__ = _Twin.__qualname__ = name
    _Twin.__doc__ = f"{name} as registered by the {toolkit} overlay of MTP; B200-native kernels underneath."
    return _Twin


class LocalRegistry:
    """Minimal stand-in for mmengine.Registry: ``register_module`` decorator and ``build(dict(type=...))``."""

    def __init__(self, name):
        self.name, self._modules = name, {}

    def register_module(self, name=None, module=None, force=False):
        def _reg(cls):
            key = name or cls.__name__
            if key in self._modules and not force:
                raise KeyError(f"{key} is already registered in {self.name}")
            self._modules[key] = cls
            return cls
        return _reg(module) if module is not None else _reg

    def get(self, key):
        return self._modules.get(key)

    def build(self, cfg: dict):
        cfg = dict(cfg)
        typ = cfg.pop("type")
        cls = self._modules.get(typ) if isinstance(typ, str) else typ
        if cls is None:
            raise KeyError(f"{typ} is not in the {self.name} registry")
        return cls(**cfg)

    def __contains__(self, key):
        return key in self._modules


MODELS = LocalRegistry("mtp_b200.MODELS")
RVSA_MTP = MODELS.register_module(module=make_class("mmseg", "RVSA_MTP"))                 # default flavour: semantic segmentation
RVSA_MTP_branches = MODELS.register_module(module=make_class("mmdet", "RVSA_MTP_branches"))
for _tk, _names in FLAVOURS.items():
    for _n in _names:
        MODELS.register_module(name=f"{_tk}.{_n}", module=make_class(_tk, _n))


def register_all(force=True):
    """Register into every importable OpenMMLab registry under the reference's keys.  Returns {toolkit: [names]}."""
    done = {}
    for tk, (modpath, attr) in _REGISTRY_PATHS.items():
        try:
            reg = getattr(importlib.import_module(modpath), attr)
        except Exception:
            continue
        base = "mmrotate" if tk == "mmrotate0.3.4" else tk
        for name in FLAVOURS[base]:
            reg.register_module(name=name, module=make_class(tk, name), force=force)
            done.setdefault(tk, []).append(name)
    return done
