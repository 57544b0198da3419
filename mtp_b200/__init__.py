"""mtp_b200 — B200-native (sm_100a) ViT + RVSA backbone for MTP multi-task pretraining.

Public surface mirrors the reference's (Multi-Task_Pretrain/backbone/vit_win_rvsa_v3_wsz7.py and the registered finetune
twins): ``ViT_Win_RVSA_V3_WSZ7``, ``vit_b_rvsa``, ``vit_l_rvsa``, ``RVSA_MTP``, ``RVSA_MTP_branches``.
"""
from .backbone import ViT_Win_RVSA_V3_WSZ7, vit_b_rvsa, vit_l_rvsa  # noqa: F401
from .preprocess import ImagePreprocess  # noqa: F401
from .registry import MODELS, RVSA_MTP, RVSA_MTP_branches, register_all  # noqa: F401

__all__ = ["ViT_Win_RVSA_V3_WSZ7", "vit_b_rvsa", "vit_l_rvsa", "RVSA_MTP", "RVSA_MTP_branches", "MODELS", "register_all", "ImagePreprocess"]
