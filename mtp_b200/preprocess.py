"""Image normalisation of ``MTP_DataPreprocessor`` (Multi-Task_Pretrain/preprocessing.py:145-187 -> mmengine ``ImgDataPreprocessor``)
for the fused uint8 input path of the backbone: the channel flip, ``.float()`` and ``(x - mean) / std`` run inside the patch gather
kernel (``mtp_patchify_u8``), so a batch travels host -> device as uint8 (a quarter of the fp32 bytes) and is never materialised as a
normalised fp32 image.  Only the image arithmetic is mirrored; the data-sample bookkeeping of the mmdet preprocessor (box types,
mask / seg-map padding, batch augments) concerns the decoders and stays with them.

    m = vit_l_rvsa(args)
    m.input_preprocess = ImagePreprocess()            # models.py:37-41 defaults
    feats = m(batch_u8)                                # uint8 (B, 3, H, W)  (layout="hwc": (B, H, W, 3))
"""
from dataclasses import dataclass, field
from typing import Sequence

import torch


@dataclass
class ImagePreprocess:
    mean: Sequence[float] = (123.675, 116.28, 103.53)      # models.py:38
    std: Sequence[float] = (58.395, 57.12, 57.375)         # models.py:39
    bgr_to_rgb: bool = True                                 # models.py:40
    rgb_to_bgr: bool = False
    layout: str = "chw"                                     # "chw": what PackDetInputs hands over; "hwc": a decoded image in memory
    out_dtype: torch.dtype = torch.float32                  # dtype of the returned NCHW maps (the reference's decoders are fp32)

    def __post_init__(self):
        if self.bgr_to_rgb and self.rgb_to_bgr:
            raise ValueError("`bgr2rgb` and `rgb2bgr` cannot be set to True at the same time")     # mmengine's own check
        if self.layout not in ("chw", "hwc"):
            raise ValueError("layout must be 'chw' or 'hwc'")
        if len(self.mean) != len(self.std):
            raise ValueError("mean and std need one entry per channel")

    @property
    def flip_channels(self) -> bool:
        return self.bgr_to_rgb or self.rgb_to_bgr

    def reference(self, x_u8: torch.Tensor) -> torch.Tensor:
        """The same arithmetic in plain torch (fp32 NCHW out) -- what mmengine computes; used by the tests as the restatement."""
        x = x_u8.permute(0, 3, 1, 2) if self.layout == "hwc" else x_u8
        if self.flip_channels:
            x = x.flip(1)
        x = x.float()
        mean = torch.tensor(self.mean, dtype=torch.float32, device=x.device).view(1, -1, 1, 1)
        std = torch.tensor(self.std, dtype=torch.float32, device=x.device).view(1, -1, 1, 1)
        return (x - mean) / std
