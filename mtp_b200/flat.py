"""Flat parameter / gradient layout shared by the fused optimizer and the gradient all-reduce (device-agnostic host logic).

Layout: [small parameters | GEMM weights], every tensor starting at a multiple of 64 elements.  "Small" parameters
(biases, LayerNorm, rel-pos tables, sampling heads, pos_embed) receive ACCUMULATED gradients and are zeroed each step;
GEMM weights are overwritten by the wgrad kernels.  The GEMM weights of the blocks are contiguous per block, so the
all-reduce buckets are plain slices that become final in backward order.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

_BIG_SUFFIXES = ("attn.qkv.weight", "attn.proj.weight", "mlp.fc1.weight", "mlp.fc2.weight")


def is_gemm_weight(name: str) -> bool:
    return name.endswith(_BIG_SUFFIXES) or name == "patch_embed.proj.weight" or \
        (name.startswith("fpn") and name.endswith(".weight") and ".ln." not in name)


def _pad64(n: int) -> int:
    return (n + 63) // 64 * 64


class FlatLayout:
    def __init__(self, named_shapes: List[Tuple[str, tuple]]):
        """named_shapes: [(name, shape)] in nn.Module.named_parameters() order."""
        def numel(s):
            n = 1
            for d in s:
                n *= d
            return n
        self.numel = {n: numel(s) for n, s in named_shapes}
        self.order = [n for n, _ in named_shapes if not is_gemm_weight(n)] + [n for n, _ in named_shapes if is_gemm_weight(n)]
        self.offsets: Dict[str, int] = {}
        total = 0
        self.small_end = 0
        for n in self.order:
            self.offsets[n] = total
            total += _pad64(self.numel[n])
            if not is_gemm_weight(n):
                self.small_end = total
        self.total = total

    def span(self, name):
        o = self.offsets[name]
        return o, o + _pad64(self.numel[name])

    def block_buckets(self, depth: int, bucket_blocks: int) -> Dict[int, Tuple[int, int]]:
        """{lowest block of the group: (lo, hi)} — the slice is final once that block's backward has run."""
        out = {}
        for hi_blk in range(depth - 1, -1, -bucket_blocks):
            lo_blk = max(0, hi_blk - bucket_blocks + 1)
            lo = self.offsets[f"blocks.{lo_blk}.attn.qkv.weight"]
            hi = self.span(f"blocks.{hi_blk}.mlp.fc2.weight")[1]
            out[lo_blk] = (lo, hi)
        return out

    def tail_ranges(self, depth: int, bucket_blocks: int) -> List[Tuple[int, int]]:
        """Everything not covered by the block buckets: the small region and the GEMM weights outside the blocks."""
        b = self.block_buckets(depth, bucket_blocks)
        lo_blocks = min(lo for lo, _ in b.values())
        hi_blocks = max(hi for _, hi in b.values())
        out = [(0, self.small_end)]
        if lo_blocks > self.small_end:
            out.append((self.small_end, lo_blocks))
        if hi_blocks < self.total:
            out.append((hi_blocks, self.total))
        return out

    def split_tail(self, depth: int, bucket_blocks: int):
        """(pyramid-weight ranges, the rest of the tail): the ConvTranspose2d weights lie behind the last block in the flat buffer and are
        final as soon as the pyramid tail has been differentiated (first, in the data-parallel trainer); the small accumulated region and
        the patch-embed weight are final only at the end of the backward."""
        tail = self.tail_ranges(depth, bucket_blocks)
        first_block = self.offsets["blocks.0.attn.qkv.weight"]
        fpn = [r for r in tail if r[0] >= self.small_end and r[0] > first_block]
        rest = [r for r in tail if r not in fpn]
        return fpn, rest
