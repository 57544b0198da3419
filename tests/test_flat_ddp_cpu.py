"""Host logic of the data-parallel path on CPU: flat layout invariants and the bucketed all-reduce with gloo (world 2)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import mtp_b200
from mtp_b200.flat import FlatLayout, is_gemm_weight

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _layout(depth=4):
    with torch.device("meta"):
        m = mtp_b200.ViT_Win_RVSA_V3_WSZ7(img_size=160, embed_dim=128, depth=depth, num_heads=2, qkv_bias=True, use_abs_pos_emb=True,
                                          interval=2, out_indices=list(range(depth))[-4:], drop_path_rate=0.1)
    return FlatLayout([(n, tuple(p.shape)) for n, p in m.named_parameters()]), m


def test_layout_is_aligned_disjoint_and_small_first():
    lay, m = _layout()
    spans = sorted(lay.span(n) for n in lay.order)
    assert spans[0][0] == 0 and spans[-1][1] == lay.total
    for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
        assert a1 == b0 and a0 % 64 == 0
    for n in lay.order:
        assert (lay.offsets[n] < lay.small_end) == (not is_gemm_weight(n))
    big = sum(p.numel() for n, p in m.named_parameters() if is_gemm_weight(n))
    assert big > 0.95 * sum(p.numel() for p in m.parameters())       # zeroing only the small region skips >95% of the bytes


@pytest.mark.parametrize("depth,bb", [(4, 1), (4, 3), (8, 4), (6, 4)])
def test_buckets_cover_everything_once(depth, bb):
    lay, _ = _layout(depth)
    ranges = list(lay.block_buckets(depth, bb).values()) + lay.tail_ranges(depth, bb)
    ranges.sort()
    assert ranges[0][0] == 0 and ranges[-1][1] == lay.total
    for (a0, a1), (b0, b1) in zip(ranges, ranges[1:]):
        assert a1 == b0
    # a bucket keyed by block b holds exactly the GEMM weights of blocks b .. b+bb-1
    for lo_blk, (lo, hi) in lay.block_buckets(depth, bb).items():
        inside = {n for n in lay.order if lo <= lay.offsets[n] < hi}
        blocks = {int(n.split(".")[1]) for n in inside}
        assert min(blocks) == lo_blk and all(n.startswith("blocks.") and is_gemm_weight(n) for n in inside)


def _worker(rank, world, port, total, ranges, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    flat = torch.arange(total, dtype=torch.float32) * (rank + 1)
    for lo, hi in ranges:                       # same call pattern as PretrainStep._allreduce_range, issue order = backward order
        dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM)
    flat *= 1.0 / world
    if rank == 0:
        q.put(flat.clone())
    dist.barrier()
    dist.destroy_process_group()


def test_bucketed_allreduce_gloo_world2():
    lay, _ = _layout(4)
    b = lay.block_buckets(4, 2)
    ranges = [b[k] for k in sorted(b, reverse=True)] + lay.tail_ranges(4, 2)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lay.total, ranges, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    want = torch.arange(lay.total, dtype=torch.float32) * 1.5       # mean of rank-scaled buffers
    assert torch.equal(got, want)


def test_pyramid_first_ordering_covers_the_flat_buffer_once():
    """Data-parallel reduction order: pyramid weights first, block buckets in backward order, small region + patch embed last -- every
    element of the flat gradient buffer is reduced exactly once."""
    import torch
    from mtp_b200 import ViT_Win_RVSA_V3_WSZ7
    from mtp_b200.flat import FlatLayout
    m = ViT_Win_RVSA_V3_WSZ7(img_size=160, embed_dim=128, depth=6, num_heads=2, interval=3, out_indices=[1, 2, 3, 5], use_abs_pos_emb=True,
                             qkv_bias=True, use_rel_pos_bias=True)
    lay = FlatLayout([(n, tuple(p.shape)) for n, p in m.named_parameters()])
    for bb in (1, 2, 4):
        fpn, rest = lay.split_tail(6, bb)
        buckets = list(lay.block_buckets(6, bb).values())
        seen = torch.zeros(lay.total, dtype=torch.int32)
        for lo, hi in fpn + buckets + rest:
            seen[lo:hi] += 1
        assert int(seen.min()) == 1 and int(seen.max()) == 1
        names = [n for n in lay.order if any(lo <= lay.offsets[n] < hi for lo, hi in fpn)]
        assert names and all(n.startswith("fpn") and n.endswith(".weight") for n in names), names
        assert (0, lay.small_end) in rest


def test_image_preprocess_reference_and_stream_split():
    import torch
    import bench
    from mtp_b200.preprocess import ImagePreprocess
    x = torch.randint(0, 256, (2, 3, 4, 5), dtype=torch.uint8)
    pre = ImagePreprocess()
    want = (x[:, [2, 1, 0]].float() - torch.tensor(pre.mean).view(1, 3, 1, 1)) / torch.tensor(pre.std).view(1, 3, 1, 1)
    assert torch.equal(pre.reference(x), want)
    hwc = ImagePreprocess(layout="hwc", bgr_to_rgb=False)
    assert torch.equal(hwc.reference(x.permute(0, 2, 3, 1).contiguous()),
                       (x.float() - torch.tensor(pre.mean).view(1, 3, 1, 1)) / torch.tensor(pre.std).view(1, 3, 1, 1))
    assert bench.split3(8) == (3, 3, 2) and bench.split3(32) == (11, 11, 10) and bench.split3(2) is None and sum(bench.split3(4)) == 4
    import pytest
    with pytest.raises(ValueError):
        ImagePreprocess(bgr_to_rgb=True, rgb_to_bgr=True)


def test_bench_reference_arm_line():
    """bench.py piece that needs no GPU: the JSON contract of the CPU reference arm (one timed step of the smallest configuration on the oracle port)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--config", "c2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "images/s" and line["value"] > 0 and line["steps"] == 1
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["e2e"]["h2d_bytes_per_step"] == 0
