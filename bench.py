#!/usr/bin/env python
"""Benchmark of the hot path: ViT-L + RVSA backbone pretrain step @224^2, bf16, synthetic data (BASELINE.json metric).

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference ...        # the reference's CPU path (oracle port) on the host cores, same metric

A "step" = one encoder call on an 8-image batch per GPU (global batch 64 at 8 GPUs, weak scaling) + synthetic heads +
full backward + gradient all-reduce (N > 1) + global-norm clip + fused AdamW.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

MODEL = dict(img_size=224, embed_dim=1024, depth=24, num_heads=16, interval=6, out_indices=[7, 11, 15, 23])
PER_GPU_BATCH = 8
FWD_GFLOP_PER_IMG = 130.20          # SURVEY.md Appendix C (oracle.algorithmic_gflop_per_image reproduces it)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH, help="images per GPU")
    ap.add_argument("--graph", type=int, default=1, help="capture the step into a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--bucket-blocks", type=int, default=2, help="transformer blocks per gradient all-reduce bucket (N > 1)")
    ap.add_argument("--comm-sms", type=int, default=16, help="SMs left to NCCL while the backward runs (N > 1)")
    ap.add_argument("--cpu-batch", type=int, default=1, help="images per CPU-reference step (bounded sample)")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """SM clock / throttle-reason sampling (NVML) on a background thread during the timed region."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake"}

    def __init__(self, gpu_index, period=0.05):
        self.gpu, self.period, self.sm, self.reasons, self.power = gpu_index, period, [], set(), []
        self._stop = threading.Event()
        self._thr = None
        self.max_sm = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu]) if vis and vis.split(",")[0].isdigit() else self.gpu
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nv = pynvml
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        except Exception:
            self._thr = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.period)

    def stop(self):
        if self._thr is None:
            return None
        self._stop.set()
        self._thr.join(timeout=1.0)
        if not self.sm:
            return None
        return {"sm_mhz": statistics.median(self.sm), "sm_max_mhz": self.max_sm, "reasons": sorted(self.reasons),
                "power_w_max": max(self.power) if self.power else None, "samples": len(self.sm)}


METRIC = "images/sec ViT-L+RVSA MTP step @224^2 bf16"
WORKLOAD = "ViT-L+RVSA backbone pretrain step @224^2: fwd + synthetic heads + bwd + grad all-reduce + clip + AdamW"


# ------------------------------------------------------------------------------------------------------ CPU reference arm
def usable_threads():
    """Thread count that actually maximises CPU throughput on this host.  Containers often expose every logical CPU of the
    machine while a cgroup quota / co-tenants make more than a fraction of them counter-productive (measured on the pool's
    B200 boxes: 128 logical CPUs, 16 threads are 60x faster than 128), so probe a GEMM instead of trusting cpu_count()."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    try:
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = max(1, min(n, int(float(q) / float(p) + 0.5)))
    except Exception:
        pass
    cands = sorted({c for c in (4, 8, 16, 32, 64, n) if c <= n})
    a = torch.randn(2048, 2048)
    times = {}
    for c in cands:
        torch.set_num_threads(c)
        a @ a
        t0 = time.perf_counter()
        for _ in range(4):
            a @ a
        times[c] = time.perf_counter() - t0
    tmin = min(times.values())
    return max(c for c in cands if times[c] <= 1.15 * tmin)      # the largest count that is still (nearly) the fastest


def cpu_reference_rate(batch, steps, warmup, threads=None):
    """The reference's CPU path (oracle port of [V], fp32) doing the same step: fwd + synthetic heads + bwd + AdamW."""
    from oracle import rvsa_oracle as O
    threads = threads or usable_threads()
    torch.set_num_threads(threads)
    cfg = O.vit_l_config(224)
    torch.manual_seed(0)
    from mtp_b200 import ViT_Win_RVSA_V3_WSZ7
    m = ViT_Win_RVSA_V3_WSZ7(img_size=224, patch_size=16, embed_dim=1024, depth=24, num_heads=16, mlp_ratio=4, qkv_bias=True,
                             use_abs_pos_emb=True, interval=6, out_indices=[7, 11, 15, 23], drop_path_rate=0.1, use_rel_pos_bias=True)
    P = {k: (v.detach().clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in m.state_dict().items()}
    del m
    params = [v for v in P.values() if v.is_floating_point()]
    opt = torch.optim.AdamW(params, lr=6e-5, weight_decay=0.05)
    x = torch.randn(batch, 3, 224, 224)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        loss = O.synthetic_loss(O.backbone_forward(P, cfg, x))
        loss.backward()
        torch.nn.utils.clip_grad_norm_([p for p in params if p.grad is not None], 5.0)
        opt.step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return batch / statistics.median(times), threads, sum(times)


def run_reference(args, rank):
    if rank != 0:
        return
    rate, threads, total = cpu_reference_rate(args.cpu_batch, args.steps, args.warmup)
    ms = 1000.0 * args.cpu_batch / rate
    sample = f"{args.cpu_batch} image(s) per step, ViT-L+RVSA @224 fwd+bwd+AdamW, fp32, {threads} threads"
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "implementation": "oracle port of the reference ([V]) on the host cores, fp32",
                       "per_step_batch": args.cpu_batch},
            "cpu_baseline": {"value": rate, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------ native arm
def build_model(dev):
    from mtp_b200 import ViT_Win_RVSA_V3_WSZ7
    torch.manual_seed(0)
    m = ViT_Win_RVSA_V3_WSZ7(img_size=224, patch_size=16, embed_dim=MODEL["embed_dim"], depth=MODEL["depth"], num_heads=MODEL["num_heads"],
                             mlp_ratio=4, qkv_bias=True, use_abs_pos_emb=True, interval=MODEL["interval"],
                             out_indices=MODEL["out_indices"], drop_path_rate=0.1, use_rel_pos_bias=True)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "rel_pos" in n:                      # zero-initialised in the reference: re-draw so the terms are exercised
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
    return m.to(dev).train()


def gemm_profile(trainer, x):
    """One eager (non-graph) step with CUDA events around every GEMM launch on the launching stream."""
    from mtp_b200 import ops
    recs = []
    abytes = [0.0]
    orig = ops.gemm

    def alg_bytes(M, N, K, out, aux=None, out2=None, **_):      # operands read once + outputs written once
        b = 2.0 * (M * K + N * K) + out.element_size() * M * N
        if aux is not None:
            b += aux.element_size() * M * N
        if out2 is not None:
            b += out2.element_size() * M * N
        return b

    def timed(A, B, M, N, K, out, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig(A, B, M, N, K, out, **kw)
        e1.record()
        recs.append((e0, e1, 2.0 * M * N * K))
        abytes[0] += alg_bytes(M, N, K, out, **kw)
        return r
    orig_dual = ops.gemm_dual

    def timed_dual(g0, g1, force_bn=0):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig_dual(g0, g1, force_bn)
        e1.record()
        recs.append((e0, e1, 2.0 * (g0["M"] * g0["N"] * g0["K"] + g1["M"] * g1["N"] * g1["K"])))
        abytes[0] += sum(alg_bytes(g["M"], g["N"], g["K"], g["out"], aux=g.get("aux"), out2=g.get("out2")) for g in (g0, g1))
        return r
    ops.gemm = timed
    ops.gemm_dual = timed_dual
    try:
        trainer._step_body(x)
        torch.cuda.synchronize()
    finally:
        ops.gemm = orig
        ops.gemm_dual = orig_dual
    ms = sum(a.elapsed_time(b) for a, b, _ in recs)
    flops = sum(f for _, _, f in recs)
    return len(recs), ms, flops, abytes[0]


def count_launches(trainer, x):
    """Kernels of libmtp_b200.so launched per step (entry point -> kernels it enqueues)."""
    from mtp_b200 import _lib
    per_call = {"mtp_rvsa_sampling_fwd": 2, "mtp_rvsa_attn_bwd": 3, "mtp_rvsa_sampling_bwd": 3}
    n = [0]
    orig = _lib.call

    def counting(name, *a):
        n[0] += per_call.get(name, 1)
        return orig(name, *a)
    _lib.call = counting           # ops / trainer / engine_bwd all call through the module attribute
    try:
        trainer._step_body(x)
        torch.cuda.synchronize()
    finally:
        _lib.call = orig
    return n[0]


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    assert torch.cuda.is_available(), "bench.py (native arm) needs a CUDA device"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist
    if world > 1:
        if args.comm_sms > 0:
            os.environ.setdefault("NCCL_MAX_CTAS", str(args.comm_sms))       # the SMs PretrainStep(comm_sms=) leaves to the all-reduce kernels
        dist.init_process_group("nccl", device_id=dev)
    from mtp_b200 import _lib
    from mtp_b200.trainer import PretrainStep
    _lib.load()
    B = args.batch
    model = build_model(dev)
    trainer = PretrainStep(model, lr=6e-5, weight_decay=0.05, max_norm=5.0, t_max=80000, use_cuda_graph=bool(args.graph),
                           bucket_blocks=args.bucket_blocks, comm_sms=args.comm_sms)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    x = torch.randn(B, 3, 224, 224, device=dev, generator=g).to(torch.bfloat16)
    x_host = x.cpu().pin_memory()

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- eager instrumented passes (also serve as warm-up for kernel attributes / allocator)
    trainer_eager_graph = trainer.use_cuda_graph
    trainer.use_cuda_graph = False
    for _ in range(2):
        trainer.step(x)
    torch.cuda.synchronize()
    launches = count_launches(trainer, x)
    n_gemm, gemm_ms, gemm_flops, gemm_alg_bytes = gemm_profile(trainer, x)
    trainer.use_cuda_graph = trainer_eager_graph

    # ---- device-resident timing
    for _ in range(max(3, args.warmup)):
        loss = trainer.step(x)
    sync_all()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = trainer.step(x)
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / args.steps
    value = world * B / (ms_step / 1e3)
    final_loss = float(loss.item())

    # ---- end to end: pinned host batch -> device -> step -> loss back on the host, every step
    for _ in range(2):
        trainer.step_from_host(x_host)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        trainer.step_from_host(x_host)
    sync_all()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([e2e_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B / (t.item() / args.steps / 1e3)

    # ---- GEMM time inside the graph-replayed step: the same step re-captured with empty GEMM launches (mtp_gemm_set_debug_mode 4);
    #      the difference is what the tcgen05 GEMM kernels cost in situ (operands in the state the step leaves them, PDL overlap
    #      included).  Run last: the training state is garbage afterwards.
    gemm_ms_graph = None
    if args.graph and world == 1:
        try:
            _lib.call("mtp_gemm_set_debug_mode", 4)
            trainer.graph = None
            for _ in range(3):
                trainer.step(x)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(args.steps):
                trainer.step(x)
            e1.record()
            torch.cuda.synchronize()
            gemm_ms_graph = ms_step - e0.elapsed_time(e1) / args.steps
        finally:
            _lib.call("mtp_gemm_set_debug_mode", 0)
        if gemm_ms_graph is not None and gemm_ms_graph > 0:
            gemm_ms_eager, gemm_ms = gemm_ms, gemm_ms_graph

    if rank == 0:
        pk = peaks()
        # DRAM traffic of the GEMM launches of one step, from the committed ncu pass (profiles/: dram__bytes_read/write.sum per launch)
        traffic = None
        try:
            tj = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r1_gemm_dram.json")))
            per_launch = (tj["dram_read_bytes"] + tj["dram_write_bytes"]) / max(1, tj["gemm_launches"])
            traffic = per_launch * n_gemm      # bytes per step over all GEMM launches (same unit of work as `achieved`)
        except Exception:
            pass
        step_tflops = 3.0 * FWD_GFLOP_PER_IMG * (value / world) / 1e3          # per GPU, training step = 3 x forward
        gemm_tflops = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
        line = {
            "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "per_gpu_batch": B, "global_batch": B * world, "tokens_per_gpu": B * 196, "parallelism": f"dp{world}",
                       "cuda_graph": bool(args.graph), "l2": "per-step working set (~5 GB of weights, activations, gradients) >> 126 MB L2; no explicit flush",
                       "loss": final_loss},
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": x_host.numel() * x_host.element_size(), "d2h_bytes_per_step": 4},
            "gpu_launches": launches * args.steps,
            "roofline": {"bound": "tensor", "kernel": "gemm_bf16_kernel (all tcgen05 GEMM launches of one step)", "achieved": gemm_tflops,
                         "peak": pk["tf_burst"], "unit": "TFLOP/s", "frac": (gemm_tflops / pk["tf_burst"]) if gemm_tflops else None,
                         "traffic": traffic, "algorithmic_bytes_per_step": gemm_alg_bytes, "peak_source": pk["src"] + " (burst cuBLAS bf16)", "gemm_launches_per_step": n_gemm,
                         "gemm_ms_per_step": gemm_ms, "gemm_share_of_step": gemm_ms / ms_step if ms_step else None,
                         "step_tflops_per_gpu": step_tflops, "step_frac_of_sustained_peak": step_tflops / pk["tf_sustained"]},
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            try:
                rate, threads, total = cpu_reference_rate(args.cpu_batch, 2, 1)
                line["cpu_baseline"] = {"value": rate, "unit": "images/s", "cores": threads, "kind": "port",
                                        "sample": f"{args.cpu_batch} image/step x 2 timed steps (+1 warm-up), ViT-L+RVSA @224 fwd+bwd+AdamW fp32 oracle, {total:.1f} s"}
            except Exception as ex:      # the baseline is informative; never lose the GPU line to it
                line["cpu_baseline"] = {"error": repr(ex)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
