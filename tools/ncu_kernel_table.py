"""Per-kernel table from `ncu --page raw --csv` exports of one `--set full` capture each (tools/profile_round.sh): duration, DRAM bytes and
achieved GB/s, tensor-pipe activity (the tcgen05-aware counters: `sm__pipe_tensor_cycles_active_realtime`, `sm__ops_path_tensor_op_hmma_*`),
TMEM pipe, achieved occupancy, registers / shared memory.      python tools/ncu_kernel_table.py gpurun_out/r2prof > profiles/r2_kernels.md"""
import csv, glob, json, os, sys

PEAKS = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}


def load(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H, U, V = rows[hdr], rows[hdr + 1], rows[hdr + 2]
    return {h: (v, u) for h, u, v in zip(H, U, V)}


def num(d, key, default=float("nan")):
    try:
        return float(d[key][0].replace(",", ""))
    except Exception:
        return default


def main(folder):
    print(f"| capture | kernel | grid x block | regs | smem/CTA KB | duration us | DRAM MB (r+w) | DRAM GB/s (of measured {PEAKS['hbm_gbs']:.0f}) | tcgen05 tensor-memory pipe active % (`sm__mem_tensor_cycles_active`, of elapsed) | "
          f"legacy tensor pipe % (`sm__pipe_tensor_cycles_active_realtime`: does not count UTCHMMA) | TMEM ld/st issue % | achieved occupancy % | L2 hit % |")
    print("|---|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
    for path in sorted(glob.glob(os.path.join(folder, "*_raw.csv"))):
        try:
            d = load(path)
        except (IndexError, OSError):
            continue          # empty export: the capture's kernel filter matched nothing
        name = d.get("Kernel Name", ("?", ""))[0].split("(")[0].replace("void ", "").replace("mtp::", "")
        dur_ns = num(d, "gpu__time_duration.sum")
        scale = {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(d.get("gpu__time_duration.sum", ("", "ns"))[1], 1.0)
        dur_us = dur_ns * scale / 1e3

        def bytes_of(key):
            v = num(d, key, 0.0)
            u = d.get(key, ("", "byte"))[1]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        dram = bytes_of("dram__bytes_read.sum") + bytes_of("dram__bytes_write.sum")
        gbs = dram / (dur_us * 1e-6) / 1e9 if dur_us == dur_us and dur_us > 0 else float("nan")
        tens = num(d, "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed")
        if tens != tens:
            tens = num(d, "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed")
        ops_ps = num(d, "sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.sum.per_second")
        opu = d.get("sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.sum.per_second", ("", ""))[1]
        # `ops` counts multiply-adds x 2 as FLOP in ncu's convention for this path; report as-is in TFLOP/s
        tf = ops_ps * {"": 1, "/second": 1, "K/second": 1e3, "M/second": 1e6, "G/second": 1e9, "T/second": 1e12}.get(opu.replace("inst", "").strip(), 1) / 1e12 if ops_ps == ops_ps else float("nan")
        tmem = num(d, "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active")
        occ = num(d, "sm__warps_active.avg.pct_of_peak_sustained_active")
        l2 = num(d, "lts__t_sector_hit_rate.pct")
        grid = d.get("Grid Size", ("?", ""))[0]
        block = d.get("Block Size", ("?", ""))[0]
        regs = d.get("launch__registers_per_thread", ("?", ""))[0]
        smem = num(d, "launch__shared_mem_per_block", 0.0)
        su = d.get("launch__shared_mem_per_block", ("", "byte"))[1]
        smem_kb = smem * {"byte/block": 1, "Kbyte/block": 1e3, "Mbyte/block": 1e6}.get(su, 1) / 1024
        print(f"| {os.path.basename(path)[:-8]} | `{name[:48]}` | {grid} x {block} | {regs} | {smem_kb:.0f} | {dur_us:.1f} | {dram / 1e6:.1f} | {gbs:.0f} ({100 * gbs / PEAKS['hbm_gbs']:.0f} %) | "
              f"{num(d, 'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'):.1f} | {tens:.1f} | {tmem:.2f} | {occ:.1f} | {l2:.1f} |")


if __name__ == "__main__":
    main(sys.argv[1])
