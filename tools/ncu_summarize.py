"""Summarise ncu CSV logs for profiles/: per-kernel shares of a launch list, and DRAM traffic of the GEMM launches.

  python tools/ncu_summarize.py launches gpurun_out/launches_X.csv            -> markdown table on stdout (last full step)
  python tools/ncu_summarize.py dram gpurun_out/gemm_dram_X.csv out.json      -> per-step DRAM bytes of the GEMM launches
"""
import collections, csv, json, sys


def read(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ki, mi, vi = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value")
    out = []
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        try:
            out.append((int(r[0]), r[ki].split("(")[0].replace("void ", "").replace("mtp::", ""), r[mi], float(r[vi].replace(",", ""))))
        except ValueError:
            pass
    return out


def launches(path):
    L = [(k, v / 1e3) for _, k, m, v in read(path) if m == "gpu__time_duration.sum"]
    ad = [i for i, (k, _) in enumerate(L) if k.startswith("adamw")]
    step = L[ad[-2] + 1:ad[-1] + 1] if len(ad) >= 2 else L
    tot = sum(v for _, v in step)
    d = collections.defaultdict(lambda: [0, 0.0])
    for k, v in step:
        d[k][0] += 1
        d[k][1] += v
    print(f"One training step = {len(step)} kernel launches, {tot / 1e3:.2f} ms summed kernel time (ncu: serialised, cold caches)\n")
    print("| kernel | launches | total µs | share | avg µs |\n|---|---:|---:|---:|---:|")
    for k, (c, t) in sorted(d.items(), key=lambda x: -x[1][1]):
        print(f"| `{k[:70]}` | {c} | {t:.0f} | {100 * t / tot:.1f}% | {t / c:.1f} |")


def dram(path, out):
    recs = collections.defaultdict(dict)
    for i, k, m, v in read(path):
        if "gemm_bf16" in k:
            recs[i][m] = v
            recs[i]["k"] = k
    ids = sorted(recs)
    n = len(ids)
    rd = sum(recs[i].get("dram__bytes_read.sum", 0.0) for i in ids)
    wr = sum(recs[i].get("dram__bytes_write.sum", 0.0) for i in ids)
    t = sum(recs[i].get("gpu__time_duration.sum", 0.0) for i in ids)
    json.dump({"gemm_launches": n, "dram_read_bytes": rd, "dram_write_bytes": wr, "time_ns": t, "source": path}, open(out, "w"), indent=1)
    print(n, "GEMM launches", rd / 1e6, "MB read", wr / 1e6, "MB written", t / 1e6, "ms")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        dram(sys.argv[2], sys.argv[3])
