#!/usr/bin/env python
"""Turn ncu outputs into the markdown tables under profiles/ (no GPU needed: reads files brought back in gpurun_out/).

  launches   <launches.csv>                    per-kernel launch counts / total / mean / share from an
                                               `ncu --metrics gpu__time_duration.sum --csv --log-file` launch list
  full       <report.ncu-rep> [name regex]     one row per captured kernel from an `ncu --set full` report: duration,
                                               DRAM bytes, pipe utilisation, issue rate, registers, top stall reasons
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict, Counter


def short(name):
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::|nnlm::|void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("(int)", "")


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path, errors="replace") if l.startswith('"'))]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    acc = OrderedDict()
    for r in rows[1:]:
        if len(r) <= iv:
            continue
        v = float(r[iv].replace(",", ""))
        u = r[iu]
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}.get(u, 1e-6)
        k = short(r[ik])
        a = acc.setdefault(k, [0, 0.0])
        a[0] += 1; a[1] += ms
    tot = sum(a[1] for a in acc.values())
    print("| kernel | launches | total ms | mean ms | share |")
    print("|---|---:|---:|---:|---:|")
    for k, (c, t) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {c} | {t:.3f} | {t / c:.4f} | {t / tot:.3f} |")
    print(f"| total | {sum(a[0] for a in acc.values())} | {tot:.3f} | | |")


WANT = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma pipe %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu (MUFU) pipe %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("lts__t_bytes.sum", "L2 bytes"),
        ("launch__registers_per_thread", "registers"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__cluster_size", "cluster"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__inst_executed.sum", "warp instructions")]


def full(path, pattern=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    ik = idx["Kernel Name"]
    seen = Counter()
    for r in rows[2:]:
        name = short(r[ik])
        if pattern and not re.search(pattern, name):
            continue
        seen[name] += 1
        if seen[name] > 2:
            continue
        print(f"\n**`{name}`** (capture {seen[name]})\n")
        print("| metric | value |")
        print("|---|---:|")
        for key, label in WANT:
            if key in idx and r[idx[key]] not in ("", "n/a"):
                print(f"| {label} | {r[idx[key]]} {units[idx[key]]} |")
        st = []
        for h in hdr:
            m = re.match(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active.ratio", h)
            if m and r[idx[h]] not in ("", "n/a"):
                st.append((float(r[idx[h]].replace(",", "")), m.group(1)))
        st.sort(reverse=True)
        print("| stalls per issue (top 5) | " + ", ".join(f"{n} {v:.2f}" for v, n in st[:5]) + " |")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
