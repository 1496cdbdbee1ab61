"""Summarise ncu output into small tracked files under profiles/.

    python tools/ncu_summary.py launches gpurun_out/x.csv profiles/r1_launches.md   # --metrics gpu__time_duration.sum ... --csv
    python tools/ncu_summary.py full gpurun_out/x.ncu-rep profiles/r1_full.md       # --set full report
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEYS = [
    ("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm % of peak"),
    ("smsp__inst_executed.sum", "warp instructions"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma pipe %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "regs/thread"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    d = defaultdict(dict)
    for row in csv.DictReader(lines):
        d[(int(row["ID"]), row["Kernel Name"])][row["Metric Name"]] = (row["Metric Value"], row["Metric Unit"])
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({src})\n\n`ncu --metrics gpu__time_duration.sum,... --clock-control none` -- per-launch times are "
                "cold-cache and serialised: compare shares, not absolutes.\n\n| id | kernel | duration | grid | block | warp instr | issue active % |\n|---|---|---|---|---|---|---|\n")
        for (i, k), m in sorted(d.items()):
            g = lambda key: " ".join(m.get(key, ("", "")))
            f.write(f"| {i} | `{k[:90]}` | {g('gpu__time_duration.sum')} | {g('launch__grid_size')} | {g('launch__block_size')} | "
                    f"{g('smsp__inst_executed.sum')} | {g('smsp__issue_active.avg.pct_of_peak_sustained_active')} |\n")


def shares(src, dst, title=""):
    """aggregate a launch list per kernel: launches, total/avg duration, share of the total"""
    lines = [l for l in open(src) if not l.startswith("==")]
    tot = defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1e3 if row["Metric Unit"].startswith("ns") else (v * 1e3 if row["Metric Unit"].startswith("ms") else v)
        t = tot[row["Kernel Name"]]
        t[0] += 1
        t[1] += v
    total = sum(t[1] for t in tot.values()) or 1.0
    with open(dst, "w") as f:
        f.write(f"# ncu launch list of {title or src}\n\n`ncu --metrics gpu__time_duration.sum --clock-control none --csv`: kernel nodes of the CUDA "
                "graphs are profiled one by one, cold-cache and serialised -- compare SHARES, not absolutes.\n\n"
                "| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|\n")
        for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k[:70]}` | {n} | {us:.1f} | {us / n:.2f} | {100 * us / total:.1f}% |\n")


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({src})\n\n")
        for row in rows[2:]:
            f.write(f"## `{row[hdr.index('Kernel Name')][:110]}`\n\n| metric | value |\n|---|---|\n")
            for key, label in KEYS:
                idx = [i for i, h in enumerate(hdr) if h == key]
                if idx:
                    f.write(f"| {label} (`{key}`) | {row[idx[0]]} {units[idx[0]]} |\n")
            stalls = [(float(row[i] or 0), h) for i, h in enumerate(hdr)
                      if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
            f.write("\nTop stall reasons (warps stalled per issue-active cycle): " +
                    ", ".join(f"{h.split('stalled_')[1].split('_per_issue')[0]} {v:.2f}" for v, h in sorted(stalls, reverse=True)[:6]) + "\n\n")


if __name__ == "__main__":
    {"launches": launches, "full": full, "shares": shares}[sys.argv[1]](*sys.argv[2:])
