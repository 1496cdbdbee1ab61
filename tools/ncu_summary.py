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
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
