"""Summarise an .ncu-rep: per kernel duration, issue utilisation, pipe mix, stall reasons and
the hottest SASS instructions (from the source page).  python tools/ncu_stalls.py rep.ncu-rep [--top N]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 12
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, data = rows[0], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum"]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
secs, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        secs.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
import re


def sig(name):
    """kernel identity that survives the different demanglings of the raw and source pages"""
    m = re.search(r"(\w+_kernel)<([^>]*)>", name)
    if not m:
        return name.split("(")[0].split("::")[-1], ()
    return m.group(1), tuple(re.findall(r"\d+", re.sub(r"\(int\)|\(bool\)", "", m.group(2))))


# the source page has one section per profiled launch, in launch order, like the raw page; pair them by
# position but verify the kernel identity (a mismatch means the pages are laid out differently)
seen = set()
for k, r in enumerate(data):
    name = r[idx["Kernel Name"]]
    key = (name, r[idx["launch__grid_size"]])
    if key in seen:
        continue
    seen.add(key)
    sec = secs[k] if k < len(secs) and sig(secs[k]["name"]) == sig(name) else next(
        (s_ for s_ in secs if sig(s_["name"]) == sig(name)), None)
    print("=" * 100)
    print(name[:110])
    for w in WANT:
        if w in idx:
            print(f"  {w:70s} {r[idx[w]]}")
    if sec is not None:
        s = sec
        h, d = s["rows"][0], s["rows"][1:]
        tot = sum(int(x[2]) for x in d) or 1
        print(f"  stall samples ({tot}):", end=" ")
        reasons = []
        for i in range(len(h)):
            if h[i].startswith("stall_") and "Not Issued" not in h[i]:
                t = sum(int(x[i]) for x in d)
                if t > 0.02 * tot:
                    reasons.append((t, h[i]))
        print(", ".join(f"{n[6:]} {100 * t / tot:.0f}%" for t, n in sorted(reasons, reverse=True)))
        mix, samp = collections.Counter(), collections.Counter()
        for x in d:
            op = [o for o in x[1].split() if not o.startswith("@")][0]
            mix[op] += int(x[5])
            samp[op] += int(x[2])
        ti = sum(mix.values()) or 1
        print("  mix:", ", ".join(f"{op} {100 * c / ti:.1f}%/{100 * samp[op] / tot:.0f}%s" for op, c in mix.most_common(top_n)))
