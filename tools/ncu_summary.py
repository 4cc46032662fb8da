#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) and, optionally, a
`--set full` report exported with `ncu -i X.ncu-rep --page raw --csv`, as markdown for profiles/.

    python tools/ncu_summary.py launches.csv [raw.csv] > profiles/rNN_x.md
"""
import collections
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank conflicts"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("## Launch list (cold-cache, serialised: compare shares)\n")
    print("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("| `%s` | %d | %.1f | %.2f | %.1f%% |" % (k[:80], a[0], a[1], a[1] / a[0], 100 * a[1] / tot))


def full(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    seen = set()
    print("\n## `ncu --set full` (first captured launch of each kernel)\n")
    for d in data:
        name = d[idx["Kernel Name"]].split("(")[0]
        if name in seen:
            continue
        seen.add(name)
        print("### `%s`\n" % name[:90])
        print("| metric | value |\n|---|---:|")
        for key, label in KEYS:
            if key in idx:
                print("| %s | %s %s |" % (label, d[idx[key]], units[idx[key]]))
        st = []
        for h, i in idx.items():
            if h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("not_issued") and d[i] not in ("", "n/a"):
                st.append((float(d[i].replace(",", "")), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
        tot = sum(s for s, _ in st) or 1.0
        top = ", ".join("%s %.0f%%" % (h, 100 * s / tot) for s, h in sorted(st, reverse=True)[:5])
        print("| top stall reasons | %s |\n" % top)


if __name__ == "__main__":
    launches(sys.argv[1])
    if len(sys.argv) > 2:
        full(sys.argv[2])
