#!/usr/bin/env python
"""Where a kernel's warps stall, from an ncu report captured with --set full --import-source on.

    python tools/ncu_source_stalls.py report.ncu-rep kernel_regex [min_samples]

Prints the stall-reason mix of the kernel's first captured launch and every SASS instruction with at
least min_samples stall samples (address, samples, executions, instruction, dominant reasons).
"""
import collections
import csv
import subprocess
import sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    min_s = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    print(rows[hi - 1][1][:100] if hi else "")
    hdr = rows[hi]
    idx = {h: i for i, h in enumerate(hdr)}
    seen = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) == len(hdr) and r[0].startswith("0x") and r[0] not in seen:
            seen[r[0]] = r
    data = list(seen.values())
    S = lambda r, k: int(r[idx[k]])
    tot = sum(S(r, "# Samples") for r in data)
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = collections.Counter()
    for r in data:
        for k in stalls:
            agg[k] += S(r, k)
    print("instructions %d, samples %d, warp-instructions executed %d" % (len(data), tot, sum(S(r, "Instructions Executed") for r in data)))
    print("stall mix: " + ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / tot) for k, v in agg.most_common(8)))
    for n, r in enumerate(data):
        s = S(r, "# Samples")
        if s >= min_s:
            why = {k[6:]: S(r, k) for k in stalls if S(r, k) > 0.15 * s}
            print("%5d %s %5d smp %7d exec  %-60s %s" % (n, r[0][-5:], s, S(r, "Instructions Executed"), r[idx["Source"]].strip()[:60], why))


if __name__ == "__main__":
    main()
