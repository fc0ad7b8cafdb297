#!/usr/bin/env python3
"""Per-source-line instruction counts of one kernel of an .ncu-rep (captured with --import-source on, built with -lineinfo):
    python tools/ncu_lines.py report.ncu-rep <kernel regex> [top N] [launches of that kernel to skip]"""
import csv
import io
import subprocess
import sys


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    skip = sys.argv[4] if len(sys.argv) > 4 else "0"
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                          "regex:" + pat, "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
    hdr = rows[hi]
    ci, cs, cst = hdr.index("Instructions Executed"), hdr.index("# Samples"), 1
    lines, tot, tots = [], 0, 0
    for r in rows[hi + 1:]:
        if len(r) < len(hdr) or r[2] != "-" or not r[0].isdigit():
            continue                                     # SASS rows carry an address in column 2
        n, s = int(r[ci] or 0), int(r[cs] or 0)
        tot += n
        tots += s
        lines.append((n, s, int(r[0]), r[cst]))
    lines.sort(reverse=True)
    print("total warp-instructions %d, stall samples %d" % (tot, tots))
    for n, s, ln, src in lines[:top]:
        print("%5.1f%% inst %5.1f%% samples  L%-4d %s" % (100.0 * n / max(tot, 1), 100.0 * s / max(tots, 1), ln, src.strip()[:130]))


if __name__ == "__main__":
    main()
