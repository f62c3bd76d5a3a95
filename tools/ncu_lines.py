#!/usr/bin/env python
"""Per-source-line warp-stall samples of one kernel launch in an .ncu-rep (ncu --page source --print-source cuda,sass)."""
import csv
import subprocess
import sys


def main():
    rep, skip = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", skip, "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    fname = ""
    lines = []
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif len(r) > 7 and r[0].isdigit():
            try:
                lines.append((int(r[6]), int(r[7]), fname, int(r[0]), r[1].strip()))
            except ValueError:
                pass
        elif len(r) >= 2 and r[0] == "Function Name":
            kern = r[1]
    tot = sum(x[0] for x in lines) or 1
    print(kern.split("(")[0], "samples", tot)
    for s, n, f, ln, src in sorted(lines, reverse=True)[:top]:
        print("%5.1f%% %9d inst  %s:%d  %s" % (100.0 * s / tot, n, f, ln, src[:110]))


if __name__ == "__main__":
    main()
