#!/usr/bin/env python
"""Top source lines of an `ncu --page source --csv --print-source cuda,sass` export (SASS rows already stripped)."""
import csv, gzip, sys
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
want = sys.argv[3] if len(sys.argv) > 3 else None
rows = list(csv.reader(gzip.open(path, "rt")))
fn = None
files = {}
cur_file = None
hdr = None
out = {}
seen = set()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        base = r[1].replace("void ", "").split("(")[0].split("<")[0]
        k = 0
        while (base, k, cur_file) in seen:
            k += 1
        seen.add((base, k, cur_file))
        fn = "%s#%d" % (base, k)
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit():
        d = dict(zip(hdr, r))
        def num(k):
            try:
                return float(d.get(k, "0").replace(",", ""))
            except ValueError:
                return 0.0
        out.setdefault(fn, []).append((num("Instructions Executed"), num("# Samples"), num("Thread Instructions Executed"), cur_file, r[0], r[1].strip()[:110]))
for f, L in out.items():
    if want and want not in f:
        continue
    tot = sum(x[0] for x in L); ts = sum(x[1] for x in L)
    print("== %s: %.0f warp-instr, %.0f samples" % (f, tot, ts))
    for x in sorted(L, key=lambda x: -x[1])[:top]:
        print("  %5.1f%% smp %5.1f%% ins thr/ins %4.1f  %s:%s  %s" % (100 * x[1] / max(ts, 1), 100 * x[0] / max(tot, 1), x[2] / max(x[0], 1), x[3], x[4], x[5]))
