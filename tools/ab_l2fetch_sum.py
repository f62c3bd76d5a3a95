#!/usr/bin/env python
"""Summarise tools/ab_l2fetch.sh: second call of every run, per kernel name: time (us) and DRAM read (MB) at each granularity."""
import collections
import csv

names = None
tab = collections.OrderedDict()
for g in (0, 32, 64, 128):
    rows = list(csv.DictReader([l for l in open("/tmp/l2f%d.csv" % g) if l.startswith('"')]))
    per = collections.OrderedDict()
    ids = sorted({int(r["ID"]) for r in rows})
    byid = collections.defaultdict(dict)
    for r in rows:
        byid[int(r["ID"])]["name"] = r["Kernel Name"].split("(")[0].replace("void ", "")[:40]
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = r["Metric Unit"]
        if r["Metric Name"].startswith("gpu__time"):
            byid[int(r["ID"])]["us"] = v / 1e3 if unit in ("ns", "nsecond") else (v if unit.startswith("us") else v * 1e3)
        else:
            byid[int(r["ID"])]["mb"] = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(unit, 1e-6)
    seq = [byid[i] for i in ids]
    starts = [i for i, k in enumerate(seq) if k["name"].startswith("seed_front_kernel")]
    call = seq[starts[-1]:]
    for k in call:
        e = per.setdefault(k["name"], [0.0, 0.0])
        e[0] += k.get("us", 0.0); e[1] += k.get("mb", 0.0)
    tab[g] = per
keys = list(tab[0].keys())
print("%-42s" % "kernel (second call, summed over its launches)" + "".join("   %4s: us      MB" % (g if g else "dflt") for g in tab))
tot = {g: [0.0, 0.0] for g in tab}
for k in keys:
    line = "%-42s" % k
    for g in tab:
        e = tab[g].get(k, [0.0, 0.0])
        tot[g][0] += e[0]; tot[g][1] += e[1]
        line += "   %9.1f %7.1f" % (e[0], e[1])
    print(line)
print("%-42s" % "total" + "".join("   %9.1f %7.1f" % (tot[g][0], tot[g][1]) for g in tab))
