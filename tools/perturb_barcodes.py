#!/usr/bin/env python
"""Make a gen_synth.py --barcodes data set harder: add whitelist entries one substitution away from existing ones
(ambiguous corrections) and put N / extra substitutions into some barcodes."""
import os
import sys

import numpy as np

d = sys.argv[1]
rng = np.random.default_rng(5)
wl = [l.strip() for l in open(os.path.join(d, "whitelist.txt"), "rb")]
extra = []
for w in wl[:len(wl) // 3]:
    b = bytearray(w)
    p = int(rng.integers(0, len(b)))
    b[p] = b"ACGT"[(b"ACGT".index(b[p]) + 1) % 4]
    extra.append(bytes(b))
open(os.path.join(d, "whitelist.txt"), "wb").write(b"\n".join(sorted(set(wl + extra))) + b"\n")
lines = open(os.path.join(d, "barcode.fq"), "rb").read().split(b"\n")
out = []
for i in range(0, len(lines) - 1, 4):
    s = bytearray(lines[i + 1])
    r = rng.random()
    if r < 0.03:
        s[int(rng.integers(0, len(s)))] = ord("N")
    elif r < 0.08:
        s[int(rng.integers(0, len(s)))] = b"ACGT"[int(rng.integers(0, 4))]
    out += [lines[i], bytes(s), lines[i + 2], lines[i + 3]]
open(os.path.join(d, "barcode.fq"), "wb").write(b"\n".join(out) + b"\n")
