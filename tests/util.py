"""Test helpers: tiny FASTA/FASTQ readers (gz-aware) and batch packing."""
import gzip
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _open(p):
    return gzip.open(p, "rb") if p.endswith(".gz") else open(p, "rb")


def read_fasta(path):
    names, seqs, cur = [], [], []
    with _open(path) as f:
        for line in f:
            line = line.rstrip(b"\r\n")
            if line.startswith(b">"):
                if names:
                    seqs.append(np.frombuffer(b"".join(cur), dtype=np.uint8))
                names.append(line[1:].split()[0].decode())
                cur = []
            else:
                cur.append(line)
    seqs.append(np.frombuffer(b"".join(cur), dtype=np.uint8))
    return names, seqs


def read_fastq(path):
    seqs = []
    with _open(path) as f:
        for i, line in enumerate(f):
            if i % 4 == 1:
                seqs.append(line.rstrip(b"\r\n"))
    return seqs


def read_fastq_records(path):
    """[(name, sequence, quality)] of a 4-line FASTQ; name = header up to the first whitespace (kseq)."""
    out = []
    with _open(path) as f:
        lines = [l.rstrip(b"\r\n") for l in f]
    for i in range(0, len(lines) - 3, 4):
        out.append((lines[i][1:].split()[0] if lines[i][1:].split() else b"", lines[i + 1], lines[i + 3]))
    return out


def pack(reads):
    off = np.zeros(len(reads) + 1, dtype=np.uint32)
    off[1:] = np.cumsum([len(r) for r in reads])
    return np.frombuffer(b"".join(reads), dtype=np.uint8), off


def load_pairs(d, r1="read1.fq.gz", r2="read2.fq.gz"):
    a, b = read_fastq(os.path.join(d, r1)), read_fastq(os.path.join(d, r2))
    s1, o1 = pack(a)
    s2, o2 = pack(b)
    return s1, o1, s2, o2
