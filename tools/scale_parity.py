#!/usr/bin/env python
"""Parity at scale, on the GPU box: the unmodified reference binary (oracle/_ref/chromap, CPU) and chromap-b200 (GPU) map the
same multi-million-pair FASTQ files against the same 3 Gbp synthetic reference + index file; the outputs must be
byte-identical.  Prints one JSON object per case.  (Test infrastructure: executes oracle/_ref.)"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def md5(path):
    h = hashlib.md5()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 24), b""):
            h.update(blk)
    return h.hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref-gbp", type=float, default=3.0)
    ap.add_argument("--n-seq", type=int, default=24)
    ap.add_argument("--batches", type=int, default=10, help="500000-pair reference batches")
    ap.add_argument("--read-len", type=int, default=50)
    ap.add_argument("--hic-batches", type=int, default=4)
    ap.add_argument("--cases", default="chip,atac,se_chip")
    ap.add_argument("--workdir", default="/dev/shm/chromap_b200_scale")
    a = ap.parse_args()
    import torch
    import chromap_b200 as cb
    dev = torch.device("cuda", 0)
    work = a.workdir
    os.makedirs(work, exist_ok=True)
    ref, offsets, seq_len = bench.gen_reference(torch, dev, int(a.ref_gbp * 1e9), a.n_seq, 11)
    m = cb.Mapper(cb.make_params("chip", max_read_length=64), device=0)  # (only builds the index; the runs below are the two binaries)
    m.upload_reference_ptr(ref.data_ptr(), offsets)
    m.build_index(bench.K_MER, bench.WINDOW)
    idx = m.download_index()
    with open(os.path.join(work, "ref.index"), "wb") as f:
        np.array([bench.K_MER, bench.WINDOW], dtype=np.int32).tofile(f)
        np.array([idx["n_keys"], idx["n_buckets"], idx["n_keys"], idx["n_keys"], int(idx["n_buckets"] * 0.77 + 0.5)], dtype=np.uint32).tofile(f)
        idx["flags"].tofile(f); idx["keys"].tofile(f); idx["vals"].tofile(f)
        np.array([len(idx["occ"])], dtype=np.uint32).tofile(f)
        idx["occ"].tofile(f)
    del idx
    href = ref.cpu().numpy()
    with open(os.path.join(work, "ref.fa"), "wb") as f:
        for i in range(a.n_seq):
            f.write(b">chr%d\n" % (i + 1))
            href[int(offsets[i]):int(offsets[i + 1])].tofile(f)
            f.write(b"\n")
    n = 500000

    def write_reads(tag, L, batches, seed0):
        for which in (0, 1):
            with open(os.path.join(work, "%sread%d.fq" % (tag, which + 1)), "wb") as f:
                for b in range(batches):
                    r = bench.gen_pairs(torch, ref, a.n_seq, seq_len, n, L, seed0 + b, dev)[which].cpu().numpy().reshape(n, L)
                    rec = np.empty((n, 2 * L + 16), dtype=np.uint8)
                    ids = np.char.zfill((np.arange(n) + b * n).astype(str), 9).astype("S9")
                    rec[:, 0] = ord("@"); rec[:, 1:10] = np.frombuffer(ids.tobytes(), dtype=np.uint8).reshape(n, 9)
                    rec[:, 10] = 10; rec[:, 11:11 + L] = r; rec[:, 11 + L] = 10; rec[:, 12 + L] = ord("+"); rec[:, 13 + L] = 10
                    rec[:, 14 + L:14 + 2 * L] = ord("I"); rec[:, 14 + 2 * L] = 10
                    rec[:, :15 + 2 * L].tofile(f)

    cases = a.cases.split(",")
    write_reads("", a.read_len, a.batches, 7000003)
    if "hic" in cases:  # BASELINE config 5: 2 x 150 bp, split alignment, pairs output
        write_reads("hic_", 150, a.hic_batches, 9000003)
    if "scatac" in cases:  # BASELINE config 4: 16 bp cell barcodes + a 737 k-entry whitelist, the same 2 x 50 bp reads
        bc_len = 16
        wl_keys, cell_keys = bench.gen_whitelist(torch, dev, 737000, 10000, bc_len, 28)
        sh = (2 * (bc_len - 1 - torch.arange(bc_len, device=dev)))[None, :]
        lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
        wl_txt = lut[((wl_keys[:, None] >> sh) & 3)].cpu().numpy()
        with open(os.path.join(work, "whitelist.txt"), "wb") as f:
            np.concatenate([wl_txt, np.full((len(wl_txt), 1), 10, dtype=np.uint8)], axis=1).tofile(f)
        with open(os.path.join(work, "barcode.fq"), "wb") as f:
            for b in range(a.batches):
                bs, bq = bench.gen_barcodes(torch, cell_keys, n, bc_len, 5000003 + b, dev)
                bs = bs.cpu().numpy().reshape(n, bc_len); bq = bq.cpu().numpy().reshape(n, bc_len)
                rec = np.empty((n, 2 * bc_len + 16), dtype=np.uint8)
                ids = np.char.zfill((np.arange(n) + b * n).astype(str), 9).astype("S9")
                rec[:, 0] = ord("@"); rec[:, 1:10] = np.frombuffer(ids.tobytes(), dtype=np.uint8).reshape(n, 9)
                rec[:, 10] = 10; rec[:, 11:11 + bc_len] = bs; rec[:, 11 + bc_len] = 10; rec[:, 12 + bc_len] = ord("+"); rec[:, 13 + bc_len] = 10
                rec[:, 14 + bc_len:14 + 2 * bc_len] = bq; rec[:, 14 + 2 * bc_len] = 10
                rec[:, :15 + 2 * bc_len].tofile(f)
    del ref, href, m
    torch.cuda.empty_cache()
    refbin = os.path.join(ROOT, "oracle", "_ref", "chromap")
    ours = os.path.join(ROOT, "chromap_b200", "bin", "chromap-b200")
    cores = os.cpu_count() or 1
    common = ["-x", os.path.join(work, "ref.index"), "-r", os.path.join(work, "ref.fa"), "-1", os.path.join(work, "read1.fq")]
    results = []
    for case in cases:
        extra = {"chip": ["--preset", "chip"], "atac": ["--preset", "atac"], "se_chip": ["--preset", "chip"], "hic": ["--preset", "hic"],
                 "scatac": ["--preset", "atac", "-b", os.path.join(work, "barcode.fq"), "--barcode-whitelist", os.path.join(work, "whitelist.txt")],
                 "default_q0": ["-q", "0", "--remove-pcr-duplicates"], "se_q0": ["-q", "0", "--remove-pcr-duplicates", "--Tn5-shift"]}[case]
        files = common + ([] if case.startswith("se_") else ["-2", os.path.join(work, "read2.fq")])
        if case == "hic":
            files = ["-x", os.path.join(work, "ref.index"), "-r", os.path.join(work, "ref.fa"), "-1", os.path.join(work, "hic_read1.fq"), "-2", os.path.join(work, "hic_read2.fq")]
        out_ref, out_ours = os.path.join(work, case + ".ref.bed"), os.path.join(work, case + ".ours.bed")
        t0 = time.time()
        # the reference's single-end loop (taskloop num_tasks(t*t), chromap.h:383) crashed at -t 128 on this box; 16 is fine
        threads = min(cores, 16) if case.startswith("se_") else cores
        r1 = subprocess.run([refbin] + extra + files + ["-o", out_ref, "-t", str(threads)], capture_output=True, text=True)
        t_ref = time.time() - t0
        t0 = time.time()
        r2 = subprocess.run([ours] + extra + files + ["-o", out_ours], capture_output=True, text=True)
        t_ours = time.time() - t0
        ok = r1.returncode == 0 and r2.returncode == 0
        res = {"case": case, "pairs": (a.hic_batches if case == "hic" else a.batches) * n, "ref_gbp": a.ref_gbp, "reference_rc": r1.returncode, "ours_rc": r2.returncode}
        if ok:
            res.update(identical=md5(out_ref) == md5(out_ours), md5_reference=md5(out_ref), md5_ours=md5(out_ours), bytes=os.path.getsize(out_ref),
                       lines=sum(1 for _ in open(out_ref, "rb")), reference_wall_s=round(t_ref, 1), ours_wall_s=round(t_ours, 1),
                       reference_mapping=[l for l in r1.stderr.splitlines() if l.startswith("Mapped all")],
                       ours_mapping=[l for l in r2.stderr.splitlines() if l.startswith("Mapped all")],
                       ours_startup=[l for l in r2.stderr.splitlines() if l.startswith("Reference and index resident") or l.startswith("Start-up:")],
                       ours_phases=[l for l in r2.stderr.splitlines() if l.startswith("Mapping phase:") or l.startswith("Mapped ") and "read pairs in" in l
                                    or l.startswith("Sorted") or l.startswith("Post") or l.startswith("Wrote") or l.startswith("Output")],
                       ours_total=[l for l in r2.stderr.splitlines() if l.startswith("Total time")],
                       reference_total=[l for l in r1.stderr.splitlines() if l.startswith("Total time")])
        else:
            res["stderr"] = (r1.stderr[-300:] + " || " + r2.stderr[-300:])
        print(json.dumps(res), flush=True)
        results.append(res)
        for p in (out_ref, out_ours):
            if os.path.exists(p):
                os.remove(p)
    for f in ("ref.index", "ref.fa", "read1.fq", "read2.fq", "hic_read1.fq", "hic_read2.fq", "barcode.fq", "whitelist.txt"):
        try:
            os.remove(os.path.join(work, f))
        except OSError:
            pass
    sys.exit(0 if all(r.get("identical") for r in results) else 1)


if __name__ == "__main__":
    main()
