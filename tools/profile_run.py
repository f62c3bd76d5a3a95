#!/usr/bin/env python
"""Small driver for ncu: synthetic reference + index on the device, then N identical map_batch calls."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref-gbp", type=float, default=3.0)
    ap.add_argument("--n-seq", type=int, default=24)
    ap.add_argument("--pairs", type=int, default=2000000)
    ap.add_argument("--calls", type=int, default=3)
    ap.add_argument("--preset", default="chip")
    a = ap.parse_args()
    import torch
    import chromap_b200 as cb
    dev = torch.device("cuda", 0)
    ref, offsets, seq_len = bench.gen_reference(torch, dev, int(a.ref_gbp * 1e9), a.n_seq, 11)
    m = cb.Mapper(cb.make_params(a.preset, max_read_length=64), device=0)
    m.upload_reference_ptr(ref.data_ptr(), offsets)
    m.build_index(17, 7)
    r1, r2, off = bench.gen_pairs(torch, ref, a.n_seq, seq_len, a.pairs, 50, 11 * 1000003, dev)
    out = torch.empty(a.pairs * 24, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    for i in range(a.calls):
        _, st = m.map_batch(r1, off, r2, off, on_device=True, n_pairs=a.pairs, out=out, out_on_device=True)
        print(i, st, {k: round(v, 3) if isinstance(v, float) else v for k, v in m.timing().items()}, flush=True)


if __name__ == "__main__":
    main()
