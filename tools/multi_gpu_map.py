#!/usr/bin/env python
"""Multi-GPU end-to-end check (run under torchrun on N GPUs): every rank maps the reference batches it owns
(batch b -> rank b mod N), the ranks run the one NCCL exchange (duplicate removal) in both its forms — cmx_dedup_exchange
(all-gather of tuples) and cmx_dedup_shuffle (range shuffle of records) — rank 0 writes the BED — which must equal the single-process result byte for byte.

  torchrun --nproc-per-node 2 tools/multi_gpu_map.py --dir tests/golden/synth_small --preset chip --batch 1000
"""
import argparse
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dir", required=True)
    ap.add_argument("--preset", default="chip")
    ap.add_argument("--batch", type=int, default=1000)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    import chromap_b200 as cb
    from chromap_b200 import distributed as cd
    from tests.util import load_pairs, read_fasta
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    names, seqs = read_fasta(os.path.join(a.dir, "ref.fa.gz"))
    s1, o1, s2, o2 = load_pairs(a.dir)
    n = len(o1) - 1
    p = cb.make_params(a.preset, max_read_length=64, batch_size=a.batch)
    m = cb.Mapper(p, device=local)
    m.upload_reference(seqs, names)
    m.build_index(17, 7)
    mine = []
    for bi, b0 in enumerate(range(0, n, a.batch)):
        if bi % world != rank:
            continue
        b1 = min(n, b0 + a.batch)
        r, _ = m.map_batch(s1[o1[b0]:o1[b1]], o1[b0:b1 + 1] - o1[b0], s2[o2[b0]:o2[b1]], o2[b0:b1 + 1] - o2[b0], first_read_id=b0)
        mine.append(r.copy())
    mine = np.concatenate(mine) if mine else np.zeros(0, dtype=cb.PE_RECORD)
    cd.init_comm(m)                                       # the library's own NCCL communicator
    surv, xst = cd.dedup_exchange(m, mine)                # pack -> ONE ncclAllGather -> sort / decide on the GPU
    final = cd.gather_and_finish(p, surv)
    part, sst = cd.dedup_shuffle(m, mine, capacity=2 * n + 4096)  # range shuffle: records travel once, post-processing on the owner
    ranged = cd.gather_ranges(part)
    if rank == 0:
        print("multi_gpu_map: exchange", xst, flush=True)
        print("multi_gpu_map: shuffle", sst, flush=True)
    if rank == 0:
        bed = m.format_bed(final)
        bed_sh = m.format_bed(ranged)
        # single-process result on this rank's GPU for comparison
        allr = []
        for b0 in range(0, n, a.batch):
            b1 = min(n, b0 + a.batch)
            r, _ = m.map_batch(s1[o1[b0]:o1[b1]], o1[b0:b1 + 1] - o1[b0], s2[o2[b0]:o2[b1]], o2[b0:b1 + 1] - o2[b0], first_read_id=b0)
            allr.append(r.copy())
        want = m.format_bed(m.postprocess(np.concatenate(allr)))
        ok = bed == want and bed_sh == want
        print("multi_gpu_map: world=%d records=%d bed_md5=%s shuffle_md5=%s single_md5=%s %s" % (world, len(final), hashlib.md5(bed).hexdigest(),
              hashlib.md5(bed_sh).hexdigest(), hashlib.md5(want).hexdigest(), "IDENTICAL" if ok else "DIFFERENT"), flush=True)
        if not ok:
            sys.exit(1)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
