"""world_size-2 gloo test of the one multi-GPU exchange step (duplicate removal): sharding records over two
ranks + dedup_exchange + gather == the single-process low-memory post-processing of the oracle."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np
import torch.distributed as dist
import chromap_b200 as cb
from chromap_b200 import distributed as cd
from oracle import oracle_py as orc
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
rng = np.random.default_rng(7)
n = 40000
recs = np.zeros(n, dtype=cb.PE_RECORD)
recs["read_id"] = np.arange(n)
recs["rid"] = rng.integers(0, 3, n)
recs["fragment_start"] = rng.integers(0, 3000, n)         # many collisions -> duplicate groups
recs["fragment_length"] = rng.integers(100, 104, n)
recs["mapq"] = rng.choice([0, 5, 30, 42, 60], n)
recs["direction"] = rng.integers(0, 2, n)
recs["is_unique"] = rng.integers(0, 2, n)
recs["num_dups"] = 1
recs["positive_alignment_length"] = 50
recs["negative_alignment_length"] = 50
for preset, kw in (("chip", {}), ("atac", {}), ("", dict(low_memory_mode=1, mapq_threshold=0))):
    p = cb.make_params(preset, **kw)
    batch = 500
    mine = recs[(recs["read_id"] // batch) %% 2 == rank]      # batch b -> rank b mod N
    surv = cd.dedup_exchange(mine, p)
    final = cd.gather_and_finish(surv, p)
    if rank == 0:
        op = orc.make_params(preset, **kw)
        want = orc.postprocess(op, recs)
        assert len(final) == len(want), (preset, len(final), len(want))
        for f in cb.PE_RECORD.names:
            assert np.array_equal(final[f], want[f]), (preset, f)
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_dedup_exchange_two_ranks_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER % dict(root=ROOT, port=29500 + os.getpid() % 400))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o
