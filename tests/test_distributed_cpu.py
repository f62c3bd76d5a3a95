"""gloo tests (world sizes 2 and 4, CPU) of the host-side logic around the one multi-GPU exchange step: sharding records by
reference batch, the tuple layout and decision rule of the exchange (tests/exchange_protocol.py, the test double of
csrc/exchange.cuh), gathering the survivors and the library's own last step (cmx_exchange_finish, host C) == the
single-process low-memory post-processing of the oracle."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np
import torch.distributed as dist
import chromap_b200 as cb
from chromap_b200 import distributed as cd
from tests import exchange_protocol as xp
from oracle import oracle_py as orc
world = %(world)d
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=world)
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
    mine = recs[np.array([cd.shard_owner(int(b), world) == rank for b in recs["read_id"] // batch])]   # batch b -> rank b mod N
    if rank == world - 1 and preset == "atac":
        mine = mine[:0]                                      # a rank without records must not break the exchange
        recs_all = recs[np.array([cd.shard_owner(int(b), world) != rank for b in recs["read_id"] // batch])]
    else:
        recs_all = recs if preset != "atac" else recs[np.array([cd.shard_owner(int(b), world) != world - 1 for b in recs["read_id"] // batch])]
    surv = xp.dedup_exchange(mine, p)
    final = cd.gather_and_finish(p, surv)                    # cmx_exchange_finish: order + deferred Tn5
    # the range-shuffle form of the same step: sample-sort partition, records travel once, ordinary post-processing on the
    # receiving rank, the ranks' parts in rank order are the run's output
    op = orc.make_params(preset, **kw)
    part = xp.dedup_shuffle(mine, p, lambda r: orc.postprocess(op, r))
    ranged = cd.gather_ranges(part)
    if rank == 0:
        want = orc.postprocess(op, recs_all)
        assert len(final) == len(want), (preset, len(final), len(want))
        assert len(ranged) == len(want), (preset, len(ranged), len(want))
        for f in cb.PE_RECORD.names:
            assert np.array_equal(final[f], want[f]), (preset, f)
            assert np.array_equal(ranged[f], want[f]), ("shuffle", preset, f)
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


@pytest.mark.parametrize("world", [2, 4])
def test_dedup_exchange_gloo(tmp_path, world):
    script = tmp_path / "w.py"
    script.write_text(WORKER % dict(root=ROOT, port=29500 + (os.getpid() + world) % 400, world=world))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o


def test_exchange_entry_points_refuse_without_a_communicator():
    import chromap_b200 as cb
    L = cb.load_library()
    for name in ("cmx_comm_unique_id", "cmx_comm_init", "cmx_comm_destroy", "cmx_dedup_exchange", "cmx_dedup_shuffle", "cmx_exchange_finish"):
        assert hasattr(L, name)
