"""Multi-GPU layer, one process per GPU: read batches are sharded over ranks (batch b -> rank b mod N, no collective
inside the hot path); the ONE exchange step is duplicate removal over the whole run (mapping_writer.h:166-376 semantics).

Two native forms of that step (include/chromap_b200.h, csrc/exchange.cuh):
  * `cmx_dedup_shuffle` — a range shuffle (sample sort): every record travels once, by grouped ncclSend / ncclRecv over
    NVLink, to the rank that owns its key range, which runs the single-GPU post-processing on what it received.  The run's
    output is the ranks' outputs in rank order; work per rank follows its share of the run.
  * `cmx_dedup_exchange` — packs this rank's records into 16-byte tuples, moves them with ONE ncclAllGather and decides on
    the GPU which of this rank's records survive (no record leaves its rank; every rank sorts all tuples).
This module only does the plumbing torch.distributed is here for: handing the NCCL unique id of the library's
communicator to the other ranks, and bringing the finished parts to the rank that writes the output.  There is no CPU
path: without a GPU the library refuses to create a context.
"""
import numpy as np
import torch.distributed as dist

from .binding import PE_RECORD, exchange_finish


def shard_owner(batch_index, world):
    """reference batch b (500 000 pairs, read ids b*500000 ..) is mapped by rank b mod N."""
    return batch_index % world


def init_comm(mapper, group=None):
    """Create the library's own NCCL communicator over the ranks of `group` (rank 0 draws the unique id, the process
    group — nccl or gloo — carries its 128 bytes to the others)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [mapper.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    mapper.comm_init(world, rank, box[0])
    return rank, world


def dedup_exchange(mapper, recs, barcode_keys=None):
    """This rank's records -> this rank's survivors (reference order, num_dups set, MAPQ-filtered, Tn5 not yet applied)
    plus the device-side timing / byte counts of the exchange.  The collective is the library's ncclAllGather."""
    return mapper.dedup_exchange(recs, barcode_keys)


def dedup_shuffle(mapper, recs, barcode_keys=None, capacity=None):
    """This rank's records -> the finished records of this rank's KEY RANGE (sample-sort shuffle over NVLink, then the
    single-GPU post-processing on the receiving rank): `cmx_dedup_shuffle`.  Work per rank follows its share of the run."""
    return mapper.dedup_shuffle(recs, barcode_keys, capacity=capacity)


def gather_ranges(part, barcode_keys=None, group=None, dst=0):
    """The ranks' key ranges one after the other in rank order = the run's output (nothing left to sort or shift)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    objs = [None] * world if rank == dst else None
    dist.gather_object((part.tobytes(), barcode_keys.tobytes() if barcode_keys is not None else None), objs, dst=dst, group=group)
    if rank != dst:
        return None
    recs = np.concatenate([np.frombuffer(o[0], dtype=PE_RECORD) for o in objs])
    if barcode_keys is not None:
        return recs, np.concatenate([np.frombuffer(o[1], dtype=np.uint64) for o in objs])
    return recs


def gather_and_finish(params, survivors, barcode_keys=None, group=None, dst=0):
    """Bring every rank's survivors to `dst`, put them in the reference's output order and apply the deferred Tn5
    shift (cmx_exchange_finish, mapping_writer.h:285-287).  Returns the final records on dst, None elsewhere."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    objs = [None] * world if rank == dst else None
    payload = (survivors.tobytes(), barcode_keys.tobytes() if barcode_keys is not None else None)
    dist.gather_object(payload, objs, dst=dst, group=group)
    if rank != dst:
        return None
    recs = np.concatenate([np.frombuffer(o[0], dtype=PE_RECORD) for o in objs]) if objs else survivors[:0]
    if barcode_keys is not None:
        bcs = np.concatenate([np.frombuffer(o[1], dtype=np.uint64) for o in objs])
        return exchange_finish(params, recs, bcs)
    return exchange_finish(params, recs)
