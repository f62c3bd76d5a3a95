"""Multi-GPU layer, one process per GPU: read batches are sharded over ranks (batch b -> rank b mod N, no collective
inside the hot path); the ONE exchange step is duplicate removal over the whole run (mapping_writer.h:166-376 semantics).

The exchange itself is native: `cmx_dedup_exchange` (include/chromap_b200.h, csrc/exchange.cuh) packs this rank's
device-resident records into 16-byte tuples, moves them with ONE ncclAllGather over NVLink and decides on the GPU
which of this rank's records survive.  This module only does the plumbing torch.distributed is here for: handing the
NCCL unique id of the library's communicator to the other ranks, and bringing the (few) survivors to the rank that
writes the output.  There is no CPU path: without a GPU the library refuses to create a context.
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
