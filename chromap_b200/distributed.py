"""Multi-GPU layer: read batches are sharded over ranks (batch b -> rank b mod N, no collective inside the
hot path); the ONE exchange step is duplicate removal over the whole run (mapping_writer.h:166-376 semantics,
i.e. every preset: --low-mem): one all-gather of 16-byte tuples {rid, start | len, mapq, dir, uniq, read_id},
then every rank decides locally which of ITS records survive and with which duplicate count.

Device-agnostic torch code: `nccl` over NVLink on GPUs, `gloo` in the CPU tests.
"""
import numpy as np
import torch
import torch.distributed as dist

from .binding import PE_RECORD

_SIGN = -(1 << 63)


def pack_tuples(recs):
    """PE records (numpy structured) -> int64 [n, 2] whose lexicographic (unsigned) order is the reference's
    record order prefixed by rid (bed_mapping.h:208-215): (rid, start, len, mapq, direction, is_unique, read_id)."""
    a = (recs["rid"].astype(np.int64) << 32) | recs["fragment_start"].astype(np.int64)
    b = (recs["fragment_length"].astype(np.uint64) << np.uint64(48)) | (recs["mapq"].astype(np.uint64) << np.uint64(42)) | \
        (recs["direction"].astype(np.uint64) << np.uint64(41)) | (recs["is_unique"].astype(np.uint64) << np.uint64(40)) | \
        recs["read_id"].astype(np.uint64)
    b = (b ^ np.uint64(1 << 63)).view(np.int64)  # flip the sign bit: signed order == unsigned order
    return torch.from_numpy(np.stack([a, b], axis=1))


def _all_gather_var(t, group=None):
    """all_gather of tensors whose first dimension differs per rank."""
    world = dist.get_world_size(group)
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes) if sizes else 0
    pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return [o[:s] for o, s in zip(out, sizes)], sizes


def dedup_exchange(recs, params, device=None, group=None):
    """recs: this rank's PE records (numpy).  Returns this rank's surviving records (numpy, pre-Tn5, num_dups set,
    MAPQ-filtered), exactly the subset of the reference's low-memory merge output that this rank owns."""
    rank = dist.get_rank(group)
    device = device or torch.device("cpu")
    tup = pack_tuples(recs).to(device)
    parts, sizes = _all_gather_var(tup, group)
    allt = torch.cat(parts, dim=0)
    owner = torch.cat([torch.full((s,), r, dtype=torch.int64, device=device) for r, s in enumerate(sizes)])
    local = torch.cat([torch.arange(s, dtype=torch.int64, device=device) for s in sizes])
    # lexicographic sort by (a, b): stable sort by b, then by a
    o1 = torch.sort(allt[:, 1], stable=True).indices
    o2 = torch.sort(allt[o1, 0], stable=True).indices
    order = o1[o2]
    a, b = allt[order, 0], allt[order, 1]
    owner, local = owner[order], local[order]
    n = a.shape[0]
    if n == 0:
        return recs[:0].copy()
    glen = (b ^ _SIGN) >> 48 & 0xFFFF           # fragment length
    gq = ((b ^ _SIGN) >> 42) & 0x3F             # mapq
    idx = torch.arange(n, device=device)
    if params.remove_pcr_duplicates:
        new_group = torch.ones(n, dtype=torch.bool, device=device)
        new_group[1:] = (a[1:] != a[:-1]) | (glen[1:] != glen[:-1])
    else:
        new_group = torch.ones(n, dtype=torch.bool, device=device)
    gid = torch.cumsum(new_group.to(torch.int64), 0) - 1
    n_groups = int(gid[-1].item()) + 1
    gsize = torch.zeros(n_groups, dtype=torch.int64, device=device).scatter_add_(0, gid, torch.ones(n, dtype=torch.int64, device=device))
    # kept record = first one, in sort order, carrying the group's maximum MAPQ (mapping_writer.h:268-270): records of a
    # group are sorted by MAPQ, so it is the start of the group's last (mapq) run
    new_run = new_group.clone()
    new_run[1:] |= gq[1:] != gq[:-1]
    run_start = torch.where(new_run, idx, torch.zeros_like(idx))
    run_start = torch.cummax(run_start, 0).values         # start index of the run each element belongs to
    last_of_group = torch.ones(n, dtype=torch.bool, device=device)
    last_of_group[:-1] = new_group[1:]
    kept_idx = run_start[last_of_group]                    # one per group, in group order
    keep_q = gq[kept_idx] >= params.mapq_threshold
    mine = (owner[kept_idx] == rank) & keep_q
    sel_local = local[kept_idx][mine].cpu().numpy()
    dups = torch.clamp(gsize, max=255)[mine].cpu().numpy().astype(np.uint8)
    out = recs[sel_local].copy()
    out["num_dups"] = dups
    return out


def gather_and_finish(survivors, params, group=None, dst=0):
    """Gather every rank's survivors on `dst`, put them in output order and apply the Tn5 shift (after sorting and
    dedup, as the low-memory merge does, mapping_writer.h:285-287).  Returns the final records on dst, None elsewhere."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    objs = [None] * world if rank == dst else None
    dist.gather_object(survivors.tobytes(), objs, dst=dst, group=group)
    if rank != dst:
        return None
    recs = np.concatenate([np.frombuffer(o, dtype=PE_RECORD) for o in objs]) if objs else survivors[:0]
    t = pack_tuples(recs)
    o1 = torch.sort(t[:, 1], stable=True).indices
    order = o1[torch.sort(t[o1, 0], stable=True).indices].numpy()
    recs = recs[order].copy()
    if params.tn5_shift:  # bed_mapping.h:225-230
        recs["fragment_start"] += 4
        recs["positive_alignment_length"] -= 4
        recs["fragment_length"] -= 9
        recs["negative_alignment_length"] -= 5
    return recs
