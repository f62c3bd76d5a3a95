"""Test double of the duplicate-removal exchange (csrc/exchange.cuh) in numpy + torch.distributed, so that the host-side
logic of the multi-GPU path — sharding, tuple layout, the decision rule, the final ordering — can be checked with `gloo` on
machines without a GPU.  Test infrastructure only: the product path is cmx_dedup_exchange (ncclAllGather + CUDA kernels)."""
import numpy as np
import torch
import torch.distributed as dist

_SIGN = -(1 << 63)


def pack_tuples(recs):
    """exchange.cuh's 16-byte tuple: a = rid << 32 | start, b = length << 48 | mapq << 40 | direction << 36 | unique << 32 | read_id."""
    a = (recs["rid"].astype(np.int64) << 32) | recs["fragment_start"].astype(np.int64)
    b = (recs["fragment_length"].astype(np.uint64) << np.uint64(48)) | (recs["mapq"].astype(np.uint64) << np.uint64(40)) | \
        (recs["direction"].astype(np.uint64) << np.uint64(36)) | (recs["is_unique"].astype(np.uint64) << np.uint64(32)) | recs["read_id"].astype(np.uint64)
    b = (b ^ np.uint64(1 << 63)).view(np.int64)  # flip the sign bit: signed order == unsigned order
    return torch.from_numpy(np.stack([a, b], axis=1))


def all_gather_padded(t, group=None):
    """the exchange's two collectives: counts (8 bytes per rank), then equally padded tuples"""
    world = dist.get_world_size(group)
    n = torch.tensor([t.shape[0]], dtype=torch.int64)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    m = max(max(sizes), 1)
    pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype)
    pad[:t.shape[0]] = t
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return [o[:s] for o, s in zip(out, sizes)], sizes


def dedup_exchange(recs, params, group=None):
    """this rank's survivors, in the reference's order, num_dups set, MAPQ-filtered, before the Tn5 shift"""
    rank = dist.get_rank(group)
    parts, sizes = all_gather_padded(pack_tuples(recs), group)
    allt = torch.cat(parts, dim=0)
    owner = torch.cat([torch.full((s,), r, dtype=torch.int64) for r, s in enumerate(sizes)])
    local = torch.cat([torch.arange(s, dtype=torch.int64) for s in sizes])
    o1 = torch.sort(allt[:, 1], stable=True).indices
    order = o1[torch.sort(allt[o1, 0], stable=True).indices]
    a, b = allt[order, 0], allt[order, 1]
    owner, local = owner[order], local[order]
    n = a.shape[0]
    if n == 0:
        return recs[:0].copy()
    glen = ((b ^ _SIGN) >> 48) & 0xFFFF
    gq = ((b ^ _SIGN) >> 40) & 0xFF
    idx = torch.arange(n)
    new_group = torch.ones(n, dtype=torch.bool)
    if params.remove_pcr_duplicates:
        new_group[1:] = (a[1:] != a[:-1]) | (glen[1:] != glen[:-1])
    gid = torch.cumsum(new_group.to(torch.int64), 0) - 1
    gsize = torch.zeros(int(gid[-1]) + 1, dtype=torch.int64).scatter_add_(0, gid, torch.ones(n, dtype=torch.int64))
    new_run = new_group.clone()
    new_run[1:] |= gq[1:] != gq[:-1]
    run_start = torch.cummax(torch.where(new_run, idx, torch.zeros_like(idx)), 0).values
    last_of_group = torch.ones(n, dtype=torch.bool)
    last_of_group[:-1] = new_group[1:]
    kept = run_start[last_of_group]  # first record of the group's highest-MAPQ run (mapping_writer.h:268-270)
    mine = (owner[kept] == rank) & (gq[kept] >= params.mapq_threshold)
    out = recs[local[kept][mine].numpy()].copy()
    if params.remove_pcr_duplicates:
        out["num_dups"] = torch.clamp(gsize, max=255)[mine].numpy().astype(np.uint8)
    return out


SH_SAMPLE = 4096
_PAD = np.uint64(0xFFFFFFFFFFFFFFFF)


def shuffle_partition(recs, group=None):
    """exchange.cuh's range partition: splitters from an all-gathered sample of SH_SAMPLE keys per rank, destination of a
    record = number of splitters <= its key word a = rid << 32 | fragment_start.  Returns (dest per record, splitters)."""
    world = dist.get_world_size(group)
    a = (recs["rid"].astype(np.uint64) << np.uint64(32)) | recs["fragment_start"].astype(np.uint64)
    n = len(a)
    take = min(n, SH_SAMPLE)
    samp = np.full(SH_SAMPLE, _PAD, dtype=np.uint64)
    if take:
        samp[:take] = a[(np.arange(take, dtype=np.uint64) * np.uint64(n)) // np.uint64(take)]
    t = torch.from_numpy(samp.view(np.int64))
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    alls = np.sort(torch.cat(out).numpy().view(np.uint64))
    valid = int(np.searchsorted(alls, _PAD, side="left"))
    split = np.array([alls[valid * j // world] if valid else 0 for j in range(1, world)], dtype=np.uint64)
    return np.searchsorted(split, a, side="right").astype(np.int64), split


def dedup_shuffle(recs, params, postprocess, group=None):
    """Test double of cmx_dedup_shuffle: records travel to the rank that owns their key range, `postprocess` (the
    single-process low-memory post-processing) runs there.  The ranks' results in rank order are the run's output."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dest, _ = shuffle_partition(recs, group)
    order = np.argsort(dest, kind="stable")
    parts = [recs[order][dest[order] == d] for d in range(world)]
    got = [None] * world
    for src in range(world):  # the grouped send / recv, spelled as one scatter per source rank
        box = [None]
        dist.scatter_object_list(box, [p.tobytes() for p in parts] if rank == src else None, src=src, group=group)
        got[src] = np.frombuffer(box[0], dtype=recs.dtype)
    mine = np.concatenate(got) if got else recs[:0]
    return postprocess(mine)
