"""Multi-GPU plumbing: one process per GPU; test users sharded with no data-path collective, or songs partitioned with one all-to-all.

The reference distributes the same way (distributed.scala:450-452: `ctx.parallelize(testUsers, slices).map(getRanks1).collect`):
every worker holds the whole train set (there: the task closure; here: a train replica in each GPU's HBM) and scores a
contiguous range of test users.  Scored pairs are independent, so the only exchange is the `collect` — here an all-gather of the
fixed-size top-k blocks (k * 12 bytes per test user) over NCCL.  The index-dependent blends (Aggregation MR:372-382, Stochastic
MR:447) need each shard's position in the global (user, song)-sorted pair list; `pair_index_bases` computes it from the CSR alone.

Works with any torch.distributed backend: `nccl` with CUDA tensors on the GPU box, `gloo` with CPU tensors in the tests.
"""
from __future__ import annotations

import numpy as np

from .dataset import Dataset


def shard_range(n_users: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced test-user range of `rank` (first n_users % world ranks get one extra user)."""
    base, rem = divmod(n_users, world)
    u0 = rank * base + min(rank, rem)
    return u0, u0 + base + (1 if rank < rem else 0)


def pair_index_bases(ds: Dataset, world: int) -> tuple[np.ndarray, int]:
    """(index of each rank's first scored pair in the global MAIN:57-59 order, total number of scored pairs)."""
    per_user = ds.S - np.diff(ds.te_ptr)                       # unlistened songs per test user (MR:109)
    prefix = np.concatenate([[0], np.cumsum(per_user)])
    starts = np.array([prefix[shard_range(ds.U, r, world)[0]] for r in range(world)], np.int64)
    return starts, int(prefix[-1])


_GATHER_BUFFERS: dict = {}


def gather_topk(song, score, length, n_users_total: int, world: int, rank: int, group=None, reuse_buffers: bool = False):
    """All-gather the per-shard top-k blocks into the full [U, k] result (the reference's `.collect`).  Accepts numpy arrays or
    torch tensors (CPU for gloo, CUDA for nccl); returns torch tensors on the same device.  Equal shards (the usual case) are gathered
    by one `all_gather_into_tensor` per array straight into the result; with reuse_buffers=True that result lives in a per-shape
    buffer that the next call overwrites (a steady-state scoring loop then allocates nothing)."""
    import torch
    import torch.distributed as dist
    song = torch.as_tensor(song)
    score = torch.as_tensor(score)
    length = torch.as_tensor(length)
    if world == 1:
        return song, score, length
    sizes = [shard_range(n_users_total, r, world) for r in range(world)]
    max_n = max(b - a for a, b in sizes)
    equal = all(b - a == max_n for a, b in sizes)

    def pad(t, fill):
        if t.shape[0] == max_n:
            return t.contiguous()
        extra = torch.full((max_n - t.shape[0],) + tuple(t.shape[1:]), fill, dtype=t.dtype, device=t.device)
        return torch.cat([t, extra]).contiguous()

    outs = []
    for t, fill in ((song, -1), (score, 0.0), (length, 0)):
        p = pad(t, fill)
        shape = (world * max_n,) + tuple(p.shape[1:])
        key = (str(p.device), p.dtype, shape)
        full = _GATHER_BUFFERS.get(key) if reuse_buffers else None
        if full is None:
            full = torch.empty(shape, dtype=p.dtype, device=p.device)
            if reuse_buffers:
                _GATHER_BUFFERS[key] = full
        dist.all_gather_into_tensor(full, p, group=group)
        if equal:
            outs.append(full)
        else:
            outs.append(torch.cat([full[r * max_n: r * max_n + (hi - lo)] for r, (lo, hi) in enumerate(sizes)]))
    return tuple(outs)


_PACKED: dict = {}


def gather_topk_packed(mr, k: int, world: int, rank: int, lib_stream, comm_stream, dst: int = 0):
    """The reference's `.collect` (distributed.scala:451-452: results go to the driver only) as ONE NCCL gather of the packed
    (song | score | len) block of every rank's shard to rank `dst`, issued on `comm_stream` so that it overlaps whatever the handle
    computes next.  The block is first copied device-to-device into a staging tensor (the library reuses its result block for the
    next model); the library stream waits for that copy, nothing waits for the gather until the caller joins comm_stream.
    Returns the list of per-rank uint8 blocks on rank dst (None elsewhere); equal shard sizes are required (pad the shards otherwise)."""
    import torch
    import torch.distributed as dist
    block, _, _ = mr.topk_packed_tensor(k)
    key = (block.numel(), world, rank)
    st = _PACKED.get(key)
    if st is None:
        recv = [[torch.empty(block.numel(), dtype=torch.uint8, device=block.device) for _ in range(world)] for _ in range(2)] if rank == dst else None
        st = _PACKED[key] = {"stage": torch.empty_like(block), "recv": recv, "turn": 0, "copied": torch.cuda.Event()}
    comm_stream.wait_stream(lib_stream)
    with torch.cuda.stream(comm_stream):
        st["stage"].copy_(block, non_blocking=True)
        st["copied"].record(comm_stream)
        bufs = None
        if rank == dst:
            bufs = st["recv"][st["turn"]]
            st["turn"] ^= 1
        dist.gather(st["stage"], gather_list=bufs, dst=dst)
    lib_stream.wait_event(st["copied"])
    return bufs


def song_window(n_songs: int, rank: int, world: int) -> tuple[int, int]:
    """Song partition of `rank` (the reference's second partitioning, distributed.scala:459-461: `ctx.parallelize(songs, n)`): a contiguous,
    balanced range of song ids.  A handle created with this window scores ALL test users against its songs only."""
    return shard_range(n_songs, rank, world)


_EXCHANGE: dict = {}


def exchange_song_partitions(song, score, length, n_users_total: int, world: int, rank: int, group=None):
    """Song partitioning's exchange step: every rank holds, for ALL test users, the ranked list of its own song window; afterwards rank r
    holds all `world` lists of the users of ITS test-user range (shard_range): one all-to-all per array (row blocks of unequal size
    allowed).  Returns (parts_song [world, n_r, k], parts_score [world, n_r, k], parts_len [world, n_r]); the receive buffers are reused
    by the next call with the same shapes.  Tensors live where the backend needs them (CUDA for nccl, CPU for gloo)."""
    import torch
    import torch.distributed as dist
    song, score, length = torch.as_tensor(song), torch.as_tensor(score), torch.as_tensor(length)
    ranges = [shard_range(n_users_total, r, world) for r in range(world)]
    n_mine = ranges[rank][1] - ranges[rank][0]
    send_rows = [b - a for a, b in ranges]
    outs = []
    for t in (song, score, length):
        shape = (world, n_mine) + tuple(t.shape[1:])
        key = (str(t.device), t.dtype, shape)
        recv = _EXCHANGE.get(key)
        if recv is None:
            recv = _EXCHANGE[key] = torch.empty(shape, dtype=t.dtype, device=t.device)
        if world == 1:
            recv[0].copy_(t)
        else:
            dist.all_to_all_single(recv.view((world * n_mine,) + tuple(t.shape[1:])), t.contiguous(), [n_mine] * world, send_rows, group=group)
        outs.append(recv)
    return tuple(outs)


def split_train_users(ds: Dataset, rank: int, world: int) -> Dataset:
    """K-split of the item-item Gram (BASELINE configs[4]): rank r keeps the train users of its contiguous range, every song.
    deg_song stays global — the cosine denominators count all listeners (MR:237)."""
    v0, v1 = shard_range(ds.T, rank, world)
    a, b = int(ds.tr_ptr[v0]), int(ds.tr_ptr[v1])
    return Dataset(v1 - v0, ds.U, ds.S, (ds.tr_ptr[v0:v1 + 1] - a).astype(np.int64), ds.tr_col[a:b], ds.te_ptr, ds.te_col, ds.lab_ptr,
                   ds.lab_col, ds.deg_tr[v0:v1], ds.deg_te, ds.deg_song, None, ds.test_users, ds.songs, dict(ds.meta, train_shard=(v0, v1)))


def reduce_scatter_rows(partial, world: int, rank: int, group=None):
    """Sum the ranks' partial int32 panels [n, ld] and leave rows [rank*n/world, (rank+1)*n/world) on each rank: one NCCL
    reduce-scatter (integer sums: bit-exact in any reduction order).  n must be a multiple of world.  gloo (CPU tests) has no
    reduce-scatter, so it all-reduces and slices."""
    import torch
    import torch.distributed as dist
    n = partial.shape[0]
    assert n % world == 0
    if world == 1:
        return partial
    if dist.get_backend(group) == "gloo":
        full = partial.clone()
        dist.all_reduce(full, group=group)
        return full[rank * (n // world):(rank + 1) * (n // world)]
    out = torch.empty((n // world, partial.shape[1]), dtype=partial.dtype, device=partial.device)
    dist.reduce_scatter_tensor(out, partial.contiguous(), group=group)
    return out
